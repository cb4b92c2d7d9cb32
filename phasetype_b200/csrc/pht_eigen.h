/*
 * pht_eigen.h -- spectral decomposition S = Q diag(evals) Q^-1 of a small real matrix, written once for the
 * device (k_model.cu: k_spectral_solve) and the host checkers.  Real spectrum: exactly that.  Complex pairs
 * a +- ib: the REAL block form S = Q B Q^-1, where the two columns of Q that belong to the pair hold the real and
 * the imaginary part of its eigenvector and B has the 2x2 block [[a, b], [-b, a]] there, so that
 * exp(xS) = Q exp(xB) Q^-1 with exp(x [[a, b], [-b, a]]) = e^{ax} [[cos bx, sin bx], [-sin bx, cos bx]] --
 * everything stays real arithmetic (evals[] holds a twice, the work array e[] holds +b, -b on return).
 *
 * Replaces, for the engine, what the reference obtains from LAPACK dgeevx + dgetrf/dgetri
 * (src/utility.c:87-129 via LJMA_eigen / LJMA_inverse).  LAPACK is a third-party dependency of the
 * reference, not part of its tree; this is the classical EISPACK route instead -- Householder reduction to
 * Hessenberg form (orthes/ortran), Francis double-shift QR with accumulation and back-substitution for the
 * eigenvectors (hqr2), then Gauss-Jordan with partial pivoting for Q^-1 -- in plain IEEE arithmetic with a
 * fixed operation order, so host and device produce identical bits.  Eigenvalue order and eigenvector
 * scaling differ from LAPACK's (columns are scaled to unit Euclidean norm, as dgeevx does, but signs and
 * order are the algorithm's own); everything downstream only uses Q f(evals) Q^-1, which is invariant.
 * No balancing step (dgeevx is called with balanc = 'B'): generator matrices are already well scaled.
 *
 * Returns 0 on success; bit 0 set if the QR iteration did not converge, bit 1 if a complex pair was found
 * (informational: the decomposition is then the block form above; the reference prints "Error: imaginary part of
 * eigenvalue" and carries on using the packed columns as if they were real eigenvectors, src/utility.c:118-121),
 * bit 2 if Q is numerically singular.
 *
 * Work space: H, V (n*n each, row-major), ort, d, e (n each).  S, Q, Qinv are column-major.
 */
#ifndef PHT_EIGEN_H
#define PHT_EIGEN_H

#include "pht_math.h"

#if defined(__CUDA_ARCH__)
#define PHT_ESQRT(x) __dsqrt_rn(x)
#else
#define PHT_ESQRT(x) __builtin_sqrt(x)
#endif
#define PHT_EABS(x) ((x) < 0.0 ? -(x) : (x))

/* (ar + i ai) / (br + i bi), Smith's scaling */
PHT_HD void pht_cdiv(double ar, double ai, double br, double bi, double *cr, double *ci) {
    if (PHT_EABS(br) > PHT_EABS(bi)) {
        const double r = bi / br, dd = br + r * bi;
        *cr = (ar + r * ai) / dd; *ci = (ai - r * ar) / dd;
    } else {
        const double r = br / bi, dd = bi + r * br;
        *cr = (r * ar + ai) / dd; *ci = (r * ai - ar) / dd;
    }
}

PHT_HD int pht_eigen_real(int nn, const double *S, double *evals, double *Q, double *Qinv,
                          double *H, double *V, double *ort, double *d, double *e) {
    const double eps = 2.220446049250313e-16;
    int status = 0;
    const int low = 0, high = nn - 1;
#define HH(i, j) H[(i) * nn + (j)]
#define VV(i, j) V[(i) * nn + (j)]
    for (int i = 0; i < nn; i++) for (int j = 0; j < nn; j++) HH(i, j) = S[i + j * nn];

    /* ---- orthes: Householder reduction to upper Hessenberg form */
    for (int m = low + 1; m <= high - 1; m++) {
        double scale = 0.0;
        for (int i = m; i <= high; i++) scale = scale + PHT_EABS(HH(i, m - 1));
        if (scale != 0.0) {
            double h = 0.0;
            for (int i = high; i >= m; i--) { ort[i] = HH(i, m - 1) / scale; h += ort[i] * ort[i]; }
            double g = PHT_ESQRT(h);
            if (ort[m] > 0) g = -g;
            h = h - ort[m] * g;
            ort[m] = ort[m] - g;
            for (int j = m; j < nn; j++) {
                double f = 0.0;
                for (int i = high; i >= m; i--) f += ort[i] * HH(i, j);
                f = f / h;
                for (int i = m; i <= high; i++) HH(i, j) -= f * ort[i];
            }
            for (int i = 0; i <= high; i++) {
                double f = 0.0;
                for (int j = high; j >= m; j--) f += ort[j] * HH(i, j);
                f = f / h;
                for (int j = m; j <= high; j++) HH(i, j) -= f * ort[j];
            }
            ort[m] = scale * ort[m];
            HH(m, m - 1) = scale * g;
        }
    }
    /* ---- ortran: accumulate the transformations */
    for (int i = 0; i < nn; i++) for (int j = 0; j < nn; j++) VV(i, j) = (i == j) ? 1.0 : 0.0;
    for (int m = high - 1; m >= low + 1; m--) {
        if (HH(m, m - 1) != 0.0) {
            for (int i = m + 1; i <= high; i++) ort[i] = HH(i, m - 1);
            for (int j = m; j <= high; j++) {
                double g = 0.0;
                for (int i = m; i <= high; i++) g += ort[i] * VV(i, j);
                g = (g / ort[m]) / HH(m, m - 1);
                for (int i = m; i <= high; i++) VV(i, j) += g * ort[i];
            }
        }
    }

    /* ---- hqr2: eigenvalues and Schur vectors by the shifted QR algorithm */
    int n = nn - 1;
    double exshift = 0.0, p = 0, q = 0, r = 0, s = 0, z = 0, t, w, x, y;
    double norm = 0.0;
    for (int i = 0; i < nn; i++)
        for (int j = (i - 1 > 0 ? i - 1 : 0); j < nn; j++) norm = norm + PHT_EABS(HH(i, j));
    int iter = 0, total_iter = 0;
    while (n >= low) {
        int l = n;
        while (l > low) {
            s = PHT_EABS(HH(l - 1, l - 1)) + PHT_EABS(HH(l, l));
            if (s == 0.0) s = norm;
            if (PHT_EABS(HH(l, l - 1)) < eps * s) break;
            l--;
        }
        if (l == n) {                                   /* one root */
            HH(n, n) = HH(n, n) + exshift;
            d[n] = HH(n, n); e[n] = 0.0;
            n--; iter = 0;
        } else if (l == n - 1) {                        /* two roots */
            w = HH(n, n - 1) * HH(n - 1, n);
            p = (HH(n - 1, n - 1) - HH(n, n)) / 2.0;
            q = p * p + w;
            z = PHT_ESQRT(PHT_EABS(q));
            HH(n, n) = HH(n, n) + exshift;
            HH(n - 1, n - 1) = HH(n - 1, n - 1) + exshift;
            x = HH(n, n);
            if (q >= 0) {                               /* real pair */
                z = (p >= 0) ? p + z : p - z;
                d[n - 1] = x + z;
                d[n] = d[n - 1];
                if (z != 0.0) d[n] = x - w / z;
                e[n - 1] = 0.0; e[n] = 0.0;
                x = HH(n, n - 1);
                s = PHT_EABS(x) + PHT_EABS(z);
                p = x / s; q = z / s;
                r = PHT_ESQRT(p * p + q * q);
                p = p / r; q = q / r;
                for (int j = n - 1; j < nn; j++) { z = HH(n - 1, j); HH(n - 1, j) = q * z + p * HH(n, j); HH(n, j) = q * HH(n, j) - p * z; }
                for (int i = 0; i <= n; i++) { z = HH(i, n - 1); HH(i, n - 1) = q * z + p * HH(i, n); HH(i, n) = q * HH(i, n) - p * z; }
                for (int i = low; i <= high; i++) { z = VV(i, n - 1); VV(i, n - 1) = q * z + p * VV(i, n); VV(i, n) = q * VV(i, n) - p * z; }
            } else {                                    /* complex pair */
                d[n - 1] = x + p; d[n] = x + p; e[n - 1] = z; e[n] = -z;
                status |= 2;
            }
            n = n - 2; iter = 0;
        } else {                                        /* no convergence yet */
            x = HH(n, n); y = 0.0; w = 0.0;
            if (l < n) { y = HH(n - 1, n - 1); w = HH(n, n - 1) * HH(n - 1, n); }
            if (iter == 10) {                           /* Wilkinson's ad hoc shift */
                exshift += x;
                for (int i = low; i <= n; i++) HH(i, i) -= x;
                s = PHT_EABS(HH(n, n - 1)) + PHT_EABS(HH(n - 1, n - 2));
                x = y = 0.75 * s;
                w = -0.4375 * s * s;
            }
            if (iter == 30) {                           /* second ad hoc shift */
                s = (y - x) / 2.0;
                s = s * s + w;
                if (s > 0) {
                    s = PHT_ESQRT(s);
                    if (y < x) s = -s;
                    s = x - w / ((y - x) / 2.0 + s);
                    for (int i = low; i <= n; i++) HH(i, i) -= s;
                    exshift += s;
                    x = y = w = 0.964;
                }
            }
            iter = iter + 1; total_iter++;
            if (total_iter > 60 * nn) { status |= 1; break; }
            int m = n - 2;
            while (m >= l) {
                z = HH(m, m);
                r = x - z; s = y - z;
                p = (r * s - w) / HH(m + 1, m) + HH(m, m + 1);
                q = HH(m + 1, m + 1) - z - r - s;
                r = HH(m + 2, m + 1);
                s = PHT_EABS(p) + PHT_EABS(q) + PHT_EABS(r);
                p = p / s; q = q / s; r = r / s;
                if (m == l) break;
                if (PHT_EABS(HH(m, m - 1)) * (PHT_EABS(q) + PHT_EABS(r)) <
                    eps * (PHT_EABS(p) * (PHT_EABS(HH(m - 1, m - 1)) + PHT_EABS(z) + PHT_EABS(HH(m + 1, m + 1))))) break;
                m--;
            }
            for (int i = m + 2; i <= n; i++) { HH(i, i - 2) = 0.0; if (i > m + 2) HH(i, i - 3) = 0.0; }
            for (int k = m; k <= n - 1; k++) {          /* double QR step on rows l:n, columns m:n */
                const int notlast = (k != n - 1);
                if (k != m) {
                    p = HH(k, k - 1); q = HH(k + 1, k - 1); r = notlast ? HH(k + 2, k - 1) : 0.0;
                    x = PHT_EABS(p) + PHT_EABS(q) + PHT_EABS(r);
                    if (x == 0.0) continue;
                    p = p / x; q = q / x; r = r / x;
                }
                s = PHT_ESQRT(p * p + q * q + r * r);
                if (p < 0) s = -s;
                if (s != 0) {
                    if (k != m) HH(k, k - 1) = -s * x;
                    else if (l != m) HH(k, k - 1) = -HH(k, k - 1);
                    p = p + s; x = p / s; y = q / s; z = r / s; q = q / p; r = r / p;
                    for (int j = k; j < nn; j++) {
                        p = HH(k, j) + q * HH(k + 1, j);
                        if (notlast) { p = p + r * HH(k + 2, j); HH(k + 2, j) = HH(k + 2, j) - p * z; }
                        HH(k, j) = HH(k, j) - p * x;
                        HH(k + 1, j) = HH(k + 1, j) - p * y;
                    }
                    const int imax = (n < k + 3) ? n : k + 3;
                    for (int i = 0; i <= imax; i++) {
                        p = x * HH(i, k) + y * HH(i, k + 1);
                        if (notlast) { p = p + z * HH(i, k + 2); HH(i, k + 2) = HH(i, k + 2) - p * r; }
                        HH(i, k) = HH(i, k) - p;
                        HH(i, k + 1) = HH(i, k + 1) - p * q;
                    }
                    for (int i = low; i <= high; i++) {
                        p = x * VV(i, k) + y * VV(i, k + 1);
                        if (notlast) { p = p + z * VV(i, k + 2); VV(i, k + 2) = VV(i, k + 2) - p * r; }
                        VV(i, k) = VV(i, k) - p;
                        VV(i, k + 1) = VV(i, k + 1) - p * q;
                    }
                }
            }
        }
    }

    /* ---- back-substitution: eigenvectors of the quasi-triangular form */
    if (norm != 0.0 && !(status & 1)) {
        for (n = nn - 1; n >= 0; n--) {
            p = d[n]; q = e[n];
            if (q > 0.0) continue;                      /* first column of a complex pair: done together with the second */
            if (q < 0.0) {
                /* complex pair in rows/columns n-1, n: column n-1 receives the real, column n the imaginary part of the
                 * eigenvector of d + i e[n-1] (the hqr2 back-substitution for a complex vector) */
                int l = n - 1;
                double cr, ci, ra, sa, vr, vi;
                if (PHT_EABS(HH(n, n - 1)) > PHT_EABS(HH(n - 1, n))) {
                    HH(n - 1, n - 1) = q / HH(n, n - 1);
                    HH(n - 1, n) = -(HH(n, n) - p) / HH(n, n - 1);
                } else {
                    pht_cdiv(0.0, -HH(n - 1, n), HH(n - 1, n - 1) - p, q, &cr, &ci);
                    HH(n - 1, n - 1) = cr; HH(n - 1, n) = ci;
                }
                HH(n, n - 1) = 0.0; HH(n, n) = 1.0;
                for (int i = n - 2; i >= 0; i--) {
                    ra = 0.0; sa = 0.0;
                    for (int j = l; j <= n; j++) { ra = ra + HH(i, j) * HH(j, n - 1); sa = sa + HH(i, j) * HH(j, n); }
                    w = HH(i, i) - p;
                    if (e[i] < 0.0) { z = w; r = ra; s = sa; }
                    else {
                        l = i;
                        if (e[i] == 0.0) {
                            pht_cdiv(-ra, -sa, w, q, &cr, &ci);
                            HH(i, n - 1) = cr; HH(i, n) = ci;
                        } else {
                            x = HH(i, i + 1); y = HH(i + 1, i);
                            vr = (d[i] - p) * (d[i] - p) + e[i] * e[i] - q * q;
                            vi = (d[i] - p) * 2.0 * q;
                            if (vr == 0.0 && vi == 0.0) vr = eps * norm * (PHT_EABS(w) + PHT_EABS(q) + PHT_EABS(x) + PHT_EABS(y) + PHT_EABS(z));
                            pht_cdiv(x * r - z * ra + q * sa, x * s - z * sa - q * ra, vr, vi, &cr, &ci);
                            HH(i, n - 1) = cr; HH(i, n) = ci;
                            if (PHT_EABS(x) > PHT_EABS(z) + PHT_EABS(q)) {
                                HH(i + 1, n - 1) = (-ra - w * HH(i, n - 1) + q * HH(i, n)) / x;
                                HH(i + 1, n) = (-sa - w * HH(i, n) - q * HH(i, n - 1)) / x;
                            } else {
                                pht_cdiv(-r - y * HH(i, n - 1), -s - y * HH(i, n), z, q, &cr, &ci);
                                HH(i + 1, n - 1) = cr; HH(i + 1, n) = ci;
                            }
                        }
                        t = PHT_EABS(HH(i, n - 1)); if (PHT_EABS(HH(i, n)) > t) t = PHT_EABS(HH(i, n));
                        if ((eps * t) * t > 1) for (int j = i; j <= n; j++) { HH(j, n - 1) = HH(j, n - 1) / t; HH(j, n) = HH(j, n) / t; }
                    }
                }
                continue;
            }
            int l = n;
            HH(n, n) = 1.0;
            for (int i = n - 1; i >= 0; i--) {
                w = HH(i, i) - p;
                r = 0.0;
                for (int j = l; j <= n; j++) r = r + HH(i, j) * HH(j, n);
                if (e[i] < 0.0) { z = w; s = r; }
                else {
                    l = i;
                    if (e[i] == 0.0) {
                        if (w != 0.0) HH(i, n) = -r / w; else HH(i, n) = -r / (eps * norm);
                    } else {
                        x = HH(i, i + 1); y = HH(i + 1, i);
                        q = (d[i] - p) * (d[i] - p) + e[i] * e[i];
                        t = (x * s - z * r) / q;
                        HH(i, n) = t;
                        if (PHT_EABS(x) > PHT_EABS(z)) HH(i + 1, n) = (-r - w * t) / x; else HH(i + 1, n) = (-s - y * t) / z;
                    }
                    t = PHT_EABS(HH(i, n));
                    if ((eps * t) * t > 1) for (int j = i; j <= n; j++) HH(j, n) = HH(j, n) / t;
                }
            }
        }
        /* back transformation to the eigenvectors of the original matrix */
        for (int j = nn - 1; j >= low; j--)
            for (int i = low; i <= high; i++) {
                z = 0.0;
                const int kmax = (j < high) ? j : high;
                for (int k = low; k <= kmax; k++) z = z + VV(i, k) * HH(k, j);
                VV(i, j) = z;
            }
    }

    /* ---- outputs: unit-norm columns, then Q^-1 by Gauss-Jordan with partial pivoting (H is free now) */
    for (int k = 0; k < nn; k++) {
        double nrm = 0.0;
        for (int i = 0; i < nn; i++) nrm += VV(i, k) * VV(i, k);
        /* the two columns of a complex pair are one complex vector: one common factor, or S [p q] = [p q] B breaks */
        if (e[k] > 0.0 && k + 1 < nn) for (int i = 0; i < nn; i++) nrm += VV(i, k + 1) * VV(i, k + 1);
        if (e[k] < 0.0 && k > 0) for (int i = 0; i < nn; i++) nrm += VV(i, k - 1) * VV(i, k - 1);
        nrm = PHT_ESQRT(nrm);
        if (!(nrm > 0.0)) nrm = 1.0;
        evals[k] = d[k];
        for (int i = 0; i < nn; i++) { const double v = VV(i, k) / nrm; Q[i + k * nn] = v; HH(i, k) = v; }
    }
    for (int i = 0; i < nn; i++) for (int j = 0; j < nn; j++) VV(i, j) = (i == j) ? 1.0 : 0.0;       /* V becomes the inverse */
    for (int c = 0; c < nn; c++) {
        int piv = c; double best = PHT_EABS(HH(c, c));
        for (int i = c + 1; i < nn; i++) { const double a = PHT_EABS(HH(i, c)); if (a > best) { best = a; piv = i; } }
        if (!(best > 0.0)) { status |= 4; continue; }
        if (piv != c) for (int j = 0; j < nn; j++) {
            double tmp = HH(c, j); HH(c, j) = HH(piv, j); HH(piv, j) = tmp;
            tmp = VV(c, j); VV(c, j) = VV(piv, j); VV(piv, j) = tmp;
        }
        const double dinv = 1.0 / HH(c, c);
        for (int j = 0; j < nn; j++) { HH(c, j) = HH(c, j) * dinv; VV(c, j) = VV(c, j) * dinv; }
        for (int i = 0; i < nn; i++) {
            if (i == c) continue;
            const double f = HH(i, c);
            if (f != 0.0) for (int j = 0; j < nn; j++) { HH(i, j) -= f * HH(c, j); VV(i, j) -= f * VV(c, j); }
        }
    }
    for (int i = 0; i < nn; i++) for (int j = 0; j < nn; j++) Qinv[i + j * nn] = VV(i, j);
#undef HH
#undef VV
    return status;
}

#endif /* PHT_EIGEN_H */
