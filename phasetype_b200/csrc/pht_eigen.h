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

/* On the device the solver is run by ONE WARP in lock step (k_spectral_solve defines PHT_EIGEN_WARP): every lane follows
 * the same control flow on the same shared-memory data -- the scalar recurrences of the algorithm are computed redundantly
 * by all 32 lanes -- and the loops whose iterations are independent (one row, one column, one diagonal element each) are
 * dealt out over the lanes.  Every element still sees exactly the operations, in exactly the order, of the sequential
 * code, so the result is the host's, bit for bit (tests/test_model_gpu.py).  Rules that keep the lanes consistent:
 * shared data is written either inside a dealt-out loop or by lane 0 alone (PHT_E_ONE), with a warp barrier between
 * any write and the reads that depend on it (PHT_E_SYNC), and loop bodies use their own temporaries, never the uniform
 * scalars.  On the host the three macros are the plain loop, nothing and nothing. */
#if defined(__CUDA_ARCH__) && defined(PHT_EIGEN_WARP)
#define PHT_E_FOR(v, lo, hi) for (int v = (lo) + (int)(threadIdx.x & 31u); v <= (hi); v += 32)
#define PHT_E_SYNC() __syncwarp()
#define PHT_E_ONE if ((threadIdx.x & 31u) == 0u)
#else
#define PHT_E_FOR(v, lo, hi) for (int v = (lo); v <= (hi); v++)
#define PHT_E_SYNC() do { } while (0)
#define PHT_E_ONE
#endif

PHT_HD int pht_eigen_real(int nn, const double *S, double *evals, double *Q, double *Qinv,
                          double *H, double *V, double *ort, double *d, double *e) {
    const double eps = 2.220446049250313e-16;
    int status = 0;
    const int low = 0, high = nn - 1;
#define HH(i, j) H[(i) * nn + (j)]
#define VV(i, j) V[(i) * nn + (j)]
    PHT_E_FOR(i, 0, nn - 1) for (int j = 0; j < nn; j++) HH(i, j) = S[i + j * nn];
    PHT_E_SYNC();

    /* ---- orthes: Householder reduction to upper Hessenberg form */
    for (int m = low + 1; m <= high - 1; m++) {
        double scale = 0.0;
        for (int i = m; i <= high; i++) scale = scale + PHT_EABS(HH(i, m - 1));
        if (scale != 0.0) {
            double h = 0.0;
            for (int i = high; i >= m; i--) { const double oi = HH(i, m - 1) / scale; h += oi * oi; }
            const double om0 = HH(m, m - 1) / scale;
            double g = PHT_ESQRT(h);
            if (om0 > 0) g = -g;
            h = h - om0 * g;
            const double om = om0 - g;
            PHT_E_SYNC();
            PHT_E_FOR(i, m, high) ort[i] = (i == m) ? om : HH(i, m - 1) / scale;
            PHT_E_SYNC();
            PHT_E_FOR(j, m, nn - 1) {
                double f = 0.0;
                for (int i = high; i >= m; i--) f += ort[i] * HH(i, j);
                f = f / h;
                for (int i = m; i <= high; i++) HH(i, j) -= f * ort[i];
            }
            PHT_E_SYNC();
            PHT_E_FOR(i, 0, high) {
                double f = 0.0;
                for (int j = high; j >= m; j--) f += ort[j] * HH(i, j);
                f = f / h;
                for (int j = m; j <= high; j++) HH(i, j) -= f * ort[j];
            }
            PHT_E_SYNC();
            PHT_E_ONE { ort[m] = scale * om; HH(m, m - 1) = scale * g; }
            PHT_E_SYNC();
        }
    }
    /* ---- ortran: accumulate the transformations */
    PHT_E_FOR(i, 0, nn - 1) for (int j = 0; j < nn; j++) VV(i, j) = (i == j) ? 1.0 : 0.0;
    PHT_E_SYNC();
    for (int m = high - 1; m >= low + 1; m--) {
        if (HH(m, m - 1) != 0.0) {
            PHT_E_FOR(i, m + 1, high) ort[i] = HH(i, m - 1);
            PHT_E_SYNC();
            PHT_E_FOR(j, m, high) {
                double g = 0.0;
                for (int i = m; i <= high; i++) g += ort[i] * VV(i, j);
                g = (g / ort[m]) / HH(m, m - 1);
                for (int i = m; i <= high; i++) VV(i, j) += g * ort[i];
            }
            PHT_E_SYNC();
        }
    }

    /* ---- hqr2: eigenvalues and Schur vectors by the shifted QR algorithm */
    int n = nn - 1;
    double exshift = 0.0, p = 0, q = 0, r = 0, s = 0, z = 0, w, x, y;
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
            const double hnn = HH(n, n) + exshift;
            PHT_E_SYNC();
            PHT_E_ONE { HH(n, n) = hnn; d[n] = hnn; e[n] = 0.0; }
            PHT_E_SYNC();
            n--; iter = 0;
        } else if (l == n - 1) {                        /* two roots */
            w = HH(n, n - 1) * HH(n - 1, n);
            p = (HH(n - 1, n - 1) - HH(n, n)) / 2.0;
            q = p * p + w;
            z = PHT_ESQRT(PHT_EABS(q));
            const double hnn = HH(n, n) + exshift, hn1 = HH(n - 1, n - 1) + exshift;
            PHT_E_SYNC();
            PHT_E_ONE { HH(n, n) = hnn; HH(n - 1, n - 1) = hn1; }
            x = hnn;
            if (q >= 0) {                               /* real pair */
                z = (p >= 0) ? p + z : p - z;
                const double d1 = x + z;
                double d2 = d1;
                if (z != 0.0) d2 = x - w / z;
                PHT_E_ONE { d[n - 1] = d1; d[n] = d2; e[n - 1] = 0.0; e[n] = 0.0; }
                x = HH(n, n - 1);
                s = PHT_EABS(x) + PHT_EABS(z);
                p = x / s; q = z / s;
                r = PHT_ESQRT(p * p + q * q);
                p = p / r; q = q / r;
                PHT_E_SYNC();
                PHT_E_FOR(j, n - 1, nn - 1) { const double zz = HH(n - 1, j); HH(n - 1, j) = q * zz + p * HH(n, j); HH(n, j) = q * HH(n, j) - p * zz; }
                PHT_E_SYNC();
                PHT_E_FOR(i, 0, n) { const double zz = HH(i, n - 1); HH(i, n - 1) = q * zz + p * HH(i, n); HH(i, n) = q * HH(i, n) - p * zz; }
                PHT_E_FOR(i, low, high) { const double zz = VV(i, n - 1); VV(i, n - 1) = q * zz + p * VV(i, n); VV(i, n) = q * VV(i, n) - p * zz; }
                PHT_E_SYNC();
            } else {                                    /* complex pair */
                PHT_E_ONE { d[n - 1] = x + p; d[n] = x + p; e[n - 1] = z; e[n] = -z; }
                PHT_E_SYNC();
                status |= 2;
            }
            n = n - 2; iter = 0;
        } else {                                        /* no convergence yet */
            x = HH(n, n); y = 0.0; w = 0.0;
            if (l < n) { y = HH(n - 1, n - 1); w = HH(n, n - 1) * HH(n - 1, n); }
            if (iter == 10) {                           /* Wilkinson's ad hoc shift */
                exshift += x;
                PHT_E_SYNC();
                PHT_E_FOR(i, low, n) HH(i, i) -= x;
                PHT_E_SYNC();
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
                    PHT_E_SYNC();
                    PHT_E_FOR(i, low, n) HH(i, i) -= s;
                    PHT_E_SYNC();
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
            PHT_E_SYNC();
            PHT_E_FOR(i, m + 2, n) { HH(i, i - 2) = 0.0; if (i > m + 2) HH(i, i - 3) = 0.0; }
            PHT_E_SYNC();
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
                    const double hk = (k != m) ? -s * x : -HH(k, k - 1);
                    PHT_E_SYNC();
                    PHT_E_ONE { if (k != m || l != m) HH(k, k - 1) = hk; }
                    p = p + s; x = p / s; y = q / s; z = r / s; q = q / p; r = r / p;
                    PHT_E_FOR(j, k, nn - 1) {
                        double pp = HH(k, j) + q * HH(k + 1, j);
                        if (notlast) { pp = pp + r * HH(k + 2, j); HH(k + 2, j) = HH(k + 2, j) - pp * z; }
                        HH(k, j) = HH(k, j) - pp * x;
                        HH(k + 1, j) = HH(k + 1, j) - pp * y;
                    }
                    PHT_E_SYNC();
                    const int imax = (n < k + 3) ? n : k + 3;
                    PHT_E_FOR(i, 0, imax) {
                        double pp = x * HH(i, k) + y * HH(i, k + 1);
                        if (notlast) { pp = pp + z * HH(i, k + 2); HH(i, k + 2) = HH(i, k + 2) - pp * r; }
                        HH(i, k) = HH(i, k) - pp;
                        HH(i, k + 1) = HH(i, k + 1) - pp * q;
                    }
                    PHT_E_FOR(i, low, high) {
                        double pp = x * VV(i, k) + y * VV(i, k + 1);
                        if (notlast) { pp = pp + z * VV(i, k + 2); VV(i, k + 2) = VV(i, k + 2) - pp * r; }
                        VV(i, k) = VV(i, k) - pp;
                        VV(i, k + 1) = VV(i, k + 1) - pp * q;
                    }
                    PHT_E_SYNC();
                }
            }
        }
    }

    /* ---- back-substitution: eigenvectors of the quasi-triangular form.  Column n's solve reads the triangle to its left,
     * which the solves of the columns to its left overwrite: inherently one after the other, so one lane does it (a few per
     * cent of the solver's work). */
    PHT_E_SYNC();
    if (norm != 0.0 && !(status & 1)) {
        PHT_E_ONE {
        double t;
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
        }
        PHT_E_SYNC();
        /* back transformation to the eigenvectors of the original matrix: row i of V only needs row i of V, and column j
         * only columns <= j, so with j running downwards inside the row the rows are independent */
        PHT_E_FOR(i, low, high) {
            for (int j = nn - 1; j >= low; j--) {
                double zz = 0.0;
                const int kmax = (j < high) ? j : high;
                for (int k = low; k <= kmax; k++) zz = zz + VV(i, k) * HH(k, j);
                VV(i, j) = zz;
            }
        }
        PHT_E_SYNC();
    }

    /* ---- outputs: unit-norm columns, then Q^-1 by Gauss-Jordan with partial pivoting (H is free now: a column of H is
     * overwritten only after the same lane has read it as V's column, and nobody else reads H here) */
    PHT_E_FOR(k, 0, nn - 1) {
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
    PHT_E_SYNC();
    PHT_E_FOR(i, 0, nn - 1) for (int j = 0; j < nn; j++) VV(i, j) = (i == j) ? 1.0 : 0.0;       /* V becomes the inverse */
    PHT_E_SYNC();
    for (int c = 0; c < nn; c++) {
        int piv = c; double best = PHT_EABS(HH(c, c));
        for (int i = c + 1; i < nn; i++) { const double a = PHT_EABS(HH(i, c)); if (a > best) { best = a; piv = i; } }
        if (!(best > 0.0)) { status |= 4; continue; }
        PHT_E_SYNC();
        if (piv != c) {
            PHT_E_FOR(j, 0, nn - 1) {
                double tmp = HH(c, j); HH(c, j) = HH(piv, j); HH(piv, j) = tmp;
                tmp = VV(c, j); VV(c, j) = VV(piv, j); VV(piv, j) = tmp;
            }
            PHT_E_SYNC();
        }
        const double dinv = 1.0 / HH(c, c);
        PHT_E_SYNC();
        PHT_E_FOR(j, 0, nn - 1) { HH(c, j) = HH(c, j) * dinv; VV(c, j) = VV(c, j) * dinv; }
        PHT_E_SYNC();
        PHT_E_FOR(i, 0, nn - 1) {
            if (i == c) continue;
            const double f = HH(i, c);
            if (f != 0.0) for (int j = 0; j < nn; j++) { HH(i, j) -= f * HH(c, j); VV(i, j) -= f * VV(c, j); }
        }
        PHT_E_SYNC();
    }
    PHT_E_FOR(i, 0, nn - 1) for (int j = 0; j < nn; j++) Qinv[i + j * nn] = VV(i, j);
    PHT_E_SYNC();
#undef HH
#undef VV
    return status;
}

#endif /* PHT_EIGEN_H */
