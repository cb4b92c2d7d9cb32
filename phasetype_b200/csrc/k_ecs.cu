/*
 * k_ecs.cu -- Aslett-Wilson exact conditional sampling (method bit 2) for sm_100a, two kernels:
 *
 *   k_ecs_exact  observations absorbed AT y (reference src/Simulate_AbsCTMC_eq_Aslett_ECS.c:205-373): at each
 *                step decide "absorb now" with probability exp(S_jj T) s_j / (e_j e^{ST} s), else draw the sojourn
 *                d in (0,T) by ARMS from log(p_j^T e^{S(T-d)} s) + S_jj d and the next state with weight
 *                P_ji (e^{S(T-d)} s)_i;
 *   k_ecs_gt     right-censored observations, absorbed AFTER y (the reference delegates them to the Aslett-DCS
 *                gt sampler, src/Simulate_AbsCTMC_gt_Aslett_DCS.c:299-418 with :184-260): before y the sojourn
 *                is either "jump beyond y" or an ARMS draw conditioned on survival, after y plain forward
 *                simulation to absorption.
 *
 * Both use a per-lane ARMS (Gilks' adaptive rejection Metropolis sampling, src/arms.c) with the reference's
 * settings (4 starting abscissae, up to 100 envelope points, convex = 1, Metropolis on, xprev = 0).  The
 * envelope is an index-linked array in the lane's local memory (typically 9-13 live points; a compact
 * shared-memory envelope was tried and lost to the occupancy it costs);
 * log-density evaluations read the lane's hoisted row vector p^T Q from a shared-memory slab and the spectrum
 * from shared memory.  Exact and censored observations are separate launches over index lists built at upload,
 * so a warp never serialises the two samplers.  All sums run in the reference's order inside one thread:
 * results are bit-identical to the host for the same spectral data.
 *
 * Roofline: FP64 issue bound (~6 density evaluations of n exp each per sojourn); 12 B of HBM per path.
 */
#include "path_common.cuh"

#define ECS_THREADS 128
#ifndef ECS_WARPS_PER_SM
#define ECS_WARPS_PER_SM 16          /* launch bound: resident warps per SM the register allocation must allow */
#endif
#define A_XEPS 0.00001
#define A_YEPS 0.1
#define A_EYEPS 0.001
#define A_YCEIL 50.
#define A_NPOINT 100
#define A_NIL (-1)

struct EcsSmem {
    double *S, *Q, *P, *Pfull, *evals, *evi, *s, *Qinv_s, *Qinv_1, *pi;
    int *cplx;                          /* the spectrum has complex pairs: block formulas (spec_* below) instead of the reference's */
    double *PQ, *W, *Z;                 /* per-lane slabs: hoisted p^T Q, scratch weights, sojourn totals */
    unsigned long long *zlo; long long *zhi; unsigned int *Nacc, *Bacc;
    __device__ __forceinline__ void carve(unsigned char *raw, int n) {
        double *d = reinterpret_cast<double *>(raw);
        S = d; d += n * n; Q = d; d += n * n; P = d; d += n * n; Pfull = d; d += n * (n + 1);
        evals = d; d += n; evi = d; d += n; s = d; d += n; Qinv_s = d; d += n; Qinv_1 = d; d += n; pi = d; d += n;
        PQ = d; d += n * ECS_THREADS; W = d; d += n * ECS_THREADS; Z = d; d += n * ECS_THREADS;
        zlo = reinterpret_cast<unsigned long long *>(d); d += n; zhi = reinterpret_cast<long long *>(d); d += n;
        Nacc = reinterpret_cast<unsigned int *>(d); Bacc = Nacc + n * n; cplx = reinterpret_cast<int *>(Bacc + n);
    }
    static size_t bytes(int n) {
        return sizeof(double) * (size_t)(3 * n * n + n * (n + 1) + 6 * n + 3 * n * ECS_THREADS + 2 * n) + sizeof(unsigned int) * (size_t)(n * n + n + 2);
    }
};

struct EcsCounters { unsigned long long jumps, evals, updates, calls, rejects, nonfinite, paths; };

/* ---- spectra with complex pairs (pht_eigen.h: real block form S = Q B Q^-1).  The reference takes the real parts
 * and carries on (src/utility.c:116-121), i.e. is silently wrong there; these two helpers are what its formulas
 * become when exp(x Lambda) is block diagonal: a pair a +- ib in positions (i, i+1) contributes
 * e^{ax} [[cos bx, sin bx], [-sin bx, cos bx]].  Checked against the analytic conditional expectations (tier 2). */
/* u^T exp(x B) v */
static __device__ __noinline__ double spec_bilinear(const double *u, int su, const double *v, const double *ev, const double *evi, int n, double x) {
    double acc = 0.0;
#pragma unroll 1
    for (int i = 0; i < n; i++) {
        const double b = evi[i], ex = pht_exp(ev[i] * x);
        if (b > 0.0 && i + 1 < n) {
            double sn, cs; sincos(b * x, &sn, &cs);
            const double u1 = u[i * su], u2 = u[(i + 1) * su], v1 = v[i], v2 = v[i + 1];
            acc += ex * (cs * (u1 * v1 + u2 * v2) + sn * (u1 * v2 - u2 * v1));
            i++;
        } else acc += (u[i * su] * ex) * v[i];
    }
    return acc;
}
/* out = exp(x B) v, out with stride so */
static __device__ __noinline__ void spec_apply(double *out, int so, const double *v, const double *ev, const double *evi, int n, double x) {
#pragma unroll 1
    for (int i = 0; i < n; i++) {
        const double b = evi[i], ex = pht_exp(ev[i] * x);
        if (b > 0.0 && i + 1 < n) {
            double sn, cs; sincos(b * x, &sn, &cs);
            const double v1 = v[i], v2 = v[i + 1];
            out[i * so] = ex * (cs * v1 + sn * v2); out[(i + 1) * so] = ex * (cs * v2 - sn * v1);
            i++;
        } else out[i * so] = ex * v[i];
    }
}

/* log-density closures.  The evaluation itself is ONE out-of-line function per sampler: ARMS calls its density from
 * five places, and five inlined copies of an n-term exp loop made the kernels' loop bodies larger than the
 * instruction cache (`no_instruction` was the top stall: profiles/r1a_ecs_ncu_full.md). */
static __device__ __noinline__ double ecs_dens_exact(const double *pq, const double *evals, const double *qinv_s, int n,
                                                     double y_t, double Sjj, double d) {      /* eq_Aslett_ECS.c:150-171 */
    double term1 = 0.0;
    const double a = y_t - d;
    /* four exponentials at a time (pht_exp4: the same bits, the dependent chains interleaved); the sum keeps its order */
#pragma unroll 1
    for (int i = 0; i < n; i += 4) {
        double e0, e1, e2, e3;
        pht_exp4(evals[i] * a, (i + 1 < n) ? evals[i + 1] * a : 0.0, (i + 2 < n) ? evals[i + 2] * a : 0.0, (i + 3 < n) ? evals[i + 3] * a : 0.0, e0, e1, e2, e3);
        term1 += (pq[i * ECS_THREADS] * e0) * qinv_s[i];
        if (i + 1 < n) term1 += (pq[(i + 1) * ECS_THREADS] * e1) * qinv_s[i + 1];
        if (i + 2 < n) term1 += (pq[(i + 2) * ECS_THREADS] * e2) * qinv_s[i + 2];
        if (i + 3 < n) term1 += (pq[(i + 3) * ECS_THREADS] * e3) * qinv_s[i + 3];
    }
    return pht_log(term1) + Sjj * d;
}
static __device__ __noinline__ double ecs_dens_gt(const double *pq, const double *evals, const double *qinv_1, int n,
                                                  double rem, double scale, double d) {        /* gt_Aslett_DCS.c:111-132 */
    const double x1 = rem - d;
    double r1 = 1.0;
    if (x1 > 0) {
        r1 = 0.0;
#pragma unroll 1
        for (int i = 0; i < n; i += 4) {
            double e0, e1, e2, e3;
            pht_exp4(x1 * evals[i], (i + 1 < n) ? x1 * evals[i + 1] : 0.0, (i + 2 < n) ? x1 * evals[i + 2] : 0.0, (i + 3 < n) ? x1 * evals[i + 3] : 0.0, e0, e1, e2, e3);
            r1 += pq[i * ECS_THREADS] * e0 * qinv_1[i];
            if (i + 1 < n) r1 += pq[(i + 1) * ECS_THREADS] * e1 * qinv_1[i + 1];
            if (i + 2 < n) r1 += pq[(i + 2) * ECS_THREADS] * e2 * qinv_1[i + 2];
            if (i + 3 < n) r1 += pq[(i + 3) * ECS_THREADS] * e3 * qinv_1[i + 3];
        }
    }
    double dens;                                     /* dexp(d, scale, log = TRUE) */
    if (scale <= 0.0) dens = pht_u2d(0x7ff8000000000000ULL);
    else if (d < 0.0) dens = -pht_u2d(0x7ff0000000000000ULL);
    else dens = (-d / scale) - pht_log(scale);
    return pht_log(r1) + dens;
}
template <bool CPLX>
struct DensExact {
    double y_t, Sjj;
    __device__ __forceinline__ double operator()(const EcsSmem &sm, int n, double d) const {
        if (CPLX) return pht_log(spec_bilinear(sm.PQ + threadIdx.x, ECS_THREADS, sm.Qinv_s, sm.evals, sm.evi, n, y_t - d)) + Sjj * d;
        return ecs_dens_exact(sm.PQ + threadIdx.x, sm.evals, sm.Qinv_s, n, y_t, Sjj, d);
    }
};
template <bool CPLX>
struct DensGt {
    double rem, scale;
    __device__ __forceinline__ double operator()(const EcsSmem &sm, int n, double d) const {
        if (CPLX) {
            const double x1 = rem - d;
            const double r1 = x1 > 0 ? spec_bilinear(sm.PQ + threadIdx.x, ECS_THREADS, sm.Qinv_1, sm.evals, sm.evi, n, x1) : 1.0;
            double dens;
            if (scale <= 0.0) dens = pht_u2d(0x7ff8000000000000ULL);
            else if (d < 0.0) dens = -pht_u2d(0x7ff0000000000000ULL);
            else dens = (-d / scale) - pht_log(scale);
            return pht_log(r1) + dens;
        }
        return ecs_dens_gt(sm.PQ + threadIdx.x, sm.evals, sm.Qinv_1, n, rem, scale, d);
    }
};

__device__ __forceinline__ double a_expshift(double y, double y0) { return (y - y0 > -2.0 * A_YCEIL) ? pht_exp(y - y0 + A_YCEIL) : 0.0; }
__device__ __forceinline__ double a_logshift(double y, double y0) { return pht_log(y) + y0 - A_YCEIL; }

/* The ARMS envelope of one lane: an index-linked list of points (x, y, cum) with links packed as
 * f | (pl + 1) << 8 | (pr + 1) << 16, as structure-of-arrays in local memory: 28 bytes per point (the reference's
 * POINT is 56).  exp(y - ymax) is not stored: cumulate and invert recompute it where they need it (same function of
 * the same numbers).  Typically 9-13 points are live; the capacity is the reference's 100. */
template <int STRIDE, int CAPACITY>
struct Env {
    enum { capacity = CAPACITY };
    double *ex, *ey, *ec; unsigned int *el;
    __device__ __forceinline__ double x(int i) const { return ex[i * STRIDE]; }
    __device__ __forceinline__ double y(int i) const { return ey[i * STRIDE]; }
    __device__ __forceinline__ double cum(int i) const { return ec[i * STRIDE]; }
    __device__ __forceinline__ unsigned lk(int i) const { return el[i * STRIDE]; }
    __device__ __forceinline__ void setx(int i, double v) { ex[i * STRIDE] = v; }
    __device__ __forceinline__ void sety(int i, double v) { ey[i * STRIDE] = v; }
    __device__ __forceinline__ void setcum(int i, double v) { ec[i * STRIDE] = v; }
    __device__ __forceinline__ void setlk(int i, unsigned v) { el[i * STRIDE] = v; }
    __device__ __forceinline__ int f(int i) const { return (int)(lk(i) & 1u); }
    __device__ __forceinline__ int pl(int i) const { return (int)((lk(i) >> 8) & 0xffu) - 1; }
    __device__ __forceinline__ int pr(int i) const { return (int)((lk(i) >> 16) & 0xffu) - 1; }
    __device__ __forceinline__ void setpl(int i, int v) { setlk(i, (lk(i) & ~0xff00u) | ((unsigned)(v + 1) << 8)); }
    __device__ __forceinline__ void setpr(int i, int v) { setlk(i, (lk(i) & ~0xff0000u) | ((unsigned)(v + 1) << 16)); }
    __device__ __forceinline__ static unsigned pack(int f, int pl, int pr) { return (unsigned)f | ((unsigned)(pl + 1) << 8) | ((unsigned)(pr + 1) << 16); }
};
typedef Env<1, A_NPOINT> EnvLocal;

/* chord intersection at envelope point q (arms.c:659-764, Metropolis on) */
template <class E>
__device__ __forceinline__ void a_meet(E &e, int q, double convex) {
    double gl = 0.0, gr = 0.0, grl = 0.0, dl = 0.0, dr = 0.0;
    const int pl = e.pl(q), pr = e.pr(q);
    bool il = false, ir = false, irl = false;
    double xpl = 0.0, ypl = 0.0, xpr = 0.0, ypr = 0.0;
    if (pl != A_NIL) { xpl = e.x(pl); ypl = e.y(pl); }
    if (pr != A_NIL) { xpr = e.x(pr); ypr = e.y(pr); }
    if (pl != A_NIL) { const int a = e.pl(e.pl(pl)); if (a != A_NIL) { gl = (ypl - e.y(a)) / (xpl - e.x(a)); il = true; } }
    if (pr != A_NIL) { const int a = e.pr(e.pr(pr)); if (a != A_NIL) { gr = (ypr - e.y(a)) / (xpr - e.x(a)); ir = true; } }
    if (pl != A_NIL && pr != A_NIL) { grl = (ypr - ypl) / (xpr - xpl); irl = true; }
    if (irl && il && (gl < grl)) gl = gl + (1.0 + convex) * (grl - gl);
    if (irl && ir && (gr > grl)) gr = gr + (1.0 + convex) * (grl - gr);
    if (il && irl) { dr = (gl - grl) * (xpr - xpl); if (dr < A_YEPS) dr = A_YEPS; }
    if (ir && irl) { dl = (grl - gr) * (xpr - xpl); if (dl < A_YEPS) dl = A_YEPS; }
    if (il && ir && irl) {
        e.setx(q, (dl * xpr + dr * xpl) / (dl + dr));
        e.sety(q, (dl * ypr + dr * ypl + dl * dr) / (dl + dr));
    } else if (il && irl) { e.setx(q, xpr); e.sety(q, ypr + dr); }
    else if (ir && irl) { e.setx(q, xpl); e.sety(q, ypl + dl); }
    else if (il) e.sety(q, ypl + gl * (e.x(q) - xpl));
    else if (ir) e.sety(q, ypr - gr * (xpr - e.x(q)));
}

/* exponentiate and integrate the envelope (arms.c:625-655, :768-790); returns ymax */
template <class E>
__device__ __forceinline__ double a_cumulate(E &e) {
    double ymax = e.y(0);
#pragma unroll 1
    for (int q = e.pr(0); q != A_NIL; q = e.pr(q)) { const double yq = e.y(q); if (yq > ymax) ymax = yq; }
    double xl = e.x(0), yl = e.y(0), eyl = a_expshift(yl, ymax), cl = 0.;
    e.setcum(0, 0.);
#pragma unroll 1
    for (int q = e.pr(0); q != A_NIL; q = e.pr(q)) {
        const double xq = e.x(q), yq = e.y(q), eyq = a_expshift(yq, ymax);
        double a;
        if (xl == xq) a = 0.;
        else if (fabs(yq - yl) < A_YEPS) a = 0.5 * (eyq + eyl) * (xq - xl);
        else a = ((eyq - eyl) / (yq - yl)) * (xq - xl);
        cl = cl + a;
        e.setcum(q, cl);
        xl = xq; yl = yq; eyl = eyq;
    }
    return ymax;
}

/* One ARMS draw on (0, xr) with the four reference abscissae xr*{1e-6, 1/3, 2/3, 1-1e-6}: arms.c:115-222 */
/* `overflow` is set (and the return value meaningless) when the envelope E cannot take the next point although the
 * reference's could. */
template <class Dens, class E>
__device__ __forceinline__ double arms_core(E &e, const SweepParams &p, uint32_t iter, PathRng &rng, const EcsSmem &sm, int n, const Dens &f,
                                            const double xinit[4], double xr, EcsCounters &c, bool &overflow) {
    const double xl = 0.0, convex = 1.0;
    overflow = false;
    const int mpoint = 9, right = mpoint - 1;
    c.calls++;
    if (xinit[0] <= xl || xinit[3] >= xr) return 0.0;                /* reference error 1003: caller uses xsamp = 0 */
    for (int i = 1; i < 4; i++) if (xinit[i] <= xinit[i - 1]) return 0.0;       /* error 1004 */
#pragma unroll 1
    for (int j = 0; j < mpoint; j++) { e.setlk(j, E::pack(j & 1, j - 1, (j == mpoint - 1) ? A_NIL : j + 1)); e.sety(j, 0.0); }
    e.setx(0, xl); e.setx(right, xr);
#pragma unroll 1
    for (int j = 1, k = 0; j < mpoint - 1; j += 2) {
        const double xv = xinit[k++], yv = f(sm, n, xv); c.evals++;
        e.setx(j, xv); e.sety(j, yv);
        if (!isfinite(yv)) c.nonfinite++;
    }
    int cpoint = mpoint;
    /* intersection points whose chords must be (re)met: up to five indices packed one per byte */
    unsigned long long todo = 0x0806040200ull; int nt = 5;
    const double xprev = 0.0; double yprev = 0.0; bool have_prev = false;
    for (;;) {
#pragma unroll 1
        for (int t = 0; t < nt; t++) a_meet(e, (int)((todo >> (8 * t)) & 0xffull), convex);
        const double ymax = a_cumulate(e);
        if (!have_prev) { yprev = f(sm, n, xprev); c.evals++; if (!isfinite(yprev)) c.nonfinite++; have_prev = true; }
        bool rebuilt = false;
        while (!rebuilt) {
            /* ---- sample from the envelope (invert, arms.c:356-420) */
            double wx, wy, wey; int wpl, wpr;
            {
                int q = right;
                const double u = rng.next(p, iter) * e.cum(q);
#pragma unroll 1
                while (e.pl(q) != A_NIL && e.cum(e.pl(q)) > u) q = e.pl(q);
                const int l = e.pl(q);
                wpl = l; wpr = q;
                const double cuml = e.cum(l), cumq = e.cum(q);
                const double prop = (u - cuml) / (cumq - cuml);
                const double xa = e.x(l), xb = e.x(q), yl = e.y(l), yr = e.y(q);
                const double eyl = a_expshift(yl, ymax), eyr = a_expshift(yr, ymax);
                if (xa == xb) { wx = xb; wy = yr; wey = eyr; }
                else if (fabs(yr - yl) < A_YEPS) {
                    if (fabs(eyr - eyl) > A_EYEPS * fabs(eyr + eyl))
                        wx = xa + ((xb - xa) / (eyr - eyl)) * (-eyl + PHT_SQRT((1. - prop) * eyl * eyl + prop * eyr * eyr));
                    else wx = xa + (xb - xa) * prop;
                    wey = ((wx - xa) / (xb - xa)) * (eyr - eyl) + eyl;
                    wy = a_logshift(wey, ymax);
                } else {
                    wx = xa + ((xb - xa) / (yr - yl)) * (-yl + a_logshift(((1. - prop) * eyl + prop * eyr), ymax));
                    wy = ((wx - xa) / (xb - xa)) * (yr - yl) + yl;
                    wey = a_expshift(wy, ymax);
                }
            }
            /* ---- rejection / Metropolis tests (arms.c:424-521) */
            const double ystar = a_logshift(rng.next(p, iter) * wey, ymax);
            const double ynew = f(sm, n, wx); c.evals++;
            if (!isfinite(ynew)) c.nonfinite++;
            if (ystar >= ynew) {
                /* reject; add the point to the envelope (update, arms.c:525-621) */
                if (cpoint <= A_NPOINT - 2) {
                    if (cpoint + 2 > (int)E::capacity) { overflow = true; return 0.0; }
                    const int q = cpoint, m = cpoint + 1;
                    bool linked = true;
                    const int fl = e.f(wpl), fr = e.f(wpr);
                    e.setx(q, wx); e.sety(q, ynew);
                    if (fl && !fr) {
                        e.setlk(m, E::pack(0, wpl, q)); e.setlk(q, E::pack(1, m, wpr));
                        e.setpr(wpl, m); e.setpl(wpr, q);
                    } else if (!fl && fr) {
                        e.setlk(m, E::pack(0, q, wpr)); e.setlk(q, E::pack(1, wpl, m));
                        e.setpl(wpr, m); e.setpr(wpl, q);
                    } else linked = false;
                    if (linked) {
                        cpoint += 2;
                        const int qpl = e.pl(q), qpr = e.pr(q);
                        const int qpll = e.pl(qpl), qprr = e.pr(qpr);
                        const int ql = (qpll != A_NIL) ? qpll : qpl;
                        const int qr = (qprr != A_NIL) ? qprr : qpr;
                        const double xql = e.x(ql), xqr = e.x(qr), xq = e.x(q);
                        if (xq < (1. - A_XEPS) * xql + A_XEPS * xqr) {
                            const double xv = (1. - A_XEPS) * xql + A_XEPS * xqr; e.setx(q, xv); e.sety(q, f(sm, n, xv)); c.evals++;
                        } else if (xq > A_XEPS * xql + (1. - A_XEPS) * xqr) {
                            const double xv = A_XEPS * xql + (1. - A_XEPS) * xqr; e.setx(q, xv); e.sety(q, f(sm, n, xv)); c.evals++;
                        }
                        /* re-intersect the chords around the new point: q.pl, q.pr, then the next intersections outwards */
                        todo = (unsigned long long)qpl | ((unsigned long long)qpr << 8); nt = 2;
                        if (qpll != A_NIL) { todo |= (unsigned long long)e.pl(qpll) << (8 * nt); nt++; }
                        if (qprr != A_NIL) { todo |= (unsigned long long)e.pr(qprr) << (8 * nt); nt++; }
                        c.updates++;
                        rebuilt = true;                      /* back to cumulate */
                    }
                }
                continue;
            }
            /* Metropolis step against the previous iterate xprev (always 0 here) */
            int ql = 0;
#pragma unroll 1
            while (e.x(e.pr(ql)) < xprev) ql = e.pr(ql);
            const int qr = e.pr(ql);
            double wgt = (xprev - e.x(ql)) / (e.x(qr) - e.x(ql));
            double zold = e.y(ql) + wgt * (e.y(qr) - e.y(ql));
            double znew = wy;
            if (yprev < zold) zold = yprev;
            if (ynew < znew) znew = ynew;
            wgt = ynew - znew - yprev + zold;
            if (wgt > 0.0) wgt = 0.0;
            wgt = (wgt > -A_YCEIL) ? pht_exp(wgt) : 0.0;
            if (rng.next(p, iter) > wgt) { c.rejects++; return xprev; }
            return wx;
        }
    }
}

/* One ARMS draw over a 100-point envelope (the reference's capacity) in the lane's local memory */
template <class Dens>
__device__ __forceinline__ double arms_draw(const SweepParams &p, uint32_t iter, PathRng &rng, const EcsSmem &sm, int n, const Dens &f,
                                            const double xinit[4], double xr, EcsCounters &c) {
    double lx[A_NPOINT], ly[A_NPOINT], lc[A_NPOINT]; unsigned int ll[A_NPOINT];
    EnvLocal e; e.ex = lx; e.ey = ly; e.ec = lc; e.el = ll;
    bool overflow;
    return arms_core(e, p, iter, rng, sm, n, f, xinit, xr, c, overflow);
}

__device__ __forceinline__ void ecs_load_model(const SweepParams &p, EcsSmem &sm, int n) {
    const ModelLayout ML = ModelLayout::make(n, p.m);
    const int tid = threadIdx.x;
    for (int i = tid; i < n * n; i += ECS_THREADS) {
        sm.S[i] = p.model[ML.S + i]; sm.Q[i] = p.model[ML.Q + i]; sm.P[i] = p.model[ML.P + i]; sm.Nacc[i] = 0u;
    }
    for (int i = tid; i < n * (n + 1); i += ECS_THREADS) sm.Pfull[i] = p.model[ML.Pfull + i];
    for (int i = tid; i < n; i += ECS_THREADS) {
        sm.evals[i] = p.model[ML.evals + i]; sm.evi[i] = p.model[ML.evals_im + i]; sm.s[i] = p.model[ML.s + i]; sm.pi[i] = p.model[ML.pi + i];
        sm.Qinv_s[i] = p.model[ML.Qinv_s + i]; sm.Qinv_1[i] = p.model[ML.Qinv_1 + i];
        sm.zlo[i] = 0ull; sm.zhi[i] = 0; sm.Bacc[i] = 0u;
    }
    if (tid == 0) { int c = 0; for (int i = 0; i < n; i++) c |= (p.model[ML.evals_im + i] != 0.0); *sm.cplx = c; }
    __syncthreads();
}

__device__ __forceinline__ void ecs_finish(const SweepParams &p, int n, EcsSmem &sm, EcsCounters &c) {
    const unsigned FULL = 0xffffffffu;
    __syncthreads();
    block_flush<ECS_THREADS>(p, n, sm.Nacc, sm.Bacc, sm.zlo, sm.zhi);
    for (int o = 16; o > 0; o >>= 1) {
        c.jumps += __shfl_down_sync(FULL, c.jumps, o); c.evals += __shfl_down_sync(FULL, c.evals, o);
        c.updates += __shfl_down_sync(FULL, c.updates, o); c.calls += __shfl_down_sync(FULL, c.calls, o);
        c.rejects += __shfl_down_sync(FULL, c.rejects, o); c.nonfinite += __shfl_down_sync(FULL, c.nonfinite, o);
        c.paths += __shfl_down_sync(FULL, c.paths, o);
    }
    if ((threadIdx.x & 31) == 0) {
        atomicAdd(&p.state->counters[PHT_CNT_JUMPS], c.jumps); atomicAdd(&p.state->counters[PHT_CNT_DENS_EVALS], c.evals);
        atomicAdd(&p.state->counters[PHT_CNT_ENV_UPDATES], c.updates); atomicAdd(&p.state->counters[PHT_CNT_ARMS_CALLS], c.calls);
        atomicAdd(&p.state->counters[PHT_CNT_METROP_REJECTS], c.rejects); atomicAdd(&p.state->counters[PHT_CNT_NONFINITE], c.nonfinite);
        atomicAdd(&p.state->counters[PHT_CNT_PATHS], c.paths);
    }
}

/* start state from pi: `while (sofar < target) sofar += pi[k++]` bounded at n-1 */
__device__ __forceinline__ int start_state(const EcsSmem &sm, int n, double target) {
    double sofar = 0.0; int k = 0;
#pragma unroll 1
    while (sofar < target && k <= n - 1) { sofar += sm.pi[k]; k++; }
    return k - 1 < 0 ? 0 : k - 1;
}

/* index list entry k -> local observation; list == nullptr means the identity */
struct ObsList { const uint32_t *idx; unsigned long long count; };

/* warp dispenser over a list */
struct ListDispenser {
    unsigned long long next, end, total; bool exhausted; unsigned long long *counter;
    __device__ __forceinline__ void init(unsigned long long total_, unsigned long long *counter_) {
        next = end = 0ull; total = total_; exhausted = (total_ == 0ull); counter = counter_;
    }
    __device__ __forceinline__ unsigned long long take(unsigned idle, bool me_idle) {
        const unsigned FULL = 0xffffffffu; const int lane = threadIdx.x & 31;
        if (next == end && !exhausted) {
            unsigned long long base = 0;
            if (lane == 0) base = atomicAdd(counter, (unsigned long long)PATH_CHUNK);
            base = __shfl_sync(FULL, base, 0);
            next = base < total ? base : total;
            end = base + PATH_CHUNK < total ? base + PATH_CHUNK : total;
            if (next == end) exhausted = true;
        }
        const unsigned avail = (unsigned)(end - next);
        const unsigned rank = __popc(idle & ((1u << lane) - 1u));
        const unsigned long long mine = (me_idle && rank < avail) ? next + rank : ~0ull;
        const unsigned cnt = __popc(idle);
        next += cnt < avail ? cnt : avail;
        return mine;
    }
};

/* ------------------------------------------------------------------ exact observations */
template <bool CPLX>
__device__ __forceinline__ void ecs_exact_body(const SweepParams &p, const ObsList &list, EcsSmem &sm, int n) {
    const int tid = threadIdx.x;
    const unsigned FULL = 0xffffffffu;
    const uint32_t iter = p.state->iter;
    constexpr bool cplx = CPLX;
    EcsCounters c = {0, 0, 0, 0, 0, 0, 0};
    ListDispenser disp; disp.init(list.count, &p.state->next_obs);
    PathRng rng; rng.seek(0);
    bool active = false;
    double y = 0.0, t = 0.0; int j = 0, B = 0; long out_idx = 0;

    for (;;) {
        unsigned idle = __ballot_sync(FULL, !active);
        if (idle && !disp.exhausted) {
            const unsigned long long k = disp.take(idle, !active);
            if (k != ~0ull) {
                const uint32_t o = list.idx ? list.idx[k] : (uint32_t)k;
                active = true; y = p.y[o]; out_idx = (long)o - p.first; t = 0.0;
                rng.seek(p.obs_rank + o * p.obs_world);
                B = start_state(sm, n, rng.next(p, iter));                              /* eq_Aslett_ECS.c:231-238 */
                j = B;
#pragma unroll 1
                for (int i = 0; i < n; i++) sm.Z[i * ECS_THREADS + tid] = 0.0;
            }
            idle = __ballot_sync(FULL, !active);
        }
        if (idle == FULL) { if (disp.exhausted) break; else continue; }
        /* Structured from here on (no `continue` out of the step): the lanes that have just been refilled, the lanes
         * that absorb and the lanes that go on to a sojourn must run each stage TOGETHER, and the explicit warp
         * barriers make the reconvergence points unmistakable for the compiler (without them the refilled lanes ran
         * the whole step body on their own: every stage showed ~15 of 32 lanes, profiles/r1_final_ecs_1e7_ncu_full.md). */
        __syncwarp();
        const bool step = active;

        /* ---- one step of the path (eq_Aslett_ECS.c:247-364) */
        const double y_t = y - t, Sjj = step ? sm.S[j + j * n] : -1.0;
        bool absorb = false;
        if (step && sm.s[j] > 0.0) {                                                            /* :251-255, probAbsorb :120-136 */
            const double num = (Sjj * y_t) + pht_log(sm.s[j]);
            double den = 0.0;
            if (cplx) den = spec_bilinear(sm.Q + j, n, sm.Qinv_s, sm.evals, sm.evi, n, y_t);
            else {
#pragma unroll 1
                for (int i = 0; i < n; i++) den += sm.Q[j + i * n] * pht_exp(sm.evals[i] * y_t) * sm.Qinv_s[i];
            }
            absorb = rng.next(p, iter) < pht_exp(num - pht_log(den));
        }
        if (absorb) {
            count_transition(p, n, sm.Nacc, out_idx, j, j);                             /* :368 */
            sm.Z[j * ECS_THREADS + tid] += y - t;                                       /* :369 */
            path_flush<ECS_THREADS>(p, n, sm.Z, sm.zlo, sm.zhi, sm.Bacc, B, out_idx);
            c.paths++; active = false;
        }
        __syncwarp();
        if (step && !absorb) {
        /* p_j = S[j,.]/(-S_jj) with p_jj = 0 (:292-295); PQ = p_j^T Q in reference-BLAS order (:160) */
#pragma unroll 1
        for (int i = 0; i < n; i++) sm.W[i * ECS_THREADS + tid] = (i == j) ? 0.0 : sm.S[j + i * n] / (-Sjj);
#pragma unroll 1
        for (int col = 0; col < n; col++) {
            double acc = 0.0;
#pragma unroll 1
            for (int i = 0; i < n; i++) acc += sm.Q[i + col * n] * sm.W[i * ECS_THREADS + tid];
            sm.PQ[col * ECS_THREADS + tid] = 0.0 + 1.0 * acc;
        }
        DensExact<CPLX> f; f.y_t = y_t; f.Sjj = Sjj;
        double xinit[4];
        xinit[0] = y_t / 1e6; xinit[1] = y_t / 3.0; xinit[2] = xinit[1] * 2.0; xinit[3] = y_t - xinit[0];   /* :315-318 */
        const double d = arms_draw(p, iter, rng, sm, n, f, xinit, y_t, c);              /* :338 */
        t += d;
        /* next state (moveMass :21-41): W_r = sum_c Q[r,c] exp(evals_c rem) (Q^-1 s)_c, accumulated over c */
        const double rem = y_t - d;
#pragma unroll 1
        for (int r = 0; r < n; r++) sm.W[r * ECS_THREADS + tid] = 0.0;
        if (cplx) spec_apply(sm.PQ + tid, ECS_THREADS, sm.Qinv_s, sm.evals, sm.evi, n, rem);      /* PQ is free now: exp(rem B) Q^-1 s */
#pragma unroll 1
        for (int col = 0; col < n; col++) {
            const double tv = cplx ? sm.PQ[col * ECS_THREADS + tid] : 1.0 * (pht_exp(sm.evals[col] * rem) * sm.Qinv_s[col]);
#pragma unroll 1
            for (int r = 0; r < n; r++) sm.W[r * ECS_THREADS + tid] += tv * sm.Q[r + col * n];
        }
        double sum = 0.0;
#pragma unroll 1
        for (int i = 0; i < n; i++) { const double v = sm.W[i * ECS_THREADS + tid] * sm.P[j + i * n]; sm.W[i * ECS_THREADS + tid] = v; sum += v; }
#pragma unroll 1
        for (int i = 0; i < n; i++) sm.W[i * ECS_THREADS + tid] = sm.W[i * ECS_THREADS + tid] / sum;
        const int k = slab_scan<ECS_THREADS>(sm.W, n, rng.next(p, iter));               /* :352-358 */
        sm.Z[j * ECS_THREADS + tid] += d;                                               /* :362 */
        count_transition(p, n, sm.Nacc, out_idx, j, k);                                 /* :363 */
        j = k; c.jumps++;
        }
    }
    ecs_finish(p, n, sm, c);
}
/* The spectrum of the sweep's generator is known on the device only: both variants are compiled into the kernel and the
 * block takes one of them (the real-spectrum variant is the reference's arithmetic, bit for bit). */
__global__ void __launch_bounds__(ECS_THREADS, ECS_WARPS_PER_SM * 32 / ECS_THREADS) k_ecs_exact(SweepParams p, ObsList list) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int n = p.n;
    EcsSmem sm; sm.carve(smem_raw, n);
    ecs_load_model(p, sm, n);
    if (*sm.cplx) ecs_exact_body<true>(p, list, sm, n); else ecs_exact_body<false>(p, list, sm, n);
}

/* ------------------------------------------------------------------ censored observations
 * MH = false: the live use (censored observations of ECS, censored = 1 throughout).  MH = true: LJMA_MHsample_Aslett
 * (eq_Aslett_DCS.c:49-143 with reverse = 0, method bit 16; nothing in the reference calls it): every observation goes
 * through this sampler with its own censoring flag -- an uncensored path stops at the first jump time beyond y
 * (gt_Aslett_DCS.c:339,381,390) -- and the chains go through the MH wrapper of path_common.cuh. */
template <bool CPLX, bool MH>
__device__ __forceinline__ void ecs_gt_body(const SweepParams &p, const ObsList &list, EcsSmem &sm, int n) {
    const int tid = threadIdx.x;
    const unsigned FULL = 0xffffffffu;
    const uint32_t iter = p.state->iter;
    constexpr bool cplx = CPLX;
    EcsCounters c = {0, 0, 0, 0, 0, 0, 0};
    ListDispenser disp; disp.init(list.count, &p.state->unit_counter);
    PathRng rng; rng.seek(0);
    MhChain mh; mh.begin(true);
    bool active = false;
    double y = 0.0, t = 0.0; int j = 0, B = 0; long out_idx = 0;

    for (;;) {
        unsigned idle = __ballot_sync(FULL, !active);
        if (idle && !disp.exhausted) {
            const unsigned long long k = disp.take(idle, !active);
            if (k != ~0ull) {
                const uint32_t o = list.idx ? list.idx[k] : (uint32_t)k;
                active = true; y = p.y[o]; out_idx = (long)o - p.first; t = 0.0;
                rng.seek(p.obs_rank + o * p.obs_world);
                if (MH) mh.begin(p.cens[o] != 0);
                B = start_state(sm, n, rng.next(p, iter));                              /* gt_Aslett_DCS.c:313-320 */
                j = B;
#pragma unroll 1
                for (int i = 0; i < n; i++) sm.Z[i * ECS_THREADS + tid] = 0.0;
            }
            idle = __ballot_sync(FULL, !active);
        }
        if (idle == FULL) { if (disp.exhausted) break; else continue; }
        __syncwarp();                       /* refilled and continuing lanes take the step together (see k_ecs_exact) */
        /* A path is a few expensive sojourns before y (survival probability, ARMS draw, spectral next-state weights) and
         * then ~1 / P(absorb) cheap ones after it (an exponential and a table scan).  One step per lane per pass of the
         * outer loop made the ARMS code run with the lanes that happened to be before y: 7.7 of 32
         * (profiles/r2f_ecs_general_4e6_before_ncu_full.md).  So: the first pass of this inner loop is the step of every lane -- all
         * of them before y, see below -- and the further passes serve only lanes that are beyond y, until their paths end:
         * at the next outer pass every active lane is before y again and the expensive code runs converged. */
        bool first = true;
        for (;;) {
        const bool go = active && (first || !(t < y));
        if (__ballot_sync(FULL, go) == 0u) break;
        if (go) {
        /* ---- one step (gt_Aslett_DCS.c:339-384; censored = 1 unless MH) */
        const double lastt = t; const int lastj = j;
        const double Sjj = sm.S[j + j * n];
        double d;
        if (t >= y) d = (1.0 / -Sjj) * (-pht_log(rng.next(p, iter)));                   /* :187-191 */
        else {
            const double x = y - t;
            /* e_j^T Q is row j of Q (the reference forms it with a dgemv over a unit vector) */
            double denom = 1.0;
            if (x > 0) {
                if (cplx) denom = spec_bilinear(sm.Q + j, n, sm.Qinv_1, sm.evals, sm.evi, n, x);
                else { denom = 0.0; for (int i = 0; i < n; i++) denom += sm.Q[j + i * n] * pht_exp(x * sm.evals[i]) * sm.Qinv_1[i]; }
            }
            if (rng.next(p, iter) < pht_exp(Sjj * (y - t)) / denom)                     /* :200-204 */
                d = y - t + (1.0 / -Sjj) * (-pht_log(rng.next(p, iter)));
            else {
#pragma unroll 1
                for (int col = 0; col < n; col++) {                                     /* P[j,.]^T Q */
                    double acc = 0.0;
#pragma unroll 1
                    for (int i = 0; i < n; i++) acc += sm.Q[i + col * n] * sm.P[j + i * n];
                    sm.PQ[col * ECS_THREADS + tid] = 0.0 + 1.0 * acc;
                }
                DensGt<CPLX> f; f.rem = y - t; f.scale = -1.0 / Sjj;
                double xinit[4];
                xinit[0] = (y - t) / 1e6; xinit[1] = (y - t) / 3.0; xinit[2] = xinit[1] * 2.0; xinit[3] = y - t - xinit[0];   /* :227-230 */
                d = arms_draw(p, iter, rng, sm, n, f, xinit, y - t, c);                 /* :250 */
            }
        }
        c.jumps++;
        t += d;
        const double target = rng.next(p, iter);                                        /* :350 */
        int k;
        if (t < y) {                                                                    /* :353-369 */
            const double x1 = y - t;
#pragma unroll 1
            for (int col = 0; col < n; col++) {
                double acc = 0.0;
#pragma unroll 1
                for (int i = 0; i < n; i++) acc += sm.Q[i + col * n] * sm.P[lastj + i * n];
                sm.PQ[col * ECS_THREADS + tid] = 0.0 + 1.0 * acc;
            }
            double r2 = 0.0;
            const bool cx = cplx;
            if (cx) {
                /* W = exp(x1 B) Q^-1 1 once; then every weight is a plain dot product with it */
                spec_apply(sm.W + tid, ECS_THREADS, sm.Qinv_1, sm.evals, sm.evi, n, x1);
#pragma unroll 1
                for (int i = 0; i < n; i++) r2 += sm.PQ[i * ECS_THREADS + tid] * sm.W[i * ECS_THREADS + tid];
            } else {
#pragma unroll 1
                for (int i = 0; i < n; i++) {
                    const double ex = pht_exp(x1 * sm.evals[i]);
                    sm.W[i * ECS_THREADS + tid] = ex;
                    r2 += sm.PQ[i * ECS_THREADS + tid] * ex * sm.Qinv_1[i];
                }
            }
            double sofar = 0.0; k = 0;
#pragma unroll 1
            while (sofar < target && k <= n - 1) {
                const double Plk = sm.P[lastj + k * n];
                if (Plk == 0.0) { k++; continue; }
                double r1 = 0.0;
#pragma unroll 1
                for (int i = 0; i < n; i++) r1 += cx ? sm.Q[k + i * n] * sm.W[i * ECS_THREADS + tid] : sm.Q[k + i * n] * sm.W[i * ECS_THREADS + tid] * sm.Qinv_1[i];
                sofar += r1 * Plk / r2; k++;
            }
            k--; if (k < 0) k = 0;
        } else {                                                                        /* :370-375 */
            double sofar = 0.0; k = 0;
#pragma unroll 1
            while (sofar < target && k <= n) { sofar += sm.Pfull[lastj + k * n]; k++; }
            k--; if (k < 0) k = 0;
        }
        const bool uncens = MH && !mh.cens;
        if (k == n || (uncens && !(t < y))) {                                           /* :379 / loop test :339, then :390-393 */
            sm.Z[lastj * ECS_THREADS + tid] += (uncens ? y : t) - lastt;
            if (!MH || mh.rec) count_transition(p, n, sm.Nacc, out_idx, lastj, lastj);
            if (!MH || mh.chain_end<true>(lastj, sm.s, p.mhit, p, iter, rng.obs)) {
                path_flush<ECS_THREADS>(p, n, sm.Z, sm.zlo, sm.zhi, sm.Bacc, B, out_idx);
                c.paths++; active = false;
            } else {
                /* the next chain of this observation: its own sub-stream, a fresh start state, an empty z */
                rng.seek_sub(mh.chain, mh.off, p, iter);
                B = start_state(sm, n, rng.next(p, iter));
                j = B; t = 0.0;
#pragma unroll 1
                for (int i = 0; i < n; i++) sm.Z[i * ECS_THREADS + tid] = 0.0;
            }
        } else {
            sm.Z[lastj * ECS_THREADS + tid] += t - lastt;                               /* :382 */
            if (!MH || mh.rec) count_transition(p, n, sm.Nacc, out_idx, lastj, k);      /* :383 */
            j = k;
        }
        }
        first = false;
        __syncwarp();
        }
    }
    ecs_finish(p, n, sm, c);
}
__global__ void __launch_bounds__(ECS_THREADS, ECS_WARPS_PER_SM * 32 / ECS_THREADS) k_ecs_gt(SweepParams p, ObsList list) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int n = p.n;
    EcsSmem sm; sm.carve(smem_raw, n);
    ecs_load_model(p, sm, n);
    if (*sm.cplx) ecs_gt_body<true, false>(p, list, sm, n); else ecs_gt_body<false, false>(p, list, sm, n);
}
__global__ void __launch_bounds__(ECS_THREADS, ECS_WARPS_PER_SM * 32 / ECS_THREADS) k_mhs_aslett(SweepParams p, ObsList list) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int n = p.n;
    EcsSmem sm; sm.carve(smem_raw, n);
    ecs_load_model(p, sm, n);
    if (*sm.cplx) ecs_gt_body<true, true>(p, list, sm, n); else ecs_gt_body<false, true>(p, list, sm, n);
}

int pht_ecs_grid_blocks(int device, int n) {
    int a = 0, b = 0, sms = 0;
    const size_t smem = EcsSmem::bytes(n);
    if (cudaFuncSetAttribute(k_ecs_exact, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return -1;
    if (cudaFuncSetAttribute(k_ecs_gt, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return -1;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&a, k_ecs_exact, ECS_THREADS, smem) != cudaSuccess) return -1;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b, k_ecs_gt, ECS_THREADS, smem) != cudaSuccess) return -1;
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device) != cudaSuccess) return -1;
    return (a < b ? a : b) * sms;
}

int pht_mhs_aslett_grid_blocks(int device, int n) {
    int a = 0, sms = 0;
    const size_t smem = EcsSmem::bytes(n);
    if (cudaFuncSetAttribute(k_mhs_aslett, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return -1;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&a, k_mhs_aslett, ECS_THREADS, smem) != cudaSuccess) return -1;
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device) != cudaSuccess) return -1;
    return a * sms;
}
cudaError_t pht_launch_mhs_aslett(const SweepParams &p, int grid_blocks, const uint32_t *idx, unsigned long long count, cudaStream_t st) {
    ObsList l; l.idx = idx; l.count = count;
    if (count) k_mhs_aslett<<<grid_blocks, ECS_THREADS, EcsSmem::bytes(p.n), st>>>(p, l);
    return cudaGetLastError();
}

cudaError_t pht_launch_ecs(const SweepParams &p, int grid_blocks, const uint32_t *idx_exact, unsigned long long n_exact,
                           const uint32_t *idx_cens, unsigned long long n_cens, cudaStream_t st) {
    const size_t smem = EcsSmem::bytes(p.n);
    ObsList le; le.idx = idx_exact; le.count = n_exact;
    ObsList lc; lc.idx = idx_cens; lc.count = n_cens;
    if (n_exact) k_ecs_exact<<<grid_blocks, ECS_THREADS, smem, st>>>(p, le);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    if (n_cens) k_ecs_gt<<<grid_blocks, ECS_THREADS, smem, st>>>(p, lc);
    return cudaGetLastError();
}
