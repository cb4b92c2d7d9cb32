/*
 * pht_oracle.c -- CPU restatement of PhaseType's Gibbs hot path (see pht_oracle.h
 * for scope, parity status and the TEST-INFRASTRUCTURE-ONLY rule).
 *
 * Every function names the reference lines it follows (paths relative to
 * /root/reference/).  The arithmetic keeps the reference's association order so
 * that, built with -ffp-contract=off and the shared exp/log of pht_math.h, its
 * results are bit-identical to the reference compiled by oracle/Makefile.
 * Deliberate differences (all flagged in DESIGN.md):
 *   - categorical scans are bounded at the last category (the reference reads
 *     past the array on round-off: src/Simulate_AbsCTMC_gt_Bladt_MHRS.c:61,101);
 *   - sojourn totals are accumulated across observations in int64 fixed point
 *     (order independent, hence identical for any GPU count) instead of the
 *     reference's sequential double sum (src/Simulate_AbsCTMC_eq_Bladt_MHRS.c:106);
 *   - random numbers come from the keyed Philox contract (pht_philox.h).
 */
#include "pht_oracle.h"
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include "../phasetype_b200/csrc/pht_philox.h"

double pho_exp(double x) { return pht_exp(x); }
double pho_log(double x) { return pht_log(x); }
double pho_exp_general(double x) { return pht_exp_general(x); }      /* the all-range path, for the fast-path equivalence test */
void pho_philox(const uint32_t c[4], const uint32_t k[2], uint32_t out[4]) {
    pht_u32x4 r = pht_philox4x32_10(c[0], c[1], c[2], c[3], k[0], k[1]);
    memcpy(out, r.v, sizeof(r.v));
}
double pho_unif_at(uint64_t seed, uint32_t iter, uint32_t obs, uint32_t sub, uint32_t d) {
    return pht_unif_at((uint32_t)seed, (uint32_t)(seed >> 32), iter, obs, sub, d);
}
double pho_rgamma_at(uint64_t seed, uint32_t iter, uint32_t sub, double shape, double scale) {
    pht_stream st; st.k0 = (uint32_t)seed; st.k1 = (uint32_t)(seed >> 32);
    pht_stream_seek(&st, iter, PHT_OBS_PARAM, sub, 0);
    return pht_rgamma(&st, shape, scale);
}

static pht_stream stream_for(uint64_t seed, uint32_t iter, uint32_t obs) {
    pht_stream st; st.k0 = (uint32_t)seed; st.k1 = (uint32_t)(seed >> 32);
    pht_stream_seek(&st, iter, obs, 0, 0);
    return st;
}
/* the reference's LJMA_GUI() once per attempt/path = advance to the next sub-stream */
static void next_substream(pht_stream *st) { pht_stream_seek(st, st->iter, st->obs, st->sub + 1, 0); }

static double runif01(pht_stream *st) { return 0.0 + (1.0 - 0.0) * pht_stream_unif(st); }
static double rexp_scale(pht_stream *st, double scale) {
    if (!isfinite(scale) || scale <= 0.0) return scale == 0.0 ? 0.0 : NAN;
    return scale * (-pht_log(pht_stream_unif(st)));
}

/* smallest k with (p[0]+...+p[k]) >= target by left-to-right accumulation,
 * bounded at `last` (reference: `while(sofar < target) sofar += p[k++]; k--`) */
static int cat_scan(const double *p, int stride, int last, double target) {
    double sofar = 0.0; int k = 0;
    while (sofar < target && k <= last) { sofar += p[(size_t)k * stride]; k++; }
    k--;
    if (k < 0) k = 0;          /* target <= 0 cannot happen with u in (0,1) */
    return k;
}

/* ------------------------------------------------------------------ a2 */
/* src/PHT_MCMC_Aslett.c:280-297 */
void pho_embedded(int n, const double *S, const double *s, double *P, double *Pfull) {
    for (int i = 0; i < n; i++) {
        double rsumfull = 0.0;
        for (int j = 0; j < n; j++) {
            double v = -S[i + j * n] / S[i + i * n];
            P[i + j * n] = v; Pfull[i + j * n] = v;
            rsumfull += v;
        }
        double rsum = rsumfull - P[i + i * n];
        Pfull[i + n * n] = -s[i] / S[i + i * n];
        rsumfull += Pfull[i + n * n];
        rsumfull -= Pfull[i + i * n];
        Pfull[i + i * n] = 0.0; P[i + i * n] = 0.0;
        for (int j = 0; j < n; j++) {
            P[i + j * n] = P[i + j * n] / rsum;
            Pfull[i + j * n] = Pfull[i + j * n] / rsumfull;
        }
        Pfull[i + n * n] = Pfull[i + n * n] / rsumfull;
    }
}

/* ------------------------------------------------------------------ a6 */
typedef struct { int B, pre; } path_ends;

/* Start distribution of the paths.  The reference fixes it to e1 (src/PHT_MCMC_Aslett.c:190-193) but every sampler takes
 * it as an argument; pho_set_pi lets a test run them from another one (n = 0 restores e1). */
static double g_pi[64]; static int g_pi_n = 0;
void pho_set_pi(const double *pi, int n) { g_pi_n = (pi && n > 0 && n <= 64) ? n : 0; for (int i = 0; i < g_pi_n; i++) g_pi[i] = pi[i]; }
static void pi_init(double *pi, int n) {
    if (g_pi_n == n) for (int i = 0; i < n; i++) pi[i] = g_pi[i];
    else { for (int i = 0; i < n; i++) pi[i] = 0.0; pi[0] = 1.0; }
}

/* One call of the rejection sampler: src/Simulate_AbsCTMC_gt_Bladt_MHRS.c:34-160.
 * `st` continues across calls exactly as R's generator would; one sub-stream per
 * attempt (LJMA_GUI at :120). */
static path_ends mhrs_draw(pht_stream *st, double y, int censored, int n, const double *pi, const double *S,
                           const double *Pfull, double *z2, int *N2, unsigned long long *cnt) {
    double t = 0.0, lastt = 0.0; int B2 = 0, j, lastj = 0;
    while (t < y) {                                                   /* :49 */
        t = 0.0;
        memset(N2, 0, sizeof(int) * (size_t)n * n);
        for (int i = 0; i < n; i++) z2[i] = 0.0;
        B2 = cat_scan(pi, 1, n - 1, runif01(st));                     /* :57-64 */
        j = B2; lastt = t; lastj = j;
        while ((t < y && j < n) || (censored && j < n)) {             /* :75 */
            t = t + rexp_scale(st, 1.0 / -S[j + j * n]);              /* :80 */
            j = cat_scan(Pfull + j, n, n, runif01(st));               /* :86-104 */
            if (cnt) cnt[PHO_C_JUMPS]++;
            if ((t < y && j < n) || (censored && j < n)) {            /* :111-118 */
                z2[lastj] += t - lastt;
                N2[lastj + j * n]++;
                lastj = j; lastt = t;
            }
        }
        next_substream(st);                                           /* :120 */
        if (cnt) cnt[PHO_C_ATTEMPTS]++;
    }
    if (!censored) z2[lastj] += y - lastt; else z2[lastj] += t - lastt;   /* :135-136 */
    N2[lastj + lastj * n]++;                                          /* :137 */
    path_ends e; e.B = B2; e.pre = lastj;
    return e;
}

/* a5: src/Simulate_AbsCTMC_eq_Bladt_MHRS.c:63-110 for one observation */
static path_ends mhrs_observation(pht_stream *st, double y, int censored, int n, const double *pi, const double *S,
                                  const double *s, const double *Pfull, int mhit, double **zc, int **Nc,
                                  double **zp, int **Np, unsigned long long *cnt) {
    path_ends cur = mhrs_draw(st, y, censored, n, pi, S, Pfull, *zc, *Nc, cnt);
    while (s[cur.pre] == 0) cur = mhrs_draw(st, y, censored, n, pi, S, Pfull, *zc, *Nc, cnt);   /* :65-68 */
    if (!censored) {                                                  /* :70 */
        for (int it = 0; it < mhit; it++) {
            path_ends prop = mhrs_draw(st, y, censored, n, pi, S, Pfull, *zp, *Np, cnt);
            while (s[prop.pre] == 0) prop = mhrs_draw(st, y, censored, n, pi, S, Pfull, *zp, *Np, cnt);
            double U = runif01(st);                                   /* :79 */
            if (U < s[prop.pre] / s[cur.pre]) {                       /* :82 */
                double *tz = *zc; *zc = *zp; *zp = tz;
                int *tN = *Nc; *Nc = *Np; *Np = tN;
                cur = prop;
            }
        }
    }
    return cur;
}

int pho_mhrs_paths(uint64_t seed, uint32_t iter, long obs0, long stride, long count,
                   const double *y, const int *cens, int n, const double *S, const double *s,
                   const double *Pfull, int mhit, int *outB, int *outN, double *outz,
                   unsigned long long *counters) {
    double *pi = (double *)calloc(n, sizeof(double));
    double *za = (double *)calloc(n, sizeof(double)), *zb = (double *)calloc(n, sizeof(double));
    int *Na = (int *)calloc((size_t)n * n, sizeof(int)), *Nb = (int *)calloc((size_t)n * n, sizeof(int));
    if (!pi || !za || !zb || !Na || !Nb) return -1;
    pi_init(pi, n);                                                   /* src/PHT_MCMC_Aslett.c:191-192 */
    for (long k = 0; k < count; k++) {
        pht_stream st = stream_for(seed, iter, (uint32_t)(obs0 + k * stride));
        double *zc = za, *zp = zb; int *Nc = Na, *Np = Nb;
        path_ends e = mhrs_observation(&st, y[k], cens[k], n, pi, S, s, Pfull, mhit, &zc, &Nc, &zp, &Np, counters);
        if (counters) counters[PHO_C_PATHS]++;
        outB[k] = e.B;
        memcpy(outN + (size_t)k * n * n, Nc, sizeof(int) * (size_t)n * n);
        memcpy(outz + (size_t)k * n, zc, sizeof(double) * (size_t)n);
    }
    free(pi); free(za); free(zb); free(Na); free(Nb);
    return 0;
}

/* ------------------------------------------------------------------ a3 / a4 */
extern void scipy_dgeevx_(const char *, const char *, const char *, const char *, const int *, double *, const int *,
                          double *, double *, double *, const int *, double *, const int *, int *, int *, double *,
                          double *, double *, double *, double *, const int *, int *, int *,
                          size_t, size_t, size_t, size_t);
extern void scipy_dgetrf_(const int *, const int *, double *, const int *, int *, int *);
extern void scipy_dgetri_(const int *, double *, const int *, int *, double *, const int *, int *);

/* src/utility.c:87-129: S = Q diag(evals) Q^-1 through LAPACK dgeevx('B','V','V','B') and an LU inverse.
 * The algorithm lives in LAPACK (third-party; here the copy inside scipy's OpenBLAS 0.3.x wheel), so this is a
 * call, not a restatement; the device solver is checked against it through invariants (tests/test_spectral*.py). */
int pho_eigen(int n, const double *S, double *evals, double *evals_im, double *Q, double *Qinv) {
    const char B = 'B', V = 'V';
    int info = 0, ilo, ihi, lwork = -1; double abnrm, wq;
    double *A = (double *)malloc(sizeof(double) * n * n), *Ql = (double *)malloc(sizeof(double) * n * n);
    double *wi = (double *)malloc(sizeof(double) * n), *scale = (double *)malloc(sizeof(double) * n);
    double *rce = (double *)malloc(sizeof(double) * n), *rcv = (double *)malloc(sizeof(double) * n);
    int *iwork = (int *)malloc(sizeof(int) * (2 * n + 2)), *ipiv = (int *)malloc(sizeof(int) * n);
    memcpy(A, S, sizeof(double) * n * n);
    scipy_dgeevx_(&B, &V, &V, &B, &n, A, &n, evals, wi, Ql, &n, Q, &n, &ilo, &ihi, scale, &abnrm, rce, rcv, &wq, &lwork, iwork, &info, 1, 1, 1, 1);
    lwork = (int)wq; if (lwork < 4 * n * n + 64) lwork = 4 * n * n + 64;
    double *work = (double *)malloc(sizeof(double) * lwork);
    scipy_dgeevx_(&B, &V, &V, &B, &n, A, &n, evals, wi, Ql, &n, Q, &n, &ilo, &ihi, scale, &abnrm, rce, rcv, work, &lwork, iwork, &info, 1, 1, 1, 1);
    if (info == 0) {
        if (evals_im) memcpy(evals_im, wi, sizeof(double) * n);
        memcpy(Qinv, Q, sizeof(double) * n * n);
        scipy_dgetrf_(&n, &n, Qinv, &n, ipiv, &info);
        if (info == 0) scipy_dgetri_(&n, Qinv, &n, ipiv, work, &lwork, &info);
    }
    free(A); free(Ql); free(wi); free(scale); free(rce); free(rcv); free(iwork); free(ipiv); free(work);
    return info;
}

/* ------------------------------------------------------------------ Brent root finder */
/* src/utility.c:233-338 (R's zeroin with f(a), f(b) supplied; Tol = 0 asks for machine precision) */
typedef double (*pho_fn)(double x, void *ctx);
static double brent_root(double ax, double bx, double fa, double fb, pho_fn f, void *ctx, int maxit, int *status,
                         unsigned long long *cnt) {
    const double EPS = 2.220446049250313e-16;
    double a = ax, b = bx, c = a, fc = fa;
    *status = 0;
    if (fa == 0.0) return a;
    if (fb == 0.0) return b;
    for (int left = maxit + 1; left > 0; left--) {
        const double prev_step = b - a;
        if (fabs(fc) < fabs(fb)) { a = b; b = c; c = a; fa = fb; fb = fc; fc = fa; }
        const double tol_act = 2 * EPS * fabs(b) + 0.0 / 2;
        double new_step = (c - b) / 2;
        if (fabs(new_step) <= tol_act || fb == 0.0) return b;
        if (fabs(prev_step) >= tol_act && fabs(fa) > fabs(fb)) {
            double p, q; const double cb = c - b;
            if (a == c) { const double t1 = fb / fa; p = cb * t1; q = 1.0 - t1; }
            else {
                q = fa / fc; const double t1 = fb / fc, t2 = fb / fa;
                p = t2 * (cb * q * (q - t1) - (b - a) * (t1 - 1.0));
                q = (q - 1.0) * (t1 - 1.0) * (t2 - 1.0);
            }
            if (p > 0.0) q = -q; else p = -p;
            if (p < (0.75 * cb * q - fabs(tol_act * q) / 2) && p < fabs(prev_step * q / 2)) new_step = p / q;
        }
        if (fabs(new_step) < tol_act) new_step = (new_step > 0.0) ? tol_act : -tol_act;
        a = b; fa = fb;
        b += new_step; fb = f(b, ctx);
        if (cnt) cnt[PHO_C_BRENT_EVALS]++;
        if ((fb > 0 && fc > 0) || (fb < 0 && fc < 0)) { c = a; fc = fa; }
    }
    *status = -1;
    return b;
}

/* ------------------------------------------------------------------ a11 / a12: DCS */
typedef struct {
    int n, j;               /* j = destination state of the jump being timed */
    double prob, Pab, T, u; /* T = y - t at the start of the jump */
    double Sll, Slj;        /* S[lastj,lastj], S[lastj,j] */
    const double *Q, *evals, *w; double *J;
    unsigned long long *cnt;
} hob_ctx;

/* sojourn-time CDF minus u: src/Simulate_AbsCTMC_gt_Hobolth_DCS.c:23-40 */
static double hob_cdf(double x, void *vp) {
    hob_ctx *h = (hob_ctx *)vp;
    const int n = h->n;
    for (int i = 0; i < n; i++) {
        if (fabs((h->evals[i] - h->Sll) / h->Sll) < 1e-13) h->J[i] = x * pht_exp(h->evals[i] * h->T);
        else h->J[i] = (pht_exp(h->evals[i] * h->T) - pht_exp((h->T - x) * h->evals[i] + h->Sll * x)) / (h->evals[i] - h->Sll);
    }
    double tmp = 0.0;
    for (int i = 0; i < n; i++) tmp += h->Q[h->j + i * n] * h->J[i] * h->w[i];
    return 1 / h->prob * h->Slj / h->Pab * tmp - h->u;
}

/* plain-loop y = A^T x in the association order of the reference BLAS (R's default libRblas dgemv 'T') */
static void gemv_t(int n, const double *A, const double *x, double *y) {
    for (int j = 0; j < n; j++) {
        double t = 0.0;
        for (int i = 0; i < n; i++) t += A[i + j * n] * x[i];
        y[j] = 0.0 + 1.0 * t;
    }
}

/* src/Simulate_AbsCTMC_eq_AslettHobolth_DCS.c:11-51: end state b with weight (pi e^{Sy})_i s_i */
static int hob_end_state(pht_stream *st, double y, int n, const double *pi, const double *Q, const double *evals,
                         const double *Qinv, const double *s, double *p, double *tmp) {
    gemv_t(n, Q, pi, p);
    for (int i = 0; i < n; i++) p[i] *= pht_exp(evals[i] * y);
    gemv_t(n, Qinv, p, tmp);
    double sum = 0.0;
    for (int i = 0; i < n; i++) { p[i] = tmp[i] * s[i]; sum += p[i]; }
    for (int i = 0; i < n; i++) p[i] = p[i] / sum;
    return cat_scan(p, 1, n - 1, runif01(st));
}

/* src/Simulate_AbsCTMC_gt_Hobolth_DCS.c:74-226.  Live caller (eq_AslettHobolth_DCS.c:124-133): bvec = e_b, i.e. the
 * exit set is the single state b and w = Q^-1 e_b (column b of Q^-1); then bvec may be NULL.  The MH variant
 * (:268-355) passes a general indicator vector bvec with w = Q^-1 bvec.  *pre = the state the chain stays in to the end. */
static int hob_path(pht_stream *st, double y, int b, const double *bvec, int n, const double *pi, const double *S, const double *Q,
                    const double *evals, const double *w, double *z, int *N, double *J, double *p,
                    int *pre, unsigned long long *cnt) {
    for (int i = 0; i < n; i++) z[i] = 0.0;
    memset(N, 0, sizeof(int) * (size_t)n * n);
    const int B = cat_scan(pi, 1, n - 1, runif01(st));                               /* :88-95 */
    double t = 0.0; int j = B;
    while (t < y) {                                                                  /* :112 */
        const int lastj = j;
        const double T = y - t, Sjj = S[j + j * n];
        double Pab = 0.0;
        for (int i = 0; i < n; i++) Pab += Q[j + i * n] * pht_exp(evals[i] * T) * w[i];     /* :118-121 */
        if (bvec ? (bvec[j] > 0.0) : (j == b)) {                                     /* :124 (b[j] > 0) */
            if (runif01(st) < pht_exp(Sjj * T) / Pab) { z[j] += T; N[j + j * n] = 1; if (pre) *pre = j; break; }   /* :125-130 */
        }
        for (int i = 0; i < n; i++) {                                                /* :137-144 */
            if (fabs((evals[i] - Sjj) / Sjj) < 1e-13) J[i] = T * pht_exp(evals[i] * T);
            else J[i] = (pht_exp(evals[i] * T) - pht_exp(Sjj * T)) / (evals[i] - Sjj);
        }
        double p_sum = 0.0;
        for (int i = 0; i < n; i++) {                                                /* :148-158 */
            if (i == j) continue;
            double tmp = 0.0;
            for (int k = 0; k < n; k++) tmp += Q[i + k * n] * J[k] * w[k];
            p[i] = S[j + i * n] / Pab * tmp;
            p_sum += p[i];
        }
        p[j] = 0.0;
        const double target = (p_sum == 0.0) ? 0.0 : 0.0 + (p_sum - 0.0) * pht_stream_unif(st);   /* runif(0, p_sum), :164 */
        j = cat_scan(p, 1, n - 1, target);
        hob_ctx h; h.n = n; h.j = j; h.prob = p[j]; h.Pab = Pab; h.T = T; h.Sll = Sjj; h.Slj = S[lastj + j * n];
        h.Q = Q; h.evals = evals; h.w = w; h.J = J; h.cnt = cnt;
        h.u = runif01(st);                                                           /* :184 */
        int status;
        double jtime = brent_root(0.0, T, -h.u, 1.0 - h.u, hob_cdf, &h, 1000, &status, cnt);   /* :189 */
        if (status != 0 && cnt) cnt[PHO_C_NONFINITE]++;
        while (t + jtime >= y) jtime = jtime / 2;                                    /* :204-206 */
        N[lastj + j * n]++; z[lastj] += jtime; t += jtime;                           /* :209-213 */
        if (cnt) cnt[PHO_C_JUMPS]++;
    }
    return B;
}

int pho_dcs_paths(uint64_t seed, uint32_t iter, long obs0, long stride, long count,
                  const double *y, int n, const double *S, const double *s,
                  const double *evals, const double *Q, const double *Qinv,
                  int *outB, int *outN, double *outz, unsigned long long *counters) {
    double *pi = (double *)calloc(n, sizeof(double)), *wk = (double *)calloc(4 * (size_t)n, sizeof(double));
    double *z = (double *)calloc(n, sizeof(double)); int *N = (int *)calloc((size_t)n * n, sizeof(int));
    if (!pi || !wk || !z || !N) return -1;
    pi_init(pi, n);
    for (long k = 0; k < count; k++) {
        pht_stream st = stream_for(seed, iter, (uint32_t)(obs0 + k * stride));
        const int b = hob_end_state(&st, y[k], n, pi, Q, evals, Qinv, s, wk, wk + n);
        const int B = hob_path(&st, y[k], b, NULL, n, pi, S, Q, evals, Qinv + (size_t)b * n, z, N, wk + 2 * n, wk + 3 * n, NULL, counters);
        if (counters) counters[PHO_C_PATHS]++;
        if (outB) {
            outB[k] = B;
            memcpy(outN + (size_t)k * n * n, N, sizeof(int) * (size_t)n * n);
            memcpy(outz + (size_t)k * n, z, sizeof(double) * (size_t)n);
        }
    }
    free(pi); free(wk); free(z); free(N);
    return 0;
}

/* ------------------------------------------------------------------ a9: ARMS (Gilks) */
/* src/arms.c restated with index links instead of pointers.  Point 0 is always the left bound and point
 * 2*ninit the right bound (new points are only ever inserted between them), which replaces the reference's
 * walks to the list ends.  Constants: src/arms.c:46-49. */
#define A_XEPS 0.00001
#define A_YEPS 0.1
#define A_EYEPS 0.001
#define A_YCEIL 50.
#define A_NPOINT 100
#define A_NIL (-1)

typedef struct { double x, y, ey, cum; int f, pl, pr; } arms_pt;
typedef struct {
    arms_pt p[A_NPOINT];
    int cpoint, right;
    double ymax, convex, xprev, yprev;
    pho_fn f; void *ctx;
    pht_stream *st; unsigned long long *cnt;
} arms_env;

static double a_expshift(double y, double y0) { return (y - y0 > -2.0 * A_YCEIL) ? pht_exp(y - y0 + A_YCEIL) : 0.0; }   /* :794-803 */
static double a_logshift(double y, double y0) { return pht_log(y) + y0 - A_YCEIL; }                                      /* :807-812 */
static double a_perfunc(arms_env *e, double x) {                                                                          /* :816-834 */
    double y = e->f(x, e->ctx);
    if (e->cnt) { e->cnt[PHO_C_DENS_EVALS]++; if (!isfinite(y)) e->cnt[PHO_C_NONFINITE]++; }
    return y;
}

/* intersection of the chords either side of q: src/arms.c:659-764 (Metropolis is always on here) */
static void a_meet(arms_env *e, int q) {
    arms_pt *p = e->p;
    double gl = 0.0, gr = 0.0, grl = 0.0, dl = 0.0, dr = 0.0;
    const int pl = p[q].pl, pr = p[q].pr;
    int il = 0, ir = 0, irl = 0;
    if (pl != A_NIL && p[p[pl].pl].pl != A_NIL) {
        const int a = p[p[pl].pl].pl;
        gl = (p[pl].y - p[a].y) / (p[pl].x - p[a].x); il = 1;
    }
    if (pr != A_NIL && p[p[pr].pr].pr != A_NIL) {
        const int a = p[p[pr].pr].pr;
        gr = (p[pr].y - p[a].y) / (p[pr].x - p[a].x); ir = 1;
    }
    if (pl != A_NIL && pr != A_NIL) { grl = (p[pr].y - p[pl].y) / (p[pr].x - p[pl].x); irl = 1; }
    if (irl && il && (gl < grl)) gl = gl + (1.0 + e->convex) * (grl - gl);
    if (irl && ir && (gr > grl)) gr = gr + (1.0 + e->convex) * (grl - gr);
    if (il && irl) { dr = (gl - grl) * (p[pr].x - p[pl].x); if (dr < A_YEPS) dr = A_YEPS; }
    if (ir && irl) { dl = (grl - gr) * (p[pr].x - p[pl].x); if (dl < A_YEPS) dl = A_YEPS; }
    if (il && ir && irl) {
        p[q].x = (dl * p[pr].x + dr * p[pl].x) / (dl + dr);
        p[q].y = (dl * p[pr].y + dr * p[pl].y + dl * dr) / (dl + dr);
    } else if (il && irl) { p[q].x = p[pr].x; p[q].y = p[pr].y + dr; }
    else if (ir && irl) { p[q].x = p[pl].x; p[q].y = p[pl].y + dl; }
    else if (il) p[q].y = p[pl].y + gl * (p[q].x - p[pl].x);
    else if (ir) p[q].y = p[pr].y - gr * (p[pr].x - p[q].x);
}

/* exponentiate and integrate the envelope: src/arms.c:625-655, area :768-790 */
static void a_cumulate(arms_env *e) {
    arms_pt *p = e->p;
    e->ymax = p[0].y;
    for (int q = p[0].pr; q != A_NIL; q = p[q].pr) if (p[q].y > e->ymax) e->ymax = p[q].y;
    for (int q = 0; q != A_NIL; q = p[q].pr) p[q].ey = a_expshift(p[q].y, e->ymax);
    p[0].cum = 0.;
    for (int q = p[0].pr; q != A_NIL; q = p[q].pr) {
        const int l = p[q].pl; double a;
        if (p[l].x == p[q].x) a = 0.;
        else if (fabs(p[q].y - p[l].y) < A_YEPS) a = 0.5 * (p[q].ey + p[l].ey) * (p[q].x - p[l].x);
        else a = ((p[q].ey - p[l].ey) / (p[q].y - p[l].y)) * (p[q].x - p[l].x);
        p[q].cum = p[l].cum + a;
    }
}

/* x with cumulative envelope probability prob: src/arms.c:356-420 */
static void a_invert(arms_env *e, double prob, arms_pt *w) {
    arms_pt *p = e->p;
    int q = e->right;
    const double u = prob * p[q].cum;
    while (p[q].pl != A_NIL && p[p[q].pl].cum > u) q = p[q].pl;
    const int l = p[q].pl;
    w->pl = l; w->pr = q; w->f = 0; w->cum = u;
    const double prop = (u - p[l].cum) / (p[q].cum - p[l].cum);
    if (p[l].x == p[q].x) { w->x = p[q].x; w->y = p[q].y; w->ey = p[q].ey; return; }
    const double xl = p[l].x, xr = p[q].x, yl = p[l].y, yr = p[q].y, eyl = p[l].ey, eyr = p[q].ey;
    if (fabs(yr - yl) < A_YEPS) {
        if (fabs(eyr - eyl) > A_EYEPS * fabs(eyr + eyl))
            w->x = xl + ((xr - xl) / (eyr - eyl)) * (-eyl + sqrt((1. - prop) * eyl * eyl + prop * eyr * eyr));
        else w->x = xl + (xr - xl) * prop;
        w->ey = ((w->x - xl) / (xr - xl)) * (eyr - eyl) + eyl;
        w->y = a_logshift(w->ey, e->ymax);
    } else {
        w->x = xl + ((xr - xl) / (yr - yl)) * (-yl + a_logshift(((1. - prop) * eyl + prop * eyr), e->ymax));
        w->y = ((w->x - xl) / (xr - xl)) * (yr - yl) + yl;
        w->ey = a_expshift(w->y, e->ymax);
    }
}

/* insert the evaluated point w: src/arms.c:525-621 */
static void a_update(arms_env *e, arms_pt *w) {
    arms_pt *p = e->p;
    if (!w->f || e->cpoint > A_NPOINT - 2) return;
    const int q = e->cpoint++, m = e->cpoint++;
    p[q].x = w->x; p[q].y = w->y; p[q].f = 1;
    p[m].f = 0;
    if (p[w->pl].f && !p[w->pr].f) {
        p[m].pl = w->pl; p[m].pr = q; p[q].pl = m; p[q].pr = w->pr;
        p[p[m].pl].pr = m; p[p[q].pr].pl = q;
    } else if (!p[w->pl].f && p[w->pr].f) {
        p[m].pr = w->pr; p[m].pl = q; p[q].pr = m; p[q].pl = w->pl;
        p[p[m].pr].pl = m; p[p[q].pl].pr = q;
    } else { e->cpoint -= 2; return; }      /* "impossible" in the reference (it prints and carries on with dangling links) */
    const int ql = (p[p[q].pl].pl != A_NIL) ? p[p[q].pl].pl : p[q].pl;
    const int qr = (p[p[q].pr].pr != A_NIL) ? p[p[q].pr].pr : p[q].pr;
    if (p[q].x < (1. - A_XEPS) * p[ql].x + A_XEPS * p[qr].x) {
        p[q].x = (1. - A_XEPS) * p[ql].x + A_XEPS * p[qr].x; p[q].y = a_perfunc(e, p[q].x);
    } else if (p[q].x > A_XEPS * p[ql].x + (1. - A_XEPS) * p[qr].x) {
        p[q].x = A_XEPS * p[ql].x + (1. - A_XEPS) * p[qr].x; p[q].y = a_perfunc(e, p[q].x);
    }
    a_meet(e, p[q].pl); a_meet(e, p[q].pr);
    if (p[p[q].pl].pl != A_NIL) a_meet(e, p[p[p[q].pl].pl].pl);
    if (p[p[q].pr].pr != A_NIL) a_meet(e, p[p[p[q].pr].pr].pr);
    a_cumulate(e);
    if (e->cnt) e->cnt[PHO_C_ENV_UPDATES]++;
}

/* rejection + Metropolis tests: src/arms.c:424-521 with metrop->on; 1 = accepted (w->x is the draw) */
static int a_test(arms_env *e, arms_pt *w) {
    arms_pt *p = e->p;
    const double u = pht_stream_unif(e->st) * w->ey;
    const double y = a_logshift(u, e->ymax);
    const double ynew = a_perfunc(e, w->x);
    if (y >= ynew) {
        w->y = ynew; w->ey = a_expshift(w->y, e->ymax); w->f = 1;
        a_update(e, w);
        return 0;
    }
    const double yold = e->yprev;
    int ql = 0;
    while (p[p[ql].pr].x < e->xprev) ql = p[ql].pr;
    const int qr = p[ql].pr;
    double wgt = (e->xprev - p[ql].x) / (p[qr].x - p[ql].x);
    double zold = p[ql].y + wgt * (p[qr].y - p[ql].y);
    double znew = w->y;
    if (yold < zold) zold = yold;
    if (ynew < znew) znew = ynew;
    wgt = ynew - znew - yold + zold;
    if (wgt > 0.0) wgt = 0.0;
    wgt = (wgt > -A_YCEIL) ? pht_exp(wgt) : 0.0;
    const double u2 = pht_stream_unif(e->st);
    if (u2 > wgt) {
        w->x = e->xprev; w->y = e->yprev; w->ey = a_expshift(w->y, e->ymax); w->f = 1; w->pl = ql; w->pr = qr;
        if (e->cnt) e->cnt[PHO_C_METROP_REJECTS]++;
    } else { e->xprev = w->x; e->yprev = ynew; }
    return 1;
}

/* one ARMS draw with the settings both callers use (ninit = 4, npoint = 100, convex = 1, Metropolis on,
 * xprev = 0): src/arms.c:115-222 + initial :226-333 */
static double arms_draw(pht_stream *st, const double xinit[4], double xl, double xr, pho_fn f, void *ctx,
                        unsigned long long *cnt) {
    arms_env e; arms_pt w;
    e.f = f; e.ctx = ctx; e.st = st; e.cnt = cnt; e.convex = 1.0;
    if (cnt) cnt[PHO_C_ARMS_CALLS]++;
    const int mpoint = 9;
    /* the reference returns error 1003/1004 (and its callers carry on with xsamp = 0) when the abscissae are not
     * strictly inside (xl, xr) and increasing */
    if (xinit[0] <= xl || xinit[3] >= xr) return 0.0;
    for (int i = 1; i < 4; i++) if (xinit[i] <= xinit[i - 1]) return 0.0;
    for (int j = 0; j < mpoint; j++) { e.p[j].pl = j - 1; e.p[j].pr = (j == mpoint - 1) ? A_NIL : j + 1; e.p[j].f = j & 1; e.p[j].y = 0.0; }
    e.p[0].x = xl; e.p[mpoint - 1].x = xr;
    for (int j = 1, k = 0; j < mpoint - 1; j += 2) { e.p[j].x = xinit[k++]; e.p[j].y = a_perfunc(&e, e.p[j].x); }
    e.right = mpoint - 1; e.cpoint = mpoint;
    for (int j = 0; j < mpoint; j += 2) a_meet(&e, j);
    a_cumulate(&e);
    e.xprev = 0.0;
    if (e.xprev < xl || e.xprev > xr) return 0.0;                    /* error 1007 */
    e.yprev = a_perfunc(&e, e.xprev);
    for (;;) {
        a_invert(&e, pht_stream_unif(st), &w);
        if (a_test(&e, &w)) return w.x;
    }
}

/* ------------------------------------------------------------------ a8: ECS, exact observation */
typedef struct { int n; double y_t, Sjj; const double *pq, *evals, *Qinv_s; } ecs_ctx;

/* log density of the sojourn up to a constant: src/Simulate_AbsCTMC_eq_Aslett_ECS.c:150-171; pq = p_j^T Q is the
 * same for every evaluation of one ARMS call, so it is formed once by the caller */
static double ecs_dens(double d, void *vp) {
    const ecs_ctx *c = (const ecs_ctx *)vp;
    double term1 = 0.0;
    for (int i = 0; i < c->n; i++) term1 += (c->pq[i] * pht_exp(c->evals[i] * (c->y_t - d))) * c->Qinv_s[i];
    return pht_log(term1) + c->Sjj * d;
}

/* src/Simulate_AbsCTMC_eq_Aslett_ECS.c:205-373 */
static int ecs_exact_path(pht_stream *st, double y, int n, const double *pi, const double *S, const double *s,
                          const double *Q, const double *evals, const double *Qinv_s, const double *P,
                          double *z, int *N, double *wk, unsigned long long *cnt) {
    double *p = wk, *pq = wk + n, *tmp = wk + 2 * n;
    for (int i = 0; i < n; i++) z[i] = 0.0;
    memset(N, 0, sizeof(int) * (size_t)n * n);
    const int B = cat_scan(pi, 1, n - 1, runif01(st));                               /* :231-238 */
    double t = 0.0; int j = B;
    for (;;) {
        const double y_t = y - t, Sjj = S[j + j * n];
        if (s[j] > 0.0) {                                                            /* :251-255, probAbsorb :120-136 */
            const double num = (Sjj * y_t) + pht_log(s[j]);
            double den = 0.0;
            for (int i = 0; i < n; i++) den += Q[j + i * n] * pht_exp(evals[i] * y_t) * Qinv_s[i];
            if (runif01(st) < pht_exp(num - pht_log(den))) break;
        }
        const int lastj = j;
        for (int i = 0; i < n; i++) p[i] = S[j + i * n] / (-Sjj);                    /* :292-295 */
        p[j] = 0.0;
        gemv_t(n, Q, p, pq);                                                         /* :160 */
        ecs_ctx c; c.n = n; c.y_t = y_t; c.Sjj = Sjj; c.pq = pq; c.evals = evals; c.Qinv_s = Qinv_s;
        double xinit[4];
        xinit[0] = y_t / 1e6; xinit[1] = y_t / 3.0; xinit[2] = xinit[1] * 2.0; xinit[3] = y_t - xinit[0];   /* :315-318 */
        const double d = arms_draw(st, xinit, 0.0, y_t, ecs_dens, &c, cnt);          /* :338 */
        t += d;
        /* next state, moveMass :21-41: weights P[j,i] * (Q e^{L (y_t-d)} Q^-1 s)_i */
        const double rem = y_t - d;
        for (int i = 0; i < n; i++) tmp[i] = pht_exp(evals[i] * rem) * Qinv_s[i];
        for (int r = 0; r < n; r++) p[r] = 0.0;
        for (int c2 = 0; c2 < n; c2++) { const double tv = 1.0 * tmp[c2]; for (int r = 0; r < n; r++) p[r] += tv * Q[r + c2 * n]; }
        double sum = 0.0;
        for (int i = 0; i < n; i++) { p[i] = p[i] * P[j + i * n]; sum += p[i]; }
        for (int i = 0; i < n; i++) p[i] = p[i] / sum;
        j = cat_scan(p, 1, n - 1, runif01(st));                                      /* :352-358 */
        z[lastj] += d; N[lastj + j * n]++;                                           /* :362-363 */
        if (cnt) cnt[PHO_C_JUMPS]++;
    }
    N[j + j * n]++; z[j] += y - t;                                                   /* :368-369 */
    return B;
}

/* ------------------------------------------------------------------ a10: Aslett-DCS gt sampler (censored observations under ECS) */
typedef struct { int n; double rem, scale; const double *pq1, *evals, *Qinv_1; } gt_ctx;

/* upper tail sum_c piQ_c exp(x evals_c) (Q^-1 1)_c: src/Simulate_AbsCTMC_gt_Aslett_DCS.c:26-43 */
static double pht_tail(int n, double x, const double *piQ, const double *evals, const double *Qinv_1) {
    if (!(x > 0)) return 1.0;
    double r = 0.0;
    for (int i = 0; i < n; i++) r += piQ[i] * pht_exp(x * evals[i]) * Qinv_1[i];
    return r;
}
/* log conditional jump density up to a constant: :111-132 (pq1 = P[j,.]^T Q hoisted) */
static double gt_dens(double d, void *vp) {
    const gt_ctx *c = (const gt_ctx *)vp;
    const double x1 = c->rem - d;
    const double r1 = (x1 > 0) ? pht_tail(c->n, x1, c->pq1, c->evals, c->Qinv_1) : 1.0;
    double dens;                                                                     /* dexp(d, scale, log = TRUE) */
    if (c->scale <= 0.0) dens = NAN; else if (d < 0.0) dens = -INFINITY; else dens = (-d / c->scale) - pht_log(c->scale);
    return pht_log(r1) + dens;
}

/* src/Simulate_AbsCTMC_gt_Aslett_DCS.c:299-418 (reverse = 0) with condjump_r_ars :184-260 */
static int gt_path(pht_stream *st, double y, int censored, int n, const double *pi, const double *S, const double *Q,
                   const double *evals, const double *Qinv_1, const double *P, const double *Pfull,
                   double *z, int *N, double *wk, int *pre, unsigned long long *cnt) {
    double *pv = wk, *pq = wk + n, *ex = wk + 2 * n;
    for (int i = 0; i < n; i++) z[i] = 0.0;
    memset(N, 0, sizeof(int) * (size_t)n * n);
    const int B = cat_scan(pi, 1, n - 1, runif01(st));                               /* :313-320 */
    double t = 0.0, lastt = 0.0; int j = B, lastj = 0;
    while (t < y || censored) {                                                      /* :339 */
        lastt = t; lastj = j;
        const double Sjj = S[j + j * n];
        double d;
        if (t >= y) d = rexp_scale(st, 1.0 / -Sjj);                                  /* :187-191 */
        else {
            const double x = y - t;
            for (int i = 0; i < n; i++) pv[i] = 0.0;
            pv[j] = 1.0;
            gemv_t(n, Q, pv, pq);                                                    /* :75 via :198 */
            const double denom = pht_tail(n, x, pq, evals, Qinv_1);
            if (runif01(st) < pht_exp(Sjj * (y - t)) / denom) d = y - t + rexp_scale(st, 1.0 / -Sjj);   /* :200-204 */
            else {
                for (int i = 0; i < n; i++) pv[i] = P[j + i * n];
                gemv_t(n, Q, pv, pq);
                gt_ctx c; c.n = n; c.rem = y - t; c.scale = -1.0 / Sjj; c.pq1 = pq; c.evals = evals; c.Qinv_1 = Qinv_1;
                double xinit[4];
                xinit[0] = (y - t) / 1e6; xinit[1] = (y - t) / 3.0; xinit[2] = xinit[1] * 2.0; xinit[3] = y - t - xinit[0];   /* :227-230 */
                d = arms_draw(st, xinit, 0.0, y - t, gt_dens, &c, cnt);              /* :250 */
            }
        }
        if (cnt) cnt[PHO_C_JUMPS]++;
        t += d;
        const double target = runif01(st);                                           /* :350 */
        double sofar = 0.0; j = 0;
        if (t < y) {                                                                 /* :353-369 */
            const double x1 = y - t;
            for (int i = 0; i < n; i++) pv[i] = P[lastj + i * n];
            gemv_t(n, Q, pv, pq);
            const double r2 = pht_tail(n, x1, pq, evals, Qinv_1);
            for (int i = 0; i < n; i++) ex[i] = pht_exp(x1 * evals[i]);
            while (sofar < target && j <= n - 1) {
                if (P[lastj + j * n] == 0.0) { j++; continue; }
                for (int i = 0; i < n; i++) pv[i] = 0.0;
                pv[j] = 1.0;
                gemv_t(n, Q, pv, pq);
                double r1 = 0.0;
                for (int i = 0; i < n; i++) r1 += pq[i] * ex[i] * Qinv_1[i];
                sofar += r1 * P[lastj + j * n] / r2; j++;
            }
            j--;
            if (j < 0) j = 0;
        } else j = cat_scan(Pfull + lastj, n, n, target);                            /* :370-375 */
        if (j == n) break;                                                           /* :379 */
        if (t < y || censored) { z[lastj] += t - lastt; N[lastj + j * n]++; }        /* :381-383 */
    }
    if (!censored) z[lastj] += y - lastt; else z[lastj] += t - lastt;                /* :390-391 */
    N[lastj + lastj * n]++;                                                          /* :392 */
    if (pre) *pre = lastj;                                                           /* :393 */
    return B;
}

/* a7: src/Simulate_AbsCTMC_eq_Aslett_ECS.c:461-479 */
int pho_ecs_paths(uint64_t seed, uint32_t iter, long obs0, long stride, long count,
                  const double *y, const int *cens, int n, const double *S, const double *s,
                  const double *P, const double *Pfull,
                  const double *evals, const double *Q, const double *Qinv_s, const double *Qinv_1,
                  int *outB, int *outN, double *outz, unsigned long long *counters) {
    double *pi = (double *)calloc(n, sizeof(double)), *wk = (double *)calloc(4 * (size_t)n, sizeof(double));
    double *z = (double *)calloc(n, sizeof(double)); int *N = (int *)calloc((size_t)n * n, sizeof(int));
    if (!pi || !wk || !z || !N) return -1;
    pi_init(pi, n);
    for (long k = 0; k < count; k++) {
        pht_stream st = stream_for(seed, iter, (uint32_t)(obs0 + k * stride));
        int B;
        if (cens[k]) B = gt_path(&st, y[k], 1, n, pi, S, Q, evals, Qinv_1, P, Pfull, z, N, wk, NULL, counters);
        else B = ecs_exact_path(&st, y[k], n, pi, S, s, Q, evals, Qinv_s, P, z, N, wk, counters);
        if (counters) counters[PHO_C_PATHS]++;
        if (outB) {
            outB[k] = B;
            memcpy(outN + (size_t)k * n * n, N, sizeof(int) * (size_t)n * n);
            memcpy(outz + (size_t)k * n, z, sizeof(double) * (size_t)n);
        }
    }
    free(pi); free(wk); free(z); free(N);
    return 0;
}

/* ------------------------------------------------------------------ f1: the two compiled-but-unreachable MH variants
 * Both wrap a direct conditional sampler of "alive at y" paths in the independence Metropolis-Hastings step of
 * MHRS (accept ratio s[p_pre] / s[c_pre]).  Every chain ends with LJMA_GUI() (gt_Hobolth_DCS.c:224,
 * gt_Aslett_DCS.c:416), the sub-stream hook of the Philox contract: chain k of an observation is its sub-stream k,
 * and the accept uniform that follows a proposal is draw 0 of the next sub-stream -- the same layout as MHRS. */
#define CHAIN(call) ((tmpB = (call)), next_substream(&st), tmpB)

/* The exit set the (non-existent) caller of LJMA_MHsample_Hobolth would pass: b_j = 1 where s_j > 0, and
 * Qinv_b = Q^-1 b formed as the reference forms such products (dgemv 'N': y_i accumulated over j,
 * eq_AslettHobolth_DCS.c:131). */
void pho_exit_set(int n, const double *s, const double *Qinv, double *bvec, double *Qinv_b) {
    for (int j = 0; j < n; j++) bvec[j] = (s[j] > 0.0) ? 1.0 : 0.0;
    for (int i = 0; i < n; i++) Qinv_b[i] = 0.0;
    for (int j = 0; j < n; j++) { const double t = 1.0 * bvec[j]; for (int i = 0; i < n; i++) Qinv_b[i] += t * Qinv[i + j * n]; }
}

/* src/Simulate_AbsCTMC_gt_Hobolth_DCS.c:268-355 (LJMA_MHsample_Hobolth).  Reproduced as written: the retry loop
 * of a PROPOSAL tests s[c_pre] (:312), which is non-zero by then, so an invalid proposal is never redrawn --
 * it is rejected by the accept ratio 0 instead. */
int pho_mhs_hobolth_paths(uint64_t seed, uint32_t iter, long obs0, long stride, long count,
                          const double *y, const int *cens, int n, const double *S, const double *s,
                          const double *evals, const double *Q, const double *Qinv, int mhit,
                          int *outB, int *outN, double *outz, unsigned long long *counters) {
    double *pi = (double *)calloc(n, sizeof(double)), *wk = (double *)calloc(4 * (size_t)n, sizeof(double));
    double *za = (double *)calloc(n, sizeof(double)), *zb = (double *)calloc(n, sizeof(double));
    int *Na = (int *)calloc((size_t)n * n, sizeof(int)), *Nb = (int *)calloc((size_t)n * n, sizeof(int));
    if (!pi || !wk || !za || !zb || !Na || !Nb) return -1;
    pi_init(pi, n);
    double *bvec = wk, *Qb = wk + n;
    pho_exit_set(n, s, Qinv, bvec, Qb);
    for (long k = 0; k < count; k++) {
        pht_stream st = stream_for(seed, iter, (uint32_t)(obs0 + k * stride));
        double *zc = za, *zp = zb; int *Nc = Na, *Np = Nb;
        int c_pre = 0, p_pre = 0, tmpB = 0;
        int cB = CHAIN(hob_path(&st, y[k], 0, bvec, n, pi, S, Q, evals, Qb, zc, Nc, wk + 2 * n, wk + 3 * n, &c_pre, counters));       /* :297 */
        while (s[c_pre] == 0.0) cB = CHAIN(hob_path(&st, y[k], 0, bvec, n, pi, S, Q, evals, Qb, zc, Nc, wk + 2 * n, wk + 3 * n, &c_pre, counters));   /* :299-301 */
        if (!cens[k]) {                                                                                      /* :308 */
            for (int it = 0; it < mhit; it++) {
                const int pB = CHAIN(hob_path(&st, y[k], 0, bvec, n, pi, S, Q, evals, Qb, zp, Np, wk + 2 * n, wk + 3 * n, &p_pre, counters));   /* :311; :312 never loops */
                const double U = runif01(&st);                                                               /* :317 */
                if (U < s[p_pre] / s[c_pre]) {                                                               /* :320 */
                    double *tz = zc; zc = zp; zp = tz; int *tN = Nc; Nc = Np; Np = tN;
                    cB = pB; c_pre = p_pre;
                }
            }
        }
        if (counters) counters[PHO_C_PATHS]++;
        outB[k] = cB;
        memcpy(outN + (size_t)k * n * n, Nc, sizeof(int) * (size_t)n * n);
        memcpy(outz + (size_t)k * n, zc, sizeof(double) * (size_t)n);
    }
    free(pi); free(wk); free(za); free(zb); free(Na); free(Nb);
    return 0;
}

/* src/Simulate_AbsCTMC_eq_Aslett_DCS.c:49-143 (LJMA_MHsample_Aslett) with reverse = 0: the gt sampler of
 * src/Simulate_AbsCTMC_gt_Aslett_DCS.c:299-418 called with the observation's own censoring flag. */
int pho_mhs_aslett_paths(uint64_t seed, uint32_t iter, long obs0, long stride, long count,
                         const double *y, const int *cens, int n, const double *S, const double *s,
                         const double *P, const double *Pfull, const double *evals, const double *Q, const double *Qinv_1,
                         int mhit, int *outB, int *outN, double *outz, unsigned long long *counters) {
    double *pi = (double *)calloc(n, sizeof(double)), *wk = (double *)calloc(4 * (size_t)n, sizeof(double));
    double *za = (double *)calloc(n, sizeof(double)), *zb = (double *)calloc(n, sizeof(double));
    int *Na = (int *)calloc((size_t)n * n, sizeof(int)), *Nb = (int *)calloc((size_t)n * n, sizeof(int));
    if (!pi || !wk || !za || !zb || !Na || !Nb) return -1;
    pi_init(pi, n);
    for (long k = 0; k < count; k++) {
        pht_stream st = stream_for(seed, iter, (uint32_t)(obs0 + k * stride));
        double *zc = za, *zp = zb; int *Nc = Na, *Np = Nb;
        int c_pre = 0, p_pre = 0, tmpB = 0;
        int cB = CHAIN(gt_path(&st, y[k], cens[k], n, pi, S, Q, evals, Qinv_1, P, Pfull, zc, Nc, wk, &c_pre, counters));            /* :93 */
        while (s[c_pre] == 0.0) cB = CHAIN(gt_path(&st, y[k], cens[k], n, pi, S, Q, evals, Qinv_1, P, Pfull, zc, Nc, wk, &c_pre, counters));   /* :95-97 */
        if (!cens[k]) {                                                                                      /* :100 */
            for (int it = 0; it < mhit; it++) {
                int pB = CHAIN(gt_path(&st, y[k], 0, n, pi, S, Q, evals, Qinv_1, P, Pfull, zp, Np, wk, &p_pre, counters));          /* :103 */
                while (s[p_pre] == 0.0) pB = CHAIN(gt_path(&st, y[k], 0, n, pi, S, Q, evals, Qinv_1, P, Pfull, zp, Np, wk, &p_pre, counters));   /* :104-106 */
                const double U = runif01(&st);                                                               /* :109 */
                if (U < s[p_pre] / s[c_pre]) {                                                               /* :112 */
                    double *tz = zc; zc = zp; zp = tz; int *tN = Nc; Nc = Np; Np = tN;
                    cB = pB; c_pre = p_pre;
                }
            }
        }
        if (counters) counters[PHO_C_PATHS]++;
        outB[k] = cB;
        memcpy(outN + (size_t)k * n * n, Nc, sizeof(int) * (size_t)n * n);
        memcpy(outz + (size_t)k * n, zc, sizeof(double) * (size_t)n);
    }
    free(pi); free(wk); free(za); free(zb); free(Na); free(Nb);
    return 0;
}
#undef CHAIN

/* ------------------------------------------------------------------ the engine's own spectral solver, on the host */
#include "../phasetype_b200/csrc/pht_eigen.h"
/* same code the device runs (pht_eigen.h); used by the chain-level oracle and checked against LAPACK through
 * invariants in tests/test_spectral.py */
int pho_eigen_native(int n, const double *S, double *evals, double *Q, double *Qinv) {
    double *w = (double *)malloc(sizeof(double) * (2 * (size_t)n * n + 3 * (size_t)n));
    if (!w) return -1;
    int rc = pht_eigen_real(n, S, evals, Q, Qinv, w, w + n * n, w + 2 * n * n, w + 2 * n * n + n, w + 2 * n * n + 2 * n);
    free(w);
    return rc;
}

/* ------------------------------------------------------------------ a1 / a13 / a14: the Gibbs driver */
int pho_choose_zbits(double sum_y) {
    if (!(sum_y > 0.0) || !isfinite(sum_y)) return 30;
    int b = 62 - (int)ceil(log2(16.0 * sum_y + 1.0));
    if (b > 52) b = 52;
    if (b < 0) b = 0;
    return b;
}

/* theta -> TT, S, s: src/PHT_MCMC_Aslett.c:209-246 (first) and :365-397 (afterwards) */
static void assemble(int n, const int *T, const double *C, const double *theta, int first, double *TT, double *S, double *s) {
    const int n1 = n + 1;
    for (int i = 0; i < n1; i++) {
        for (int j = 0; j < n1; j++) {
            const int v = T[i + j * n1];
            if (j != i || v != 0) TT[i + j * n1] = v ? theta[v - 1] * C[i + j * n1] : 0.0;
        }
        double acc = 0.0;
        if (first) { for (int j = 0; j < n1; j++) if (T[i + j * n1] != 0) acc -= TT[i + j * n1]; }
        else       { for (int j = n; j >= 0; j--) if (T[i + j * n1] != 0) acc -= TT[i + j * n1]; }
        if (i < n || first) TT[i + i * n1] = acc;
    }
    for (int i = 0; i < n; i++) {
        for (int j = 0; j < n; j++) S[i + j * n] = TT[i + j * n1];
        s[i] = TT[i + n * n1];
    }
}

static int method_pick(int method) {        /* dispatch priority of src/PHT_MCMC_Aslett.c:325-337 */
    if (method & PHO_MHRS) return PHO_MHRS;
    if (method & PHO_DCS) return PHO_DCS;
    if (method & PHO_ECS) return PHO_ECS;
    /* the engine's extension: the two variants the reference compiles but never dispatches (section 8(f)1) */
    if (method & PHO_MHS_HOBOLTH) return PHO_MHS_HOBOLTH;
    if (method & PHO_MHS_ASLETT) return PHO_MHS_ASLETT;
    return 0;
}

int pho_sweep_stats(uint64_t seed, uint32_t iter, int first, int mhit, int method, int n, int m, const int *T, const double *C,
                    const double *theta, const double *y, long l, const int *censored, int rank, int world,
                    int zbits, long long *Nacc, long long *Bacc, long long *zfix, unsigned long long *counters) {
    (void)m;
    const int n1 = n + 1, which = method_pick(method);
    if (!which) return -2;
    long cnt = 0;
    for (long i = rank; i < l; i += world) cnt++;
    double *TT = (double *)calloc((size_t)n1 * n1, sizeof(double)), *S = (double *)calloc((size_t)n * n, sizeof(double));
    double *s = (double *)calloc(n, sizeof(double)), *P = (double *)calloc((size_t)n * n, sizeof(double));
    double *Pfull = (double *)calloc((size_t)n * n1, sizeof(double));
    double *yl = (double *)malloc(sizeof(double) * (size_t)(cnt + 1)); int *cl = (int *)malloc(sizeof(int) * (size_t)(cnt + 1));
    int *B = (int *)malloc(sizeof(int) * (size_t)(cnt + 1)), *N = (int *)malloc(sizeof(int) * (size_t)(cnt + 1) * n * n);
    double *z = (double *)malloc(sizeof(double) * (size_t)(cnt + 1) * n);
    double *ev = (double *)calloc(n, sizeof(double)), *Q = (double *)calloc((size_t)n * n, sizeof(double)), *Qi = (double *)calloc((size_t)n * n, sizeof(double));
    double *Qs = (double *)calloc(n, sizeof(double)), *Q1 = (double *)calloc(n, sizeof(double));
    int rc = 0;
    long k = 0;
    for (long i = rank; i < l; i += world, k++) { yl[k] = y[i]; cl[k] = censored[i]; }
    assemble(n, T, C, theta, first, TT, S, s);
    pho_embedded(n, S, s, P, Pfull);
    if (which != PHO_MHRS) {
        rc = pho_eigen_native(n, S, ev, Q, Qi);
        for (int i = 0; i < n; i++) {             /* Q^-1 s, Q^-1 1 in reference-BLAS order (src/PHT_MCMC_Aslett.c:331-332) */
            double a = 0.0, b = 0.0;
            for (int j = 0; j < n; j++) { a += (1.0 * s[j]) * Qi[i + j * n]; b += (1.0 * 1.0) * Qi[i + j * n]; }
            Qs[i] = a; Q1[i] = b;
        }
    }
    if (rc == 0) {
        if (which == PHO_MHRS) rc = pho_mhrs_paths(seed, iter, rank, world, cnt, yl, cl, n, S, s, Pfull, mhit, B, N, z, counters);
        else if (which == PHO_DCS) rc = pho_dcs_paths(seed, iter, rank, world, cnt, yl, n, S, s, ev, Q, Qi, B, N, z, counters);
        else if (which == PHO_MHS_HOBOLTH) rc = pho_mhs_hobolth_paths(seed, iter, rank, world, cnt, yl, cl, n, S, s, ev, Q, Qi, mhit, B, N, z, counters);
        else if (which == PHO_MHS_ASLETT) rc = pho_mhs_aslett_paths(seed, iter, rank, world, cnt, yl, cl, n, S, s, P, Pfull, ev, Q, Q1, mhit, B, N, z, counters);
        else rc = pho_ecs_paths(seed, iter, rank, world, cnt, yl, cl, n, S, s, P, Pfull, ev, Q, Qs, Q1, B, N, z, counters);
    }
    if (rc == 0) {
        const double zs = ldexp(1.0, zbits);
        memset(Nacc, 0, sizeof(long long) * (size_t)n * n); memset(Bacc, 0, sizeof(long long) * n); memset(zfix, 0, sizeof(long long) * n);
        for (k = 0; k < cnt; k++) {
            Bacc[B[k]]++;
            for (int c = 0; c < n * n; c++) Nacc[c] += N[(size_t)k * n * n + c];
            for (int i = 0; i < n; i++) { const double v = z[(size_t)k * n + i]; if (v != 0.0) zfix[i] += llrint(v * zs); }
        }
    }
    free(TT); free(S); free(s); free(P); free(Pfull); free(yl); free(cl); free(B); free(N); free(z);
    free(ev); free(Q); free(Qi); free(Qs); free(Q1);
    return rc;
}

/* gather + conjugate Gamma draw: src/PHT_MCMC_Aslett.c:340-366 (prepend lists = reverse insertion order) */
int pho_update(uint64_t seed, uint32_t iter, int n, int m, const double *nu, const double *zeta, const int *T,
               const double *C, int zbits, const long long *Nacc, const long long *zfix, double *theta_new) {
    const int n1 = n + 1;
    const double zscale = ldexp(1.0, -zbits);
    for (int v = 0; v < m; v++) {
        long long Nsum = 0; double zsum = 0.0;
        for (int i = n; i >= 0; i--)
            for (int j = n; j >= 0; j--) {
                if (T[i + j * n1] != v + 1) continue;
                Nsum += (j == n) ? Nacc[i + i * n] : Nacc[i + j * n];
                const double zi = (double)zfix[i] * zscale;
                zsum += zi / C[i + j * n1];
            }
        theta_new[v] = pho_rgamma_at(seed, iter, (uint32_t)v, nu[v] + (double)Nsum, 1.0 / (zeta[v] + zsum));
    }
    return 0;
}

int pho_gibbs(uint64_t seed, int it, int mhit, int method, int n, int m, const double *nu, const double *zeta,
              const int *T, const double *C, const double *y, long l, const int *censored,
              const double *start, double *res, unsigned long long *counters) {
    double *theta = (double *)malloc(sizeof(double) * m), *next = (double *)malloc(sizeof(double) * m);
    long long *Nacc = (long long *)malloc(sizeof(long long) * (size_t)n * n), *Bacc = (long long *)malloc(sizeof(long long) * n);
    long long *zfix = (long long *)malloc(sizeof(long long) * n);
    if (start[0] < 0) {                                        /* src/PHT_MCMC_Aslett.c:195-207 */
        uint32_t k = 0;
        for (int i = 0; i < m; i++) theta[i] = (nu[i] > 1) ? (nu[i] - 1.0) / zeta[i] : pho_rgamma_at(seed, 0, k++, nu[i], 1.0 / zeta[i]);
    } else for (int i = 0; i < m; i++) theta[i] = start[i];
    for (int i = 0; i < m; i++) res[0 + (size_t)i * it] = theta[i];
    double sum_y = 0.0;
    for (long i = 0; i < l; i++) sum_y += y[i];
    const int zbits = pho_choose_zbits(sum_y);
    int rc = 0;
    for (int iter = 1; iter < it && rc == 0; iter++) {         /* :268 */
        rc = pho_sweep_stats(seed, (uint32_t)iter, iter == 1, mhit, method, n, m, T, C, theta, y, l, censored, 0, 1, zbits,
                             Nacc, Bacc, zfix, counters);
        if (rc) break;
        pho_update(seed, (uint32_t)iter, n, m, nu, zeta, T, C, zbits, Nacc, zfix, next);
        for (int i = 0; i < m; i++) { theta[i] = next[i]; res[iter + (size_t)i * it] = next[i]; }
    }
    free(theta); free(next); free(Nacc); free(Bacc); free(zfix);
    return rc;
}
