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
    pi[0] = 1.0;                                                      /* src/PHT_MCMC_Aslett.c:191-192 */
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
