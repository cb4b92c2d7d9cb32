/*
 * rshim.c -- R API stand-in used to build and drive the UNMODIFIED reference C
 * (sources stay under /root/reference/src; nothing is copied) as a checker.
 * TEST INFRASTRUCTURE ONLY: linked into oracle/_ref/libphtref.so, never into
 * the product library.
 *
 * What it defines (SURVEY.md section 8(c) lists every call site):
 *   - unif_rand/exp_rand/runif/rexp/dexp/rgamma on the engine's Philox stream
 *     contract (phasetype_b200/csrc/pht_philox.h), so the reference samplers
 *     and the CUDA kernels consume identical uniforms;
 *   - R_FlushConsole as the sub-stream hook: the reference calls LJMA_GUI()
 *     once per MHRS rejection attempt (src/Simulate_AbsCTMC_gt_Bladt_MHRS.c:120)
 *     and once per finished path in ECS/DCS (src/Simulate_AbsCTMC_eq_Aslett_ECS.c:371,
 *     src/Simulate_AbsCTMC_gt_Aslett_DCS.c:416, src/Simulate_AbsCTMC_gt_Hobolth_DCS.c:224);
 *   - dgemv/dgemm as plain loops in the association order of the Fortran
 *     reference BLAS (R's default libRblas); dgeevx/dgetrf/dgetri forwarded to
 *     the LAPACK in scipy's OpenBLAS;
 *   - Rprintf/REprintf (counted, silent unless PHT_SHIM_VERBOSE), R_alloc.
 */
#include <R.h>
#include <Rmath.h>
#include <R_ext/Lapack.h>
#undef exp
#undef log
#include <stdint.h>
#include "../../phasetype_b200/csrc/pht_philox.h"

/* ---------------------------------------------------------------- math */
#ifdef PHT_SHIM_LIBM
#define SHIM_EXP(x) exp(x)
#define SHIM_LOG(x) log(x)
#else
double phtshim_exp(double x) { return pht_exp(x); }
double phtshim_log(double x) { return pht_log(x); }
#define SHIM_EXP(x) pht_exp(x)
#define SHIM_LOG(x) pht_log(x)
#endif

/* ---------------------------------------------------------------- RNG state */
enum { MODE_SEQ = 0, MODE_KEYED = 1, MODE_GIBBS = 2 };
static struct {
    pht_stream st;
    int mode;
    /* MODE_GIBBS bookkeeping (see phtshim_gibbs_mode) */
    long l, m, skip_flush, paths_done, gammas_done;
    uint32_t iter;
    /* event counters */
    unsigned long long n_unif, n_flush, n_print, n_gamma;
} G;

void phtshim_seed(uint64_t seed) {
    memset(&G, 0, sizeof(G));
    G.st.k0 = (uint32_t)seed; G.st.k1 = (uint32_t)(seed >> 32);
    G.mode = MODE_SEQ;
    pht_stream_seek(&G.st, 0, 0, 0, 0);
}
/* position on (iter, obs), substream 0, draw 0; flushes advance the substream */
void phtshim_key(uint32_t iter, uint32_t obs) {
    G.mode = MODE_KEYED;
    pht_stream_seek(&G.st, iter, obs, 0, 0);
}
/* Drive a whole reference LJMA_Gibbs run (ECS / DCS only: one flush per path).
 * skip_flush = number of LJMA_GUI() calls before the first path (3 when silent:
 * src/PHT_MCMC_Aslett.c:188,264,266).  Path p of sweep `iter` uses stream
 * (iter, obs=p); the m rgamma calls closing the sweep use (iter, PARAM, v). */
void phtshim_gibbs_mode(long l, long m, long skip_flush) {
    G.mode = MODE_GIBBS; G.l = l; G.m = m; G.skip_flush = skip_flush;
    G.paths_done = 0; G.gammas_done = 0; G.iter = 1;
    pht_stream_seek(&G.st, 1, 0, 0, 0);
}
void phtshim_counters(unsigned long long *out) {
    out[0] = G.n_unif; out[1] = G.n_flush; out[2] = G.n_print; out[3] = G.n_gamma;
}

void GetRNGstate(void) {}
void PutRNGstate(void) {}
void R_CheckUserInterrupt(void) {}

void R_FlushConsole(void) {
    G.n_flush++;
    if (G.mode == MODE_KEYED) {
        pht_stream_seek(&G.st, G.st.iter, G.st.obs, G.st.sub + 1, 0);
    } else if (G.mode == MODE_GIBBS) {
        if (G.skip_flush > 0) { if (--G.skip_flush == 0) G.gammas_done = 0; return; }
        G.paths_done++;
        pht_stream_seek(&G.st, G.iter, (uint32_t)G.paths_done, 0, 0);
    }
}

double unif_rand(void) { G.n_unif++; return pht_stream_unif(&G.st); }
double exp_rand(void) { return -SHIM_LOG(unif_rand()); }
double norm_rand(void) { return pht_norm_polar(&G.st); }

double runif(double a, double b) {
    if (a == b) return a;
    return a + (b - a) * unif_rand();
}
double rexp(double scale) {
    if (!isfinite(scale) || scale <= 0.0) return scale == 0.0 ? 0.0 : NAN;
    return scale * exp_rand();
}
double dexp(double x, double scale, int give_log) {
    if (scale <= 0.0) return NAN;
    if (x < 0.0) return give_log ? -INFINITY : 0.0;
    return give_log ? (-x / scale) - SHIM_LOG(scale) : SHIM_EXP(-x / scale) / scale;
}
double rgamma(double shape, double scale) {
    G.n_gamma++;
    if (G.mode == MODE_GIBBS) {
        /* start-value draws happen before any path: iteration 0 */
        uint32_t it = (G.paths_done == 0 && G.iter == 1 && G.skip_flush > 0) ? 0u : G.iter;
        pht_stream gs = G.st;
        pht_stream_seek(&gs, it, PHT_OBS_PARAM, (uint32_t)G.gammas_done, 0);
        double r = pht_rgamma(&gs, shape, scale);
        /* iteration 0: the k-th start draw uses substream k (only parameters with
         * nu <= 1 draw, src/PHT_MCMC_Aslett.c:197-201) */
        if (++G.gammas_done == G.m && it != 0) {
            G.gammas_done = 0; G.iter++; G.paths_done = 0;
            pht_stream_seek(&G.st, G.iter, 0, 0, 0);
        }
        return r;
    }
    return pht_rgamma(&G.st, shape, scale);
}

/* ---------------------------------------------------------------- console + memory */
static int verbose(void) {
    static int v = -1;
    if (v < 0) v = getenv("PHT_SHIM_VERBOSE") != NULL;
    return v;
}
void Rprintf(const char *fmt, ...) {
    G.n_print++;
    if (verbose()) { va_list ap; va_start(ap, fmt); vfprintf(stdout, fmt, ap); va_end(ap); }
}
void REprintf(const char *fmt, ...) {
    G.n_print++;
    if (verbose()) { va_list ap; va_start(ap, fmt); vfprintf(stderr, fmt, ap); va_end(ap); }
}

typedef struct blk_ { struct blk_ *next; } blk;
static blk *allocs = NULL;
char *R_alloc(size_t n, int size) {
    blk *b = (blk *)calloc(1, sizeof(blk) + 16 + n * (size_t)size);
    if (!b) { fprintf(stderr, "phtshim: out of memory\n"); abort(); }
    b->next = allocs; allocs = b;
    return (char *)b + 16;       /* sizeof(blk) <= 16: keeps 16-byte alignment */
}
void phtshim_release(void) {
    while (allocs) { blk *n = allocs->next; free(allocs); allocs = n; }
}

/* ---------------------------------------------------------------- BLAS (reference order) */
#ifdef PHT_SHIM_LIBM
extern void scipy_dgemv_(const char *, const int *, const int *, const double *, const double *, const int *, const double *,
                         const int *, const double *, double *, const int *, size_t);
extern void scipy_dgemm_(const char *, const char *, const int *, const int *, const int *, const double *, const double *,
                         const int *, const double *, const int *, const double *, double *, const int *, size_t, size_t);
void phtshim_dgemv(const char *trans, const int *m, const int *n, const double *alpha,
                   const double *a, const int *lda, const double *x, const int *incx,
                   const double *beta, double *y, const int *incy) {
    scipy_dgemv_(trans, m, n, alpha, a, lda, x, incx, beta, y, incy, 1);
}
void phtshim_dgemm(const char *transa, const char *transb, const int *m, const int *n, const int *k,
                   const double *alpha, const double *a, const int *lda, const double *b, const int *ldb,
                   const double *beta, double *c, const int *ldc) {
    scipy_dgemm_(transa, transb, m, n, k, alpha, a, lda, b, ldb, beta, c, ldc, 1, 1);
}
#else
void phtshim_dgemv(const char *trans, const int *m, const int *n, const double *alpha,
                   const double *a, const int *lda, const double *x, const int *incx,
                   const double *beta, double *y, const int *incy) {
    const int M = *m, N = *n, LDA = *lda, IX = *incx, IY = *incy;
    const int tr = (*trans == 'T' || *trans == 't' || *trans == 'C' || *trans == 'c');
    const int leny = tr ? N : M;
    for (int i = 0; i < leny; i++) y[i * IY] = (*beta == 0.0) ? 0.0 : *beta * y[i * IY];
    if (!tr) {
        for (int j = 0; j < N; j++) {
            double t = *alpha * x[j * IX];
            for (int i = 0; i < M; i++) y[i * IY] += t * a[i + (size_t)j * LDA];
        }
    } else {
        for (int j = 0; j < N; j++) {
            double t = 0.0;
            for (int i = 0; i < M; i++) t += a[i + (size_t)j * LDA] * x[i * IX];
            y[j * IY] += *alpha * t;
        }
    }
}
void phtshim_dgemm(const char *transa, const char *transb, const int *m, const int *n, const int *k,
                   const double *alpha, const double *a, const int *lda, const double *b, const int *ldb,
                   const double *beta, double *c, const int *ldc) {
    if (*transa != 'N' || *transb != 'N') { fprintf(stderr, "phtshim_dgemm: only NN is implemented\n"); abort(); }
    for (int j = 0; j < *n; j++) {
        for (int i = 0; i < *m; i++) c[i + (size_t)j * *ldc] = (*beta == 0.0) ? 0.0 : *beta * c[i + (size_t)j * *ldc];
        for (int l = 0; l < *k; l++) {
            double t = *alpha * b[l + (size_t)j * *ldb];
            for (int i = 0; i < *m; i++) c[i + (size_t)j * *ldc] += t * a[i + (size_t)l * *lda];
        }
    }
}

#endif

/* ---------------------------------------------------------------- LAPACK forwards */
extern void scipy_dgeevx_(const char *, const char *, const char *, const char *, const int *, double *, const int *,
                          double *, double *, double *, const int *, double *, const int *, int *, int *, double *,
                          double *, double *, double *, double *, const int *, int *, int *,
                          size_t, size_t, size_t, size_t);
extern void scipy_dgetrf_(const int *, const int *, double *, const int *, int *, int *);
extern void scipy_dgetri_(const int *, double *, const int *, int *, double *, const int *, int *);

void phtshim_dgeevx(const char *balanc, const char *jobvl, const char *jobvr, const char *sense,
                    const int *n, double *a, const int *lda, double *wr, double *wi,
                    double *vl, const int *ldvl, double *vr, const int *ldvr,
                    int *ilo, int *ihi, double *scale, double *abnrm,
                    double *rconde, double *rcondv, double *work, const int *lwork,
                    int *iwork, int *info) {
    scipy_dgeevx_(balanc, jobvl, jobvr, sense, n, a, lda, wr, wi, vl, ldvl, vr, ldvr, ilo, ihi, scale, abnrm,
                  rconde, rcondv, work, lwork, iwork, info, 1, 1, 1, 1);
}
void phtshim_dgetrf(const int *m, const int *n, double *a, const int *lda, int *ipiv, int *info) {
    scipy_dgetrf_(m, n, a, lda, ipiv, info);
}
void phtshim_dgetri(const int *n, double *a, const int *lda, int *ipiv, double *work, const int *lwork, int *info) {
    scipy_dgetri_(n, a, lda, ipiv, work, lwork, info);
}
