/*
 * ref_harness.c -- thin drivers around the UNMODIFIED reference entry points,
 * compiled together with /root/reference/src/*.c and shim/rshim.c into
 * oracle/_ref/libphtref.so.  TEST INFRASTRUCTURE ONLY.
 *
 * Each phtref_*_paths() call positions the shim's Philox stream on
 * (iteration, global observation index) and invokes the reference's own
 * per-method loop with *m = 1, returning per-observation (B, N, z).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* prototypes copied from the reference's public headers' meaning (signatures only) */
void LJMA_Gibbs(int *it, int *mhit, int *method, int *n, int *m, double *nu, double *zeta, int *T, double *C,
                double *y, int *l, int *censored, double *start, int *silent, double *res);
void LJMA_MHsample_Bladt(double *y, int *censored, int *m, double *pi, double *S, double *s, double *Pfull, int *n,
                         int *iter, double *res_z, int *res_B, int *res_N, double *workD, int *workI);
void LJMA_MHsample_Aslett2(double *y, int *censored, int *m, double *pi, double *S, double *s, double *Q,
                           double *evals, double *Qinv_s, double *Qinv_1, double *P, double *Pfull, int *n,
                           double *res_z, int *res_B, int *res_N, double *workD, int *workI);
void LJMA_MHsample_Hobolth2(double *y, int *censored, int *m, double *pi, double *S, double *s, double *Q,
                            double *evals, double *Qinv_b, double *bvec, double *Qinv, int *n, int *iter,
                            double *res_z, int *res_B, int *res_N, double *workD, int *workI);
void LJMA_MHsample_Hobolth(double *y, int *censored, int *m, double *pi, double *S, double *s, double *Q, double *evals,
                           double *Qinv_b, double *b, double *Qinv, int *n, int *iter, double *res_z, int *res_B, int *res_N,
                           double *workD, int *workI);
void LJMA_MHsample_Aslett(double *y, int *censored, int *m, double *pi, double *S, double *s, double *P, double *Pfull,
                          double *Q, double *evals, double *Qinv_1, double *Qinv, int *n, int *iter, int *reverse,
                          double *res_z, int *res_B, int *res_N, double *workD, int *workI);
int LJMA_eigen(int *n, double *S, double *evals, double *Q, double *Qinv, double *workD, int *workI);
void LJMA_LAPACKspace(int *n);
void LJMA_LAPACKspaceFree(void);
extern int LJMA_counter;

void phtshim_seed(uint64_t seed);
void phtshim_key(uint32_t iter, uint32_t obs);
void phtshim_gibbs_mode(long l, long m, long skip_flush);
void phtshim_counters(unsigned long long *out);
void phtshim_release(void);

#define WORK 100000   /* same fixed workspace as src/PHT_MCMC_Aslett.c:159-175 */

typedef struct {
    double *workD; int *workI; double *z; int *N, *B; double *pi;
} scratch;

/* start distribution handed to the reference's samplers (they all take it as an argument); default e1 */
static double g_pi[64]; static int g_pi_n = 0;
void phtref_set_pi(const double *pi, int n) { g_pi_n = (pi && n > 0 && n <= 64) ? n : 0; for (int i = 0; i < g_pi_n; i++) g_pi[i] = pi[i]; }

static int scratch_init(scratch *w, int n) {
    w->workD = (double *)calloc(WORK, sizeof(double));
    w->workI = (int *)calloc(WORK, sizeof(int));
    w->z = (double *)calloc(n, sizeof(double));
    w->N = (int *)calloc((size_t)n * n, sizeof(int));
    w->B = (int *)calloc(n, sizeof(int));
    w->pi = (double *)calloc(n, sizeof(double));
    if (!w->workD || !w->workI || !w->z || !w->N || !w->B || !w->pi) return -1;
    if (g_pi_n == n) for (int i = 0; i < n; i++) w->pi[i] = g_pi[i];
    else w->pi[0] = 1.0;       /* src/PHT_MCMC_Aslett.c:191-192 */
    return 0;
}
static void scratch_free(scratch *w) {
    free(w->workD); free(w->workI); free(w->z); free(w->N); free(w->B); free(w->pi);
}
static void emit(const scratch *w, int n, long k, int *outB, int *outN, double *outz) {
    int b = 0;
    for (int i = 0; i < n; i++) if (w->B[i]) b = i;
    outB[k] = b;
    memcpy(outN + (size_t)k * n * n, w->N, sizeof(int) * (size_t)n * n);
    memcpy(outz + (size_t)k * n, w->z, sizeof(double) * (size_t)n);
}

/* counters out: [0] uniforms, [1] flushes (= MHRS attempts), [2] prints, [3] gammas, [4] LJMA_counter (jumps) */
static void counters_out(unsigned long long *c) {
    if (!c) return;
    phtshim_counters(c);
    c[4] = (unsigned long long)LJMA_counter;
}

int phtref_mhrs_paths(uint64_t seed, uint32_t iter, long obs0, long stride, long count,
                      const double *y, const int *cens, int n, double *S, double *s, double *Pfull, int mhit,
                      int *outB, int *outN, double *outz, unsigned long long *counters) {
    scratch w; if (scratch_init(&w, n)) return -1;
    phtshim_seed(seed); LJMA_counter = 0;
    int one = 1;
    for (long k = 0; k < count; k++) {
        double yk = y[k]; int ck = cens[k];
        phtshim_key(iter, (uint32_t)(obs0 + k * stride));
        LJMA_MHsample_Bladt(&yk, &ck, &one, w.pi, S, s, Pfull, &n, &mhit, w.z, w.B, w.N, w.workD, w.workI);
        if (outB) emit(&w, n, k, outB, outN, outz);
    }
    counters_out(counters);
    scratch_free(&w);
    return 0;
}

int phtref_ecs_paths(uint64_t seed, uint32_t iter, long obs0, long stride, long count,
                     const double *y, const int *cens, int n, double *S, double *s, double *P, double *Pfull,
                     double *evals, double *Q, double *Qinv_s, double *Qinv_1,
                     int *outB, int *outN, double *outz, unsigned long long *counters) {
    scratch w; if (scratch_init(&w, n)) return -1;
    phtshim_seed(seed); LJMA_counter = 0;
    int one = 1;
    for (long k = 0; k < count; k++) {
        double yk = y[k]; int ck = cens[k];
        phtshim_key(iter, (uint32_t)(obs0 + k * stride));
        LJMA_MHsample_Aslett2(&yk, &ck, &one, w.pi, S, s, Q, evals, Qinv_s, Qinv_1, P, Pfull, &n,
                              w.z, w.B, w.N, w.workD, w.workI);
        if (outB) emit(&w, n, k, outB, outN, outz);
    }
    counters_out(counters);
    scratch_free(&w);
    return 0;
}

int phtref_dcs_paths(uint64_t seed, uint32_t iter, long obs0, long stride, long count,
                     const double *y, const int *cens, int n, double *S, double *s,
                     double *evals, double *Q, double *Qinv,
                     int *outB, int *outN, double *outz, unsigned long long *counters) {
    scratch w; if (scratch_init(&w, n)) return -1;
    double *Qinv_b = (double *)calloc(n, sizeof(double)), *bvec = (double *)calloc(n, sizeof(double));
    phtshim_seed(seed); LJMA_counter = 0;
    int one = 1, mhit = 1;
    for (long k = 0; k < count; k++) {
        double yk = y[k]; int ck = 0;      /* DCS ignores the flag (and prints a warning per obs if set) */
        (void)cens;
        phtshim_key(iter, (uint32_t)(obs0 + k * stride));
        LJMA_MHsample_Hobolth2(&yk, &ck, &one, w.pi, S, s, Q, evals, Qinv_b, bvec, Qinv, &n, &mhit,
                               w.z, w.B, w.N, w.workD, w.workI);
        if (outB) emit(&w, n, k, outB, outN, outz);
    }
    counters_out(counters);
    free(Qinv_b); free(bvec);
    scratch_free(&w);
    return 0;
}

/* The two MH variants nothing in the reference calls (src/Simulate_AbsCTMC_gt_Hobolth_DCS.c:268,
 * src/Simulate_AbsCTMC_eq_Aslett_DCS.c:49): driven here one observation at a time like the live ones.  The exit set
 * handed to the Hobolth variant is b_j = [s_j > 0] with Qinv_b = Q^-1 b in dgemv-'N' order (the caller's choice;
 * the oracle and the engine make the same one). */
int phtref_mhs_hobolth_paths(uint64_t seed, uint32_t iter, long obs0, long stride, long count,
                             const double *y, const int *cens, int n, double *S, double *s,
                             double *evals, double *Q, double *Qinv, int mhit,
                             int *outB, int *outN, double *outz, unsigned long long *counters) {
    scratch w; if (scratch_init(&w, n)) return -1;
    double *Qinv_b = (double *)calloc(n, sizeof(double)), *bvec = (double *)calloc(n, sizeof(double));
    for (int j = 0; j < n; j++) bvec[j] = (s[j] > 0.0) ? 1.0 : 0.0;
    for (int j = 0; j < n; j++) { const double t = 1.0 * bvec[j]; for (int i = 0; i < n; i++) Qinv_b[i] += t * Qinv[i + j * n]; }
    phtshim_seed(seed); LJMA_counter = 0;
    int one = 1;
    for (long k = 0; k < count; k++) {
        double yk = y[k]; int ck = cens[k];
        phtshim_key(iter, (uint32_t)(obs0 + k * stride));
        LJMA_MHsample_Hobolth(&yk, &ck, &one, w.pi, S, s, Q, evals, Qinv_b, bvec, Qinv, &n, &mhit, w.z, w.B, w.N, w.workD, w.workI);
        if (outB) emit(&w, n, k, outB, outN, outz);
    }
    counters_out(counters);
    free(Qinv_b); free(bvec);
    scratch_free(&w);
    return 0;
}

int phtref_mhs_aslett_paths(uint64_t seed, uint32_t iter, long obs0, long stride, long count,
                            const double *y, const int *cens, int n, double *S, double *s, double *P, double *Pfull,
                            double *evals, double *Q, double *Qinv_1, double *Qinv, int mhit,
                            int *outB, int *outN, double *outz, unsigned long long *counters) {
    scratch w; if (scratch_init(&w, n)) return -1;
    phtshim_seed(seed); LJMA_counter = 0;
    int one = 1, reverse = 0;
    for (long k = 0; k < count; k++) {
        double yk = y[k]; int ck = cens[k];
        phtshim_key(iter, (uint32_t)(obs0 + k * stride));
        LJMA_MHsample_Aslett(&yk, &ck, &one, w.pi, S, s, P, Pfull, Q, evals, Qinv_1, Qinv, &n, &mhit, &reverse,
                             w.z, w.B, w.N, w.workD, w.workI);
        if (outB) emit(&w, n, k, outB, outN, outz);
    }
    counters_out(counters);
    scratch_free(&w);
    return 0;
}

/* spectral decomposition exactly as the reference obtains it (src/utility.c:87-129) */
int phtref_eigen(int n, double *S, double *evals, double *Q, double *Qinv) {
    double *workD = (double *)calloc(WORK, sizeof(double));
    int *workI = (int *)calloc(WORK, sizeof(int));
    LJMA_LAPACKspace(&n);
    int rc = LJMA_eigen(&n, S, evals, Q, Qinv, workD, workI);
    LJMA_LAPACKspaceFree();
    free(workD); free(workI);
    return rc;
}

/* The reference's whole Gibbs routine.  keyed != 0 (ECS/DCS only) maps path p of
 * sweep i onto Philox stream (i, p) through the flush hook; otherwise the run
 * consumes one sequential stream (valid for every method, tier-3 use). */
int phtref_gibbs(uint64_t seed, int keyed, int it, int mhit, int method, int n, int m, double *nu, double *zeta,
                 int *T, double *C, double *y, int l, int *censored, double *start, double *res) {
    int silent = 1;
    phtshim_seed(seed);
    if (keyed) phtshim_gibbs_mode(l, m, 3);
    LJMA_Gibbs(&it, &mhit, &method, &n, &m, nu, zeta, T, C, y, &l, censored, start, &silent, res);
    phtshim_release();
    return 0;
}
