/*
 * gibbs_host.c -- the drop-in native entry point, in C like the reference's.
 *
 * Same symbol and 15-pointer signature as reference src/PHT_MCMC_Aslett.c:104 (argument
 * meaning :72-103); R reaches it through the unchanged registration table
 * (src/Registrations.c:6-14).  Everything inside the iteration loop (:268-405) runs on the
 * GPU through the engine ABI of include/pht_b200.h; this file only does the once-per-call
 * work: start values (:195-207), seed, sharding, upload, sweep batches, result scatter,
 * console text.
 *
 * Several GPUs: R calls the routine once, in one process (R/phtMCMC2.R:73), so the fan-out
 * happens here: one engine and one host thread per device, observation i -> device i mod G,
 * and the engines' exchange windows attached to one another (peer memory over NVLink) for the
 * per-sweep all-reduce of the statistics and the global MHRS tail.  Every device draws the
 * same parameters from the shared key, so rank 0's rows are the result.
 *
 * There is no CPU path: if an engine cannot be created, or the device raises its error word,
 * the routine reports through Rprintf and fills the rows it could not produce with NA, so
 * that nothing downstream can mistake them for samples.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <stdint.h>
#include <pthread.h>
#include <time.h>
#include "../../include/pht_b200.h"
#include "pht_philox.h"

void Rprintf(const char *, ...);
void R_FlushConsole(void);
void GetRNGstate(void);
void PutRNGstate(void);
double unif_rand(void);

#define MAX_GPUS 16

/* comma-separated doubles from the environment: exactly n of them, else NULL */
static double *env_vector(const char *name, int n) {
    const char *s = getenv(name);
    if (s == NULL || *s == 0 || n < 1) return NULL;
    double *v = (double *)malloc(sizeof(double) * (size_t)n);
    if (!v) return NULL;
    int k = 0; char *end = NULL;
    while (k < n) { v[k] = strtod(s, &end); if (end == s) break; k++; s = end; while (*s == ',' || *s == ' ') s++; if (*s == 0) break; }
    if (k != n) { Rprintf("Warning (LJMA_Gibbs): %s needs %d comma-separated values; ignored\n", name, n); free(v); return NULL; }
    return v;
}

static uint64_t env_u64(const char *name, int *found) {
    const char *s = getenv(name);
    *found = (s != NULL && *s != 0);
    return *found ? strtoull(s, NULL, 0) : 0ULL;
}

/* R's NA_real_: a quiet NaN with payload 1954 (arithmetic.c) */
static double na_real(void) { const uint64_t u = 0x7FF00000000007A2ULL; double d; memcpy(&d, &u, 8); return d; }

/* what the host threads share */
typedef struct {
    int world, IT, M, sweeps, batch, silent, use_nccl;
    const pht_config *base;
    const double *y; const int *censored; long l;
    const double *theta;
    double *res;
    unsigned char nccl_id[128];
    unsigned char handles[MAX_GPUS * PHT_PEER_HANDLE_BYTES];
    pthread_barrier_t bar;
    volatile int failed;                 /* some rank could not go on: everybody stops at the next barrier */
    volatile int done_rows;              /* rows of res produced so far (rank 0) */
    char err[512];
    pthread_mutex_t err_lock;
    int devices[MAX_GPUS];
    double *ys[MAX_GPUS]; int *cs[MAX_GPUS]; int ys_ok;       /* the shards, dealt out by the rank threads together */
    /* inference of the start distribution (PHT_B200_BETA): prior, initial value, and the draws (IT x n, row-major) */
    int n; const double *beta, *pi0; double *pi_chain;
} shared_t;

typedef struct { shared_t *sh; int rank; } rank_arg;

static void set_error(shared_t *sh, int rank, const char *msg) {
    pthread_mutex_lock(&sh->err_lock);
    if (!sh->failed) snprintf(sh->err, sizeof(sh->err), "GPU %d: %s", sh->devices[rank], msg);
    sh->failed = 1;
    pthread_mutex_unlock(&sh->err_lock);
}

static double now_ms(void) { struct timespec t; clock_gettime(CLOCK_MONOTONIC, &t); return t.tv_sec * 1e3 + t.tv_nsec * 1e-6; }
#define STAGE(what) do { if (timing) { const double t_ = now_ms(); fprintf(stderr, "[LJMA_Gibbs rank %d] %-24s %8.3f ms\n", rank, what, t_ - t_stage); t_stage = t_; } } while (0)

/* deal observations [a, b) of the caller's vectors out to the shards: observation i -> shard i mod world, position i / world */
#define DEAL_THREADS 8
typedef struct { shared_t *sh; long a, b; } deal_arg;
static void *deal_main(void *argp) {
    deal_arg *d = (deal_arg *)argp; shared_t *sh = d->sh; const int world = sh->world;
    int r = (int)(d->a % world); long k = d->a / world;
    for (long i = d->a; i < d->b; i++) {
        sh->ys[r][k] = sh->y[i]; sh->cs[r][k] = sh->censored[i];
        if (++r == world) { r = 0; k++; }
    }
    return NULL;
}

/* One rank = one device.  Every collective step is bracketed by the thread barrier, and a rank that has failed keeps
 * walking through the barriers without touching its engine, so nobody waits for it in vain. */
static void *rank_main(void *argp) {
    rank_arg *ra = (rank_arg *)argp; shared_t *sh = ra->sh; const int rank = ra->rank, world = sh->world;
    pht_engine *eng = NULL;
    const int timing = getenv("PHT_B200_TIMING") != NULL && (rank == 0 || rank == world - 1);
    double t_stage = now_ms();
    /* this rank's shard: observations rank, rank + world, ... */
    const long l_local = sh->l > rank ? (sh->l - rank + world - 1) / world : 0;
    const double *yl = sh->y; const int *cl = sh->censored;
    if (world > 1) {
        /* Every thread deals ITS contiguous slice of the caller's vectors out to all the shards (one sequential read of
         * the input in total, instead of every rank striding through all of it), then the threads meet. */
        if (sh->ys_ok) {
            /* (with few ranks a slice is tens of megabytes: helper threads share it, DEAL_THREADS in all) */
            const long a = (long)((double)sh->l * rank / world), b = rank == world - 1 ? sh->l : (long)((double)sh->l * (rank + 1) / world);
            int helpers = DEAL_THREADS / world; if (helpers < 1) helpers = 1; if (helpers > 8) helpers = 8;
            deal_arg da[8]; pthread_t dth[8]; int started[8];
            for (int h = 0; h < helpers; h++) {
                da[h].sh = sh; da[h].a = a + (b - a) * h / helpers; da[h].b = (h == helpers - 1) ? b : a + (b - a) * (h + 1) / helpers;
                started[h] = (h > 0) && pthread_create(&dth[h], NULL, deal_main, &da[h]) == 0;
            }
            deal_main(&da[0]);
            for (int h = 1; h < helpers; h++) { if (started[h]) pthread_join(dth[h], NULL); else deal_main(&da[h]); }
        } else set_error(sh, rank, "out of memory");
        pthread_barrier_wait(&sh->bar);
        yl = sh->ys[rank]; cl = sh->cs[rank];
    }
    STAGE("shard gather");
    pht_config cfg = *sh->base; cfg.rank = rank; cfg.world = world; cfg.device = sh->devices[rank];
    if (!sh->failed && pht_engine_create(&eng, &cfg, yl, cl, l_local) != 0) set_error(sh, rank, pht_last_error());
    STAGE("engine create");
    if (world > 1) {
        /* The engines' exchange windows are attached to one another (direct peer pointers: one process): the per-sweep
         * all-reduce of the statistics and the global MHRS tail both run through them, inside the sweep's own kernels.
         * PHT_B200_NCCL=1 routes the all-reduce through an NCCL communicator instead (its set-up costs ~1 s on 8 GPUs). */
        if (sh->use_nccl && rank == 0 && !sh->failed && pht_comm_unique_id(sh->nccl_id) != 0) set_error(sh, rank, pht_last_error());
        if (eng && pht_engine_peer_handle(eng, sh->handles + (size_t)rank * PHT_PEER_HANDLE_BYTES) != 0) set_error(sh, rank, pht_last_error());
        pthread_barrier_wait(&sh->bar);
        if (sh->use_nccl && !sh->failed && pht_engine_comm_init(eng, sh->nccl_id) != 0) set_error(sh, rank, pht_last_error());
        if (!sh->failed && pht_engine_peer_attach(eng, sh->handles) != 0) set_error(sh, rank, pht_last_error());
        pthread_barrier_wait(&sh->bar);
    }
    STAGE("peer windows");
    if (!sh->failed && (sh->beta || sh->pi0) && pht_engine_set_pi(eng, sh->pi0, sh->beta) != 0) set_error(sh, rank, pht_last_error());
    if (!sh->failed && pht_engine_set_theta(eng, sh->theta, 1u) != 0) set_error(sh, rank, pht_last_error());
    double *rows = NULL;
    if (rank == 0) {
        rows = (double *)malloc(sizeof(double) * (size_t)(sh->batch > 0 ? sh->batch : 1) * (size_t)(sh->M > 0 ? sh->M : 1));
        if (!rows) set_error(sh, rank, "out of memory");
    }
    for (int done = 0; done < sh->sweeps; ) {
        if (world > 1) pthread_barrier_wait(&sh->bar);
        if (sh->failed) break;
        const int k = (sh->sweeps - done < sh->batch) ? sh->sweeps - done : sh->batch;
        /* (a device error raised on one rank reaches every rank through the all-reduced statistics block, so all of
         * them return from this call, with or without an error of their own) */
        if (pht_engine_run(eng, k, rows) != 0) { set_error(sh, rank, pht_last_error()); }
        if (world > 1) pthread_barrier_wait(&sh->bar);
        if (sh->failed) break;
        done += k;
        if (rank == 0) {
            for (int r = 0; r < k; r++)
                for (int v = 0; v < sh->M; v++) sh->res[(size_t)(1 + done - k + r) + (size_t)v * sh->IT] = rows[(size_t)r * sh->M + v];
            if (sh->pi_chain && pht_engine_pi_rows(eng, k, sh->pi_chain + (size_t)(1 + done - k) * sh->n) != 0) set_error(sh, rank, pht_last_error());
            sh->done_rows = 1 + done;
            if (!sh->silent) {
                Rprintf("\rProcessing iteration %d of %d (%.1lf%%)\r", done + 1, sh->IT, (100.0 * (done + 1)) / sh->IT); R_FlushConsole();
            }
        }
    }
    free(rows);
    STAGE("sweeps");
    if (eng) pht_engine_destroy(eng);
    STAGE("engine destroy");
    return NULL;
}

void LJMA_Gibbs(int *it, int *mhit, int *method, int *n, int *m, double *nu, double *zeta,
                int *T, double *C, double *y, int *l, int *censored, double *start,
                int *silent, double *res) {
    const int IT = *it, M = *m;
    int found = 0;
    GetRNGstate();
    /* the chain's Philox key: PHT_B200_SEED, else 64 bits from R's generator so set.seed() governs the run */
    uint64_t seed = env_u64("PHT_B200_SEED", &found);
    if (!found) {
        const uint64_t hi = (uint64_t)(unif_rand() * 4294967296.0), lo = (uint64_t)(unif_rand() * 4294967296.0);
        seed = (hi << 32) | (lo & 0xffffffffULL);
    } else if (*start >= 0 && getenv("PHT_B200_SEED_EXACT") == NULL) {
        /* A fixed key and given start values: a resumed run (R/phtMCMC2.R resume=) would otherwise replay the random
         * streams of the run it continues.  Mix the start vector into the key (FNV-1a over its bytes). */
        uint64_t h = 0xcbf29ce484222325ULL;
        const unsigned char *b = (const unsigned char *)start;
        for (size_t i = 0; i < sizeof(double) * (size_t)M; i++) { h ^= b[i]; h *= 0x100000001b3ULL; }
        seed ^= h;
    }
    PutRNGstate();

    Rprintf("Setting up Gibbs run ...\n"); R_FlushConsole();

    /* start values: prior mode when nu > 1, else a prior draw (src/PHT_MCMC_Aslett.c:195-207);
     * the k-th draw uses parameter sub-stream k of sweep 0 */
    double *theta = (double *)malloc(sizeof(double) * (size_t)(M > 0 ? M : 1));
    if (!theta) { Rprintf("Error (LJMA_Gibbs): out of memory\n"); return; }
    if (*start < 0) {
        uint32_t k = 0;
        for (int i = 0; i < M; i++) {
            if (nu[i] > 1) theta[i] = (nu[i] - 1.0) / zeta[i];
            else {
                pht_stream st; st.k0 = (uint32_t)seed; st.k1 = (uint32_t)(seed >> 32);
                pht_stream_seek(&st, 0u, PHT_OBS_PARAM, k++, 0u);
                theta[i] = pht_rgamma(&st, nu[i], 1.0 / zeta[i]);
            }
        }
    } else {
        for (int i = 0; i < M; i++) theta[i] = start[i];
    }
    for (int i = 0; i < M; i++) res[0 + (size_t)i * IT] = theta[i];

    /* (only the magnitude of the total matters -- it sizes the fixed-point scale of the sojourn totals -- so four
     * independent partial sums are as good as one chain of 1e7 dependent additions, and four times faster) */
    double sum_y = 0.0;
    { double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0; int i = 0;
      for (; i + 3 < *l; i += 4) { a0 += y[i]; a1 += y[i + 1]; a2 += y[i + 2]; a3 += y[i + 3]; }
      for (; i < *l; i++) a0 += y[i];
      sum_y = (a0 + a1) + (a2 + a3); }

    pht_config cfg; memset(&cfg, 0, sizeof(cfg));
    cfg.n = *n; cfg.m = M; cfg.method = *method; cfg.mhit = *mhit;
    cfg.T = T; cfg.C = C; cfg.nu = nu; cfg.zeta = zeta;
    cfg.seed = seed;
    cfg.zbits = pht_choose_zbits(sum_y);
    { int f2; uint64_t z = env_u64("PHT_B200_ZBITS", &f2); if (f2 && z <= 52) cfg.zbits = (int)z; }
    cfg.mhrs_cap = (int)env_u64("PHT_B200_MHRS_CAP", &found);
    cfg.use_graph = 1;
    { int f2; uint64_t g = env_u64("PHT_B200_GRAPH", &f2); if (f2) cfg.use_graph = (int)g; }

    /* devices: PHT_B200_GPUS of them starting at PHT_B200_DEVICE; default: every visible device the data can keep
     * busy (one per 2^19 observations; a few hundred thousand paths are microseconds of one B200) */
    shared_t sh; memset(&sh, 0, sizeof(sh));
    const int visible = pht_device_count();
    const int first_dev = (int)env_u64("PHT_B200_DEVICE", &found);
    int gpus = (int)env_u64("PHT_B200_GPUS", &found);
    if (!found || gpus < 1) { gpus = (int)(((long)*l + (1L << 19) - 1) >> 19); if (gpus < 1) gpus = 1; }
    if (gpus > visible - first_dev) gpus = visible - first_dev;
    if (gpus > MAX_GPUS) gpus = MAX_GPUS;
    if (gpus < 1) gpus = 1;               /* no device: engine creation reports it below */
    for (int r = 0; r < gpus; r++) sh.devices[r] = first_dev + r;

    Rprintf("Starting phase-type MCMC sampler ...\n\nBegining processing ..."); R_FlushConsole();
    if (*silent) {
        Rprintf(" silent processing selected, there will be no further feedback until MCMC run complete"); R_FlushConsole();
    }
    { const char *ev = getenv("PHT_B200_NCCL"); sh.use_nccl = (ev && *ev && *ev != '0'); }
    sh.world = gpus; sh.IT = IT; sh.M = M; sh.sweeps = IT - 1; sh.silent = *silent;
    /* progress text as the reference prints it (:273), once per batch of sweeps instead of per sweep */
    sh.batch = sh.sweeps;
    if (!*silent) { sh.batch = sh.sweeps / 100; if (sh.batch < 1) sh.batch = 1; }
    sh.base = &cfg; sh.y = y; sh.censored = censored; sh.l = (long)*l; sh.theta = theta; sh.res = res;
    sh.done_rows = 1;
    /* start distribution: PHT_B200_BETA switches the Dirichlet update on, PHT_B200_PI0 sets the initial value */
    double *beta = env_vector("PHT_B200_BETA", *n), *pi0 = env_vector("PHT_B200_PI0", *n);
    sh.n = *n; sh.beta = beta; sh.pi0 = pi0;
    if (beta) {
        sh.pi_chain = (double *)calloc((size_t)IT * (size_t)*n, sizeof(double));
        if (sh.pi_chain) for (int i = 0; i < *n; i++) sh.pi_chain[i] = pi0 ? pi0[i] : (i == 0 ? 1.0 : 0.0);
    }
    pthread_mutex_init(&sh.err_lock, NULL);
    rank_arg args[MAX_GPUS];
    if (gpus == 1) {
        args[0].sh = &sh; args[0].rank = 0;
        rank_main(&args[0]);
    } else {
        pthread_t th[MAX_GPUS];
        pthread_barrier_init(&sh.bar, NULL, (unsigned)gpus);
        sh.ys_ok = 1;
        for (int r = 0; r < gpus; r++) {
            const long lr = sh.l > r ? (sh.l - r + gpus - 1) / gpus : 0;
            sh.ys[r] = (double *)malloc(sizeof(double) * (size_t)(lr > 0 ? lr : 1));
            sh.cs[r] = (int *)malloc(sizeof(int) * (size_t)(lr > 0 ? lr : 1));
            if (!sh.ys[r] || !sh.cs[r]) sh.ys_ok = 0;
        }
        int started = 0;
        for (int r = 0; r < gpus; r++) {
            args[r].sh = &sh; args[r].rank = r;
            if (pthread_create(&th[r], NULL, rank_main, &args[r]) != 0) break;
            started++;
        }
        if (started < gpus) {
            /* cannot happen short of resource exhaustion; the started threads would wait at the barrier forever */
            Rprintf("\nCRITICAL ERROR: could not start %d host threads\n", gpus);
            abort();
        }
        for (int r = 0; r < gpus; r++) pthread_join(th[r], NULL);
        pthread_barrier_destroy(&sh.bar);
        for (int r = 0; r < gpus; r++) { free(sh.ys[r]); free(sh.cs[r]); }
    }
    if (sh.failed) {
        /* mirrors the reference's print-and-carry-on error style (e.g. :334-337), but the rows that were not produced
         * are NA, not the zeros R pre-filled */
        Rprintf("\nCRITICAL ERROR: %s\n", sh.err);
        const double na = na_real();
        for (int r = sh.done_rows; r < IT; r++) for (int v = 0; v < M; v++) res[(size_t)r + (size_t)v * IT] = na;
    }
    if (sh.pi_chain && !sh.failed) {
        const char *path = getenv("PHT_B200_PI_OUT");
        FILE *f = (path && *path) ? fopen(path, "w") : NULL;
        if (f) {
            for (int r = 0; r < IT; r++) { for (int i = 0; i < *n; i++) fprintf(f, i ? ",%.17g" : "%.17g", sh.pi_chain[(size_t)r * *n + i]); fprintf(f, "\n"); }
            fclose(f);
        } else if (path && *path) Rprintf("Warning (LJMA_Gibbs): cannot write %s\n", path);
    }
    free(sh.pi_chain); free(beta); free(pi0);
    pthread_mutex_destroy(&sh.err_lock);
    free(theta);

    Rprintf("\n\nCompleted MCMC run, returning results ...\n"); R_FlushConsole();
}
