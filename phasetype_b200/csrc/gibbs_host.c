/*
 * gibbs_host.c -- the drop-in native entry point, in C like the reference's.
 *
 * Same symbol and 15-pointer signature as reference src/PHT_MCMC_Aslett.c:104 (argument
 * meaning :72-103); R reaches it through the unchanged registration table
 * (src/Registrations.c:6-14).  Everything inside the iteration loop (:268-405) runs on the
 * GPU through the engine ABI of include/pht_b200.h; this file only does the once-per-call
 * work: start values (:195-207), seed, upload, sweep batches, result scatter, console text.
 * There is no CPU path: if the engine cannot be created the routine reports through
 * Rprintf and returns with rows 1.. of `res` left as R pre-filled them (zeros).
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <stdint.h>
#include "../../include/pht_b200.h"
#include "pht_philox.h"

void Rprintf(const char *, ...);
void R_FlushConsole(void);
void GetRNGstate(void);
void PutRNGstate(void);
double unif_rand(void);

static uint64_t env_u64(const char *name, int *found) {
    const char *s = getenv(name);
    *found = (s != NULL && *s != 0);
    return *found ? strtoull(s, NULL, 0) : 0ULL;
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

    double sum_y = 0.0;
    for (int i = 0; i < *l; i++) sum_y += y[i];

    pht_config cfg; memset(&cfg, 0, sizeof(cfg));
    cfg.n = *n; cfg.m = M; cfg.method = *method; cfg.mhit = *mhit;
    cfg.T = T; cfg.C = C; cfg.nu = nu; cfg.zeta = zeta;
    cfg.seed = seed;
    cfg.device = (int)env_u64("PHT_B200_DEVICE", &found);
    cfg.rank = 0; cfg.world = 1;
    cfg.zbits = pht_choose_zbits(sum_y);
    cfg.mhrs_cap = (int)env_u64("PHT_B200_MHRS_CAP", &found);
    cfg.use_graph = 1;
    { int f2; uint64_t g = env_u64("PHT_B200_GRAPH", &f2); if (f2) cfg.use_graph = (int)g; }

    pht_engine *eng = NULL;
    if (pht_engine_create(&eng, &cfg, y, censored, (long)*l) != 0) {
        /* mirrors the reference's print-and-continue error style (e.g. :334-337) */
        Rprintf("CRITICAL ERROR: %s\n\n", pht_last_error());
        free(theta);
        return;
    }
    Rprintf("Starting phase-type MCMC sampler ...\n\nBegining processing ..."); R_FlushConsole();
    if (*silent) {
        Rprintf(" silent processing selected, there will be no further feedback until MCMC run complete"); R_FlushConsole();
    }

    int ok = pht_engine_set_theta(eng, theta, 1u) == 0;
    const int sweeps = IT - 1;
    /* progress text as the reference prints it (:273), once per batch of sweeps instead of per sweep */
    int batch = sweeps;
    if (!*silent) { batch = sweeps / 100; if (batch < 1) batch = 1; }
    double *rows = (double *)malloc(sizeof(double) * (size_t)(batch > 0 ? batch : 1) * (size_t)(M > 0 ? M : 1));
    if (!rows) ok = 0;
    for (int done = 0; ok && done < sweeps; ) {
        const int k = (sweeps - done < batch) ? sweeps - done : batch;
        if (pht_engine_run(eng, k, rows) != 0) { ok = 0; break; }
        for (int r = 0; r < k; r++)
            for (int v = 0; v < M; v++) res[(size_t)(1 + done + r) + (size_t)v * IT] = rows[(size_t)r * M + v];
        done += k;
        if (!*silent) {
            Rprintf("\rProcessing iteration %d of %d (%.1lf%%)\r", done + 1, IT, (100.0 * (done + 1)) / IT); R_FlushConsole();
        }
    }
    if (!ok) Rprintf("\nCRITICAL ERROR: %s\n", pht_last_error());
    free(rows); free(theta);
    pht_engine_destroy(eng);

    Rprintf("\n\nCompleted MCMC run, returning results ...\n"); R_FlushConsole();
}
