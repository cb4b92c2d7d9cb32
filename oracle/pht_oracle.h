/*
 * pht_oracle.h -- CPU restatement of PhaseType's Gibbs hot path (the checker).
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product (phasetype_b200/, include/)
 * includes, links or calls this; only tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py do.
 *
 * Parity status: the reference's own tests pin no numbers,
 * so this restatement is pinned against the reference ITSELF: oracle/_ref/
 * (the unmodified reference C built by oracle/Makefile with the R stand-in in
 * oracle/shim/) must agree bit for bit on per-observation (B, N, z); see
 * tests/test_oracle.py and the committed vectors in tests/golden/.
 *
 * Conventions: column-major matrices X[i + j*n] as in the reference; `seed`,
 * `iter`, global observation index and sub-stream address the Philox contract
 * in phasetype_b200/csrc/pht_philox.h.
 */
#ifndef PHT_ORACLE_H
#define PHT_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum { PHO_MHRS = 1, PHO_ECS = 2, PHO_DCS = 4,
       PHO_MHS_HOBOLTH = 8,     /* LJMA_MHsample_Hobolth, src/Simulate_AbsCTMC_gt_Hobolth_DCS.c:268-355 (no caller in the reference) */
       PHO_MHS_ASLETT = 16 };   /* LJMA_MHsample_Aslett, src/Simulate_AbsCTMC_eq_Aslett_DCS.c:49-143 (no caller in the reference) */

/* event counters (define the algorithmic work W of BASELINE.md section 4) */
enum {
    PHO_C_PATHS = 0,       /* paths accumulated */
    PHO_C_ATTEMPTS,        /* MHRS rejection attempts */
    PHO_C_JUMPS,           /* MHRS jump-steps / ECS sojourns / DCS jumps */
    PHO_C_DENS_EVALS,      /* ARMS log-density evaluations */
    PHO_C_ENV_UPDATES,     /* ARMS envelope insertions */
    PHO_C_BRENT_EVALS,     /* DCS sojourn-CDF evaluations inside the root finder */
    PHO_C_ARMS_CALLS,
    PHO_C_METROP_REJECTS,  /* ARMS Metropolis rejections (sojourn returned as 0) */
    PHO_C_NONFINITE,       /* non-finite log densities seen */
    PHO_C_COUNT = 16
};

double pho_exp(double x);
double pho_log(double x);
double pho_exp_general(double x);
void pho_philox(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);
double pho_unif_at(uint64_t seed, uint32_t iter, uint32_t obs, uint32_t sub, uint32_t d);
double pho_rgamma_at(uint64_t seed, uint32_t iter, uint32_t sub, double shape, double scale);

/* start distribution used by the *_paths functions (default e1 as in the reference; n = 0 restores it) */
void pho_set_pi(const double *pi, int n);

/* a2: embedded jump chain, src/PHT_MCMC_Aslett.c:280-297 */
void pho_embedded(int n, const double *S, const double *s, double *P, double *Pfull);

/* a3/a4: spectral data, src/utility.c:87-129 (LAPACK dgeevx + dgetrf/dgetri) */
int pho_eigen(int n, const double *S, double *evals, double *evals_im, double *Q, double *Qinv);

/* a5/a6: MHRS, per-observation statistics for observations obs0 + k*stride */
int pho_mhrs_paths(uint64_t seed, uint32_t iter, long obs0, long stride, long count,
                   const double *y, const int *cens, int n, const double *S, const double *s,
                   const double *Pfull, int mhit, int *outB, int *outN, double *outz,
                   unsigned long long *counters);

/* a11/a12: DCS (Aslett-Hobolth), `cens` ignored as in the reference */
int pho_dcs_paths(uint64_t seed, uint32_t iter, long obs0, long stride, long count,
                  const double *y, int n, const double *S, const double *s,
                  const double *evals, const double *Q, const double *Qinv,
                  int *outB, int *outN, double *outz, unsigned long long *counters);

/* a7-a10: ECS with the Aslett-DCS gt sampler for censored observations */
int pho_ecs_paths(uint64_t seed, uint32_t iter, long obs0, long stride, long count,
                  const double *y, const int *cens, int n, const double *S, const double *s,
                  const double *P, const double *Pfull,
                  const double *evals, const double *Q, const double *Qinv_s, const double *Qinv_1,
                  int *outB, int *outN, double *outz, unsigned long long *counters);

/* f1: the two MH variants the reference compiles but never calls; exit set b_j = [s_j > 0], Qinv_b = Q^-1 b */
void pho_exit_set(int n, const double *s, const double *Qinv, double *bvec, double *Qinv_b);
int pho_mhs_hobolth_paths(uint64_t seed, uint32_t iter, long obs0, long stride, long count,
                          const double *y, const int *cens, int n, const double *S, const double *s,
                          const double *evals, const double *Q, const double *Qinv, int mhit,
                          int *outB, int *outN, double *outz, unsigned long long *counters);
int pho_mhs_aslett_paths(uint64_t seed, uint32_t iter, long obs0, long stride, long count,
                         const double *y, const int *cens, int n, const double *S, const double *s,
                         const double *P, const double *Pfull, const double *evals, const double *Q, const double *Qinv_1,
                         int mhit, int *outB, int *outN, double *outz, unsigned long long *counters);

/* fixed-point scale used for the sojourn totals (same rule as the engine) */
int pho_choose_zbits(double sum_y);

/* a1/a13/a14: the whole Gibbs routine with the LJMA_Gibbs argument meaning
 * (src/PHT_MCMC_Aslett.c:72-103) plus the seed and the shard (rank, world) --
 * world > 1 returns the packed partial statistics of the shard instead of
 * updating, see pho_sweep_stats. */
int pho_gibbs(uint64_t seed, int it, int mhit, int method, int n, int m, const double *nu, const double *zeta,
              const int *T, const double *C, const double *y, long l, const int *censored,
              const double *start, double *res, unsigned long long *counters);

/* one sweep's sufficient statistics for the shard {obs : obs % world == rank}:
 * Nacc[n*n] and Bacc[n] as int64, zfix[n] as int64 fixed point with zbits
 * fractional bits; theta is the current parameter vector (length m). */
int pho_sweep_stats(uint64_t seed, uint32_t iter, int first, int mhit, int method, int n, int m, const int *T, const double *C,
                    const double *theta, const double *y, long l, const int *censored, int rank, int world,
                    int zbits, long long *Nacc, long long *Bacc, long long *zfix, unsigned long long *counters);
/* `first` != 0: the generator is assembled from start values (diagonal summed over ascending columns,
 * src/PHT_MCMC_Aslett.c:215,229); 0: after an update (descending, :389-393) */
int pho_eigen_native(int n, const double *S, double *evals, double *Q, double *Qinv);

/* conjugate update from (allreduced) statistics: writes theta_new (length m) */
int pho_update(uint64_t seed, uint32_t iter, int n, int m, const double *nu, const double *zeta, const int *T,
               const double *C, int zbits, const long long *Nacc, const long long *zfix, double *theta_new);

#ifdef __cplusplus
}
#endif
#endif
