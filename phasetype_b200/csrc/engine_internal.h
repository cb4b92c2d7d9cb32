/*
 * engine_internal.h -- data layout shared by the engine (engine.cu) and its kernels.
 *
 * HBM layout (all per engine = per GPU):
 *   y[l_local]      fp64   observations of this rank's shard (global index rank + k*world)
 *   cens[l_local]   uint8  right-censoring flags (the ABI's int32 is repacked at upload)
 *   model[]         fp64   the per-sweep model block, offsets from ModelLayout (< 70 KB at n = 32);
 *                          rebuilt on the device every sweep, copied to shared memory by the path kernels
 *   stats[]         int64  N (n*n) | B (n) | z fixed point (n): the only data that crosses NVLink
 *   state           DevState: sweep index, work counters, event counters, error word
 *   items/pend/...  MHRS tail work lists (see k_mhrs.cu)
 */
#ifndef PHT_ENGINE_INTERNAL_H
#define PHT_ENGINE_INTERNAL_H

#include <stdint.h>
#include <cuda_runtime.h>
#include "../../include/pht_b200.h"
#include "pht_philox.h"

#define PHT_NMAX PHT_MAX_PHASES

/* offsets in doubles inside the model block */
struct ModelLayout {
    int n, m;
    int S, s, scale, P, Pfull, cum, pi, evals, Q, Qinv, Qinv_s, Qinv_1, TT, theta, total;
    __host__ __device__ static ModelLayout make(int n, int m) {
        ModelLayout L; int o = 0;
        L.n = n; L.m = m;
        L.S = o; o += n * n;
        L.s = o; o += n;
        L.scale = o; o += n;                 /* 1.0 / -S[j,j]: the rexp() scale of state j */
        L.P = o; o += n * n;
        L.Pfull = o; o += n * (n + 1);
        L.cum = o; o += (n + 1) * (n + 1);   /* row j<n: running sums of Pfull[j,0..n]; row n: running sums of pi */
        L.pi = o; o += n;
        L.evals = o; o += n;
        L.Q = o; o += n * n;
        L.Qinv = o; o += n * n;
        L.Qinv_s = o; o += n;
        L.Qinv_1 = o; o += n;
        L.TT = o; o += (n + 1) * (n + 1);
        L.theta = o; o += m;
        L.total = o;
        return L;
    }
};

/* stats block: int64 [ N: n*n | B: n | zfix: n ] */
__host__ __device__ inline int stats_len(int n) { return n * n + 2 * n; }

/* one MHRS observation handed from the lane phase to the cooperative tail */
struct TailItem {
    uint32_t obs_local;   /* index into y/cens */
    uint32_t a;           /* first attempt index of the current search */
    uint32_t cur_a;       /* attempt index of the accepted draw so far */
    uint32_t flags;       /* bit0 have_cur, bit1 cur_off, bit2 off (for attempt a), bits 8..15 cur_pre, bits 16..31 proposals done */
};

struct DevState {
    uint32_t iter;             /* index of the sweep about to run (Philox counter word 3) */
    uint32_t first_assembly;   /* 1: diagonals summed in ascending column order (start values) */
    uint32_t res_row;          /* next row of the device result buffer */
    int error;                 /* sticky error word */
    unsigned long long next_obs;      /* lane-phase work counter */
    /* MHRS tail */
    unsigned long long unit_counter;
    uint32_t n_items;          /* items appended by the lane phase */
    uint32_t n_pend[2];        /* pending list sizes (double buffered) */
    uint32_t n_done;           /* finished items awaiting replay */
    uint32_t any_fail;         /* MHRS tail: some pending observation found no surviving attempt in the current round */
    unsigned long long counters[PHT_CNT_COUNT];
};

struct SweepParams {
    /* data */
    const double *y; const uint8_t *cens; long l_local; uint32_t obs_rank, obs_world;
    /* model + state */
    double *model; long long *stats; DevState *state;
    int n, m, mhit, zbits;
    uint32_t k0, k1;           /* Philox key */
    pht_roundkeys rk;          /* its ten round keys (constant-bank operands of the path kernels) */
    /* MHRS tail lists */
    TailItem *items; uint32_t *pend0, *pend1, *done; unsigned long long *found; uint32_t item_cap;
    int mhrs_cap;
    /* per-observation recording (parity mode); all NULL in production */
    int *outB, *outN; double *outz; long first, count;
};

struct UpdateParams {
    double *model; long long *stats; DevState *state; double *res; int res_rows;
    const int *T; const double *C; const double *nu; const double *zeta;
    const int *var_ptr; const int *cell_i; const int *cell_j;   /* CSR: parameter -> cells in reference insertion order */
    int n, m, zbits; uint32_t k0, k1;
};

/* kernel launchers (each returns the cudaError of the launch) */
cudaError_t pht_launch_assemble(const UpdateParams &p, cudaStream_t st);
cudaError_t pht_launch_update(const UpdateParams &p, cudaStream_t st);
cudaError_t pht_launch_mhrs(const SweepParams &p, int grid_blocks, cudaStream_t st);
cudaError_t pht_launch_dcs(const SweepParams &p, int grid_blocks, cudaStream_t st);
int pht_dcs_grid_blocks(int device, int n);
/* ECS: exact and right-censored observations are separate launches over index lists (nullptr = identity) */
cudaError_t pht_launch_ecs(const SweepParams &p, int grid_blocks, const uint32_t *idx_exact, unsigned long long n_exact,
                           const uint32_t *idx_cens, unsigned long long n_cens, cudaStream_t st);
int pht_ecs_grid_blocks(int device, int n);
/* spectral data of the sweep: inject != nullptr copies host-supplied (evals | Q | Qinv) instead of solving on the device */
cudaError_t pht_launch_spectral(const UpdateParams &p, const double *inject, cudaStream_t st);
int pht_mhrs_grid_blocks(int device, int n);
size_t pht_mhrs_smem_bytes(int n);

#endif
