/*
 * k_mhrs.cu -- Bladt MHRS path sampler for sm_100a (method bit 1).
 *
 * What it computes (reference: per-observation independence Metropolis-Hastings over
 * rejection-sampled paths, src/Simulate_AbsCTMC_eq_Bladt_MHRS.c:63-110 calling
 * src/Simulate_AbsCTMC_gt_Bladt_MHRS.c:34-160): for every observation y, forward-simulate
 * the unconditioned chain until an attempt survives past y, optionally MH-swap against
 * `mhit` further draws, and add the accepted path's (B, N, z) to the sweep statistics.
 *
 * How it is organised on the GPU (nothing like the reference's loop nest):
 *   - Every rejection attempt is its own Philox sub-stream (pht_philox.h), so attempts
 *     are independent work items and "the first attempt that survives" is well defined
 *     whatever order they run in.
 *   - SEARCH / REPLAY split: while searching, a lane tracks only (t, state); the 26 of 27
 *     attempts that fail never touch N or z.  The accepted attempt is replayed once from
 *     its (attempt index, offset) with recording on.  No per-attempt zeroing of N2/z2.
 *   - One jump-step is one Philox block: {state uniform of this jump, exponential uniform
 *     of the next}; the attempt's first step uses the same code with the start
 *     distribution as the scan row, and a FAILED attempt restarts the lane on the next
 *     sub-stream inside the same branch-free step, so fresh, running, failing, searching
 *     and replaying lanes execute one instruction stream.  Only the rare events (a
 *     surviving attempt, the end of a replay, a hand-over to the tail) leave it.
 *   - The categorical scan `while (sofar < target) sofar += p[k++]` is answered from
 *     precomputed running sums: a 64-bucket guide table indexed by the top 6 random bits
 *     gives a lower bound of the answer, one or two comparisons finish it (same result as
 *     the sequential scan because the running sums are non-decreasing).
 *   - Lane phase: persistent warps, one observation per lane, refilled from a global
 *     counter in warp-sized chunks as lanes finish (attempt counts are geometric with a
 *     heavy tail).  A lane gives up after `cap` attempts and appends the
 *     observation to the tail list.
 *   - Tail phase (same cooperative launch, grid barriers between rounds): all lanes of the
 *     GPU search each pending observation's attempts in parallel, 32-attempt runs handed
 *     out from a global counter, chunk size doubling per round; atomicMin keeps the first
 *     surviving attempt; one thread per observation then advances its MH state machine.
 *   - Statistics: N, B in shared-memory integer atomics; z per path in a per-lane shared
 *     slab (bit-identical to the reference's z2), then added as int64 fixed point, so the
 *     sweep totals do not depend on scheduling or on the number of GPUs.
 *
 * Roofline: FP64/issue bound (one log + ~n compares per jump-step, 12 B of HBM per path).
 */
#include <cooperative_groups.h>
#include "engine_internal.h"
#include "pht_philox.h"

namespace cg = cooperative_groups;

#define MHRS_THREADS 256
#ifndef MHRS_MIN_BLOCKS
#define MHRS_MIN_BLOCKS 3                /* 80 registers: 24 warps per SM */
#endif
#define RUN_LEN 4u                  /* attempts per tail work unit (short: a round ends when its slowest run does) */
#define GUIDE 64                    /* buckets of the scan guide table */
#define TAIL_K0 1024u               /* attempts per pending observation in tail round 0 */
#define TAIL_KMAX (1u << 24)
#define OBS_CHUNK 64u               /* observations a warp takes from the global counter at once */
#define FOUND_NONE 0xFFFFFFFFFFFFFFFFull

enum { IDLE = 0, SEARCH = 1, REPLAY = 2 };

struct Smem {
    double *scale, *s, *cum, *z2;
    long long *zacc; unsigned int *Nacc, *Bacc;
    unsigned char *guide;            /* (n+1) x GUIDE lower bounds of the scan result */
};

__device__ __forceinline__ unsigned long long gtimer() {
    unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t;
}

__device__ __forceinline__ Smem carve(unsigned char *raw, int n) {
    Smem sm; double *d = reinterpret_cast<double *>(raw);
    sm.scale = d; d += n;
    sm.s = d; d += n;
    sm.cum = d; d += (n + 1) * (n + 1);
    sm.z2 = d; d += n * MHRS_THREADS;
    sm.zacc = reinterpret_cast<long long *>(d); d += n;
    sm.Nacc = reinterpret_cast<unsigned int *>(d);
    sm.Bacc = sm.Nacc + n * n;
    sm.guide = reinterpret_cast<unsigned char *>(sm.Bacc + n);
    return sm;
}
size_t pht_mhrs_smem_bytes(int n) {
    return sizeof(double) * (size_t)(2 * n + (n + 1) * (n + 1) + n * MHRS_THREADS + n) + sizeof(unsigned int) * (size_t)(n * n + n) + (size_t)(n + 1) * GUIDE;
}

struct Lane {
    double y, t, lastt, spare;
    uint32_t spare_hi;          /* high random word behind `spare` (guide-table bucket) */
    uint32_t obs_local, obs_global;
    uint32_t a, b;              /* attempt (= sub-stream) and next Philox block inside it */
    uint32_t a_end;             /* tail phase: first attempt index beyond this lane's run */
    uint32_t cur_a, tries;
    int j, B;
    int mode;
    int cur_pre, kprop;
    bool cens, odd, fresh, off, have_cur, cur_off;
};

/* position the lane on the first draw of attempt L.a; off = the MH accept uniform of the
 * previous proposal occupies draw 0 of this sub-stream (src/Simulate_AbsCTMC_eq_Bladt_MHRS.c:79) */
__device__ __forceinline__ void begin_attempt(Lane &L, const SweepParams &p, uint32_t iter) {
    L.fresh = true; L.b = 0; L.odd = false;
    if (L.off) {
        pht_u32x4 r = pht_philox4x32_10(0u, L.a, L.obs_global, iter, p.k0, p.k1);
        L.spare = pht_u01(r.v[2], r.v[3]); L.spare_hi = r.v[3]; L.odd = true; L.b = 1;
    }
}

/* log of a uniform in (0,1): the normal, positive branch of pht_log (bit-identical on that domain) */
__device__ __forceinline__ double log_unit(double x) {
    const double LN2_HI = 6.93147180369123816490e-01, LN2_LO = 1.90821492927058770002e-10;
    const double L1 = 6.666666666666735130e-01, L2 = 3.999999999940941908e-01, L3 = 2.857142874366239149e-01,
                 L4 = 2.222219843214978396e-01, L5 = 1.818357216161805012e-01, L6 = 1.531383769920937332e-01,
                 L7 = 1.479819860511658591e-01;
    const uint64_t ux = pht_d2u(x);
    uint32_t hx = (uint32_t)(ux >> 32);
    hx += 0x3ff00000u - 0x3fe6a09eu;
    const int e = (int)(hx >> 20) - 0x3ff;
    hx = (hx & 0x000fffffu) + 0x3fe6a09eu;
    const double m = pht_u2d(((uint64_t)hx << 32) | (ux & 0xffffffffULL));
    const double f = m - 1.0;
    const double hfsq = 0.5 * f * f;
    const double s = f / (2.0 + f);
    const double z = s * s;
    const double w = z * z;
    const double t1 = w * PHT_FMA(w, PHT_FMA(w, L6, L4), L2);
    const double t2 = z * PHT_FMA(w, PHT_FMA(w, PHT_FMA(w, L7, L5), L3), L1);
    const double R = t2 + t1;
    const double dk = (double)e;
    return dk * LN2_HI - ((hfsq - (s * (hfsq + R) + dk * LN2_LO)) - f);
}

/* One jump-step.  Returns true when the lane needs the (divergent) event handler: a searching attempt that
 * survived, a replay that finished, or a failed attempt that must not simply restart (`stop_on_fail`, or the
 * lane's run of attempts [a, a_end) is used up).  A failed attempt otherwise restarts on sub-stream a+1 right
 * here.  No control flow around the expensive parts (Philox, scan, log): every lane of the warp runs the same
 * instructions and differs only in selects. */
__device__ __forceinline__ bool jump_step(Lane &L, const SweepParams &p, const Smem &sm, uint32_t iter, int n, bool stop_on_fail,
                                          unsigned long long &c_attempts) {
    pht_u32x4 r = pht_philox4x32_10(L.b, L.a, L.obs_global, iter, p.k0, p.k1);
    L.b++;
    const double f = pht_u01(r.v[0], r.v[1]), g = pht_u01(r.v[2], r.v[3]);
    const double uA = L.odd ? L.spare : f;               /* start state / next state */
    const uint32_t hiA = L.odd ? L.spare_hi : r.v[1];
    const double uB = L.odd ? f : g;                     /* next exponential */
    L.spare = g; L.spare_hi = r.v[3];
    const bool fresh = L.fresh;
    const int row = fresh ? n : L.j;
    const int last = fresh ? n - 1 : n;
    const double *c = sm.cum + row * (n + 1);
    /* reference scan `while (sofar < target) sofar += p[k++]` = number of running sums below the target among the
     * first `last` (the sums are non-decreasing).  The guide entry counts the sums <= bucket floor < uA. */
    int k = sm.guide[row * GUIDE + (hiA >> 26)];
    while (k < last && c[k] < uA) k++;
    const bool cont = (k < n) && (L.t < L.y || L.cens);             /* gt_Bladt_MHRS.c:75,111 */
    const bool ended = !fresh && !cont;
    const bool advance = !fresh && cont;
    if (advance && L.mode == REPLAY) {
        sm.z2[L.j * MHRS_THREADS + threadIdx.x] += L.t - L.lastt;                  /* :112 */
        if (p.outN != nullptr) p.outN[(size_t)(L.obs_local - p.first) * n * n + L.j + k * n]++;
        else atomicAdd(&sm.Nacc[L.j + k * n], 1u);                                 /* :113 */
    }
    const double tb = fresh ? 0.0 : L.t;
    L.lastt = fresh ? 0.0 : (advance ? L.t : L.lastt);
    L.j = ended ? L.j : k;
    L.B = fresh ? k : L.B;
    const double tn = tb + sm.scale[L.j] * (-log_unit(uB));         /* :80, rexp(1/-S_jj) */
    L.t = ended ? L.t : tn;
    /* an ended searching attempt survives when it reached y in a state that can exit (eq_Bladt_MHRS.c:66,74) */
    const bool searching = L.mode == SEARCH;
    const bool ok = (L.t >= L.y) && (sm.s[L.j] != 0.0);
    const bool failed = ended && searching && !ok;
    c_attempts += (ended && searching) ? 1ull : 0ull;
    L.tries += failed ? 1u : 0u;
    const bool restart = failed && !stop_on_fail && (L.a + 1u < L.a_end);
    /* restart on the next sub-stream (begin_attempt with off = false) */
    L.a += failed ? 1u : 0u;
    L.off = failed ? false : L.off;
    L.fresh = restart;
    L.b = restart ? 0u : L.b;
    L.odd = restart ? false : L.odd;
    return ended && !restart;
}

/* close a replayed path: gt_Bladt_MHRS.c:135-137, then eq_Bladt_MHRS.c:104-110 */
__device__ __forceinline__ void finish_replay(Lane &L, const SweepParams &p, const Smem &sm, int n, long out_idx) {
    const int tid = threadIdx.x;
    sm.z2[L.j * MHRS_THREADS + tid] += (L.cens ? L.t : L.y) - L.lastt;
    if (p.outB != nullptr) {
        p.outB[out_idx] = L.B;
        p.outN[out_idx * n * n + L.j + L.j * n]++;
        for (int i = 0; i < n; i++) p.outz[out_idx * n + i] = sm.z2[i * MHRS_THREADS + tid];
    } else {
        atomicAdd(&sm.Nacc[L.j + L.j * n], 1u);
        atomicAdd(&sm.Bacc[L.B], 1u);
        const double zs = pht_u2d((uint64_t)(1023 + p.zbits) << 52);
        for (int i = 0; i < n; i++) {
            const double v = sm.z2[i * MHRS_THREADS + tid];
            if (v != 0.0) {
                if (!(v * zs < 4.0e18)) atomicOr(&p.state->error, 2);
                atomicAdd(reinterpret_cast<unsigned long long *>(&sm.zacc[i]), (unsigned long long)__double2ll_rn(v * zs));
            }
        }
    }
}

__device__ __forceinline__ void start_replay(Lane &L, const SweepParams &p, const Smem &sm, uint32_t iter, int n) {
    L.mode = REPLAY; L.a = L.cur_a; L.off = L.cur_off; L.a_end = 0xFFFFFFFFu;
    begin_attempt(L, p, iter);
    for (int i = 0; i < n; i++) sm.z2[i * MHRS_THREADS + threadIdx.x] = 0.0;
}

__device__ __forceinline__ uint32_t pack_flags(const Lane &L) {
    return (L.have_cur ? 1u : 0u) | (L.cur_off ? 2u : 0u) | (L.off ? 4u : 0u) |
           ((uint32_t)(L.cur_pre & 0xff) << 8) | ((uint32_t)L.kprop << 16);
}

__global__ void __launch_bounds__(MHRS_THREADS, MHRS_MIN_BLOCKS) k_mhrs_sweep(SweepParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    cg::grid_group grid = cg::this_grid();
    const int n = p.n, tid = threadIdx.x, lane = tid & 31;
    const unsigned FULL = 0xffffffffu;
    const ModelLayout ML = ModelLayout::make(n, p.m);
    Smem sm = carve(smem_raw, n);
    const uint32_t iter = p.state->iter;

    for (int i = tid; i < n; i += MHRS_THREADS) { sm.scale[i] = p.model[ML.scale + i]; sm.s[i] = p.model[ML.s + i]; sm.zacc[i] = 0; sm.Bacc[i] = 0u; }
    for (int i = tid; i < (n + 1) * (n + 1); i += MHRS_THREADS) sm.cum[i] = p.model[ML.cum + i];
    for (int i = tid; i < n * n; i += MHRS_THREADS) sm.Nacc[i] = 0u;
    __syncthreads();
    /* guide[row][b] = #{ i < last(row) : cum[row][i] <= b / GUIDE }: every uniform of bucket b is > b / GUIDE */
    for (int e = tid; e < (n + 1) * GUIDE; e += MHRS_THREADS) {
        const int row = e / GUIDE, b = e % GUIDE, last = (row == n) ? n - 1 : n;
        const double floor_u = (double)b * (1.0 / GUIDE);
        int k = 0;
        while (k < last && sm.cum[row * (n + 1) + k] <= floor_u) k++;
        sm.guide[e] = (unsigned char)k;
    }
    __syncthreads();

    unsigned long long c_attempts = 0, c_jumps = 0, c_paths = 0, c_deferred = 0;
    const unsigned long long t_start = gtimer();
    const bool per_obs = p.outB != nullptr;
    const unsigned long long obs_begin = per_obs ? (unsigned long long)p.first : 0ull;
    const unsigned long long obs_end = per_obs ? (unsigned long long)(p.first + p.count) : (unsigned long long)p.l_local;
    const uint32_t cap = (uint32_t)p.mhrs_cap;

    /* ---------------------------------------------------------------- lane phase */
    {
        Lane L; L.mode = IDLE; L.tries = 0; L.a_end = 0xFFFFFFFFu;
        unsigned long long chunk_next = 0, chunk_end = 0;      /* warp-uniform */
        bool exhausted = false;
        for (;;) {
            unsigned idle = __ballot_sync(FULL, L.mode == IDLE);
            if (idle && !exhausted) {
                if (chunk_next == chunk_end) {
                    unsigned long long base = 0;
                    if (lane == 0) base = obs_begin + atomicAdd(&p.state->next_obs, (unsigned long long)OBS_CHUNK);
                    base = __shfl_sync(FULL, base, 0);
                    chunk_next = base < obs_end ? base : obs_end;
                    chunk_end = base + OBS_CHUNK < obs_end ? base + OBS_CHUNK : obs_end;
                    if (chunk_next == chunk_end) exhausted = true;
                }
                const unsigned avail = (unsigned)(chunk_end - chunk_next);
                const unsigned rank = __popc(idle & ((1u << lane) - 1u));
                if (L.mode == IDLE && rank < avail) {
                    L.obs_local = (uint32_t)(chunk_next + rank);
                    L.obs_global = p.obs_rank + L.obs_local * p.obs_world;
                    L.y = p.y[L.obs_local]; L.cens = p.cens[L.obs_local] != 0;
                    L.mode = SEARCH; L.a = 0; L.off = false; L.have_cur = false; L.kprop = 0; L.tries = 0;
                    L.cur_a = 0; L.cur_off = false; L.cur_pre = 0; L.a_end = 0xFFFFFFFFu;
                    begin_attempt(L, p, iter);
                }
                const unsigned taken = __popc(idle) < avail ? __popc(idle) : avail;
                chunk_next += taken;
                idle = __ballot_sync(FULL, L.mode == IDLE);
            }
            if (idle == FULL) { if (exhausted) break; else continue; }
            if (L.mode == IDLE) continue;

            /* a failed attempt restarts inside the step unless the lane is due to hand the observation over: after
             * `cap` attempts, or as soon as the observation stream has run dry (a lone lane grinding through
             * attempts would hold the whole grid at the barrier) */
            const bool stop_on_fail = cap != 0u && (L.tries + 1u >= cap || exhausted);
            const bool event = jump_step(L, p, sm, iter, n, stop_on_fail, c_attempts);
            c_jumps++;
            if (!event) continue;

            if (L.mode == REPLAY) {
                finish_replay(L, p, sm, n, (long)L.obs_local - p.first);
                c_paths++; L.mode = IDLE;
                continue;
            }
            /* SEARCH: an attempt just ended (gt_Bladt_MHRS.c:49 decides whether it survives) */
            const bool ok = (L.t >= L.y) && (sm.s[L.j] != 0.0);            /* eq_Bladt_MHRS.c:66,74 */
            if (!ok) {
                /* (the step already moved the lane to attempt a+1, off = false) */
                if (stop_on_fail) {
                    const uint32_t idx = atomicAdd(&p.state->n_items, 1u);
                    if (idx < p.item_cap) {
                        TailItem it; it.obs_local = L.obs_local; it.a = L.a; it.cur_a = L.cur_a; it.flags = pack_flags(L);
                        p.items[idx] = it; p.found[idx] = FOUND_NONE;
                    } else atomicOr(&p.state->error, 4);
                    c_deferred++; L.mode = IDLE;
                } else begin_attempt(L, p, iter);
                continue;
            }
            const int pre = L.j;
            if (!L.have_cur) {
                L.have_cur = true; L.cur_a = L.a; L.cur_off = L.off; L.cur_pre = pre;
                if (L.cens || p.mhit == 0) { start_replay(L, p, sm, iter, n); continue; }      /* :70 */
                L.a++; L.off = false; begin_attempt(L, p, iter);
                continue;
            }
            /* a valid proposal: accept test with draw 0 of the next sub-stream (:79-82) */
            pht_u32x4 r = pht_philox4x32_10(0u, L.a + 1u, L.obs_global, iter, p.k0, p.k1);
            const double U = pht_u01(r.v[0], r.v[1]);
            if (U < sm.s[pre] / sm.s[L.cur_pre]) { L.cur_a = L.a; L.cur_off = L.off; L.cur_pre = pre; }
            L.kprop++;
            if (L.kprop >= p.mhit) { start_replay(L, p, sm, iter, n); continue; }
            /* next proposal: sub-stream a+1 from draw 1 (the spare half of the block just computed) */
            L.a++; L.off = true;
            L.fresh = true; L.b = 1; L.odd = true; L.spare = pht_u01(r.v[2], r.v[3]); L.spare_hi = r.v[3];
        }
    }

    /* ---------------------------------------------------------------- tail phase */
    const bool timekeeper = (blockIdx.x == 0 && tid == 0);
    grid.sync();
    unsigned long long t_mark = 0;
    if (timekeeper) { t_mark = gtimer(); atomicAdd(&p.state->counters[PHT_CNT_NS_LANE], t_mark - t_start); }
    const uint32_t n_items = p.state->n_items < p.item_cap ? p.state->n_items : p.item_cap;
    if (n_items != 0u) {
        const unsigned long long gtid = (unsigned long long)blockIdx.x * MHRS_THREADS + tid;
        const unsigned long long gsize = (unsigned long long)gridDim.x * MHRS_THREADS;
        for (unsigned long long i = gtid; i < n_items; i += gsize) p.pend0[i] = (uint32_t)i;
        if (gtid == 0) { p.state->n_pend[0] = n_items; p.state->n_pend[1] = 0u; p.state->unit_counter = 0ull; p.state->n_done = 0u; }
        grid.sync();
        uint32_t K = TAIL_K0; int cur = 0; unsigned rounds = 0;
        for (;;) {
            const uint32_t P = p.state->n_pend[cur];
            if (P == 0u) break;
            const uint32_t *pend = cur ? p.pend1 : p.pend0;
            uint32_t *pend_next = cur ? p.pend0 : p.pend1;
            /* run length: RUN_LEN attempts per work unit while there is plenty of work, down to single attempts when
             * only a few heavy observations remain */
            uint32_t rl = RUN_LEN;
            while (rl > 1u && (unsigned long long)P * (K / rl) < 2ull * gsize) rl >>= 1;
            const unsigned long long rpi = K / rl, total_runs = (unsigned long long)P * rpi;
            /* --- search: lanes take runs of attempts; the first surviving attempt wins */
            {
                Lane L; L.mode = IDLE; L.tries = 0;
                uint32_t item = 0; bool out_of_runs = false;
                for (;;) {
                    unsigned idle = __ballot_sync(FULL, L.mode == IDLE);
                    if (idle && !out_of_runs) {
                        unsigned long long base = 0;
                        if (lane == 0) base = atomicAdd(&p.state->unit_counter, (unsigned long long)__popc(idle));
                        base = __shfl_sync(FULL, base, 0);
                        if (base >= total_runs) out_of_runs = true;
                        const unsigned long long run = base + __popc(idle & ((1u << lane) - 1u));
                        if (L.mode == IDLE && run < total_runs) {
                            /* chunk-major order: the first chunk of every pending observation is handed out before
                             * any second chunk, so later chunks are mostly skipped once an earlier one has survived */
                            item = pend[run % P];
                            const TailItem it = p.items[item];
                            L.obs_local = it.obs_local; L.obs_global = p.obs_rank + it.obs_local * p.obs_world;
                            L.y = p.y[it.obs_local]; L.cens = p.cens[it.obs_local] != 0;
                            L.a = it.a + (uint32_t)(run / P) * rl; L.a_end = L.a + rl;
                            L.off = (L.a == it.a) && (it.flags & 4u);
                            /* skip the run when an earlier attempt has already survived (checked once per run) */
                            if ((__ldcg(&p.found[item]) >> 8) >= (unsigned long long)L.a) { L.mode = SEARCH; begin_attempt(L, p, iter); }
                        }
                        idle = __ballot_sync(FULL, L.mode == IDLE);
                    }
                    if (idle == FULL) { if (out_of_runs) break; else continue; }
                    if (L.mode == IDLE) continue;
                    const bool event = jump_step(L, p, sm, iter, n, false, c_attempts);
                    c_jumps++;
                    if (!event) continue;
                    /* the run's last attempt failed, or an attempt survived */
                    if ((L.t >= L.y) && (sm.s[L.j] != 0.0))
                        atomicMin(&p.found[item], ((unsigned long long)L.a << 8) | (unsigned long long)L.j);
                    L.mode = IDLE;
                }
            }
            grid.sync();
            /* --- advance each pending observation's MH state machine (eq_Bladt_MHRS.c:65-101) */
            for (unsigned long long i = gtid; i < P; i += gsize) {
                const uint32_t item = pend[i];
                TailItem it = p.items[item];
                const unsigned long long f = p.found[item];
                bool done = false;
                bool dropped = false;
                if (f == FOUND_NONE) {
                    if (it.a > 0xF0000000u - K) { atomicOr(&p.state->error, 8); dropped = true; }   /* survival probability ~ 0 */
                    it.a += K; it.flags &= ~4u;
                }
                else {
                    const uint32_t a = (uint32_t)(f >> 8); const int pre = (int)(f & 0xffull);
                    const bool off = (a == it.a) && (it.flags & 4u);
                    const bool cens = p.cens[it.obs_local] != 0;
                    bool have_cur = it.flags & 1u; bool cur_off = it.flags & 2u;
                    int cur_pre = (int)((it.flags >> 8) & 0xffu); uint32_t kprop = it.flags >> 16;
                    bool next_off = false;
                    if (!have_cur) {
                        have_cur = true; it.cur_a = a; cur_off = off; cur_pre = pre;
                        done = cens || p.mhit == 0;
                    } else {
                        const uint32_t og = p.obs_rank + it.obs_local * p.obs_world;
                        pht_u32x4 r = pht_philox4x32_10(0u, a + 1u, og, iter, p.k0, p.k1);
                        const double U = pht_u01(r.v[0], r.v[1]);
                        if (U < sm.s[pre] / sm.s[cur_pre]) { it.cur_a = a; cur_off = off; cur_pre = pre; }
                        kprop++; done = (int)kprop >= p.mhit; next_off = true;
                    }
                    it.a = a + 1u;
                    it.flags = (have_cur ? 1u : 0u) | (cur_off ? 2u : 0u) | (next_off ? 4u : 0u) |
                               ((uint32_t)(cur_pre & 0xff) << 8) | (kprop << 16);
                    p.found[item] = FOUND_NONE;
                }
                p.items[item] = it;
                if (dropped) continue;
                if (done) p.done[atomicAdd(&p.state->n_done, 1u)] = item;
                else pend_next[atomicAdd(&p.state->n_pend[cur ^ 1], 1u)] = item;
            }
            if (gtid == 0) { p.state->n_pend[cur] = 0u; p.state->unit_counter = 0ull; }
            grid.sync();
            cur ^= 1; rounds++;
            K = (K < TAIL_KMAX) ? K * 2u : K;
        }
        if (timekeeper) { const unsigned long long t = gtimer(); atomicAdd(&p.state->counters[PHT_CNT_NS_TAIL], t - t_mark); t_mark = t; }
        /* --- replay the accepted attempt of every tail observation */
        {
            const uint32_t n_done = p.state->n_done;
            for (unsigned long long i = gtid; i < n_done; i += gsize) {
                const TailItem it = p.items[p.done[i]];
                Lane L; L.tries = 0;
                L.obs_local = it.obs_local; L.obs_global = p.obs_rank + it.obs_local * p.obs_world;
                L.y = p.y[it.obs_local]; L.cens = p.cens[it.obs_local] != 0;
                L.cur_a = it.cur_a; L.cur_off = it.flags & 2u;
                start_replay(L, p, sm, iter, n);
                unsigned long long dummy = 0;
                while (!jump_step(L, p, sm, iter, n, false, dummy)) c_jumps++;
                c_jumps++;
                finish_replay(L, p, sm, n, (long)L.obs_local - p.first);
                c_paths++;
            }
        }
        if (gtid == 0) atomicAdd(&p.state->counters[PHT_CNT_TAIL_ROUNDS], (unsigned long long)rounds);
        if (timekeeper) atomicAdd(&p.state->counters[PHT_CNT_NS_REPLAY], gtimer() - t_mark);
    }

    /* ---------------------------------------------------------------- block -> global statistics */
    __syncthreads();
    if (!per_obs) {
        unsigned long long *gN = reinterpret_cast<unsigned long long *>(p.stats);
        for (int i = tid; i < n * n; i += MHRS_THREADS) if (sm.Nacc[i]) atomicAdd(&gN[i], (unsigned long long)sm.Nacc[i]);
        for (int i = tid; i < n; i += MHRS_THREADS) {
            if (sm.Bacc[i]) atomicAdd(&gN[n * n + i], (unsigned long long)sm.Bacc[i]);
            if (sm.zacc[i]) atomicAdd(&gN[n * n + n + i], (unsigned long long)sm.zacc[i]);
        }
    }
    for (int o = 16; o > 0; o >>= 1) {
        c_attempts += __shfl_down_sync(FULL, c_attempts, o); c_jumps += __shfl_down_sync(FULL, c_jumps, o);
        c_paths += __shfl_down_sync(FULL, c_paths, o); c_deferred += __shfl_down_sync(FULL, c_deferred, o);
    }
    if (lane == 0) {
        atomicAdd(&p.state->counters[PHT_CNT_ATTEMPTS], c_attempts); atomicAdd(&p.state->counters[PHT_CNT_JUMPS], c_jumps);
        atomicAdd(&p.state->counters[PHT_CNT_PATHS], c_paths); atomicAdd(&p.state->counters[PHT_CNT_DEFERRED], c_deferred);
    }
}

int pht_mhrs_grid_blocks(int device, int n) {
    int per_sm = 0, sms = 0;
    const size_t smem = pht_mhrs_smem_bytes(n);
    if (cudaFuncSetAttribute(k_mhrs_sweep, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return -1;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_mhrs_sweep, MHRS_THREADS, smem) != cudaSuccess) return -1;
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device) != cudaSuccess) return -1;
    return per_sm * sms;
}

cudaError_t pht_launch_mhrs(const SweepParams &p, int grid_blocks, cudaStream_t st) {
    SweepParams q = p;
    void *args[] = { &q };
    return cudaLaunchCooperativeKernel((const void *)k_mhrs_sweep, dim3(grid_blocks), dim3(MHRS_THREADS), args,
                                       pht_mhrs_smem_bytes(p.n), st);
}
