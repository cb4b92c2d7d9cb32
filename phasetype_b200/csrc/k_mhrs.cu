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
 *   - SEARCH / REPLAY split: while searching, a lane tracks only (t, state); the 21 of 22
 *     attempts that fail never touch N or z.  The accepted attempt is replayed once from
 *     its (attempt index, offset) with recording on.  No per-attempt zeroing of N2/z2.
 *   - One jump-step (walk_step) is one Philox block: {state uniform of this jump,
 *     exponential uniform of the next}; the attempt's first step uses the same code with
 *     the start distribution as the scan row, and a failed attempt restarts the lane on
 *     the next sub-stream with a handful of selects, so fresh, running and failing lanes
 *     execute one instruction stream.  Only the rare events (a surviving attempt, a
 *     hand-over to the tail) leave it.
 *   - The categorical scan `while (sofar < target) sofar += p[k++]` is answered from
 *     precomputed running sums: a 64-bucket guide table indexed by the top 6 random bits
 *     gives a lower bound of the answer, one or two comparisons finish it (same result as
 *     the sequential scan because the running sums are non-decreasing).
 *   - Lane phase: persistent warps, one observation per lane, refilled from a global
 *     counter in warp-sized chunks (prefetched into shared memory) as lanes finish.
 *     Accepted attempts go to a per-warp ring in shared memory; when the ring holds two
 *     warps' worth, the whole warp replays them together, so the recording code runs
 *     converged instead of on one or two lanes at a time.  A lane gives up after `cap`
 *     attempts (attempt counts are geometric with a heavy tail) and appends
 *     the observation to the tail list.
 *   - Tail phase (same cooperative launch, grid barriers between rounds): the attempts of
 *     every pending observation are searched by all warps of the GPU.  A warp takes a pool
 *     of consecutive attempts of ONE observation from a global counter (pool size shrinks
 *     as the round drains), its lanes draw attempt indices from the pool with a ballot (no
 *     memory traffic), the lowest surviving attempt wins through atomicMin; one thread per
 *     observation then advances its MH state machine.  Attempts per observation double per
 *     round.
 *   - Statistics: N, B in shared-memory integer atomics; z per path in a per-lane shared
 *     slab (bit-identical to the reference's z2), then added as int64 fixed point, so the
 *     sweep totals do not depend on scheduling or on the number of GPUs.
 *
 * Roofline: FP64/issue bound (one log + one Philox block per jump-step, 9 B of HBM per path).
 */
#include <cooperative_groups.h>
#include "engine_internal.h"
#include "pht_philox.h"

namespace cg = cooperative_groups;

#ifndef MHRS_THREADS
#define MHRS_THREADS 256
#endif
#define MHRS_WARPS (MHRS_THREADS / 32)
#ifndef MHRS_MIN_BLOCKS
#define MHRS_MIN_BLOCKS 3                /* 80 registers: 24 warps per SM */
#endif
#define GUIDE 64                    /* buckets of the scan guide table */
#define OBS_CHUNK 64u               /* observations a warp takes from the global counter at once */
#ifndef END_CAP
#define END_CAP 8u                   /* attempts a lane still tries once no observations are left */
#endif
#define RING 96                     /* per-warp ring of accepted attempts awaiting replay */
#define RING_TRIGGER 64
#ifndef TAIL_CH
#define TAIL_CH 128u                /* attempts per tail chunk (a pool never crosses a chunk) */
#endif
#ifndef TAIL_K0
#define TAIL_K0 512u                /* attempts per pending observation in tail round 0 */
#endif
#define TAIL_KMAX (1u << 24)
#ifndef TAIL_GROWTH
#define TAIL_GROWTH 2u               /* attempts per pending observation grow by this factor per round */
#endif
#ifndef FOUND_PERIOD
#define FOUND_PERIOD 8u               /* steps between looks at the pool observation's `found` word (power of two) */
#endif
#define POOL_MIN 32u
#ifndef POOL_MAX
#define POOL_MAX 256u
#endif
#define FOUND_NONE 0xFFFFFFFFFFFFFFFFull

struct Smem {
    double *scale, *cum, *z2;
    long long *zacc; unsigned int *Nacc, *Bacc;
    double *ybuf; double *ring_y; uint32_t *ring_obs, *ring_a; unsigned int *ring_n;
    unsigned char *cbuf, *ring_fl, *guide;
};

__device__ __forceinline__ unsigned long long gtimer() {
    unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t;
}

__host__ __device__ inline size_t mhrs_smem_bytes_of(int n) {
    size_t o = 0;
    o += sizeof(double) * (size_t)(n + (n + 1) * (n + 1) + n * MHRS_THREADS + n + MHRS_WARPS * OBS_CHUNK + MHRS_WARPS * RING);
    o += sizeof(unsigned int) * (size_t)(n * n + n + 2 * MHRS_WARPS * RING + MHRS_WARPS);
    o += (size_t)(MHRS_WARPS * OBS_CHUNK + MHRS_WARPS * RING + (n + 1) * GUIDE);
    return (o + 15) & ~(size_t)15;
}
__device__ __forceinline__ Smem carve(unsigned char *raw, int n) {
    Smem sm; double *d = reinterpret_cast<double *>(raw);
    sm.scale = d; d += n;
    sm.cum = d; d += (n + 1) * (n + 1);
    sm.z2 = d; d += n * MHRS_THREADS;
    sm.zacc = reinterpret_cast<long long *>(d); d += n;
    sm.ybuf = d; d += MHRS_WARPS * OBS_CHUNK;
    sm.ring_y = d; d += MHRS_WARPS * RING;
    unsigned int *u = reinterpret_cast<unsigned int *>(d);
    sm.Nacc = u; u += n * n;
    sm.Bacc = u; u += n;
    sm.ring_obs = u; u += MHRS_WARPS * RING;
    sm.ring_a = u; u += MHRS_WARPS * RING;
    sm.ring_n = u; u += MHRS_WARPS;
    unsigned char *c = reinterpret_cast<unsigned char *>(u);
    sm.cbuf = c; c += MHRS_WARPS * OBS_CHUNK;
    sm.ring_fl = c; c += MHRS_WARPS * RING;
    sm.guide = c;
    return sm;
}
size_t pht_mhrs_smem_bytes(int n) { return mhrs_smem_bytes_of(n); }

/* Philox4x32-10 block (b, a, observation, sweep) with the round keys taken from the kernel parameters */
__device__ __forceinline__ pht_u32x4 philox_block(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, const SweepParams &p) {
    PHT_PHILOX_ROUND(p.rk.k[0], p.rk.k[1])   PHT_PHILOX_ROUND(p.rk.k[2], p.rk.k[3])   PHT_PHILOX_ROUND(p.rk.k[4], p.rk.k[5])
    PHT_PHILOX_ROUND(p.rk.k[6], p.rk.k[7])   PHT_PHILOX_ROUND(p.rk.k[8], p.rk.k[9])   PHT_PHILOX_ROUND(p.rk.k[10], p.rk.k[11])
    PHT_PHILOX_ROUND(p.rk.k[12], p.rk.k[13]) PHT_PHILOX_ROUND(p.rk.k[14], p.rk.k[15]) PHT_PHILOX_ROUND(p.rk.k[16], p.rk.k[17])
    PHT_PHILOX_ROUND(p.rk.k[18], p.rk.k[19])
    pht_u32x4 out; out.v[0] = c0; out.v[1] = c1; out.v[2] = c2; out.v[3] = c3;
    return out;
}

/* the chain of one attempt: sub-stream a, next Philox block b, time t of the pending exit from state j */
struct Walk {
    double t, spare;
    uint32_t spare_hi, a, b;
    int j;
    bool odd, fresh;
};
/* recording state of a replayed path */
struct Rec { double lastt; int B; long out_idx; };

/* position the walk on the first draw of attempt a; off = the MH accept uniform of the previous proposal
 * occupies draw 0 of this sub-stream (src/Simulate_AbsCTMC_eq_Bladt_MHRS.c:79) */
__device__ __forceinline__ void walk_begin(Walk &w, uint32_t a, bool off, uint32_t obs_global, const SweepParams &p, uint32_t iter) {
    w.a = a; w.fresh = true; w.b = 0; w.odd = false; w.t = 0.0; w.j = 0;
    if (off) {
        pht_u32x4 r = philox_block(0u, a, obs_global, iter, p);
        w.spare = pht_u01(r.v[2], r.v[3]); w.spare_hi = r.v[3]; w.odd = true; w.b = 1;
    }
}

/* log of a uniform in (0,1): always a positive normal number, so the core of pht_log applies directly */
__device__ __forceinline__ double log_unit(double x) { return pht_log_core(pht_d2u(x), 0); }

/* One jump-step of a walk; returns true when the attempt ended on this step (w.t, w.j are then the exit time
 * and the state occupied at the end).  No control flow around the expensive parts (Philox, scan, log). */
template <bool RECORD>
__device__ __forceinline__ bool walk_step(Walk &w, double y, bool cens, uint32_t obs_global, const SweepParams &p, const Smem &sm,
                                          uint32_t iter, int n, Rec &rec) {
    pht_u32x4 r = philox_block(w.b, w.a, obs_global, iter, p);
    w.b++;
    const double f = pht_u01(r.v[0], r.v[1]), g = pht_u01(r.v[2], r.v[3]);
    const double uA = w.odd ? w.spare : f;               /* start state / next state */
    const uint32_t hiA = w.odd ? w.spare_hi : r.v[1];
    const double uB = w.odd ? f : g;                     /* next exponential */
    w.spare = g; w.spare_hi = r.v[3];
    const bool fresh = w.fresh;
    const int row = fresh ? n : w.j;
    const int last = fresh ? n - 1 : n;
    const double *c = sm.cum + row * (n + 1);
    /* reference scan `while (sofar < target) sofar += p[k++]` = number of running sums below the target among the
     * first `last` (the sums are non-decreasing).  The guide entry counts the sums <= bucket floor < uA. */
    int k = sm.guide[row * GUIDE + (hiA >> 26)];
    while (k < last && c[k] < uA) k++;
    const bool cont = (k < n) && (w.t < y || cens);                /* gt_Bladt_MHRS.c:75,111 */
    const bool ended = !fresh && !cont;
    const bool advance = !fresh && cont;
    if (RECORD) {
        if (advance) {
            sm.z2[w.j * MHRS_THREADS + threadIdx.x] += w.t - rec.lastt;            /* :112 */
            if (p.outN != nullptr) p.outN[(size_t)rec.out_idx * n * n + w.j + k * n]++;
            else atomicAdd(&sm.Nacc[w.j + k * n], 1u);                             /* :113 */
        }
        rec.lastt = fresh ? 0.0 : (advance ? w.t : rec.lastt);
        rec.B = fresh ? k : rec.B;
    }
    const double tb = fresh ? 0.0 : w.t;
    const int jn = ended ? w.j : k;
    const double tn = tb + sm.scale[jn] * (-log_unit(uB));         /* :80, rexp(1/-S_jj) */
    w.t = ended ? w.t : tn;
    w.j = jn;
    w.fresh = false;
    return ended;
}

/* close a replayed path: gt_Bladt_MHRS.c:135-137, then eq_Bladt_MHRS.c:104-110 */
__device__ __forceinline__ void finish_replay(const Walk &w, const Rec &rec, double y, bool cens, const SweepParams &p, const Smem &sm, int n) {
    const int tid = threadIdx.x;
    sm.z2[w.j * MHRS_THREADS + tid] += (cens ? w.t : y) - rec.lastt;
    if (p.outB != nullptr) {
        p.outB[rec.out_idx] = rec.B;
        p.outN[rec.out_idx * n * n + w.j + w.j * n]++;
        for (int i = 0; i < n; i++) { p.outz[rec.out_idx * n + i] = sm.z2[i * MHRS_THREADS + tid]; sm.z2[i * MHRS_THREADS + tid] = 0.0; }
    } else {
        atomicAdd(&sm.Nacc[w.j + w.j * n], 1u);
        atomicAdd(&sm.Bacc[rec.B], 1u);
        const double zs = pht_u2d((uint64_t)(1023 + p.zbits) << 52);
        for (int i = 0; i < n; i++) {
            const double v = sm.z2[i * MHRS_THREADS + tid];
            if (v != 0.0) {
                if (!(v * zs < 4.0e18)) atomicOr(&p.state->error, 2);
                atomicAdd(reinterpret_cast<unsigned long long *>(&sm.zacc[i]), (unsigned long long)__double2ll_rn(v * zs));
                sm.z2[i * MHRS_THREADS + tid] = 0.0;
            }
        }
    }
}

/* replay one accepted attempt start to end (the lane's z2 slab is all zero on entry and on exit) */
__device__ __forceinline__ unsigned replay_path(uint32_t obs_local, uint32_t a, bool off, double y, bool cens, const SweepParams &p,
                                                const Smem &sm, uint32_t iter, int n) {
    Walk w; Rec rec; rec.lastt = 0.0; rec.B = 0; rec.out_idx = (long)obs_local - p.first;
    const uint32_t og = p.obs_rank + obs_local * p.obs_world;
    walk_begin(w, a, off, og, p, iter);
    unsigned steps = 1;
    while (!walk_step<true>(w, y, cens, og, p, sm, iter, n, rec)) steps++;
    finish_replay(w, rec, y, cens, p, sm, n);
    return steps;
}

/* the whole warp replays the accepted attempts waiting in its ring */
__device__ __forceinline__ void replay_session(const SweepParams &p, const Smem &sm, uint32_t iter, int n, int warp,
                                               unsigned &c_jumps, unsigned &c_paths) {
    const unsigned FULL = 0xffffffffu; const int lane = threadIdx.x & 31;
    __syncwarp();
    const unsigned count = sm.ring_n[warp];
    unsigned next = 0;
    Walk w; Rec rec; double y = 0.0; bool cens = false, active = false; uint32_t og = 0;
    rec.lastt = 0.0; rec.B = 0; rec.out_idx = 0;
    for (;;) {
        unsigned idle = __ballot_sync(FULL, !active);
        if (idle && next < count) {
            const unsigned rank = __popc(idle & ((1u << lane) - 1u));
            if (!active && next + rank < count) {
                const unsigned e = warp * RING + next + rank;
                const uint32_t ol = sm.ring_obs[e]; const unsigned fl = sm.ring_fl[e];
                y = sm.ring_y[e]; cens = fl & 1u; og = p.obs_rank + ol * p.obs_world;
                rec.lastt = 0.0; rec.B = 0; rec.out_idx = (long)ol - p.first;
                walk_begin(w, sm.ring_a[e], (fl & 2u) != 0u, og, p, iter);
                active = true;
            }
            const unsigned cnt = __popc(idle);
            next = next + cnt < count ? next + cnt : count;
            idle = __ballot_sync(FULL, !active);
        }
        if (idle == FULL) break;
        __syncwarp();
        if (active) {
            const bool ended = walk_step<true>(w, y, cens, og, p, sm, iter, n, rec);
            c_jumps++;
            if (ended) { finish_replay(w, rec, y, cens, p, sm, n); c_paths++; active = false; }
        }
    }
    __syncwarp();
    if (lane == 0) sm.ring_n[warp] = 0u;
    __syncwarp();
}

__device__ __forceinline__ uint32_t pack_flags(bool have_cur, bool cur_off, bool off, int cur_pre, int kprop) {
    return (have_cur ? 1u : 0u) | (cur_off ? 2u : 0u) | (off ? 4u : 0u) | ((uint32_t)(cur_pre & 0xff) << 8) | ((uint32_t)kprop << 16);
}

__global__ void __launch_bounds__(MHRS_THREADS, MHRS_MIN_BLOCKS) k_mhrs_sweep(SweepParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    cg::grid_group grid = cg::this_grid();
    const int n = p.n, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned FULL = 0xffffffffu;
    const ModelLayout ML = ModelLayout::make(n, p.m);
    Smem sm = carve(smem_raw, n);
    const uint32_t iter = p.state->iter;

    for (int i = tid; i < n; i += MHRS_THREADS) { sm.scale[i] = p.model[ML.scale + i]; sm.zacc[i] = 0; sm.Bacc[i] = 0u; }
    for (int i = tid; i < (n + 1) * (n + 1); i += MHRS_THREADS) sm.cum[i] = p.model[ML.cum + i];
    for (int i = tid; i < n * n; i += MHRS_THREADS) sm.Nacc[i] = 0u;
    for (int i = 0; i < n; i++) sm.z2[i * MHRS_THREADS + tid] = 0.0;
    if (tid < MHRS_WARPS) sm.ring_n[tid] = 0u;
    /* states that can exit: bit j of smask (the accept tests of eq_Bladt_MHRS.c:66,74 need s[j] != 0) */
    uint32_t smask = 0u;
    for (int i = 0; i < n; i++) smask |= (p.model[ML.s + i] != 0.0) ? (1u << i) : 0u;
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

    unsigned c_attempts = 0, c_jumps = 0, c_paths = 0, c_deferred = 0;      /* per thread and sweep: far below 2^32 */
    const unsigned long long t_start = gtimer();
    const bool per_obs = p.outB != nullptr;
    const unsigned long long obs_begin = per_obs ? (unsigned long long)p.first : 0ull;
    const unsigned long long obs_end = per_obs ? (unsigned long long)(p.first + p.count) : (unsigned long long)p.l_local;
    const uint32_t cap = (uint32_t)p.mhrs_cap;
    const double *s_model = p.model + ML.s;
    Rec norec; norec.lastt = 0.0; norec.B = 0; norec.out_idx = 0;

    /* ---------------------------------------------------------------- lane phase */
    {
        Walk w; w.t = 0.0; w.spare = 0.0; w.spare_hi = 0u; w.a = 0u; w.b = 0u; w.j = 0; w.odd = false; w.fresh = true;
        bool active = false, cens = false, off = false, have_cur = false, cur_off = false;
        bool keep = false;          /* the tail lists were full when this lane tried to hand its observation over */
        double y = 0.0; uint32_t obs_local = 0, og = 0, tries = 0, cur_a = 0; int cur_pre = 0, kprop = 0;
        unsigned long long chunk_base = 0; unsigned chunk_next = 0, chunk_len = 0;      /* warp-uniform */
        bool exhausted = false, dry = false;      /* this warp found the stream empty / some warp did */
        unsigned it_count = 0;
        for (;;) {
            /* every 64 steps, look whether the observation stream has run dry elsewhere: a warp whose lanes are all
             * busy would otherwise never notice and grind on to `cap` while the rest of the grid waits */
            if (((++it_count) & 63u) == 0u && !dry) {
                unsigned long long nx = 0;
                if (lane == 0) nx = __ldcg(&p.state->next_obs);
                dry = obs_begin + __shfl_sync(FULL, nx, 0) >= obs_end;
            }
            unsigned idle = __ballot_sync(FULL, !active);
            if (idle && !exhausted) {
                if (chunk_next == chunk_len) {
                    unsigned long long base = 0;
                    if (lane == 0) base = obs_begin + atomicAdd(&p.state->next_obs, (unsigned long long)OBS_CHUNK);
                    base = __shfl_sync(FULL, base, 0);
                    chunk_base = base < obs_end ? base : obs_end;
                    const unsigned long long end = base + OBS_CHUNK < obs_end ? base + OBS_CHUNK : obs_end;
                    chunk_len = (unsigned)(end > chunk_base ? end - chunk_base : 0ull); chunk_next = 0;
                    if (chunk_len == 0u) exhausted = true;
                    /* prefetch the chunk's observations (coalesced) into this warp's staging buffer */
                    __syncwarp();
                    for (unsigned q = lane; q < chunk_len; q += 32u) {
                        sm.ybuf[warp * OBS_CHUNK + q] = p.y[chunk_base + q];
                        sm.cbuf[warp * OBS_CHUNK + q] = p.cens[chunk_base + q];
                    }
                    __syncwarp();
                }
                const unsigned avail = chunk_len - chunk_next;
                const unsigned rank = __popc(idle & ((1u << lane) - 1u));
                if (!active && rank < avail) {
                    const unsigned q = chunk_next + rank;
                    obs_local = (uint32_t)(chunk_base + q);
                    og = p.obs_rank + obs_local * p.obs_world;
                    y = sm.ybuf[warp * OBS_CHUNK + q]; cens = sm.cbuf[warp * OBS_CHUNK + q] != 0;
                    active = true; off = false; have_cur = false; kprop = 0; tries = 0; cur_a = 0; cur_off = false; cur_pre = 0;
                    keep = false;
                    walk_begin(w, 0u, false, og, p, iter);
                }
                const unsigned cnt = __popc(idle);
                chunk_next += cnt < avail ? cnt : avail;
                idle = __ballot_sync(FULL, !active);
            }
            if (idle == FULL) { if (exhausted) break; else continue; }
            __syncwarp();                   /* lanes that have just taken an observation step together with the rest */
            if (active) {
                const bool ended = walk_step<false>(w, y, cens, og, p, sm, iter, n, norec);
                c_jumps++;
                /* an ended attempt survives when it reached y in a state that can exit (eq_Bladt_MHRS.c:66,74) */
                const bool ok = (w.t >= y) && ((smask >> w.j) & 1u);
                const bool failed = ended && !ok;
                c_attempts += ended ? 1u : 0u;
                tries += failed ? 1u : 0u;
                /* a failed attempt restarts on the next sub-stream right here, unless the lane is due to hand the
                 * observation over: after `cap` attempts, or after END_CAP once the observation stream has run dry
                 * (a lone lane grinding through attempts would hold the whole grid at the barrier) */
                const bool handover = cap != 0u && !keep && (tries >= cap || ((exhausted || dry) && tries >= END_CAP));
                const bool restart = failed && !handover;
                w.a += failed ? 1u : 0u;
                off = failed ? false : off;
                w.fresh = restart; w.b = restart ? 0u : w.b; w.odd = restart ? false : w.odd;
                if (ended && !restart) {
                    bool accepted = false;
                    if (!ok) {
                        const uint32_t idx = atomicAdd(&p.state->n_items, 1u);
                        if (idx < p.item_cap) {
                            TailItem it; it.obs_local = obs_local; it.a = w.a; it.cur_a = cur_a;
                            it.flags = pack_flags(have_cur, cur_off, false, cur_pre, kprop);
                            p.items[idx] = it; p.found[idx] = FOUND_NONE;
                            c_deferred++; active = false;
                        } else {
                            /* no room: the lane keeps the observation and goes on with its next attempt */
                            keep = true; walk_begin(w, w.a, false, og, p, iter);
                        }
                    } else if (!have_cur) {
                        have_cur = true; cur_a = w.a; cur_off = off; cur_pre = w.j;
                        if (cens || p.mhit == 0) accepted = true;                              /* :70 */
                        else { off = false; walk_begin(w, w.a + 1u, false, og, p, iter); }
                    } else {
                        /* a valid proposal: accept test with draw 0 of the next sub-stream (:79-82) */
                        pht_u32x4 r = philox_block(0u, w.a + 1u, og, iter, p);
                        const double U = pht_u01(r.v[0], r.v[1]);
                        if (U < s_model[w.j] / s_model[cur_pre]) { cur_a = w.a; cur_off = off; cur_pre = w.j; }
                        kprop++;
                        if (kprop >= p.mhit) accepted = true;
                        else {
                            /* next proposal: sub-stream a+1 from draw 1 (the spare half of the block just computed) */
                            off = true; w.a++; w.fresh = true; w.b = 1; w.odd = true; w.t = 0.0;
                            w.spare = pht_u01(r.v[2], r.v[3]); w.spare_hi = r.v[3];
                        }
                    }
                    if (accepted) {
                        const unsigned e = warp * RING + atomicAdd(&sm.ring_n[warp], 1u);
                        sm.ring_y[e] = y; sm.ring_obs[e] = obs_local; sm.ring_a[e] = cur_a;
                        sm.ring_fl[e] = (unsigned char)((cens ? 1u : 0u) | (cur_off ? 2u : 0u));
                        active = false;
                    }
                }
            }
            __syncwarp();
            if (sm.ring_n[warp] >= RING_TRIGGER) replay_session(p, sm, iter, n, warp, c_jumps, c_paths);
        }
        replay_session(p, sm, iter, n, warp, c_jumps, c_paths);
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
        const unsigned long long nwarps = gsize / 32ull;
        for (unsigned long long i = gtid; i < n_items; i += gsize) p.pend0[i] = (uint32_t)i;
        if (gtid == 0) { p.state->n_pend[0] = n_items; p.state->n_pend[1] = 0u; p.state->unit_counter = 0ull; p.state->n_done = 0u; p.state->any_fail = 0u; }
        grid.sync();
        uint32_t K = TAIL_K0; int cur = 0; unsigned rounds = 0;
        for (;;) {
            const uint32_t P = p.state->n_pend[cur];
            if (P == 0u) break;
            const uint32_t *pend = cur ? p.pend1 : p.pend0;
            uint32_t *pend_next = cur ? p.pend0 : p.pend1;
            /* The round's work is K attempts for each of the P pending observations, numbered as units in chunk-major
             * order: unit u belongs to chunk u / TAIL_CH; chunk c covers attempts [ (c / P) TAIL_CH, +TAIL_CH ) beyond
             * the start attempt of observation pend[c % P].  So the first chunk of every observation is handed out
             * before any second chunk, and later chunks are mostly skipped once an earlier one has survived. */
            const unsigned long long total_units = (unsigned long long)P * K;
            /* --- search: each warp works through pools of consecutive attempts of one observation; a lane keeps the
             * observation of the pool it drew its attempt from, so the warp moves on to the next pool while slower lanes
             * are still finishing attempts of the previous one */
            {
                Walk w; w.t = 0.0; w.spare = 0.0; w.spare_hi = 0u; w.a = 0u; w.b = 0u; w.j = 0; w.odd = false; w.fresh = true;
                bool active = false;
                uint32_t my_item = 0, my_og = 0; double my_y = 0.0; bool my_cens = false;
                unsigned steps = 0;
                /* warp-uniform: the units this warp holds, and the pool being handed out */
                unsigned long long u_next = 0, u_end = 0, last_base = 0;
                bool out_of_units = false;
                uint32_t item = 0, og = 0, a_first = 0, pool_next = 0, pool_end = 0; double y = 0.0; bool cens = false, first_off = false;
                for (;;) {
                    unsigned idle = __ballot_sync(FULL, !active);
                    if (idle && pool_next >= pool_end && !out_of_units) {
                        /* take the next pool (skipping pools that an earlier survivor made moot) */
                        bool have = false;
                        while (!have) {
                            if (u_next >= u_end) {
                                const unsigned long long remaining = total_units > last_base ? total_units - last_base : 0ull;
                                unsigned long long want = remaining / (2ull * nwarps);
                                want = want < POOL_MIN ? POOL_MIN : (want > POOL_MAX ? POOL_MAX : want);
                                unsigned long long base = 0;
                                if (lane == 0) base = atomicAdd(&p.state->unit_counter, want);
                                base = __shfl_sync(FULL, base, 0);
                                last_base = base;
                                if (base >= total_units) { out_of_units = true; break; }
                                u_next = base; u_end = base + want < total_units ? base + want : total_units;
                            }
                            const unsigned long long chunk = u_next / TAIL_CH;
                            const unsigned long long chunk_end_u = (chunk + 1ull) * TAIL_CH;
                            const unsigned long long stop = u_end < chunk_end_u ? u_end : chunk_end_u;
                            item = pend[chunk % P];
                            const TailItem it = p.items[item];
                            a_first = it.a; first_off = (it.flags & 4u) != 0u;
                            pool_next = it.a + (uint32_t)(chunk / P) * TAIL_CH + (uint32_t)(u_next - chunk * TAIL_CH);
                            pool_end = pool_next + (uint32_t)(stop - u_next);
                            u_next = stop;
                            if ((__ldcg(&p.found[item]) >> 8) >= (unsigned long long)pool_next) {
                                have = true;
                                y = p.y[it.obs_local]; cens = p.cens[it.obs_local] != 0; og = p.obs_rank + it.obs_local * p.obs_world;
                            } else pool_end = pool_next;
                        }
                    }
                    if (idle == FULL && pool_next >= pool_end) break;          /* nothing in flight, nothing left to hand out */
                    /* Every FOUND_PERIOD steps, look whether some other warp has found a surviving attempt of the pool's
                     * observation below this pool: the rest of the pool is then moot (when few observations are pending,
                     * all warps of the GPU hold pools of the same ones, and without this look they would each finish
                     * theirs: measured 2.1x the necessary jump-steps at 1e6 observations). */
                    if (((++steps) & (FOUND_PERIOD - 1u)) == 0u && pool_next < pool_end) {
                        unsigned long long f = 0;
                        if (lane == 0) f = __ldcg(&p.found[item]);
                        f = __shfl_sync(FULL, f, 0);
                        const uint32_t fa = (uint32_t)(f >> 8);
                        if (f != FOUND_NONE && fa < pool_end) {
                            pool_end = pool_end < fa ? pool_end : fa;
                            pool_next = pool_next < pool_end ? pool_next : pool_end;
                            if (active && my_item == item && w.a > fa) active = false;
                            idle = __ballot_sync(FULL, !active);
                        }
                    }
                    if (idle && pool_next < pool_end) {
                        /* idle lanes draw the next attempt indices of the pool (ballot only, no memory traffic) */
                        const uint32_t avail = pool_end - pool_next;
                        const uint32_t rank = __popc(idle & ((1u << lane) - 1u));
                        if (!active && rank < avail) {
                            const uint32_t a = pool_next + rank;
                            my_item = item; my_y = y; my_cens = cens; my_og = og;
                            walk_begin(w, a, first_off && a == a_first, og, p, iter);
                            active = true;
                        }
                        const uint32_t cnt = __popc(idle);
                        pool_next += cnt < avail ? cnt : avail;
                    }
                    bool survived = false;
                    __syncwarp();           /* lanes that have just drawn an attempt step together with the rest */
                    if (active) {
                        const bool ended = walk_step<false>(w, my_y, my_cens, my_og, p, sm, iter, n, norec);
                        c_jumps++;
                        c_attempts += ended ? 1u : 0u;
                        survived = ended && (w.t >= my_y) && ((smask >> w.j) & 1u);
                        active = !ended;
                    }
                    const unsigned smk = __ballot_sync(FULL, survived);
                    if (smk) {
                        /* every survivor competes for "lowest surviving attempt" of its observation; the first one in
                         * the warp also calls off the later attempts of the same observation that are in flight here */
                        if (survived) atomicMin(&p.found[my_item], ((unsigned long long)w.a << 8) | (unsigned long long)w.j);
                        const int src = __ffs(smk) - 1;
                        const uint32_t s_item = __shfl_sync(FULL, my_item, src), s_a = __shfl_sync(FULL, w.a, src);
                        if (active && my_item == s_item && w.a > s_a) active = false;
                        if (s_item == item && pool_next < pool_end) {
                            pool_end = pool_end < s_a ? pool_end : s_a;
                            pool_next = pool_next < pool_end ? pool_next : pool_end;
                        }
                    }
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
                    if (!dropped) atomicOr(&p.state->any_fail, 1u);
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
                        pht_u32x4 r = philox_block(0u, a + 1u, og, iter, p);
                        const double U = pht_u01(r.v[0], r.v[1]);
                        if (U < s_model[pre] / s_model[cur_pre]) { it.cur_a = a; cur_off = off; cur_pre = pre; }
                        kprop++; done = (int)kprop >= p.mhit; next_off = true;
                    }
                    it.a = a + 1u;
                    it.flags = pack_flags(have_cur, cur_off, next_off, cur_pre, (int)kprop);
                    p.found[item] = FOUND_NONE;
                }
                p.items[item] = it;
                if (dropped) continue;
                if (done) p.done[atomicAdd(&p.state->n_done, 1u)] = item;
                else pend_next[atomicAdd(&p.state->n_pend[cur ^ 1], 1u)] = item;
            }
            if (gtid == 0) { p.state->n_pend[cur] = 0u; p.state->unit_counter = 0ull; }
            grid.sync();
            /* the attempts per observation grow only when some observation needed more than this round offered: with
             * mhit > 1 every pending observation comes back round after round for its next proposal, and a K that kept
             * doubling would bury the round in pools to skip */
            const bool grow = p.state->any_fail != 0u;
            grid.sync();
            if (gtid == 0) p.state->any_fail = 0u;
            cur ^= 1; rounds++;
            K = (grow && K < TAIL_KMAX) ? K * TAIL_GROWTH : K;
        }
        if (timekeeper) { const unsigned long long t = gtimer(); atomicAdd(&p.state->counters[PHT_CNT_NS_TAIL], t - t_mark); t_mark = t; }
        /* --- replay the accepted attempt of every tail observation */
        {
            const uint32_t n_done = p.state->n_done;
            for (unsigned long long i = gtid; i < n_done; i += gsize) {
                const TailItem it = p.items[p.done[i]];
                c_jumps += replay_path(it.obs_local, it.cur_a, (it.flags & 2u) != 0u, p.y[it.obs_local], p.cens[it.obs_local] != 0, p, sm, iter, n);
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
    unsigned long long w_attempts = c_attempts, w_jumps = c_jumps, w_paths = c_paths, w_deferred = c_deferred;
    for (int o = 16; o > 0; o >>= 1) {
        w_attempts += __shfl_down_sync(FULL, w_attempts, o); w_jumps += __shfl_down_sync(FULL, w_jumps, o);
        w_paths += __shfl_down_sync(FULL, w_paths, o); w_deferred += __shfl_down_sync(FULL, w_deferred, o);
    }
    if (lane == 0) {
        atomicAdd(&p.state->counters[PHT_CNT_ATTEMPTS], w_attempts); atomicAdd(&p.state->counters[PHT_CNT_JUMPS], w_jumps);
        atomicAdd(&p.state->counters[PHT_CNT_PATHS], w_paths); atomicAdd(&p.state->counters[PHT_CNT_DEFERRED], w_deferred);
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
