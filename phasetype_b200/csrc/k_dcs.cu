/*
 * k_dcs.cu -- Aslett-Hobolth direct conditional sampling (method bit 4) for sm_100a.
 *
 * What it computes, per observation y (reference src/Simulate_AbsCTMC_eq_AslettHobolth_DCS.c:119-146 and
 * src/Simulate_AbsCTMC_gt_Hobolth_DCS.c:74-226): draw the state b the chain is absorbed from with weight
 * (pi e^{Sy})_i s_i, then sample the path from pi to b at time y directly (Hobolth & Stone): at each jump
 * either stay in b to the end, or draw the next state with weight S_ji * int_0^T e^{S_jj x} [e^{S(T-x)}]_{ib} dx
 * and the sojourn by inverting its CDF with Brent's method (src/utility.c:233-338).  exp{xS} is evaluated
 * spectrally, Q diag(exp(x evals)) Q^-1, exactly like the reference; the censoring flag is ignored (:132-133).
 *
 * GPU organisation.  One observation per lane, persistent warps refilled from a global counter.  The work of a
 * path is a chain of UNITS that all start with the same expensive thing -- n exponentials exp(alpha ev_i + beta)
 * of the spectrum with lane-specific (alpha, beta):
 *      NEW    alpha = y             end state b and start state of a new path
 *      JUMP   alpha = T = y - t     E_i = exp(ev_i T): stay test, jump weights, next state, root-finder set-up
 *      BRENT  alpha = T - x, beta = S_jj x     one evaluation of the sojourn CDF inside Brent's iteration
 * Every lane is a little state machine that executes one unit per loop iteration: the exponential batch runs
 * converged over the whole warp whatever the lanes are doing, and only the short kind-specific tails diverge.
 * (Running whole Brent loops in lock step instead left 7 of 32 lanes active: profiles/r1a_dcs_ncu_full.md.)
 * Brent's bookkeeping between two evaluations is one shared stage for lanes coming from JUMP and from BRENT.
 * The O(n^2) part of a JUMP unit (P_ab, the J_i with their divisions, the n jump-weight dot products) is served by
 * the whole warp: the per-lane vectors live in shared-memory slabs, so a group of G >= n lanes takes one requesting
 * lane each, lane i of the group computes term i, and the sums the reference accumulates in index order are
 * accumulated in index order with shuffles.
 * Quantities that do not depend on the observation are tabulated once per block: ev_i - S_jj, the degenerate
 * test |(ev_i - S_jj)/S_jj| < 1e-13 (gt_Hobolth_DCS.c:29,139) and pi^T Q.  All sums run in the reference's index
 * order inside one thread, so every decision and every z is bit-identical to the host.
 *
 * Roofline: FP64 issue bound: n exp per unit, ~11.5 units per jump; 8 B of HBM per path.
 */
#include "path_common.cuh"

#ifndef DCS_WARPS_PER_SM
#define DCS_WARPS_PER_SM 20          /* launch bound: resident warps per SM the register allocation must allow */
#endif

enum { K_IDLE = 0, K_NEW = 1, K_JUMP = 2, K_BRENT = 3 };

template <int THREADS>
struct DcsSmem {
    double *S, *Q, *Qinv, *evals, *evi, *s, *pi, *PIQ, *D;
    double *X, *E, *P, *Z;
    unsigned long long *zlo; long long *zhi; unsigned int *Nacc, *Bacc, *deg;
    __device__ __forceinline__ void carve(unsigned char *raw, int n) {
        double *d = reinterpret_cast<double *>(raw);
        /* Qinv and D are read with a lane-varying COLUMN (b) resp. ROW (j) and a uniform other index: their leading
         * dimension is padded to an odd number of doubles so that those reads spread over the banks (with ld = n = 8
         * they were 4-way conflicts: 3.8e9 per sweep, profiles/r1_final_dcs_1e7_ncu_full.md) */
        const int ld = n | 1;
        /* (Qinv has one more column, n: Q^-1 b of the MH variant's exit set, so that "column b" reads work for both) */
        S = d; d += n * n; Q = d; d += n * n; Qinv = d; d += (n + 1) * ld; D = d; d += n * ld;
        evals = d; d += n; evi = d; d += n; s = d; d += n; pi = d; d += n; PIQ = d; d += n;
        X = d; d += n * THREADS; E = d; d += n * THREADS; P = d; d += n * THREADS; Z = d; d += n * THREADS;
        zlo = reinterpret_cast<unsigned long long *>(d); d += n; zhi = reinterpret_cast<long long *>(d); d += n;
        Nacc = reinterpret_cast<unsigned int *>(d); Bacc = Nacc + n * n; deg = Bacc + n;
    }
    static size_t bytes(int n) {
        return sizeof(double) * (size_t)(2 * n * n + (2 * n + 1) * (n | 1) + 5 * n + 4 * n * THREADS + 2 * n) + sizeof(unsigned int) * (size_t)(n * n + 2 * n);
    }
};

/* MH = false: the live DCS sampler.  MH = true: LJMA_MHsample_Hobolth (gt_Hobolth_DCS.c:268-355, method bit 8; nothing
 * in the reference calls it): no end-state draw, the chain is conditioned on being in the exit set {j : s_j > 0} at y
 * (w = Q^-1 b, kept as column n of the Qinv table), and the chains go through the MH wrapper of path_common.cuh. */
template <int THREADS, bool MH, bool CPLX>
__global__ void __launch_bounds__(THREADS, DCS_WARPS_PER_SM * 32 / THREADS) k_dcs_sweep(SweepParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int n = p.n, tid = threadIdx.x, ld = n | 1;
    const unsigned FULL = 0xffffffffu;
    const ModelLayout ML = ModelLayout::make(n, p.m);
    DcsSmem<THREADS> sm; sm.carve(smem_raw, n);
    const uint32_t iter = p.state->iter;
    /* complex pairs in the spectrum: the reference's formulas (CPLX = false) are not valid; the CPLX = true instance of
     * this kernel takes the sweep (both are launched, the spectrum is known on the device only) */
    { bool cplx = false; for (int i = 0; i < n; i++) cplx = cplx || (p.model[ML.evals_im + i] != 0.0);
      if (cplx != CPLX) return; }
    /* CPLX: which eigen-indices open (pfirst) and close (psecond) a pair a -+ ib; warp-uniform bit masks */
    uint32_t pfirst = 0u, psecond = 0u;
    if (CPLX) {
        for (int i = 0; i + 1 < n; i++)
            if (!((psecond >> i) & 1u) && p.model[ML.evals_im + i] > 0.0) { pfirst |= 1u << i; psecond |= 1u << (i + 1); }
    }
    const uint32_t ppair = pfirst | psecond;
    for (int i = tid; i < n * n; i += THREADS) {
        sm.S[i] = p.model[ML.S + i]; sm.Q[i] = p.model[ML.Q + i]; sm.Qinv[(i % n) + (i / n) * ld] = p.model[ML.Qinv + i]; sm.Nacc[i] = 0u;
    }
    for (int i = tid; i < n; i += THREADS) {
        sm.evals[i] = p.model[ML.evals + i]; sm.evi[i] = p.model[ML.evals_im + i]; sm.s[i] = p.model[ML.s + i]; sm.pi[i] = p.model[ML.pi + i];
        sm.zlo[i] = 0ull; sm.zhi[i] = 0; sm.Bacc[i] = 0u;
    }
    __syncthreads();
    uint32_t bmask = 0u;                /* MH: the exit set */
    if (MH) {
        for (int i = 0; i < n; i++) bmask |= (sm.s[i] > 0.0) ? (1u << i) : 0u;
        /* Q^-1 b in dgemv-'N' order: y_i accumulated over the columns j (eq_AslettHobolth_DCS.c:131) */
        for (int i = tid; i < n; i += THREADS) {
            double acc = 0.0;
            for (int jj = 0; jj < n; jj++) acc += (1.0 * (((bmask >> jj) & 1u) ? 1.0 : 0.0)) * sm.Qinv[i + jj * ld];
            sm.Qinv[i + n * ld] = acc;
        }
    }
    /* observation-independent tables: D[j][i] = ev_i - S_jj, the degenerate flags, and pi^T Q in reference-BLAS order */
    for (int e = tid; e < n * n; e += THREADS) {
        const int j = e / n, i = e % n;
        if (!CPLX) sm.D[j * ld + i] = sm.evals[i] - sm.S[j + j * n];
        else {
            /* RECIPROCALS (nothing to be bit-identical to here): 1 / (ev_i - S_jj) for a real eigenvalue; for a pair in
             * (i, i + 1) the complex 1 / (lam - S_jj) at lam = a - ib, real part in slot i, imaginary part in slot i + 1 */
            const int base = ((psecond >> i) & 1u) ? i - 1 : i;
            const double dr = sm.evals[base] - sm.S[j + j * n];
            if ((ppair >> i) & 1u) { const double bi = sm.evi[base], den = dr * dr + bi * bi; sm.D[j * ld + i] = (base == i ? dr : bi) / den; }
            else sm.D[j * ld + i] = 1.0 / dr;
        }
    }
    for (int c = tid; c < n; c += THREADS) {
        double acc = 0.0;
        for (int i = 0; i < n; i++) acc += sm.Q[i + c * n] * sm.pi[i];
        sm.PIQ[c] = 0.0 + 1.0 * acc;
        const double Sjj = sm.S[c + c * n];
        unsigned m = 0u;
        for (int i = 0; i < n; i++) m |= (fabs((sm.evals[i] - Sjj) / Sjj) < 1e-13) ? (1u << i) : 0u;
        sm.deg[c] = m & ~ppair;
    }
    __syncthreads();

    unsigned c_jumps = 0, c_evals = 0, c_paths = 0, c_fail = 0;
    Dispenser disp; disp.init(p);
    PathRng rng; rng.seek(0);
    MhChain mh; mh.begin(false);
    int kind = K_IDLE;
    double y = 0.0, t = 0.0, T = 0.0, alpha = 0.0, beta = 0.0, u = 0.0, coef = 0.0;
    double ba = 0.0, bb = 0.0, bc = 0.0, fa = 0.0, fb = 0.0, fc = 0.0;
    int j = 0, b = 0, k = 0, B = 0, left = 0; long out_idx = 0; bool conv = false;
    const double EPS = 2.220446049250313e-16;

    /* sub-warp groups of G >= n lanes serve the JUMP units cooperatively: lane i of a group handles eigen-index i */
    int G = 4; while (G < n) G <<= 1;
    const int lane = tid & 31, gi = lane & (G - 1), gbase = lane & ~(G - 1), ngroups = 32 / G, mygroup = lane / G;
    const int wtid = tid & ~31;               /* slab column of lane 0 of this warp */

    for (;;) {
        unsigned idle = __ballot_sync(FULL, kind == K_IDLE);
        if (idle && !disp.exhausted) {
            const unsigned long long o = disp.take(p, idle, kind == K_IDLE);
            if (o != ~0ull) {
                kind = K_NEW; y = p.y[o]; out_idx = (long)o - p.first; t = 0.0;
                rng.seek(p.obs_rank + (uint32_t)o * p.obs_world);
                if (MH) mh.begin(p.cens[o] != 0);
                alpha = y; beta = 0.0;
            }
            idle = __ballot_sync(FULL, kind == K_IDLE);
        }
        if (idle == FULL) { if (disp.exhausted) break; else continue; }
        const int kind0 = kind;
        __syncwarp();                       /* refilled lanes and the rest take the unit together */

        /* ---- the unit's exponential batch: exp(alpha ev_i + beta), into E for a JUMP unit, else into X */
        if (kind0 != K_IDLE) {
            double *dst = (kind0 == K_JUMP ? sm.E : sm.X) + tid;
            if (!CPLX) {
                /* four at a time: the same bits as pht_exp, the dependent chains interleaved (pht_math.h) */
#pragma unroll 1
                for (int i = 0; i < n; i += 4) {
                    double e0, e1, e2, e3;
                    pht_exp4(alpha * sm.evals[i] + beta, (i + 1 < n) ? alpha * sm.evals[i + 1] + beta : 0.0,
                             (i + 2 < n) ? alpha * sm.evals[i + 2] + beta : 0.0, (i + 3 < n) ? alpha * sm.evals[i + 3] + beta : 0.0, e0, e1, e2, e3);
                    dst[i * THREADS] = e0;
                    if (i + 1 < n) dst[(i + 1) * THREADS] = e1;
                    if (i + 2 < n) dst[(i + 2) * THREADS] = e2;
                    if (i + 3 < n) dst[(i + 3) * THREADS] = e3;
                }
            } else {
                /* exp(alpha lam + beta) at lam = a - ib: real part in slot i, imaginary part in slot i + 1 */
                /* (the exponentials four at a time, interleaved -- a pair's second slot gets one it does not need -- then the
                 * rotations of the pairs) */
#pragma unroll 1
                for (int i = 0; i < n; i += 4) {
                    double e0, e1, e2, e3;
                    pht_exp4(alpha * sm.evals[i] + beta, (i + 1 < n) ? alpha * sm.evals[i + 1] + beta : 0.0,
                             (i + 2 < n) ? alpha * sm.evals[i + 2] + beta : 0.0, (i + 3 < n) ? alpha * sm.evals[i + 3] + beta : 0.0, e0, e1, e2, e3);
                    dst[i * THREADS] = e0;
                    if (i + 1 < n) dst[(i + 1) * THREADS] = e1;
                    if (i + 2 < n) dst[(i + 2) * THREADS] = e2;
                    if (i + 3 < n) dst[(i + 3) * THREADS] = e3;
                }
                {
                    uint32_t pf = pfirst;
#pragma unroll 1
                    while (pf) {
                        const int i = __ffs(pf) - 1; pf &= pf - 1u;
                        const double ex = dst[i * THREADS];
                        double sn, cs; sincos(alpha * sm.evi[i], &sn, &cs);
                        dst[i * THREADS] = ex * cs; dst[(i + 1) * THREADS] = -(ex * sn);
                    }
                }
            }
        }
        __syncwarp();

        /* ---- JUMP and NEW units, served by the whole warp (gt_Hobolth_DCS.c:112-159, eq_AslettHobolth_DCS.c:16-40):
         * up to 32/G requesting lanes at a time; group lane i computes term i of P_ab, J_i and the jump weight of
         * candidate state i (JUMP), or component i of the end-state weights (NEW); sums that the reference accumulates
         * in index order are accumulated in index order over the group's lanes */
        double sv_Pab = 0.0, sv_eS = 0.0, sv_psum = 0.0;
        unsigned need = __ballot_sync(FULL, kind0 == K_JUMP || (!MH && kind0 == K_NEW));
        while (need) {
            int r = -1; unsigned served = 0u;
            {
                unsigned m = need;
#pragma unroll 1
                for (int g = 0; g < ngroups && m; g++) { const int bit = __ffs(m) - 1; m &= m - 1u; served |= 1u << bit; if (g == mygroup) r = bit; }
            }
            const int rr = r < 0 ? 0 : r;
            const int j_r = __shfl_sync(FULL, j, rr), b_r = __shfl_sync(FULL, b, rr), kind_r = __shfl_sync(FULL, kind0, rr);
            const double T_r = __shfl_sync(FULL, T, rr);
            const int col = wtid + rr;
            const bool work = (r >= 0) && (gi < n);
            const bool wj = work && kind_r == K_JUMP, wn = work && kind_r == K_NEW;
            double val1 = 0.0, eS = 0.0, Ei = 0.0;
            if (wj) {
                const double Sjj = sm.S[j_r + j_r * n];
                eS = pht_exp(Sjj * T_r);
                Ei = sm.E[gi * THREADS + col];
                if (!CPLX || !((ppair >> gi) & 1u)) val1 = sm.Q[j_r + gi * n] * Ei * sm.Qinv[gi + b_r * ld];  /* :118-121 */
                else if ((pfirst >> gi) & 1u) {
                    /* the pair's term of u^T exp(T B) v: Re(E) (u1 v1 + u2 v2) - Im(E) (u1 v2 - u2 v1) */
                    const double u1 = sm.Q[j_r + gi * n], u2 = sm.Q[j_r + (gi + 1) * n];
                    const double v1 = sm.Qinv[gi + b_r * ld], v2 = sm.Qinv[gi + 1 + b_r * ld];
                    val1 = Ei * (u1 * v1 + u2 * v2) - sm.E[(gi + 1) * THREADS + col] * (u1 * v2 - u2 * v1);
                }
            }
            if (!CPLX) {
                if (wn) sm.X[gi * THREADS + col] = sm.PIQ[gi] * sm.X[gi * THREADS + col];                    /* p = (pi^T Q) o exp(ev y) */
            } else {
                /* p = (pi^T Q) exp(y B): a pair's two components mix, so read both before anyone writes */
                double nx = 0.0;
                if (wn) {
                    if ((ppair >> gi) & 1u) {
                        const int base = ((psecond >> gi) & 1u) ? gi - 1 : gi;
                        const double r1 = sm.PIQ[base], r2 = sm.PIQ[base + 1];
                        const double xr = sm.X[base * THREADS + col], xi = sm.X[(base + 1) * THREADS + col];
                        nx = (base == gi) ? r1 * xr + r2 * xi : r2 * xr - r1 * xi;
                    } else nx = sm.PIQ[gi] * sm.X[gi * THREADS + col];
                }
                __syncwarp();
                if (wn) sm.X[gi * THREADS + col] = nx;
            }
            __syncwarp();
            if (wn) {                                                                                        /* (p^T Q^-1)_i s_i */
                double acc = 0.0;
#pragma unroll 1
                for (int q = 0; q < n; q++) acc += sm.Qinv[q + gi * ld] * sm.X[q * THREADS + col];
                val1 = (0.0 + 1.0 * acc) * sm.s[gi];
            }
            double sum1 = 0.0;                                                                               /* P_ab, or the weights' total */
#pragma unroll 4
            for (int q = 0; q < n; q++) sum1 += __shfl_sync(FULL, val1, gbase + q);
            if (wj) {
                const unsigned dg = sm.deg[j_r];
                if (!CPLX) sm.X[gi * THREADS + col] = ((dg >> gi) & 1u) ? T_r * Ei : (Ei - eS) / sm.D[j_r * ld + gi];    /* :137-144 */
                else if (!((ppair >> gi) & 1u))
                    sm.X[gi * THREADS + col] = (((dg >> gi) & 1u) ? T_r * Ei : (Ei - eS) * sm.D[j_r * ld + gi]) * sm.Qinv[gi + b_r * ld];
                else {
                    /* component gi of J(T) v: J = (e^{lam T} - e^{S_jj T}) / (lam - S_jj) times v1 + i v2, as complex numbers */
                    const int base = ((psecond >> gi) & 1u) ? gi - 1 : gi;
                    const double Nr = sm.E[base * THREADS + col] - eS, Ni = sm.E[(base + 1) * THREADS + col];
                    const double DR = sm.D[j_r * ld + base], DI = sm.D[j_r * ld + base + 1];
                    const double Jr = Nr * DR - Ni * DI, Ji = Nr * DI + Ni * DR;
                    const double v1 = sm.Qinv[base + b_r * ld], v2 = sm.Qinv[base + 1 + b_r * ld];
                    sm.X[gi * THREADS + col] = (base == gi) ? Jr * v1 - Ji * v2 : Jr * v2 + Ji * v1;
                }
            }
            if (wn) sm.P[gi * THREADS + col] = val1 / sum1;
            __syncwarp();
            double v = 0.0;
            if (wj) {
                if (gi != j_r) {                                                                              /* :148-159 */
                    double tmp = 0.0;
#pragma unroll 1
                    for (int q = 0; q < n; q++) tmp += CPLX ? sm.Q[gi + q * n] * sm.X[q * THREADS + col]
                                                             : sm.Q[gi + q * n] * sm.X[q * THREADS + col] * sm.Qinv[q + b_r * ld];
                    v = sm.S[j_r + gi * n] / sum1 * tmp;
                }
                sm.P[gi * THREADS + col] = v;
            }
            double p_sum = 0.0;
#pragma unroll 4
            for (int q = 0; q < n; q++) p_sum += __shfl_sync(FULL, v, gbase + q);      /* the term of state j is +0.0 */
            __syncwarp();
            /* hand the three scalars back to the requesting lanes: find the group that served this lane, read its lane 0 */
            int server = -1;
#pragma unroll 1
            for (int g = 0; g < ngroups; g++) if (__shfl_sync(FULL, r, g * G) == lane) server = g * G;
            {
                const int src = server < 0 ? 0 : server;
                const double a0 = __shfl_sync(FULL, sum1, src), a1 = __shfl_sync(FULL, eS, src), a2 = __shfl_sync(FULL, p_sum, src);
                if (server >= 0) { sv_Pab = a0; sv_eS = a1; sv_psum = a2; }
            }
            need &= ~served;
        }

        bool advance = false, enter = false, flush = false;
        if (kind0 == K_BRENT) {
            /* ---- one evaluation of the sojourn CDF at x = bb (gt_Hobolth_DCS.c:23-40), then Brent's bracket update */
            const unsigned dg = sm.deg[j];
            double tmp = 0.0;
            if (!CPLX) {
#pragma unroll 1
                for (int i = 0; i < n; i++) {
                    const double Ei = sm.E[i * THREADS + tid];
                    const double Ji = ((dg >> i) & 1u) ? bb * Ei : (Ei - sm.X[i * THREADS + tid]) / sm.D[j * ld + i];
                    tmp += sm.Q[k + i * n] * Ji * sm.Qinv[i + b * ld];
                }
            } else {
                /* Q[k, :] J(x) v with J's pair block (e^{lam T} - e^{lam (T - x) + S_jj x}) / (lam - S_jj) in complex arithmetic */
#pragma unroll 1
                for (int i = 0; i < n; i++) {
                    const double Nr = sm.E[i * THREADS + tid] - sm.X[i * THREADS + tid];
                    if ((pfirst >> i) & 1u) {
                        const double Ni = sm.E[(i + 1) * THREADS + tid] - sm.X[(i + 1) * THREADS + tid];
                        const double DR = sm.D[j * ld + i], DI = sm.D[j * ld + i + 1];
                        const double Jr = Nr * DR - Ni * DI, Ji = Nr * DI + Ni * DR;
                        const double v1 = sm.Qinv[i + b * ld], v2 = sm.Qinv[i + 1 + b * ld];
                        tmp += sm.Q[k + i * n] * (Jr * v1 - Ji * v2) + sm.Q[k + (i + 1) * n] * (Jr * v2 + Ji * v1);
                        i++;
                    } else {
                        const double Ji = ((dg >> i) & 1u) ? bb * sm.E[i * THREADS + tid] : Nr * sm.D[j * ld + i];
                        tmp += sm.Q[k + i * n] * Ji * sm.Qinv[i + b * ld];
                    }
                }
            }
            fb = coef * tmp - u;
            c_evals++;
            if ((fb > 0 && fc > 0) || (fb < 0 && fc < 0)) { bc = ba; fc = fa; }
            advance = true;
        } else if (kind0 == K_JUMP) {
            /* ---- the requesting lane's own part of the jump: stay test, next state, root-finder set-up (:124-190) */
            const double Pab = sv_Pab, eS = sv_eS, p_sum = sv_psum;
            bool stay = false;
            if (MH ? (((bmask >> j) & 1u) != 0u) : (j == b)) stay = rng.next(p, iter) < eS / Pab;   /* :124-132 */
            if (stay) {
                sm.Z[j * THREADS + tid] += T;
                if (!MH || mh.rec) count_transition(p, n, sm.Nacc, out_idx, j, j);
                flush = true;
            } else {
                const double target = (p_sum == 0.0) ? 0.0 : 0.0 + (p_sum - 0.0) * rng.next(p, iter);   /* runif(0, p_sum), :164 */
                k = slab_scan<THREADS>(sm.P, n, target);
                const double prob = sm.P[k * THREADS + tid], Slk = sm.S[j + k * n];
                u = rng.next(p, iter);                                                         /* :184 */
                coef = 1 / prob * Slk / Pab;                                                   /* the constant factor of :39 */
                /* Brent's zeroin on [0, T] with f(0) = -u, f(T) = 1-u, Tol = 0, Maxit = 1000 (utility.c:233-338) */
                ba = 0.0; bb = T; bc = 0.0; fa = -u; fb = 1.0 - u; fc = -u;
                conv = (fa == 0.0) || (fb == 0.0);
                if (fa == 0.0) bb = ba;
                left = 1001;
                advance = true;
            }
        } else if (kind0 == K_NEW) {
            /* ---- new path: end state b from the weights the warp just formed (eq_AslettHobolth_DCS.c:41-50), then the
             * start state (gt_Hobolth_DCS.c:88-95) */
            if (MH) b = n; else b = slab_scan<THREADS>(sm.P, n, rng.next(p, iter));
            {
                const double target = rng.next(p, iter);
                double sofar = 0.0; int q = 0;
#pragma unroll 1
                while (sofar < target && q <= n - 1) { sofar += sm.pi[q]; q++; }
                B = q - 1 < 0 ? 0 : q - 1;
            }
            j = B;
#pragma unroll 1
            for (int i = 0; i < n; i++) sm.Z[i * THREADS + tid] = 0.0;
            enter = true;
        }

        if (advance) {
            /* ---- Brent's bookkeeping up to the next evaluation point (the loop head of utility.c:263-331) */
            bool done = conv;
            if (!done && left == 0) { c_fail++; done = true; }
            if (!done) {
                left--;
                const double prev_step = bb - ba;
                if (fabs(fc) < fabs(fb)) { ba = bb; bb = bc; bc = ba; fa = fb; fb = fc; fc = fa; }
                const double tol_act = 2 * EPS * fabs(bb) + 0.0 / 2;
                double new_step = (bc - bb) / 2;
                if (fabs(new_step) <= tol_act || fb == 0.0) done = true;
                else {
                    if (fabs(prev_step) >= tol_act && fabs(fa) > fabs(fb)) {
                        double pp, qq; const double cb = bc - bb;
                        if (ba == bc) { const double t1 = fb / fa; pp = cb * t1; qq = 1.0 - t1; }
                        else {
                            qq = fa / fc; const double t1 = fb / fc, t2 = fb / fa;
                            pp = t2 * (cb * qq * (qq - t1) - (bb - ba) * (t1 - 1.0));
                            qq = (qq - 1.0) * (t1 - 1.0) * (t2 - 1.0);
                        }
                        if (pp > 0.0) qq = -qq; else pp = -pp;
                        if (pp < (0.75 * cb * qq - fabs(tol_act * qq) / 2) && pp < fabs(prev_step * qq / 2)) new_step = pp / qq;
                    }
                    if (fabs(new_step) < tol_act) new_step = (new_step > 0.0) ? tol_act : -tol_act;
                    ba = bb; fa = fb;
                    bb += new_step;
                    kind = K_BRENT; alpha = T - bb; beta = sm.S[j + j * n] * bb;                /* argument of :33 */
                }
            }
            if (done) {
                double jtime = bb;
                int guard = 0;
#pragma unroll 1
                while (t + jtime >= y && guard++ < 2000) jtime = jtime / 2;                     /* :204-206 */
                if (!MH || mh.rec) count_transition(p, n, sm.Nacc, out_idx, j, k);              /* :209 */
                sm.Z[j * THREADS + tid] += jtime;                                               /* :210 */
                t += jtime; j = k;
                c_jumps++;
                enter = true;
            }
        }
        if (enter) {
            /* ---- next unit is a JUMP from (j, t), unless the clock has run out (the reference prints an error there) */
            if (!(t < y)) flush = true;
            else { T = y - t; alpha = T; beta = 0.0; kind = K_JUMP; }
        }
        if (flush) {
            /* MH: the chain that just ended stayed in j to the end (res_pre, :128); the wrapper says what runs next */
            if (!MH || mh.chain_end<false>(j, sm.s, p.mhit, p, iter, rng.obs)) {
                path_flush<THREADS>(p, n, sm.Z, sm.zlo, sm.zhi, sm.Bacc, B, out_idx);
                c_paths++; kind = K_IDLE;
            } else {
                rng.seek_sub(mh.chain, mh.off, p, iter);
                kind = K_NEW; t = 0.0; alpha = y; beta = 0.0;
            }
        }
    }

    __syncthreads();
    block_flush<THREADS>(p, n, sm.Nacc, sm.Bacc, sm.zlo, sm.zhi);
    unsigned long long w_jumps = c_jumps, w_evals = c_evals, w_paths = c_paths, w_fail = c_fail;
    for (int o = 16; o > 0; o >>= 1) {
        w_jumps += __shfl_down_sync(FULL, w_jumps, o); w_evals += __shfl_down_sync(FULL, w_evals, o);
        w_paths += __shfl_down_sync(FULL, w_paths, o); w_fail += __shfl_down_sync(FULL, w_fail, o);
    }
    if ((tid & 31) == 0) {
        atomicAdd(&p.state->counters[PHT_CNT_JUMPS], w_jumps); atomicAdd(&p.state->counters[PHT_CNT_BRENT_EVALS], w_evals);
        atomicAdd(&p.state->counters[PHT_CNT_PATHS], w_paths); atomicAdd(&p.state->counters[PHT_CNT_NONFINITE], w_fail);
    }
}

static int dcs_threads(int n) { return n <= 16 ? 128 : 64; }

/* the two instances (real spectrum: the reference's arithmetic; complex pairs: block formulas) share one grid size:
 * the smaller of their occupancies (the CPLX body needs a few more registers) */
template <int THREADS, bool MH>
static cudaError_t dcs_occupancy(int n, int *per_sm) {
    const size_t smem = DcsSmem<THREADS>::bytes(n);
    int a = 0, b = 0;
    cudaError_t e = cudaFuncSetAttribute(k_dcs_sweep<THREADS, MH, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(k_dcs_sweep<THREADS, MH, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e == cudaSuccess) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&a, k_dcs_sweep<THREADS, MH, false>, THREADS, smem);
    if (e == cudaSuccess) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b, k_dcs_sweep<THREADS, MH, true>, THREADS, smem);
    *per_sm = a < b ? a : b;
    return e;
}
int pht_dcs_grid_blocks(int device, int n, bool mh) {
    int per_sm = 0, sms = 0; cudaError_t e;
    if (dcs_threads(n) == 128) e = mh ? dcs_occupancy<128, true>(n, &per_sm) : dcs_occupancy<128, false>(n, &per_sm);
    else e = mh ? dcs_occupancy<64, true>(n, &per_sm) : dcs_occupancy<64, false>(n, &per_sm);
    if (e != cudaSuccess) return -1;
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device) != cudaSuccess) return -1;
    return per_sm * sms;
}

template <int THREADS, bool MH>
static cudaError_t dcs_launch_pair(const SweepParams &p, int grid_blocks, cudaStream_t st) {
    /* the sweep's spectrum is known on the device only: exactly one of the two instances does the work, the other returns */
    k_dcs_sweep<THREADS, MH, false><<<grid_blocks, THREADS, DcsSmem<THREADS>::bytes(p.n), st>>>(p);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    k_dcs_sweep<THREADS, MH, true><<<grid_blocks, THREADS, DcsSmem<THREADS>::bytes(p.n), st>>>(p);
    return cudaGetLastError();
}
cudaError_t pht_launch_dcs(const SweepParams &p, int grid_blocks, cudaStream_t st, bool mh) {
    if (dcs_threads(p.n) == 128) return mh ? dcs_launch_pair<128, true>(p, grid_blocks, st) : dcs_launch_pair<128, false>(p, grid_blocks, st);
    return mh ? dcs_launch_pair<64, true>(p, grid_blocks, st) : dcs_launch_pair<64, false>(p, grid_blocks, st);
}
