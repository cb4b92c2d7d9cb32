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
 * GPU organisation: one observation per lane, persistent warps refilled from a global counter.  The model
 * (S, Q, Q^-1, evals, s) lives in shared memory; each lane owns four shared-memory slabs of n doubles
 * (E_i = exp(evals_i T) reused by every CDF evaluation of the jump, J_i, the jump weights p_i, and the path's
 * sojourn totals z_i).  All sums run in the reference's index order inside one thread, so every decision and
 * every z is bit-identical to the host; lanes run the Brent loop in lock step until the slowest converges.
 *
 * Roofline: FP64 issue bound: ~n exp per CDF evaluation, ~10 evaluations per jump; 8 B of HBM per path.
 */
#include "path_common.cuh"

template <int THREADS>
struct DcsSmem {
    double *S, *Q, *Qinv, *evals, *s, *pi;
    double *E, *J, *P, *Z;
    long long *zacc; unsigned int *Nacc, *Bacc;
    __device__ __forceinline__ void carve(unsigned char *raw, int n) {
        double *d = reinterpret_cast<double *>(raw);
        S = d; d += n * n; Q = d; d += n * n; Qinv = d; d += n * n;
        evals = d; d += n; s = d; d += n; pi = d; d += n;
        E = d; d += n * THREADS; J = d; d += n * THREADS; P = d; d += n * THREADS; Z = d; d += n * THREADS;
        zacc = reinterpret_cast<long long *>(d); d += n;
        Nacc = reinterpret_cast<unsigned int *>(d); Bacc = Nacc + n * n;
    }
    static size_t bytes(int n) {
        return sizeof(double) * (size_t)(3 * n * n + 3 * n + 4 * n * THREADS + n) + sizeof(unsigned int) * (size_t)(n * n + n);
    }
};

/* sojourn-time CDF minus u at x (gt_Hobolth_DCS.c:23-40); E_i = exp(evals_i T) comes from the slab */
template <int THREADS>
__device__ __forceinline__ double hob_cdf(const DcsSmem<THREADS> &sm, int n, int k, int wcol, double x, double T,
                                          double Sll, double Slk, double prob, double Pab, double u) {
    const int tid = threadIdx.x;
    double tmp = 0.0;
    for (int i = 0; i < n; i++) {
        const double ev = sm.evals[i];
        const double Ei = sm.E[i * THREADS + tid];
        double Ji;
        if (fabs((ev - Sll) / Sll) < 1e-13) Ji = x * Ei;
        else Ji = (Ei - pht_exp((T - x) * ev + Sll * x)) / (ev - Sll);
        tmp += sm.Q[k + i * n] * Ji * sm.Qinv[i + wcol * n];
    }
    return 1 / prob * Slk / Pab * tmp - u;
}

template <int THREADS>
__global__ void __launch_bounds__(THREADS) k_dcs_sweep(SweepParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int n = p.n, tid = threadIdx.x;
    const unsigned FULL = 0xffffffffu;
    const ModelLayout ML = ModelLayout::make(n, p.m);
    DcsSmem<THREADS> sm; sm.carve(smem_raw, n);
    const uint32_t iter = p.state->iter;
    for (int i = tid; i < n * n; i += THREADS) {
        sm.S[i] = p.model[ML.S + i]; sm.Q[i] = p.model[ML.Q + i]; sm.Qinv[i] = p.model[ML.Qinv + i]; sm.Nacc[i] = 0u;
    }
    for (int i = tid; i < n; i += THREADS) {
        sm.evals[i] = p.model[ML.evals + i]; sm.s[i] = p.model[ML.s + i]; sm.pi[i] = p.model[ML.pi + i];
        sm.zacc[i] = 0; sm.Bacc[i] = 0u;
    }
    __syncthreads();

    unsigned long long c_jumps = 0, c_evals = 0, c_paths = 0, c_fail = 0;
    Dispenser disp; disp.init(p);
    PathRng rng; rng.seek(0);
    bool active = false;
    double y = 0.0, t = 0.0; int j = 0, b = 0, B = 0; long out_idx = 0;
    const double EPS = 2.220446049250313e-16;

    for (;;) {
        unsigned idle = __ballot_sync(FULL, !active);
        if (idle && !disp.exhausted) {
            const unsigned long long o = disp.take(p, idle, !active);
            if (o != ~0ull) {
                /* ---- new path: end state b (eq_AslettHobolth_DCS.c:11-51), then the start state */
                active = true; y = p.y[o]; out_idx = (long)o - p.first; t = 0.0;
                rng.seek(p.obs_rank + (uint32_t)o * p.obs_world);
                double *pv = sm.P + tid, *tv = sm.J + tid;                 /* p and tmp of the reference */
                for (int c = 0; c < n; c++) {                              /* p = pi^T Q, reference-BLAS order */
                    double acc = 0.0;
                    for (int i = 0; i < n; i++) acc += sm.Q[i + c * n] * sm.pi[i];
                    pv[c * THREADS] = (0.0 + 1.0 * acc) * pht_exp(sm.evals[c] * y);
                }
                double sum = 0.0;
                for (int c = 0; c < n; c++) {                              /* tmp = p^T Q^-1, then times s */
                    double acc = 0.0;
                    for (int i = 0; i < n; i++) acc += sm.Qinv[i + c * n] * pv[i * THREADS];
                    const double v = (0.0 + 1.0 * acc) * sm.s[c];
                    tv[c * THREADS] = v; sum += v;
                }
                for (int c = 0; c < n; c++) tv[c * THREADS] = tv[c * THREADS] / sum;
                b = slab_scan<THREADS>(sm.J, n, rng.next(p, iter));
                /* start state from pi (gt_Hobolth_DCS.c:88-95) */
                {
                    const double target = rng.next(p, iter);
                    double sofar = 0.0; int k = 0;
                    while (sofar < target && k <= n - 1) { sofar += sm.pi[k]; k++; }
                    B = k - 1 < 0 ? 0 : k - 1;
                }
                j = B;
                for (int i = 0; i < n; i++) sm.Z[i * THREADS + tid] = 0.0;
            }
            idle = __ballot_sync(FULL, !active);
        }
        if (idle == FULL) { if (disp.exhausted) break; else continue; }
        if (!active) continue;

        /* ---- one jump of the path (gt_Hobolth_DCS.c:112-214) */
        bool finished = !(t < y);                                          /* loop exit without the stay step: reference prints an error */
        int k = j; double jtime = 0.0;
        if (!finished) {
            const double T = y - t, Sjj = sm.S[j + j * n];
            double Pab = 0.0;
            for (int i = 0; i < n; i++) {
                const double Ei = pht_exp(sm.evals[i] * T);
                sm.E[i * THREADS + tid] = Ei;
                Pab += sm.Q[j + i * n] * Ei * sm.Qinv[i + b * n];                      /* :118-121 */
            }
            if (j == b) {                                                              /* :124-132 */
                if (rng.next(p, iter) < pht_exp(Sjj * T) / Pab) {
                    sm.Z[j * THREADS + tid] += T;
                    count_transition(p, n, sm.Nacc, out_idx, j, j);
                    finished = true;
                }
            }
            if (!finished) {
                const double eS = pht_exp(Sjj * T);
                for (int i = 0; i < n; i++) {                                          /* :137-144 */
                    const double ev = sm.evals[i], Ei = sm.E[i * THREADS + tid];
                    sm.J[i * THREADS + tid] = (fabs((ev - Sjj) / Sjj) < 1e-13) ? T * Ei : (Ei - eS) / (ev - Sjj);
                }
                double p_sum = 0.0;
                for (int i = 0; i < n; i++) {                                          /* :148-159 */
                    double v = 0.0;
                    if (i != j) {
                        double tmp = 0.0;
                        for (int q = 0; q < n; q++) tmp += sm.Q[i + q * n] * sm.J[q * THREADS + tid] * sm.Qinv[q + b * n];
                        v = sm.S[j + i * n] / Pab * tmp;
                        p_sum += v;
                    }
                    sm.P[i * THREADS + tid] = v;
                }
                const double target = (p_sum == 0.0) ? 0.0 : 0.0 + (p_sum - 0.0) * rng.next(p, iter);   /* runif(0, p_sum), :164 */
                k = slab_scan<THREADS>(sm.P, n, target);
                const double prob = sm.P[k * THREADS + tid], Slk = sm.S[j + k * n];
                const double u = rng.next(p, iter);                                    /* :184 */
                /* Brent's zeroin on [0, T] with f(0) = -u, f(T) = 1-u, Tol = 0, Maxit = 1000 (utility.c:233-338) */
                double ba = 0.0, bb = T, bc = 0.0, fa = -u, fb = 1.0 - u, fc = -u;
                bool conv = (fa == 0.0) || (fb == 0.0);
                if (fa == 0.0) bb = ba;
                int left = 1001;
                while (!conv && left > 0) {
                    left--;
                    const double prev_step = bb - ba;
                    if (fabs(fc) < fabs(fb)) { ba = bb; bb = bc; bc = ba; fa = fb; fb = fc; fc = fa; }
                    const double tol_act = 2 * EPS * fabs(bb) + 0.0 / 2;
                    double new_step = (bc - bb) / 2;
                    if (fabs(new_step) <= tol_act || fb == 0.0) { conv = true; break; }
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
                    fb = hob_cdf<THREADS>(sm, n, k, b, bb, T, Sjj, Slk, prob, Pab, u);
                    c_evals++;
                    if ((fb > 0 && fc > 0) || (fb < 0 && fc < 0)) { bc = ba; fc = fa; }
                }
                if (!conv) c_fail++;
                jtime = bb;
                int guard = 0;
                while (t + jtime >= y && guard++ < 2000) jtime = jtime / 2;             /* :204-206 */
                count_transition(p, n, sm.Nacc, out_idx, j, k);                         /* :209 */
                sm.Z[j * THREADS + tid] += jtime;                                       /* :210 */
                t += jtime; j = k;
                c_jumps++;
            }
        }
        if (finished) {
            path_flush<THREADS>(p, n, sm.Z, sm.zacc, sm.Bacc, B, out_idx);
            c_paths++; active = false;
        }
    }

    __syncthreads();
    block_flush<THREADS>(p, n, sm.Nacc, sm.Bacc, sm.zacc);
    for (int o = 16; o > 0; o >>= 1) {
        c_jumps += __shfl_down_sync(FULL, c_jumps, o); c_evals += __shfl_down_sync(FULL, c_evals, o);
        c_paths += __shfl_down_sync(FULL, c_paths, o); c_fail += __shfl_down_sync(FULL, c_fail, o);
    }
    if ((tid & 31) == 0) {
        atomicAdd(&p.state->counters[PHT_CNT_JUMPS], c_jumps); atomicAdd(&p.state->counters[PHT_CNT_BRENT_EVALS], c_evals);
        atomicAdd(&p.state->counters[PHT_CNT_PATHS], c_paths); atomicAdd(&p.state->counters[PHT_CNT_NONFINITE], c_fail);
    }
}

static int dcs_threads(int n) { return n <= 16 ? 128 : 64; }

int pht_dcs_grid_blocks(int device, int n) {
    int per_sm = 0, sms = 0; cudaError_t e;
    if (dcs_threads(n) == 128) {
        const size_t smem = DcsSmem<128>::bytes(n);
        e = cudaFuncSetAttribute(k_dcs_sweep<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e == cudaSuccess) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_dcs_sweep<128>, 128, smem);
    } else {
        const size_t smem = DcsSmem<64>::bytes(n);
        e = cudaFuncSetAttribute(k_dcs_sweep<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e == cudaSuccess) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_dcs_sweep<64>, 64, smem);
    }
    if (e != cudaSuccess) return -1;
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device) != cudaSuccess) return -1;
    return per_sm * sms;
}

cudaError_t pht_launch_dcs(const SweepParams &p, int grid_blocks, cudaStream_t st) {
    if (dcs_threads(p.n) == 128) k_dcs_sweep<128><<<grid_blocks, 128, DcsSmem<128>::bytes(p.n), st>>>(p);
    else k_dcs_sweep<64><<<grid_blocks, 64, DcsSmem<64>::bytes(p.n), st>>>(p);
    return cudaGetLastError();
}
