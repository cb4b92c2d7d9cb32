/*
 * k_peak.cu -- FP64 issue-rate microbenchmark (the roofline denominator of BASELINE.md
 * section 4): eight independent dependent-FMA chains per thread, enough resident warps
 * to saturate the FP64 pipe of every SM.  Reports thread-level FMA instructions per second.
 */
#include <cuda_runtime.h>
#include "../../include/pht_b200.h"

__global__ void __launch_bounds__(256) k_fma_chain(double *out, int iters, double a, double b) {
    double x0 = threadIdx.x * 1e-9, x1 = x0 + 1e-3, x2 = x0 + 2e-3, x3 = x0 + 3e-3,
           x4 = x0 + 4e-3, x5 = x0 + 5e-3, x6 = x0 + 6e-3, x7 = x0 + 7e-3;
    for (int i = 0; i < iters; i++) {
        x0 = __fma_rn(x0, a, b); x1 = __fma_rn(x1, a, b); x2 = __fma_rn(x2, a, b); x3 = __fma_rn(x3, a, b);
        x4 = __fma_rn(x4, a, b); x5 = __fma_rn(x5, a, b); x6 = __fma_rn(x6, a, b); x7 = __fma_rn(x7, a, b);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
}

extern "C" int pht_fp64_fma_rate(int device, double *fma_per_s) {
    if (!fma_per_s) return -1;
    if (cudaSetDevice(device) != cudaSuccess) return -1;
    int sms = 0;
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device) != cudaSuccess) return -1;
    const int blocks = sms * 8, threads = 256, iters = 1 << 16;
    double *d = nullptr;
    if (cudaMalloc(&d, sizeof(double) * blocks * threads) != cudaSuccess) return -1;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int rep = 0; rep < 5; rep++) {
        cudaEventRecord(e0);
        k_fma_chain<<<blocks, threads>>>(d, iters, 0.999999, 1e-7);
        cudaEventRecord(e1);
        if (cudaEventSynchronize(e1) != cudaSuccess) { cudaFree(d); return -1; }
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (rep > 0 && ms < best) best = ms;
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(d);
    *fma_per_s = (double)blocks * threads * (double)iters * 8.0 / (best * 1e-3);
    return 0;
}
