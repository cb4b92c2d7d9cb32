/*
 * k_sort.cu -- observation layout for the MHRS sweep: y in decreasing order.
 *
 * The number of rejection attempts an observation needs is Geometric(survival(y)), so its cost is known in
 * advance up to noise.  The lane phase of k_mhrs.cu walks the observations in layout order; with the expensive
 * ones first, the stream ends on the cheapest observations and the grid drains in microseconds instead of
 * handing a hundred thousand half-done observations to the tail.  Paths are keyed by the GLOBAL observation
 * index (perm carries it along), so the layout changes nothing in the results.
 * Done once per engine, on the device (CUB radix sort: 64-bit keys, 32-bit payload).
 */
#include <cub/device/device_radix_sort.cuh>
#include "engine_internal.h"

__global__ void k_iota(uint32_t *v, long l) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < l) v[i] = (uint32_t)i;
}
__global__ void k_gather_u8(const uint8_t *src, const uint32_t *perm, uint8_t *dst, long l) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < l) dst[i] = src[perm[i]];
}

/* ys[i] = y[perm[i]] non-increasing, cs[i] = cens[perm[i]]; all pointers device memory of length l */
cudaError_t pht_sort_by_y_desc(const double *y, const uint8_t *cens, long l, double *ys, uint8_t *cs, uint32_t *perm, cudaStream_t st) {
    if (l <= 0) return cudaSuccess;
    uint32_t *idx = nullptr; void *tmp = nullptr; size_t tmp_bytes = 0;
    cudaError_t e = pht_dev_alloc((void **)&idx, sizeof(uint32_t) * (size_t)l, st);
    if (e != cudaSuccess) return e;
    const int threads = 256; const unsigned blocks = (unsigned)((l + threads - 1) / threads);
    k_iota<<<blocks, threads, 0, st>>>(idx, l);
    e = cub::DeviceRadixSort::SortPairsDescending(nullptr, tmp_bytes, y, ys, idx, perm, (int)l, 0, 64, st);
    if (e == cudaSuccess) e = pht_dev_alloc(&tmp, tmp_bytes ? tmp_bytes : 1, st);
    if (e == cudaSuccess) e = cub::DeviceRadixSort::SortPairsDescending(tmp, tmp_bytes, y, ys, idx, perm, (int)l, 0, 64, st);
    if (e == cudaSuccess) { k_gather_u8<<<blocks, threads, 0, st>>>(cens, perm, cs, l); e = cudaGetLastError(); }
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    pht_dev_free(tmp, st);
    pht_dev_free(idx, st);
    return e;
}
