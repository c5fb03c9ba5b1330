// scan.cuh — device-wide exclusive prefix sum (reduce / recursive scan of tile sums / apply), in place.
// Used for bucket/bin offsets, stream compaction and output offsets on every stage of the path.
#pragma once
#include "common.cuh"

namespace sb200 {

constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ITEMS = 16;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;

template<class T>
__device__ __forceinline__ T warp_inclusive_scan(T v) {
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        T o = __shfl_up_sync(0xffffffffu, v, d);
        if ((threadIdx.x & 31) >= d) v += o;
    }
    return v;
}

// exclusive scan of one value per thread across the block; returns the exclusive prefix, *total = block sum
template<class T, int THREADS>
__device__ __forceinline__ T block_exclusive_scan(T v, T *total, T *smem /* THREADS/32 + 1 */) {
    int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    T inc = warp_inclusive_scan(v);
    if (lane == 31) smem[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        T s = lane < THREADS / 32 ? smem[lane] : T(0);
        T si = warp_inclusive_scan(s);
        if (lane < THREADS / 32) smem[lane] = si - s;
        if (lane == THREADS / 32 - 1) smem[THREADS / 32] = si;
    }
    __syncthreads();
    T res = smem[warp] + inc - v;
    *total = smem[THREADS / 32];
    __syncthreads();
    return res;
}

template<class T>
__global__ void __launch_bounds__(SCAN_THREADS) scan_reduce_kernel(const T *__restrict__ in, uint64_t n, T *__restrict__ sums) {
    __shared__ T sm[SCAN_THREADS / 32 + 1];
    uint64_t base = (uint64_t) blockIdx.x * SCAN_TILE;
    T acc = 0;
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; ++i) {
        uint64_t idx = base + (uint64_t) i * SCAN_THREADS + threadIdx.x;
        if (idx < n) acc += in[idx];
    }
    T total;
    block_exclusive_scan<T, SCAN_THREADS>(acc, &total, sm);
    if (threadIdx.x == 0) sums[blockIdx.x] = total;
}

// offsets == nullptr: single tile, offset 0.  data is scanned in place; if total != nullptr and this is the last
// block, the grand total is written there.
template<class T>
__global__ void __launch_bounds__(SCAN_THREADS) scan_apply_kernel(T *__restrict__ data, uint64_t n, const T *__restrict__ offsets,
                                                                 T *__restrict__ total_out) {
    __shared__ T sm[SCAN_THREADS / 32 + 1];
    uint64_t base = (uint64_t) blockIdx.x * SCAN_TILE + (uint64_t) threadIdx.x * SCAN_ITEMS;
    T v[SCAN_ITEMS];
    T acc = 0;
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; ++i) {
        v[i] = (base + i < n) ? data[base + i] : T(0);
        acc += v[i];
    }
    T total;
    T pre = block_exclusive_scan<T, SCAN_THREADS>(acc, &total, sm);
    T off = offsets ? offsets[blockIdx.x] : T(0);
    pre += off;
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; ++i) {
        if (base + i < n) data[base + i] = pre;
        pre += v[i];
    }
    if (total_out && blockIdx.x == gridDim.x - 1 && threadIdx.x == 0) *total_out = off + total;
}

// In-place exclusive scan of data[0..n); optionally writes the total to *total_dev (device pointer).
template<class T>
void exclusive_scan(sb200_ctx *ctx, T *data, uint64_t n, T *total_dev) {
    if (n == 0) {
        if (total_dev) CUDA_CHECK(cudaMemsetAsync(total_dev, 0, sizeof(T), ctx->stream));
        return;
    }
    unsigned nb = div_up(n, SCAN_TILE);
    if (nb == 1) {
        LAUNCH(ctx, scan_apply_kernel<T>, 1, SCAN_THREADS, 0, data, n, (const T *) nullptr, total_dev);
        return;
    }
    DevBuf<T> sums(ctx, nb);
    LAUNCH(ctx, scan_reduce_kernel<T>, nb, SCAN_THREADS, 0, data, n, sums.p);
    exclusive_scan<T>(ctx, sums.p, nb, nullptr);
    LAUNCH(ctx, scan_apply_kernel<T>, nb, SCAN_THREADS, 0, data, n, sums.p, total_dev);
}

}  // namespace sb200
