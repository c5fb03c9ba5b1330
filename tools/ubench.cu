// tools/ubench.cu — micro-benchmarks behind the design choices of grouphash.cuh / partition.cuh (measurement aid, not part of the library):
// shared-memory atomics on spread addresses (the group kernel's probe + count), L2 atomics on a few 10^4 counters (the partition's
// histogram and cursors) and scattered 16-byte stores behind an atomic cursor (the partition's write pass).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo tools/ubench.cu -o tools/ubench
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e_)); exit(1); } } while (0)

__device__ __forceinline__ uint32_t lcg(uint32_t &s) { s = s * 1664525u + 1013904223u; return s >> 8; }

// MODE 0: atomicAdd, 1: atomicCAS (always succeeds on 0 -> value pattern is irrelevant for timing), 2: plain load + store, 3: load only
template<int MODE>
__global__ void __launch_bounds__(512) smem_atomics_kernel(int iters, int table_words, uint32_t *sink) {
    extern __shared__ uint32_t sm[];
    for (int i = threadIdx.x; i < table_words; i += blockDim.x) sm[i] = 0;
    __syncthreads();
    uint32_t s = threadIdx.x * 2654435761u + blockIdx.x * 97u + 1u, acc = 0;
    for (int it = 0; it < iters; ++it) {
        const uint32_t a = lcg(s) % (uint32_t) table_words;
        if (MODE == 0) acc += atomicAdd(&sm[a], 1u);
        else if (MODE == 1) acc += atomicCAS(&sm[a], (uint32_t) it, (uint32_t) it + 1u);
        else if (MODE == 2) { const uint32_t v = sm[a]; sm[a] = v + 1u; acc += v; }
        else acc += sm[a];
    }
    if (acc == 0xdeadbeefu) sink[0] = acc;
}

// MODE 0: red (result unused), 1: atom (result used)
template<int MODE>
__global__ void __launch_bounds__(256) gmem_atomics_kernel(int iters, uint32_t n_counters, uint32_t *counters, uint32_t *sink) {
    uint32_t s = (blockIdx.x * blockDim.x + threadIdx.x) * 2654435761u + 12345u, acc = 0;
    for (int it = 0; it < iters; ++it) {
        const uint32_t a = lcg(s) % n_counters;
        if (MODE == 0) atomicAdd(&counters[a], 1u);
        else acc += atomicAdd(&counters[a], 1u);
    }
    if (acc == 0xdeadbeefu) sink[0] = acc;
}

// the write pass in miniature: claim a slot of a random group, store 16 (or 32) bytes there
template<int W>
__global__ void __launch_bounds__(256) scatter_kernel(int iters, uint32_t n_groups, uint32_t *cursors, uint64_t *out) {
    uint32_t s = (blockIdx.x * blockDim.x + threadIdx.x) * 2654435761u + 777u;
    for (int it = 0; it < iters; it += 4) {
        uint32_t g[4], p[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) g[i] = lcg(s) % n_groups;
#pragma unroll
        for (int i = 0; i < 4; ++i) p[i] = atomicAdd(&cursors[g[i]], 1u);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            ulonglong2 v = make_ulonglong2(g[i], p[i]);
            ulonglong2 *dst = reinterpret_cast<ulonglong2 *>(out + (size_t) p[i] * W);
#pragma unroll
            for (int j = 0; j < W / 2; ++j) dst[j] = v;
        }
    }
}

__global__ void fill_cursors(uint32_t *c, uint32_t n_groups, uint32_t per_group) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_groups) c[i] = i * per_group;
}

template<class F>
static float timed(F f, int reps = 5) {
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    f();
    CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < reps; ++r) {
        cudaEventRecord(a);
        f();
        cudaEventRecord(b);
        CK(cudaEventSynchronize(b));
        float ms;
        cudaEventElapsedTime(&ms, a, b);
        if (ms < best) best = ms;
    }
    return best;
}

int main() {
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    const int sms = prop.multiProcessorCount;
    int clk_khz = 0;
    cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
    printf("device %s, %d SMs, max clock %.0f MHz\n", prop.name, sms, clk_khz / 1e3);
    uint32_t *sink;
    CK(cudaMalloc(&sink, 64));

    // ---- shared-memory atomics: 2 CTAs x 512 threads per SM, 8192-word table (32 KB) ----
    {
        const int iters = 4096, words = 8192, threads = 512, ctas = sms * 2;
        const double ops = (double) iters * threads * ctas;
        const char *names[4] = {"ATOMS.ADD spread", "ATOMS.CAS spread", "LDS + STS spread", "LDS spread"};
        for (int mode = 0; mode < 4; ++mode) {
            float ms = timed([&] {
                if (mode == 0) smem_atomics_kernel<0><<<ctas, threads, words * 4>>>(iters, words, sink);
                if (mode == 1) smem_atomics_kernel<1><<<ctas, threads, words * 4>>>(iters, words, sink);
                if (mode == 2) smem_atomics_kernel<2><<<ctas, threads, words * 4>>>(iters, words, sink);
                if (mode == 3) smem_atomics_kernel<3><<<ctas, threads, words * 4>>>(iters, words, sink);
            });
            printf("%-20s %8.3f ms  %7.1f Gop/s chip  %6.2f ns per warp-instruction per SM\n", names[mode], ms, ops / ms / 1e6,
                   ms * 1e6 / (ops / 32 / sms));
        }
    }
    // ---- L2 atomics on n counters ----
    {
        const int iters = 256, threads = 256, ctas = sms * 32;
        const double ops = (double) iters * threads * ctas;
        uint32_t *counters;
        CK(cudaMalloc(&counters, 1u << 24));
        for (uint32_t n : {8u, 1024u, 40960u, 163840u, 1u << 22}) {
            CK(cudaMemset(counters, 0, 1u << 24));
            float ms0 = timed([&] { gmem_atomics_kernel<0><<<ctas, threads>>>(iters, n, counters, sink); });
            float ms1 = timed([&] { gmem_atomics_kernel<1><<<ctas, threads>>>(iters, n, counters, sink); });
            printf("L2 atomics on %8u counters: RED %8.3f ms %7.1f Gop/s   ATOM %8.3f ms %7.1f Gop/s\n", n, ms0, ops / ms0 / 1e6, ms1, ops / ms1 / 1e6);
        }
        cudaFree(counters);
    }
    // ---- scattered stores behind an atomic cursor: 291 M records of 16 B into 40960 groups ----
    for (int W : {2, 4}) {
        const uint32_t n_groups = 40960;
        const int threads = 256, ctas = sms * 16, iters = 480;
        const uint64_t total = (uint64_t) iters * threads * ctas;
        const uint32_t per_group = (uint32_t) (total / n_groups * 5 / 4 + 4096);
        uint32_t *cursors;
        uint64_t *out;
        CK(cudaMalloc(&cursors, n_groups * 4));
        CK(cudaMalloc(&out, (size_t) n_groups * per_group * W * 8));
        float best = 1e30f;
        for (int r = 0; r < 3; ++r) {
            fill_cursors<<<(n_groups + 255) / 256, 256>>>(cursors, n_groups, per_group);
            float ms = timed([&] {
                fill_cursors<<<(n_groups + 255) / 256, 256>>>(cursors, n_groups, per_group);
                if (W == 2) scatter_kernel<2><<<ctas, threads>>>(iters, n_groups, cursors, out);
                else scatter_kernel<4><<<ctas, threads>>>(iters, n_groups, cursors, out);
            }, 1);
            if (ms < best) best = ms;
        }
        printf("scatter %llu records of %d B into %u groups: %8.3f ms  %7.1f GB/s written\n", (unsigned long long) total, W * 8, n_groups, best,
               (double) total * W * 8 / best / 1e6);
        cudaFree(cursors); cudaFree(out);
    }
    return 0;
}
