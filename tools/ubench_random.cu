// ubench_random.cu — dependent random 128-byte-line reads over buffers of growing size: what a lookup walk step costs when the
// walk blocks (mphf.cuh) of a job no longer fit L2 / the TLBs.   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/ubench_random tools/ubench_random.cu
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e_)); exit(1); } } while (0)

__device__ __forceinline__ uint64_t mix(uint64_t x) {
    x ^= x >> 33; x *= 0xff51afd7ed558ccdULL; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL; x ^= x >> 33;
    return x;
}

// every thread chases `steps` dependent reads: the next line index is a hash of the word just read (+ ~150 ALU instructions of filler
// when `work` is set, the hash cost of a real step)
__global__ void __launch_bounds__(256) chase_kernel(const uint64_t *__restrict__ buf, uint64_t n_lines, int steps, int work, uint64_t *sink) {
    uint64_t h = mix((uint64_t) blockIdx.x * blockDim.x + threadIdx.x + 1);
    for (int s = 0; s < steps; ++s) {
        const uint64_t line = __umul64hi(h, n_lines);
        const uint64_t v = __ldg(buf + line * 16 + (h & 3));
        h = mix(h ^ v);
        for (int w = 0; w < work; ++w) h = mix(h + w);
    }
    if (h == 0x1234567) *sink = h;
}

int main() {
    int dev = 0, sms = 0;
    CK(cudaSetDevice(dev));
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    uint64_t *sink;
    CK(cudaMalloc(&sink, 8));
    const double sizes_gb[] = {0.0625, 0.25, 0.5, 1, 2, 4, 8, 16, 32};
    for (int work = 0; work <= 12; work += 12) {
        for (double gb : sizes_gb) {
            const uint64_t bytes = (uint64_t) (gb * (1ull << 30)), n_lines = bytes / 128;
            uint64_t *buf;
            CK(cudaMalloc(&buf, bytes));
            CK(cudaMemset(buf, 0x5a, bytes));
            const int steps = 64, blocks = sms * 6 * 8;
            cudaEvent_t e0, e1;
            CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
            float best = 1e30f;
            for (int rep = 0; rep < 3; ++rep) {
                CK(cudaEventRecord(e0));
                chase_kernel<<<blocks, 256>>>(buf, n_lines, steps, work, sink);
                CK(cudaEventRecord(e1));
                CK(cudaEventSynchronize(e1));
                float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
                if (ms < best) best = ms;
            }
            const double acc = (double) blocks * 256 * steps;
            printf("work %2d  buffer %7.3f GB: %8.3f ms  %6.2f G line reads/s  %7.1f GB/s of 128-byte lines\n", work, gb, best, acc / best / 1e6, acc * 128 / best / 1e6);
            CK(cudaFree(buf));
        }
    }
    return 0;
}
