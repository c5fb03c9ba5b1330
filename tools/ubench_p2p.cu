// ubench_p2p.cu — what a kernel on GPU 0 can push into GPU 1's HBM over NVLink, against the copy engine:
//   st16      every thread stores 16 B, consecutive threads consecutive addresses (what sp_flush does)
//   bulk      cp.async.bulk shared::cta -> global (peer) of `chunk` bytes per elected thread, waited with bulk_group
//   memcpy    cudaMemcpyPeerAsync (copy engine)
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/ubench_p2p tools/ubench_p2p.cu
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e_)); exit(1); } } while (0)

__global__ void __launch_bounds__(512) st16_kernel(uint4 *__restrict__ dst, uint64_t n16, int burst_gap) {
    const uint64_t nth = (uint64_t) gridDim.x * blockDim.x;
    uint4 v = make_uint4(threadIdx.x, blockIdx.x, 3, 4);
    for (uint64_t i = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += nth) {
        dst[i] = v;
        v.x += 1;
        for (int g = 0; g < burst_gap; ++g) v.y = v.y * 1664525u + 1013904223u;   // filler ALU work between stores
    }
}

__global__ void __launch_bounds__(128) bulk_kernel(uint8_t *__restrict__ dst, uint64_t bytes, uint32_t chunk) {
    extern __shared__ __align__(128) unsigned char sm[];
    for (uint32_t i = threadIdx.x; i < chunk / 4; i += blockDim.x) reinterpret_cast<uint32_t *>(sm)[i] = i + blockIdx.x;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (threadIdx.x == 0) {
        const uint32_t s = (uint32_t) __cvta_generic_to_shared(sm);
        int inflight = 0;
        for (uint64_t off = (uint64_t) blockIdx.x * chunk; off + chunk <= bytes; off += (uint64_t) gridDim.x * chunk) {
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst + off), "r"(s), "r"(chunk) : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            if (++inflight >= 8) { asm volatile("cp.async.bulk.wait_group.read 4;" ::: "memory"); inflight = 4; }
        }
        asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
}

int main() {
    int n = 0;
    CK(cudaGetDeviceCount(&n));
    if (n < 2) { printf("needs 2 GPUs\n"); return 0; }
    int can = 0;
    CK(cudaDeviceCanAccessPeer(&can, 0, 1));
    printf("peer access 0 -> 1: %d\n", can);
    CK(cudaSetDevice(1));
    const uint64_t bytes = 2ull << 30;
    uint8_t *remote, *local_src, *local_dst;
    CK(cudaMalloc(&remote, bytes));
    CK(cudaSetDevice(0));
    CK(cudaDeviceEnablePeerAccess(1, 0));
    CK(cudaMalloc(&local_src, bytes));
    CK(cudaMalloc(&local_dst, bytes));
    CK(cudaMemset(local_src, 1, bytes));
    int sms = 0;
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    auto timeit = [&](const char *name, auto fn) {
        float best = 1e30f;
        for (int rep = 0; rep < 3; ++rep) {
            CK(cudaEventRecord(e0));
            fn();
            CK(cudaEventRecord(e1));
            CK(cudaEventSynchronize(e1));
            float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
            if (ms < best) best = ms;
        }
        CK(cudaGetLastError());
        printf("%-44s %8.3f ms  %7.1f GB/s\n", name, best, bytes / best / 1e6);
    };
    char nm[128];
    for (int gap : {0, 8, 64}) {
        for (int per_sm : {1, 2, 4}) {
            snprintf(nm, sizeof nm, "st16 -> peer   %d CTAs/SM x 512, gap %d", per_sm, gap);
            timeit(nm, [&] { st16_kernel<<<sms * per_sm, 512>>>((uint4 *) remote, bytes / 16, gap); });
        }
    }
    timeit("st16 -> local  4 CTAs/SM x 512, gap 0", [&] { st16_kernel<<<sms * 4, 512>>>((uint4 *) local_dst, bytes / 16, 0); });
    for (uint32_t chunk : {2048u, 8192u, 32768u}) {
        CK(cudaFuncSetAttribute(bulk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) chunk));
        for (int per_sm : {1, 2, 4}) {
            snprintf(nm, sizeof nm, "bulk smem -> peer  %d CTAs/SM, %u B chunks", per_sm, chunk);
            timeit(nm, [&] { bulk_kernel<<<sms * per_sm, 128, chunk>>>(remote, bytes, chunk); });
        }
    }
    timeit("bulk smem -> local 2 CTAs/SM, 8192 B chunks", [&] { CK(cudaFuncSetAttribute(bulk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 8192)); bulk_kernel<<<sms * 2, 128, 8192>>>(local_dst, bytes, 8192); });
    timeit("cudaMemcpyPeerAsync (copy engine)", [&] { CK(cudaMemcpyPeerAsync(remote, 1, local_src, 0, bytes, 0)); });
    cudaStream_t s2[4];
    for (auto &s : s2) CK(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
    timeit("4 x cudaMemcpyAsync on 4 streams", [&] {
        for (int i = 0; i < 4; ++i) CK(cudaMemcpyAsync(remote + i * (bytes / 4), local_src + i * (bytes / 4), bytes / 4, cudaMemcpyDefault, s2[i]));
        for (int i = 0; i < 4; ++i) CK(cudaStreamSynchronize(s2[i]));
    });
    return 0;
}
