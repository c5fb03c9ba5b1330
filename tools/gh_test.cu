// Standalone harness for group_hash_kernel (grouphash.cuh): random groups vs a CPU sort/unique.  Debug tool, not part of the library.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo --expt-relaxed-constexpr -I spades_for_blackbird_b200/csrc -I include tools/gh_test.cu -o tools/gh_test
//   (profiles/r2_racecheck_gh_test.log: compute-sanitizer --tool racecheck on this harness)
#include "grouphash.cuh"
#include <algorithm>
#include <cstdio>
#include <map>
#include <random>
#include <vector>
using namespace sb200;

int main(int argc, char **argv) {
    const int G = argc > 1 ? atoi(argv[1]) : 10, N = argc > 2 ? atoi(argv[2]) : 4600, D = argc > 3 ? atoi(argv[3]) : 1300;
    const int reps = argc > 4 ? atoi(argv[4]) : 5;
    constexpr int W = 2;
    std::mt19937_64 rng(12345);
    std::vector<uint64_t> recs;
    std::vector<ChunkRange> ranges;
    std::vector<std::map<std::pair<uint64_t, uint64_t>, uint32_t>> truth(G);
    for (int g = 0; g < G; ++g) {
        std::vector<std::pair<uint64_t, uint64_t>> keys(D);
        for (auto &k : keys) { k.first = rng(); k.second = rng() & ((1ULL << 48) - 1); }
        for (int i = 0; i < D / 10; ++i) { keys[i].first = keys[D - 1 - i].first; }   // tag ties
        uint32_t s = recs.size() / W;
        for (int i = 0; i < N; ++i) {
            auto &k = keys[(i < D) ? i : rng() % (D / 4 + 1)];
            recs.push_back(k.first); recs.push_back(k.second);
            truth[g][k]++;
        }
        // shuffle the group
        for (int i = N - 1; i > 0; --i) {
            int j = rng() % (i + 1);
            std::swap(recs[(s + i) * W], recs[(s + j) * W]); std::swap(recs[(s + i) * W + 1], recs[(s + j) * W + 1]);
        }
        ranges.push_back(ChunkRange{s, s + (uint32_t) N});
    }
    uint64_t *d_recs, *d_out; ChunkRange *d_ranges; uint32_t *d_gu, *d_ctrl, *d_cnt;
    size_t n = recs.size() / W;
    cudaMalloc(&d_recs, n * W * 8); cudaMalloc(&d_out, n * W * 8); cudaMalloc(&d_ranges, G * sizeof(ChunkRange));
    cudaMalloc(&d_gu, (G + 1) * 4); cudaMalloc(&d_ctrl, 16); cudaMalloc(&d_cnt, n * 4);
    cudaMemcpy(d_recs, recs.data(), n * W * 8, cudaMemcpyHostToDevice);
    cudaMemcpy(d_ranges, ranges.data(), G * sizeof(ChunkRange), cudaMemcpyHostToDevice);
    auto kern = group_hash_kernel<W, 1, uint16_t, false>;
    size_t smem = group_hash_smem<uint16_t>();
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem);
    int total_bad = 0;
    for (int rep = 0; rep < reps; ++rep) {
        cudaMemset(d_out, 0xEE, n * W * 8); cudaMemset(d_cnt, 0, n * 4); cudaMemset(d_ctrl, 0, 16);
        kern<<<G, HashCfg::THREADS, smem>>>(d_recs, d_ranges, d_gu, d_ctrl, d_out, d_cnt, 64, ~0ULL, 0, (const uint8_t *) nullptr, GroupParts{nullptr, nullptr, 1u, (uint32_t) G});
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("rep %d: CUDA error %s\n", rep, cudaGetErrorString(e)); return 1; }
        std::vector<uint64_t> out(n * W); std::vector<uint32_t> gu(G), cnt(n), ctrl(4);
        cudaMemcpy(out.data(), d_out, n * W * 8, cudaMemcpyDeviceToHost); cudaMemcpy(gu.data(), d_gu, G * 4, cudaMemcpyDeviceToHost);
        cudaMemcpy(cnt.data(), d_cnt, n * 4, cudaMemcpyDeviceToHost); cudaMemcpy(ctrl.data(), d_ctrl, 16, cudaMemcpyDeviceToHost);
        int bad = 0;
        for (int g = 0; g < G; ++g) {
            if (gu[g] != truth[g].size()) { printf("rep %d group %d: U %u want %zu\n", rep, g, gu[g], truth[g].size()); ++bad; continue; }
            size_t i = ranges[g].s;
            int gb = 0, first = -1;
            for (auto &kv : truth[g]) {
                if (out[i * W] != kv.first.first || out[i * W + 1] != kv.first.second || cnt[i] != kv.second) {
                    ++gb; if (first < 0) first = (int) (i - ranges[g].s);
                    if (gb <= 6 && rep == 0 && g == 0) printf("  row %zu got %016llx %016llx x%u want %016llx %016llx x%u\n", i - ranges[g].s, (unsigned long long) out[i * W],
                        (unsigned long long) out[i * W + 1], cnt[i], (unsigned long long) kv.first.first, (unsigned long long) kv.first.second, kv.second);
                }
                ++i;
            }
            if (gb) { printf("rep %d group %d: %d bad rows of %zu, first %d\n", rep, g, gb, truth[g].size(), first); ++bad; }
        }
        printf("rep %d: %d bad groups, ctrl %u %u %x %u\n", rep, bad, ctrl[0], ctrl[1], ctrl[2], ctrl[3]);
        total_bad += bad;
    }
    return total_bad != 0;
}
