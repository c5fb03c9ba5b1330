// coverage.cu — the coverage map over the (k+1)-mers and the per-edge coverage of the condensed graph.
//
// Replaces (reference file:line):
//   CoverageHashMapBuilder::BuildIndex / FillCoverageFromStream      C/utils/ph_map/coverage_hash_map_builder.hpp:15-54
//       PerfectHashMap<RtSeq, uint32_t>: a KMerIndex over the (k+1)-mer storage + ++count[idx(x)] for every minimal valid window of a
//       SECOND pass over the fwd+RC read streams.  On the GPU the multiplicities fell out of the counting sort (run lengths, a
//       self-reverse-complement (k+1)-mer doubled: count.cu), so the map is the same BooPHF index built over the (k+1)-mers plus the
//       counts moved from file order to index order — no second pass over the reads.
//   GraphCoverageFiller::FillCoverageFromEdges / FillCoverageAndFlankingFromPHM   C/assembly_graph/graph_support/coverage_filling.hpp:45-95
//       per edge: sum of the map's values over the edge's (k+1)-mers (CoverageIndex raw coverage = GFA KC:i, DP:f = KC / length) and the
//       same over its first `averaging_range` (k+1)-mers (FlankingCoverage; the conjugate edge's flank is this edge's tail).
#include "../../include/sb200.h"
#include "common.cuh"
#include "kmer_ops.cuh"
#include "kmer_set.cuh"
#include "mphf.cuh"
#include "radix_sort.cuh"

struct sb200_covmap {
    sb200_ctx *ctx = nullptr;
    sb200_mphf *mphf = nullptr;       // KMerIndex over the (k+1)-mers
    DevBuf<uint32_t> values;          // PerfectHashMap::data_: multiplicity per (k+1)-mer in index order
    uint64_t size = 0;
    unsigned K = 0, words = 0;
    ~sb200_covmap() { delete mphf; }
};

namespace sb200 {

sb200_mphf *mphf_build(sb200_ctx *ctx, const sb200_kmers *ks, const uint64_t *global_sizes);
MphfDev mphf_dev(const sb200_mphf *m);

__global__ void __launch_bounds__(256) counts_from_place_kernel(const uint32_t *__restrict__ place, const uint32_t *__restrict__ pc_scan,
                                                               const uint64_t *__restrict__ bits, uint64_t n, const uint32_t *__restrict__ counts,
                                                               uint32_t *__restrict__ values) {
    uint64_t i = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t g = place[i];
    const uint32_t w = g >> 6;
    const uint32_t id = __ldg(pc_scan + w) + (uint32_t) __popcll(__ldg(bits + w) & ((1ULL << (g & 63u)) - 1ULL));
    values[id] = counts[i];
}

template<int W>
__global__ void __launch_bounds__(256) counts_by_lookup_kernel(MphfDev m, const uint64_t *__restrict__ recs, uint64_t n, const uint32_t *__restrict__ counts,
                                                              uint32_t *__restrict__ values) {
    uint64_t i = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint64_t r[W];
    load_rec<W>(recs, i, r);
    values[mphf_lookup<W>(m, r)] = counts[i];
}

// one warp per sequence, lane = window (stride 32): canonical (k+1)-mer -> index -> multiplicity
template<int W>
__global__ void __launch_bounds__(256) unitig_coverage_kernel(MphfDev m, const uint32_t *__restrict__ values, const uint64_t *__restrict__ words,
                                                             const uint64_t *__restrict__ word_off, const uint32_t *__restrict__ len, uint64_t n_seq, int K,
                                                             uint32_t avg_range, unsigned long long *__restrict__ kc,
                                                             unsigned long long *__restrict__ flank) {
    const int lane = threadIdx.x & 31;
    const uint64_t s = ((uint64_t) blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (s >= n_seq) return;
    const uint32_t l = len[s];
    unsigned long long sum = 0, fs = 0, fe = 0;
    if (l >= (uint32_t) K) {
        const uint32_t nwin = l - (uint32_t) K + 1, nw = (l + 31) >> 5;
        const uint32_t range = avg_range < nwin ? avg_range : nwin;
        const uint64_t *seq = words + word_off[s];
        for (uint32_t p = lane; p < nwin; p += 32) {
            uint64_t x[W], c[W];
            kmer_window<W>(seq, nw, p, K, x);
            kmer_canonical<W>(x, K, c);
            const uint64_t id = mphf_lookup<W>(m, c);
            const unsigned long long v = id == ~0ULL ? 0ULL : (unsigned long long) __ldg(values + id);
            sum += v;
            if (p < range) fs += v;
            if (p >= nwin - range) fe += v;
        }
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        sum += __shfl_down_sync(0xffffffffu, sum, d);
        fs += __shfl_down_sync(0xffffffffu, fs, d);
        fe += __shfl_down_sync(0xffffffffu, fe, d);
    }
    if (lane == 0) {
        kc[s] = sum;
        if (flank) { flank[2 * s] = fs; flank[2 * s + 1] = fe; }
    }
}

static sb200_covmap *covmap_build(sb200_ctx *ctx, const sb200_kmers *kp) {
    SB200_REQUIRE(kp->counts.p != nullptr, "this k-mer set carries no multiplicities");
    sb200_covmap *c = new sb200_covmap();
    c->ctx = ctx; c->size = kp->size; c->K = kp->k; c->words = kp->words;
    try {
        c->mphf = mphf_build(ctx, kp, nullptr);
        c->values.alloc(ctx, kp->size + 1);
        const sb200_mphf *m = c->mphf;
        if (kp->size == 0) {
        } else if (m->place.p && m->pc_scan.p && !ctx->no_place) {
            LAUNCH(ctx, counts_from_place_kernel, div_up(kp->size, 256), 256, 0, m->place.p, m->pc_scan.p, m->bits.p, kp->size, kp->counts.p, c->values.p);
        } else {
            MphfDev d = mphf_dev(m);
            switch (kp->words) {
                case 1: LAUNCH(ctx, counts_by_lookup_kernel<1>, div_up(kp->size, 256), 256, 0, d, kp->data.p, kp->size, kp->counts.p, c->values.p); break;
                case 2: LAUNCH(ctx, counts_by_lookup_kernel<2>, div_up(kp->size, 256), 256, 0, d, kp->data.p, kp->size, kp->counts.p, c->values.p); break;
                case 3: LAUNCH(ctx, counts_by_lookup_kernel<3>, div_up(kp->size, 256), 256, 0, d, kp->data.p, kp->size, kp->counts.p, c->values.p); break;
                default: LAUNCH(ctx, counts_by_lookup_kernel<4>, div_up(kp->size, 256), 256, 0, d, kp->data.p, kp->size, kp->counts.p, c->values.p); break;
            }
        }
        CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    } catch (...) {
        delete c;
        throw;
    }
    return c;
}

static void unitigs_coverage(sb200_ctx *ctx, const sb200_covmap *c, const sb200_unitigs *u, uint32_t avg_range, uint64_t *kc_out, uint64_t *flank_out) {
    SB200_REQUIRE(u->k + 1 == c->K, "the coverage map was built over (k+1)-mers of another k");
    const uint64_t n = u->count;
    if (n == 0) return;
    DevBuf<unsigned long long> kc(ctx, n), fl;
    if (flank_out) fl.alloc(ctx, 2 * n);
    MphfDev d = mphf_dev(c->mphf);
    const unsigned grid = div_up(n * 32, 256);
    switch (c->words) {
        case 1: LAUNCH(ctx, unitig_coverage_kernel<1>, grid, 256, 0, d, c->values.p, u->words.p, u->word_off.p, u->len.p, n, (int) c->K, avg_range, kc.p, fl.p); break;
        case 2: LAUNCH(ctx, unitig_coverage_kernel<2>, grid, 256, 0, d, c->values.p, u->words.p, u->word_off.p, u->len.p, n, (int) c->K, avg_range, kc.p, fl.p); break;
        case 3: LAUNCH(ctx, unitig_coverage_kernel<3>, grid, 256, 0, d, c->values.p, u->words.p, u->word_off.p, u->len.p, n, (int) c->K, avg_range, kc.p, fl.p); break;
        default: LAUNCH(ctx, unitig_coverage_kernel<4>, grid, 256, 0, d, c->values.p, u->words.p, u->word_off.p, u->len.p, n, (int) c->K, avg_range, kc.p, fl.p); break;
    }
    CUDA_CHECK(cudaMemcpyAsync(kc_out, kc.p, n * 8, cudaMemcpyDeviceToHost, ctx->stream));
    if (flank_out) CUDA_CHECK(cudaMemcpyAsync(flank_out, fl.p, 2 * n * 8, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
}

}  // namespace sb200

template<class F>
static int guarded(sb200_ctx *ctx, F &&f) {
    try {
        CUDA_CHECK(cudaSetDevice(ctx->device));
        f();
        return 0;
    } catch (const sb200_error &e) {
        ctx->last_error = e.what();
        cudaGetLastError();
        return e.code;
    } catch (const std::exception &e) {
        ctx->last_error = e.what();
        return 3;
    }
}

extern "C" {

int sb200_coverage_map_build(sb200_ctx *ctx, const sb200_kmers *kpomers, sb200_covmap **out) {
    *out = nullptr;
    return guarded(ctx, [&] { *out = sb200::covmap_build(ctx, kpomers); });
}
const sb200_mphf *sb200_coverage_map_index(const sb200_covmap *c) { return c->mphf; }
uint64_t sb200_coverage_map_size(const sb200_covmap *c) { return c->size; }
int sb200_coverage_map_values_download(const sb200_covmap *c, uint32_t *values_out) {
    return guarded(c->ctx, [&] {
        CUDA_CHECK(cudaMemcpyAsync(values_out, c->values.p, c->size * 4, cudaMemcpyDeviceToHost, c->ctx->stream));
        CUDA_CHECK(cudaStreamSynchronize(c->ctx->stream));
    });
}
void sb200_coverage_map_free(sb200_covmap *c) {
    if (!c) return;
    cudaSetDevice(c->ctx->device);
    delete c;
}
int sb200_unitigs_coverage(sb200_ctx *ctx, const sb200_covmap *c, const sb200_unitigs *u, uint32_t averaging_range, uint64_t *kc_out, uint64_t *flank_out) {
    return guarded(ctx, [&] { sb200::unitigs_coverage(ctx, c, u, averaging_range, kc_out, flank_out); });
}

}  // extern "C"
