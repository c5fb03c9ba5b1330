// count.cu — reads -> k-mer instances -> sorted-unique per-bucket k-mer sets with multiplicities.
//
// Replaces, for one GPU (reference file:line):
//   DeBruijnReadKMerSplitter::Split / FillBufferFromSequence     C/utils/kmer_mph/kmer_splitters.hpp:25-41,109-133
//   KMerSortingSplitter::DumpBuffers                             C/utils/kmer_mph/kmer_splitter.hpp:120-167
//   KMerDiskCounter::Count / MergeKMers                          C/utils/kmer_mph/kmer_index_builder.hpp:241-365
//   DeBruijnKMerKMerSplitter::FillBufferFromKMers                C/utils/kmer_mph/kmer_splitters.hpp:159-176
// The reference streams every read and its reverse complement and keeps a window iff it IsMinimal(); here each
// forward window emits min(x, rc(x)) once (a self-reverse-complement window counts twice, exactly as both of the
// reference's streams would count it), so half of the work and none of the RC stream exists.
#include <memory>

#include "common.cuh"
#include "kmer_ops.cuh"
#include "kmer_set.cuh"
#include "radix_sort.cuh"
#include "scan.cuh"
#include "segsort.cuh"
#include "grouphash.cuh"
#include "partition.cuh"
#include "staged_partition.cuh"

namespace sb200 {

enum ExtractMode { MODE_CANON_RC = 0, MODE_ALL_RC = 1, MODE_CANON_FWD = 2, MODE_ALL_FWD = 3 };

__global__ void window_count_kernel(const uint32_t *__restrict__ len, uint64_t n_reads, uint32_t K, uint32_t mult,
                                    uint64_t *__restrict__ cnt) {
    uint64_t r = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (r < n_reads) {
        uint32_t l = len[r];
        cnt[r] = l >= K ? (uint64_t) (l - K + 1) * mult : 0;
    }
}

// One warp per read; lane = window position (stride 32), so the records of a read leave as consecutive,
// fully coalesced W x 8 byte stores.  The packed read (<= a few dozen words) is served from L1 after first touch.
template<int W>
__global__ void __launch_bounds__(256) extract_reads_kernel(const uint64_t *__restrict__ words, const uint64_t *__restrict__ word_off,
                                                           const uint32_t *__restrict__ len, uint64_t n_reads, int K, int mode,
                                                           const uint64_t *__restrict__ out_off, uint64_t *__restrict__ out) {
    const int lane = threadIdx.x & 31;
    const uint64_t warps_total = (uint64_t) gridDim.x * (blockDim.x >> 5);
    for (uint64_t r = (uint64_t) blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); r < n_reads; r += warps_total) {
        uint32_t l = len[r];
        if (l < (uint32_t) K) continue;
        uint32_t nwin = l - K + 1;
        const uint64_t *seq = words + word_off[r];
        uint32_t nw = (l + 31) >> 5;
        uint64_t o = out_off[r];
        for (uint32_t p = lane; p < nwin; p += 32) {
            uint64_t x[W], y[W];
            kmer_window<W>(seq, nw, p, K, x);
            if (mode == MODE_ALL_FWD) {
                store_rec<W>(out, o + p, x);
            } else if (mode == MODE_ALL_RC) {
                kmer_rc<W>(x, K, y);
                store_rec<W>(out, o + 2ull * p, x);
                store_rec<W>(out, o + 2ull * p + 1, y);
            } else {
                bool minimal = kmer_canonical<W>(x, K, y);
                if (mode == MODE_CANON_FWD && !minimal) {
#pragma unroll
                    for (int j = 0; j < W; ++j) y[j] = ~0ULL;   // marker: all-T is never minimal, sorts last
                }
                store_rec<W>(out, o + p, y);
            }
        }
    }
}

// kmer_splitters.hpp:159-176: every stored (k+1)-mer x contributes canon(x[0..k)) and canon(x[1..k]) (its RC
// contributes the same two canonical k-mers, so add_rc only duplicates and is folded away).
// pshift >= 0: every candidate also carries, in the padding bits [pshift, pshift + 3) of its last word, the InOutMask bit this
// (k+1)-mer contributes to it — AddOutgoing(next nucleotide) for the prefix k-mer, AddIncoming(previous nucleotide) for the
// suffix k-mer, mirrored when the k-mer is stored as its reverse complement (kmer_extension_index_builder.hpp:44-59,
// kmer_extension_index.hpp:92-106) — so that the masks fall out of the k-mer sort without a single MPHF lookup.
template<int WS, int W>
__global__ void __launch_bounds__(256) derive_kernel(const uint64_t *__restrict__ kp, uint64_t n, int k, int pshift, uint64_t *__restrict__ out) {
    uint64_t i = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint64_t x[WS], a[W], c[W];
    load_rec<WS>(kp, i, x);
    const uint32_t pnucl = kmer_base(x, 0), nnucl = kmer_base_w<WS>(x, k);
    kmer_subwindow<WS, W>(x, 0, k, a);
    bool minimal = kmer_canonical<W>(a, k, c);
    if (pshift >= 0) c[W - 1] |= (uint64_t) (minimal ? nnucl : 7u - nnucl) << pshift;
    store_rec<W>(out, 2 * i, c);
    kmer_subwindow<WS, W>(x, 1, k, a);
    minimal = kmer_canonical<W>(a, k, c);
    if (pshift >= 0) c[W - 1] |= (uint64_t) (minimal ? pnucl + 4u : 3u - pnucl) << pshift;
    store_rec<W>(out, 2 * i + 1, c);
}

// payload bits off (the LSD fallback sorts whole words)
template<int W>
__global__ void strip_payload_kernel(uint64_t *__restrict__ recs, uint64_t n, uint64_t lw_keep) {
    uint64_t i = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) recs[i * W + (W - 1)] &= lw_keep;
}

// ---- unique --------------------------------------------------------------------------------------------------------
constexpr int UQ_THREADS = 256;
constexpr int UQ_ITEMS = 8;
constexpr int UQ_TILE = UQ_THREADS * UQ_ITEMS;

template<int W>
__device__ __forceinline__ bool is_head(const uint64_t *__restrict__ recs, uint64_t i) {
    if (i == 0) return true;
    uint64_t a[W], b[W];
    load_rec<W>(recs, i, a);
    load_rec<W>(recs, i - 1, b);
    return !kmer_eq<W>(a, b);
}

template<int W>
__global__ void __launch_bounds__(UQ_THREADS) unique_count_kernel(const uint64_t *__restrict__ recs, uint64_t n, uint32_t *__restrict__ tile_cnt) {
    __shared__ uint32_t sm[UQ_THREADS / 32 + 1];
    uint64_t base = (uint64_t) blockIdx.x * UQ_TILE;
    uint32_t c = 0;
#pragma unroll
    for (int it = 0; it < UQ_ITEMS; ++it) {
        uint64_t i = base + (uint64_t) it * UQ_THREADS + threadIdx.x;
        if (i < n && is_head<W>(recs, i)) ++c;
    }
    uint32_t total;
    block_exclusive_scan<uint32_t, UQ_THREADS>(c, &total, sm);
    if (threadIdx.x == 0) tile_cnt[blockIdx.x] = total;
}

// writes unique records and the position of each run head in the sorted array (head_pos has U+1 entries; the last
// one, = n, is written by the host wrapper)
template<int W>
__global__ void __launch_bounds__(UQ_THREADS) unique_write_kernel(const uint64_t *__restrict__ recs, uint64_t n,
                                                                 const uint32_t *__restrict__ tile_off, uint64_t *__restrict__ out,
                                                                 uint32_t *__restrict__ head_pos) {
    __shared__ uint32_t sm[UQ_THREADS / 32 + 1];
    uint64_t base = (uint64_t) blockIdx.x * UQ_TILE + (uint64_t) threadIdx.x * UQ_ITEMS;
    bool h[UQ_ITEMS];
    uint32_t c = 0;
#pragma unroll
    for (int it = 0; it < UQ_ITEMS; ++it) {
        uint64_t i = base + it;
        h[it] = (i < n) && is_head<W>(recs, i);
        c += h[it];
    }
    uint32_t total;
    uint32_t pos = block_exclusive_scan<uint32_t, UQ_THREADS>(c, &total, sm) + tile_off[blockIdx.x];
#pragma unroll
    for (int it = 0; it < UQ_ITEMS; ++it) {
        if (h[it]) {
            uint64_t r[W];
            load_rec<W>(recs, base + it, r);
            store_rec<W>(out, pos, r);
            head_pos[pos] = (uint32_t) (base + it);
            ++pos;
        }
    }
}

// counts[i] = run length; self-reverse-complement records count twice in the fwd+RC canonical mode
// (coverage_hash_map_builder.hpp:31-36: both streams see the window and both copies are minimal)
template<int W>
__global__ void counts_kernel(const uint32_t *__restrict__ head_pos, const uint64_t *__restrict__ recs, uint64_t u, int K,
                              int double_palindromes, uint32_t *__restrict__ counts) {
    uint64_t i = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= u) return;
    uint32_t c = head_pos[i + 1] - head_pos[i];
    if (double_palindromes) {
        uint64_t x[W], r[W];
        load_rec<W>(recs, i, x);
        kmer_rc<W>(x, K, r);
        if (kmer_eq<W>(x, r)) c *= 2;
    }
    counts[i] = c;
}

// bucket_starts[b] = index of the first record of bucket b (records are sorted by bucket)
template<int W>
__global__ void bucket_starts_kernel(const uint64_t *__restrict__ recs, uint64_t u, uint32_t B, uint64_t *__restrict__ starts) {
    uint64_t i = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= u) return;
    uint64_t r[W];
    load_rec<W>(recs, i, r);
    uint32_t b = kmer_bucket<W>(r, B);
    if (i == 0) {
        for (uint32_t t = 0; t <= b; ++t) starts[t] = 0;
    } else {
        uint64_t q[W];
        load_rec<W>(recs, i - 1, q);
        uint32_t pb = kmer_bucket<W>(q, B);
        for (uint32_t t = pb + 1; t <= b; ++t) starts[t] = i;
    }
    if (i == u - 1)
        for (uint32_t t = b + 1; t <= B; ++t) starts[t] = u;
}

template<int W>
__global__ void last_is_marker_kernel(const uint64_t *__restrict__ recs, uint64_t u, uint32_t *__restrict__ flag) {
    uint64_t r[W];
    load_rec<W>(recs, u - 1, r);
    bool m = true;
#pragma unroll
    for (int j = 0; j < W; ++j) m &= (r[j] == ~0ULL);
    *flag = m ? 1u : 0u;
}

template<int W>
__global__ void double_palindromes_kernel(const uint64_t *__restrict__ recs, uint64_t u, int K, uint32_t *__restrict__ counts) {
    uint64_t i = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= u) return;
    uint64_t x[W], r[W];
    load_rec<W>(recs, i, x);
    kmer_rc<W>(x, K, r);
    if (kmer_eq<W>(x, r)) counts[i] *= 2;
}

// The same table from the group offsets of the grouped path: group g = bucket << p | prefix, so bucket b starts where its first
// group starts (no pass over the records, no hash per record)
__global__ void bucket_starts_from_groups_kernel(const uint32_t *__restrict__ group_off, int p, uint32_t B, uint32_t first_bucket, uint32_t n_owned,
                                                 uint64_t u, uint64_t *__restrict__ starts) {
    uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b > B) return;
    uint64_t v = u;                                             // buckets behind the owned range, and the end of the table
    if (b < first_bucket) v = 0;                                // buckets ahead of the owned range are empty
    else if (b < first_bucket + n_owned) v = group_off[(uint64_t) (b - first_bucket) << p];
    starts[b] = v > u ? u : v;                                  // a dropped marker record was the last record of the last bucket
}

template<int W>
static void finish_tables(sb200_ctx *ctx, sb200_kmers *s, uint32_t B, const uint32_t *group_off = nullptr, int p = 0, uint32_t first_bucket = 0,
                          uint32_t n_owned = 0) {
    s->bucket_starts.alloc(ctx, (uint64_t) B + 1);
    if (group_off) LAUNCH(ctx, bucket_starts_from_groups_kernel, div_up((uint64_t) B + 1, 256), 256, 0, group_off, p, B, first_bucket, n_owned, s->size,
                          s->bucket_starts.p);
    else LAUNCH(ctx, bucket_starts_kernel<W>, div_up(s->size, 256), 256, 0, s->data.p, s->size, B, s->bucket_starts.p);
    s->bucket_starts_host.resize((size_t) B + 1);
    ctx->fetch(s->bucket_starts_host.data(), s->bucket_starts.p, ((size_t) B + 1) * 8);
    CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
}

template<int W>
static sb200_kmers *finish_set_lsd(sb200_ctx *ctx, DevBuf<uint64_t> &inst, uint64_t n, int K, uint32_t B, bool want_counts,
                                   bool double_palindromes, bool drop_marker);

template<int W, int MODE, typename IdxT, bool SEG>
static void launch_group_hash_inst(sb200_ctx *ctx, uint32_t n_groups, const uint64_t *grouped, const ChunkRange *ranges, uint32_t *group_unique,
                                   uint32_t *ctrl, uint64_t *out, uint32_t *out_cnt, int shift2, uint64_t lw_keep, int pshift, const uint8_t *pay,
                                   const GroupParts &parts) {
    const size_t smem = group_hash_smem<IdxT>();
    auto group_hash_kernel_ = group_hash_kernel<W, MODE, IdxT, SEG>;
    CUDA_CHECK(cudaFuncSetAttribute(group_hash_kernel_, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
    LAUNCH(ctx, group_hash_kernel_, n_groups, HashCfg::THREADS, smem, grouped, ranges, group_unique, ctrl, out, out_cnt, shift2, lw_keep, pshift, pay, parts);
}

// parts == nullptr: every group is a contiguous range of `grouped`; otherwise the groups lie in parts->n_src runs (sharded path)
template<int W, typename IdxT>
static void launch_group_hash(sb200_ctx *ctx, int mode, uint32_t n_groups, const uint64_t *grouped, const ChunkRange *ranges, uint32_t *group_unique,
                              uint32_t *ctrl, uint64_t *out, uint32_t *out_cnt, int shift2, uint64_t lw_keep, int pshift, const uint8_t *pay,
                              const GroupParts *parts = nullptr) {
    const GroupParts none{nullptr, nullptr, 1u, n_groups};
#define SB200_GH(MODE)                                                                                                                              \
    do {                                                                                                                                           \
        if (parts) launch_group_hash_inst<W, MODE, IdxT, true>(ctx, n_groups, grouped, ranges, group_unique, ctrl, out, out_cnt, shift2, lw_keep, pshift, pay, *parts); \
        else launch_group_hash_inst<W, MODE, IdxT, false>(ctx, n_groups, grouped, ranges, group_unique, ctrl, out, out_cnt, shift2, lw_keep, pshift, pay, none);    \
    } while (0)
    if (mode == 1) SB200_GH(1);
    else if (mode == 2) SB200_GH(2);
    else SB200_GH(0);
#undef SB200_GH
}

// Deduplication, order, counts and tables of instances that are already GROUPED by the composite key (bucket - first_bucket) << p |
// top p value bits: `a` holds the n grouped records, starts[g] the first record of group g (n_groups + 1 entries), `b` is a free buffer
// of the same size that receives the unique records.  Both buffers are consumed.
//   one CTA per group deduplicates by hashing, orders the distinct records and packs them (+ count / + OR-ed mask bits) at the front of
//   the group's own range of `b` (grouphash.cuh); a scan of the per-group unique counts + seg_compact_kernel closes the gaps.
// pshift >= 0 (derived sets only, never together with counts): records carry a mask-bit payload in their padding (derive kernels);
// pay != nullptr: the payload lies in a byte array beside the records instead (k-mers that fill their last word).  The set then comes
// with masks_file.
template<int W>
static sb200_kmers *finish_grouped(sb200_ctx *ctx, DevBuf<uint64_t> &a, DevBuf<uint64_t> &b, uint64_t n, const uint32_t *starts, uint32_t n_groups, int p,
                                   int K, uint32_t B, bool want_counts, bool double_palindromes, bool drop_marker, int pshift, const uint8_t *pay,
                                   uint32_t first_bucket, uint32_t n_owned) {
    using Cfg = SegCfg<W>;
    const uint64_t lw_keep = last_word_mask(K);
    const bool masks_mode = (pshift >= 0 || pay != nullptr) && !want_counts;
    const int top = (W == 1) ? 2 * K : 64;
    const int dshift = top - p - 8;
    uint64_t *grouped = a.p, *other = b.p;
    DevBuf<ChunkRange> ranges(ctx, n_groups);
    DevBuf<uint32_t> ctrl(ctx, 4);   // [0] fail flag
    ctrl.zero();
    LAUNCH(ctx, group_ranges_kernel, div_up(n_groups, 256), 256, 0, starts, n_groups, ranges.p);
    ctx->trace_point("  group bounds");

    DevBuf<uint32_t> group_unique(ctx, (uint64_t) n_groups + 1);
    DevBuf<uint32_t> cnt_full;   // multiplicity (or OR-ed mask bits) of the unique record at the same position of `other`
    if (want_counts || masks_mode) cnt_full.alloc(ctx, n);
    if (ctx->group_chunk && !pay) {   // A/B and cross-check: the sorting kernel
        size_t smem = seg_chunk_smem<W>();
        if (want_counts) {
            auto group_chunk_kernel_ = group_chunk_kernel<W, 1>;
            CUDA_CHECK(cudaFuncSetAttribute(group_chunk_kernel_, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
            LAUNCH(ctx, group_chunk_kernel_, n_groups, Cfg::THREADS, smem, grouped, ranges.p, group_unique.p, ctrl.p, other, cnt_full.p, dshift, lw_keep, 0);
        } else if (masks_mode) {
            auto group_chunk_kernel_ = group_chunk_kernel<W, 2>;
            CUDA_CHECK(cudaFuncSetAttribute(group_chunk_kernel_, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
            LAUNCH(ctx, group_chunk_kernel_, n_groups, Cfg::THREADS, smem, grouped, ranges.p, group_unique.p, ctrl.p, other, cnt_full.p, dshift, lw_keep, pshift);
        } else {
            auto group_chunk_kernel_ = group_chunk_kernel<W, 0>;
            CUDA_CHECK(cudaFuncSetAttribute(group_chunk_kernel_, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
            LAUNCH(ctx, group_chunk_kernel_, n_groups, Cfg::THREADS, smem, grouped, ranges.p, group_unique.p, ctrl.p, other, (uint32_t *) nullptr, dshift,
                   lw_keep, 0);
        }
    } else {
        // hash deduplication of every group (grouphash.cuh); groups of 65535+ records go to the 64-bit-slot instance
        const int shift2 = dshift + 8;
        const int mode = want_counts ? 1 : masks_mode ? 2 : 0;
        uint32_t *cnt_out = (mode == 0) ? nullptr : cnt_full.p;
        const int psh = (mode == 2) ? pshift : 0;
        launch_group_hash<W, uint16_t>(ctx, mode, n_groups, grouped, ranges.p, group_unique.p, ctrl.p, other, cnt_out, shift2, lw_keep, psh, pay);
        uint32_t flags[2] = {0, 0};
        ctx->fetch(flags, ctrl.p, 8);
        if (flags[1] && !flags[0])
            launch_group_hash<W, uint32_t>(ctx, mode, n_groups, grouped, ranges.p, group_unique.p, ctrl.p, other, cnt_out, shift2, lw_keep, psh, pay);
    }
    // per-group unique counts -> offsets; the total sizes the result
    DevBuf<uint32_t> total32(ctx, 1);
    exclusive_scan<uint32_t>(ctx, group_unique.p, n_groups, total32.p);
    uint32_t u32 = 0, failed = 0;
    ctx->fetch(&u32, total32.p, 4);
    ctx->fetch(&failed, ctrl.p, 4);
    ctx->trace_point("  group kernel");
    if (failed) {   // a group beyond what the group kernel orders in shared memory (massively repeated k-mer): redo with the generic path
        b.release();
        cnt_full.release();
        if (pshift >= 0) LAUNCH(ctx, strip_payload_kernel<W>, div_up(n, 256), 256, 0, a.p, n, lw_keep);   // masks then come from lookups (ext.cu)
        return finish_set_lsd<W>(ctx, a, n, K, B, want_counts, double_palindromes, drop_marker);
    }
    uint64_t u = u32;

    sb200_kmers *s = new sb200_kmers();
    s->ctx = ctx; s->k = (unsigned) K; s->words = W; s->num_buckets = B; s->instances = n;
    s->data.alloc(ctx, u * W);   // right-sized: the instance-sized ping-pong buffers go back to the allocator
    if (want_counts) {
        s->counts.alloc(ctx, u);
        LAUNCH(ctx, (seg_compact_kernel<W, 1>), n_groups, 256, 0, other, cnt_full.p, ranges.p, group_unique.p, n_groups, u32, s->data.p, s->counts.p,
               (double_palindromes && (K % 2 == 0)) ? K : 0);
    } else if (masks_mode) {
        s->masks_file.alloc(ctx, u + 4);
        LAUNCH(ctx, (seg_compact_kernel<W, 2>), n_groups, 256, 0, other, cnt_full.p, ranges.p, group_unique.p, n_groups, u32, s->data.p,
               reinterpret_cast<uint32_t *>(s->masks_file.p));
    } else {
        LAUNCH(ctx, (seg_compact_kernel<W, 0>), n_groups, 256, 0, other, (const uint32_t *) nullptr, ranges.p, group_unique.p, n_groups, u32,
               s->data.p, (uint32_t *) nullptr);
    }
    if (drop_marker) {
        uint32_t flag = 0;
        DevBuf<uint32_t> fl(ctx, 1);
        LAUNCH(ctx, last_is_marker_kernel<W>, 1, 1, 0, s->data.p, u, fl.p);
        ctx->fetch(&flag, fl.p, 4);
        if (flag) --u;
        if (u == 0) {
            delete s;
            SB200_REQUIRE(false, "No kmers were extracted from reads. Check the read lengths and k-mer length settings");
        }
    }
    s->size = u;
    finish_tables<W>(ctx, s, B, group_unique.p, p, first_bucket, n_owned);   // group_unique holds the exclusive scan of the per-group unique counts
    a.release();
    b.release();
    ctx->trace_point("  shrink + tables");
    return s;
}

// composite-key width: p value bits below the bucket so that a group is ~ Cfg::TARGET records
template<int W>
static int choose_prefix_bits(uint64_t n, uint32_t n_owned, int K, int target = 0) {
    using Cfg = SegCfg<W>;
    if (target == 0) target = Cfg::TARGET;
    const int top = (W == 1) ? 2 * K : 64;
    int bbits = 0;
    while ((1ull << bbits) < n_owned) ++bbits;
    const int pmax = std::min(24 - std::min(bbits, 24), top - 8);
    int p = 0;
    while (p < pmax && (double) n / (double) ((uint64_t) n_owned << p) > (double) target) ++p;
    return p;
}

// ---- producer-fused grouping in two staged passes (staged_partition.cuh) -----------------------------------------------------------
struct StagedPlan {
    int p = 0, s = 0;
    uint32_t n_groups = 0, n_coarse = 0;
};

// false: the staged partition does not apply (the fine-group histogram would not fit the shared memory of a count CTA)
template<int W>
static bool staged_plan(sb200_ctx *ctx, uint64_t n_est, uint32_t B, int K, StagedPlan &pl) {
    if (B > SP_MAX_GROUPS) return false;
    // the hashing group kernel does not stage the group, so 3- and 4-word records take the same group size as 1- and 2-word ones
    pl.p = choose_prefix_bits<W>(n_est, B, K, ctx->group_chunk ? 0 : 7168);
    while (pl.p > 0 && ((uint64_t) B << pl.p) > SP_MAX_GROUPS) --pl.p;   // larger groups: the group kernel takes them in rounds
    pl.n_groups = (uint32_t) ((uint64_t) B << pl.p);
    // the clamp must not leave groups far beyond what the group kernel dedups in one round (GH_UMAX distinct records): an input of that size
    // (config 5 on one GPU: 633 M instances, 432 M distinct) takes the counting passes, whose group count is not bounded by shared memory
    if (n_est / std::max<uint64_t>(pl.n_groups, 1) > 10752 && !ctx->group_chunk) return false;
    const uint32_t max_coarse = std::min<uint32_t>(SP_MAX_BINS, SpCfg<W>::CAP / 8);   // >= 8 records per run of a full tile
    pl.s = 0;
    while ((((pl.n_groups - 1) >> pl.s) + 1) > max_coarse) ++pl.s;
    if ((1u << pl.s) > (uint32_t) SP_MAX_BINS) return false;
    pl.n_coarse = ((pl.n_groups - 1) >> pl.s) + 1;
    return true;
}

// Tiles of pass 2 from a segment list (begin / count / first fine group of every run of records that shares a coarse bin)
struct StagedTiles {
    DevBuf<SpTile> tiles;
    DevBuf<uint32_t> tile_off;   // n_seg + 1: exclusive scan of the tiles per segment; the last entry is the number of tiles
    uint32_t n_seg = 0;
};

template<int W>
static void staged_tiles(sb200_ctx *ctx, const uint32_t *seg_begin, const uint32_t *seg_cnt, const uint32_t *seg_gbase, uint32_t n_seg, uint64_t n,
                         StagedTiles &t) {
    constexpr uint32_t CAP2 = SpCfg<W>::CAP2;
    t.n_seg = n_seg;
    t.tile_off.alloc(ctx, (uint64_t) n_seg + 1);
    t.tiles.alloc(ctx, n / CAP2 + n_seg + 1);   // ceil(c_i / CAP2) summed over the segments is at most n / CAP2 + n_seg
    LAUNCH(ctx, sp_seg_tiles_kernel, div_up((uint64_t) n_seg + 1, 256), 256, 0, seg_cnt, n_seg, CAP2, t.tile_off.p);
    exclusive_scan<uint32_t>(ctx, t.tile_off.p, (uint64_t) n_seg + 1, nullptr);
    LAUNCH(ctx, sp_tile_fill_kernel, div_up(n / CAP2 + n_seg + 1, 256), 256, 0, seg_begin, seg_cnt, seg_gbase, t.tile_off.p, n_seg, CAP2, t.tiles.p);
}

// After the count pass: group starts, cursors of both passes, tiles of pass 2.  hist (n_groups + 1 counts) becomes the fine starts.
struct StagedTables {
    DevBuf<uint32_t> cur1, cur2, seg;
    StagedTiles tiles;
    uint64_t n = 0;
};

__global__ void sp_coarse_cursors_kernel(const uint32_t *__restrict__ fine_start, int s, uint32_t n_coarse, uint32_t *__restrict__ cur1) {
    const uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c < n_coarse) cur1[c] = fine_start[(uint64_t) c << s];
}

template<int W>
static void staged_tables(sb200_ctx *ctx, const StagedPlan &pl, DevBuf<uint32_t> &hist, StagedTables &t) {
    DevBuf<uint32_t> total_dev(ctx, 1);
    exclusive_scan<uint32_t>(ctx, hist.p, (uint64_t) pl.n_groups + 1, total_dev.p);   // hist[g] = first record of group g, hist[n_groups] = n
    t.cur1.alloc(ctx, pl.n_coarse);
    t.cur2.alloc(ctx, (uint64_t) pl.n_groups + 1);
    t.seg.alloc(ctx, 3 * (uint64_t) pl.n_coarse);
    LAUNCH(ctx, sp_coarse_cursors_kernel, div_up(pl.n_coarse, 256), 256, 0, hist.p, pl.s, pl.n_coarse, t.cur1.p);
    CUDA_CHECK(cudaMemcpyAsync(t.cur2.p, hist.p, ((size_t) pl.n_groups + 1) * 4, cudaMemcpyDeviceToDevice, ctx->stream));
    LAUNCH(ctx, sp_coarse_segments_kernel, div_up(pl.n_coarse, 256), 256, 0, hist.p, pl.s, pl.n_coarse, pl.n_groups, t.seg.p, t.seg.p + pl.n_coarse,
           t.seg.p + 2 * (uint64_t) pl.n_coarse);
    uint32_t n32 = 0;
    ctx->fetch(&n32, total_dev.p, 4);
    t.n = n32;
    staged_tiles<W>(ctx, t.seg.p, t.seg.p + pl.n_coarse, t.seg.p + 2 * (uint64_t) pl.n_coarse, pl.n_coarse, t.n, t.tiles);
}

// pass 2: mid (runs of records per coarse bin, listed as tiles) -> out (grouped by fine group)
template<int W>
static void staged_pass2(sb200_ctx *ctx, int s, const StagedTiles &tl, uint32_t *cur2, const GroupSel &gs, uint64_t lw_keep, const uint64_t *mid,
                         const uint8_t *mid_pay, uint64_t *out, uint8_t *out_pay) {
    constexpr int CAP2 = SpCfg<W>::CAP2;
    const size_t smem = 2 * (size_t) CAP2 * W * 8 + (size_t) CAP2 * 4 + 2 * (size_t) CAP2 + (size_t) (3 * SP_MAX_BINS + 1) * 4 + 64;
    auto sp_scatter_fine_kernel_ = sp_scatter_fine_kernel<W>;
    CUDA_CHECK(cudaFuncSetAttribute(sp_scatter_fine_kernel_, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
    LAUNCH(ctx, sp_scatter_fine_kernel_, (unsigned) ctx->num_sms * 2, SP_THREADS, smem, mid, mid_pay, tl.tiles.p, tl.tile_off.p + tl.n_seg, gs, lw_keep, s,
           cur2, out, out_pay);
}

// Sort + unique + counts + bucket table of UNGROUPED instances (records that came through an exchange, sb200_count_records): one or two
// stable counting passes group them by the composite key (radix_sort.cuh), then finish_grouped.  `inst` (n x W) is consumed.
template<int W>
static sb200_kmers *finish_set(sb200_ctx *ctx, DevBuf<uint64_t> &inst, uint64_t n, int K, uint32_t B, bool want_counts,
                               bool double_palindromes, bool drop_marker, int pshift = -1, uint32_t first_bucket = 0, uint32_t n_owned = 0) {
    // [first_bucket, first_bucket + n_owned): the buckets the records can lie in (sharded path: this GPU's share; 0 = all B).  The group
    // key counts buckets from first_bucket, so that its 16 bits (two counting passes) go to the value prefix instead of empty buckets.
    if (n_owned == 0) { first_bucket = 0; n_owned = B; }
    SB200_REQUIRE(n > 0, "No kmers were extracted from reads. Check the read lengths and k-mer length settings");
    const uint64_t lw_keep = last_word_mask(K);
    if (2 * K < 24) {
        if (pshift >= 0) LAUNCH(ctx, strip_payload_kernel<W>, div_up(n, 256), 256, 0, inst.p, n, lw_keep);
        return finish_set_lsd<W>(ctx, inst, n, K, B, want_counts, double_palindromes, drop_marker);
    }
    const int top = (W == 1) ? 2 * K : 64;
    int bbits = 0;
    while ((1ull << bbits) < n_owned) ++bbits;
    const int p = choose_prefix_bits<W>(n, n_owned, K);
    const uint32_t n_groups = (uint32_t) ((uint64_t) n_owned << p);
    DigitSel gk{-3, 0, B, drop_marker ? 1 : 0, p, top, lw_keep, first_bucket};

    DevBuf<uint64_t> scratch(ctx, n * W);
    ctx->trace_point("  instances ready");
    uint64_t *grouped = radix_sort_passes<W>(ctx, inst.p, scratch.p, n, composite_passes(bbits + p, p, top, B, drop_marker, lw_keep, first_bucket));
    ctx->trace_point("  group passes");
    if (grouped != inst.p) std::swap(inst, scratch);   // `inst` holds the grouped records, `scratch` is the free ping-pong buffer
    DevBuf<uint32_t> starts(ctx, (uint64_t) n_groups + 1);
    LAUNCH(ctx, group_bounds_kernel<W>, div_up((uint64_t) n_groups + 1, 128), 128, 0, inst.p, n, gk, n_groups, starts.p);
    return finish_grouped<W>(ctx, inst, scratch, n, starts.p, n_groups, p, K, B, want_counts, double_palindromes, drop_marker, pshift, nullptr,
                             first_bucket, n_owned);
}

// First version of the path: full LSD sort of every significant byte.  Kept for very short k-mers (2K < 24 bits), as
// the fallback of oversize segments, and as the reference point for the profile in profiles/.
template<int W>
static sb200_kmers *finish_set_lsd(sb200_ctx *ctx, DevBuf<uint64_t> &inst, uint64_t n, int K, uint32_t B, bool want_counts,
                                   bool double_palindromes, bool drop_marker) {
    SB200_REQUIRE(n > 0, "No kmers were extracted from reads. Check the read lengths and k-mer length settings");
    DevBuf<uint64_t> scratch(ctx, n * W);
    uint64_t *sorted = radix_sort_records<W>(ctx, inst.p, scratch.p, n, K, B, drop_marker);

    unsigned tiles = div_up(n, UQ_TILE);
    DevBuf<uint32_t> tile_cnt(ctx, tiles);
    DevBuf<uint32_t> total_dev(ctx, 2);
    LAUNCH(ctx, unique_count_kernel<W>, tiles, UQ_THREADS, 0, sorted, n, tile_cnt.p);
    exclusive_scan<uint32_t>(ctx, tile_cnt.p, tiles, total_dev.p);
    uint32_t u32 = 0;
    ctx->fetch(&u32, total_dev.p, 4);
    CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    uint64_t u = u32;

    sb200_kmers *s = new sb200_kmers();
    s->ctx = ctx; s->k = (unsigned) K; s->words = W; s->num_buckets = B; s->instances = n;
    s->data.alloc(ctx, u * W);
    DevBuf<uint32_t> head_pos(ctx, u + 1);
    LAUNCH(ctx, unique_write_kernel<W>, tiles, UQ_THREADS, 0, sorted, n, tile_cnt.p, s->data.p, head_pos.p);
    uint32_t n32 = (uint32_t) n;
    CUDA_CHECK(cudaMemcpyAsync(head_pos.p + u, &n32, 4, cudaMemcpyHostToDevice, ctx->stream));
    if (drop_marker) {
        LAUNCH(ctx, last_is_marker_kernel<W>, 1, 1, 0, s->data.p, u, total_dev.p + 1);
        uint32_t flag = 0;
        ctx->fetch(&flag, total_dev.p + 1, 4);
        CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
        if (flag) {
            --u;
            if (u == 0) {
                delete s;
                SB200_REQUIRE(false, "No kmers were extracted from reads. Check the read lengths and k-mer length settings");
            }
        }
    }
    s->size = u;
    if (want_counts) {
        s->counts.alloc(ctx, u);
        LAUNCH(ctx, counts_kernel<W>, div_up(u, 256), 256, 0, head_pos.p, s->data.p, u, K, double_palindromes ? 1 : 0, s->counts.p);
    }
    s->bucket_starts.alloc(ctx, (uint64_t) B + 1);
    LAUNCH(ctx, bucket_starts_kernel<W>, div_up(u, 256), 256, 0, s->data.p, u, B, s->bucket_starts.p);
    s->bucket_starts_host.resize((size_t) B + 1);
    ctx->fetch(s->bucket_starts_host.data(), s->bucket_starts.p, ((size_t) B + 1) * 8);
    CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    inst.release();
    return s;
}

// Materialise every instance in read order, then group with counting passes.
template<int W>
static sb200_kmers *count_reads_legacy_w(sb200_ctx *ctx, const sb200_reads *rd, int K, int canonical_only, int add_rc, uint32_t B) {
    int mode = canonical_only ? (add_rc ? MODE_CANON_RC : MODE_CANON_FWD) : (add_rc ? MODE_ALL_RC : MODE_ALL_FWD);
    uint32_t mult = (mode == MODE_ALL_RC) ? 2 : 1;
    DevBuf<uint64_t> off(ctx, rd->n_reads + 1);
    DevBuf<uint64_t> total_dev(ctx, 1);
    LAUNCH(ctx, window_count_kernel, div_up(rd->n_reads, 256), 256, 0, rd->len.p, rd->n_reads, (uint32_t) K, mult, off.p);
    exclusive_scan<uint64_t>(ctx, off.p, rd->n_reads, total_dev.p);
    uint64_t n = 0;
    ctx->fetch(&n, total_dev.p, 8);
    CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    SB200_REQUIRE(n > 0, "No kmers were extracted from reads. Check the read lengths and k-mer length settings");
    DevBuf<uint64_t> inst(ctx, n * W);
    unsigned warps_per_block = 8;
    unsigned grid = (unsigned) std::min<uint64_t>((rd->n_reads + warps_per_block - 1) / warps_per_block, (uint64_t) ctx->num_sms * 32);
    LAUNCH(ctx, extract_reads_kernel<W>, grid, 256, 0, rd->words.p, rd->word_off.p, rd->len.p, rd->n_reads, K, mode, off.p, inst.p);
    return finish_set<W>(ctx, inst, n, K, B, true, mode == MODE_CANON_RC, mode == MODE_CANON_FWD);
}

// SB200_ATOMIC_PARTITION=1: reads -> grouped instances in one trip through HBM (partition.cuh), then deduplication / order / counts per
// group.  Measured on B200 (profiles/r2a_*): 291 M scattered 16-byte stores behind 291 M L2 atomics run at the L2's REQUEST rate
// (~80-110 G requests/s: 7.0 ms count + 12.8 ms write for S2 + S4 against 11.6 ms of coalesced counting passes), so the path that
// stages bin runs in shared memory stays the default; this one is kept as an independent implementation for the cross-checks.
template<int W>
static sb200_kmers *count_reads_w(sb200_ctx *ctx, const sb200_reads *rd, int K, int canonical_only, int add_rc, uint32_t B) {
    if (ctx->counting_passes || 2 * K < 24) return count_reads_legacy_w<W>(ctx, rd, K, canonical_only, add_rc, B);
    const int mode = canonical_only ? (add_rc ? PART_CANON : PART_MINIMAL_ONLY) : PART_FWD;
    const bool both = !canonical_only && add_rc;   // spades-kmercount: every k-mer of both strands
    // upper bound of the instance count (positions are 32-bit) and the estimate that sizes the groups
    const uint64_t bases = rd->n_bases ? rd->n_bases : rd->n_words * 32;
    const uint64_t shorter = rd->n_reads * (uint64_t) (K - 1);
    const uint64_t n_est = std::max<uint64_t>(bases > shorter ? bases - shorter : 1, 1) * (both ? 2 : 1);
    SB200_REQUIRE(bases * (both ? 2 : 1) < (1ull << 32), "more than 2^32-1 k-mer instances on one GPU: shard the input");
    StagedPlan pl;
    if (!ctx->atomic_partition) {
        if (!staged_plan<W>(ctx, n_est, B, K, pl)) return count_reads_legacy_w<W>(ctx, rd, K, canonical_only, add_rc, B);
        const GroupSel gs{B, 0u, pl.p, (W == 1) ? 2 * K : 64};
        DevBuf<uint32_t> hist(ctx, (uint64_t) pl.n_groups + 1);
        hist.zero();
        auto sp_count_reads_kernel_ = sp_count_reads_kernel<W>;
        auto sp_scatter_reads_kernel_ = sp_scatter_reads_kernel<W>;
        const size_t smem_c = (size_t) pl.n_groups * 4, smem_s = sp_stage_bytes<W>(SpCfg<W>::CAP, false);
        CUDA_CHECK(cudaFuncSetAttribute(sp_count_reads_kernel_, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem_c));
        CUDA_CHECK(cudaFuncSetAttribute(sp_scatter_reads_kernel_, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem_s));
        const unsigned grid_c = (unsigned) ctx->num_sms, grid_s = (unsigned) ctx->num_sms * 2;
        LAUNCH(ctx, sp_count_reads_kernel_, grid_c, SPC_THREADS, smem_c, rd->words.p, rd->word_off.p, rd->len.p, rd->n_reads, K, mode, gs, 0, pl.n_groups, hist.p);
        if (both) LAUNCH(ctx, sp_count_reads_kernel_, grid_c, SPC_THREADS, smem_c, rd->words.p, rd->word_off.p, rd->len.p, rd->n_reads, K, (int) PART_REV, gs, 0, pl.n_groups, hist.p);
        StagedTables t;
        staged_tables<W>(ctx, pl, hist, t);
        const uint64_t n = t.n;
        SB200_REQUIRE(n > 0, "No kmers were extracted from reads. Check the read lengths and k-mer length settings");
        DevBuf<uint64_t> inst(ctx, n * W), mid(ctx, n * W);
        // a tile takes `rounds` chunks per warp: as many as fit its capacity when every chunk is full
        const uint32_t max_nwin = rd->max_len >= (uint32_t) K ? rd->max_len - (uint32_t) K + 1 : 128u;
        const uint32_t chunk_cap = 32u * std::min<uint32_t>(PART_RUN, (max_nwin + 31u) / 32u);
        const int rounds = std::max<int>(1, SpCfg<W>::CAP / (int) (SP_WARPS * chunk_cap));
        LAUNCH(ctx, sp_scatter_reads_kernel_, grid_s, SP_THREADS, smem_s, rd->words.p, rd->word_off.p, rd->len.p, rd->n_reads, K, mode, gs, pl.s, pl.n_coarse,
               rounds, t.cur1.p, mid.p, SpPeers());
        if (both) LAUNCH(ctx, sp_scatter_reads_kernel_, grid_s, SP_THREADS, smem_s, rd->words.p, rd->word_off.p, rd->len.p, rd->n_reads, K, (int) PART_REV, gs, pl.s,
                         pl.n_coarse, rounds, t.cur1.p, mid.p, SpPeers());
        staged_pass2<W>(ctx, pl.s, t.tiles, t.cur2.p, gs, ~0ULL, mid.p, nullptr, inst.p, nullptr);
        ctx->trace_point("  instances partitioned (staged)");
        return finish_grouped<W>(ctx, inst, mid, n, hist.p, pl.n_groups, pl.p, K, B, true, mode == PART_CANON, false, -1, nullptr, 0, B);
    }
    const int p = choose_prefix_bits<W>(n_est, B, K);
    const uint32_t n_groups = (uint32_t) ((uint64_t) B << p);
    const GroupSel gs{B, 0u, p, (W == 1) ? 2 * K : 64};
    DevBuf<uint32_t> hist(ctx, (uint64_t) n_groups + 1), cursor(ctx, (uint64_t) n_groups + 1), total_dev(ctx, 1);
    hist.zero();
    const unsigned grid = (unsigned) std::min<uint64_t>((rd->n_reads + 7) / 8, (uint64_t) ctx->num_sms * 32);
    auto count_pass = partition_reads_kernel<W, false>;
    auto write_pass = partition_reads_kernel<W, true>;
    LAUNCH(ctx, count_pass, grid, 256, 0, rd->words.p, rd->word_off.p, rd->len.p, rd->n_reads, K, mode, gs, hist.p, (uint64_t *) nullptr);
    if (both) LAUNCH(ctx, count_pass, grid, 256, 0, rd->words.p, rd->word_off.p, rd->len.p, rd->n_reads, K, (int) PART_REV, gs, hist.p, (uint64_t *) nullptr);
    exclusive_scan<uint32_t>(ctx, hist.p, (uint64_t) n_groups + 1, total_dev.p);   // hist[g] = first record of group g, hist[n_groups] = n
    CUDA_CHECK(cudaMemcpyAsync(cursor.p, hist.p, ((size_t) n_groups + 1) * 4, cudaMemcpyDeviceToDevice, ctx->stream));
    uint32_t n32 = 0;
    ctx->fetch(&n32, total_dev.p, 4);
    const uint64_t n = n32;
    SB200_REQUIRE(n > 0, "No kmers were extracted from reads. Check the read lengths and k-mer length settings");
    DevBuf<uint64_t> inst(ctx, n * W), other(ctx, n * W);
    LAUNCH(ctx, write_pass, grid, 256, 0, rd->words.p, rd->word_off.p, rd->len.p, rd->n_reads, K, mode, gs, cursor.p, inst.p);
    if (both) LAUNCH(ctx, write_pass, grid, 256, 0, rd->words.p, rd->word_off.p, rd->len.p, rd->n_reads, K, (int) PART_REV, gs, cursor.p, inst.p);
    ctx->trace_point("  instances partitioned");
    return finish_grouped<W>(ctx, inst, other, n, hist.p, n_groups, p, K, B, true, mode == PART_CANON, false, -1, nullptr, 0, B);
}

sb200_kmers *count_reads(sb200_ctx *ctx, const sb200_reads *rd, unsigned K, int canonical_only, int add_rc, unsigned B) {
    SB200_REQUIRE(K >= 1 && K <= 128, "K out of range [1,128]");
    SB200_REQUIRE(B >= 1 && B <= 65536, "num_buckets out of range [1,65536]");
    switch ((K + 31) / 32) {
        case 1: return count_reads_w<1>(ctx, rd, (int) K, canonical_only, add_rc, B);
        case 2: return count_reads_w<2>(ctx, rd, (int) K, canonical_only, add_rc, B);
        case 3: return count_reads_w<3>(ctx, rd, (int) K, canonical_only, add_rc, B);
        default: return count_reads_w<4>(ctx, rd, (int) K, canonical_only, add_rc, B);
    }
}

template<int WS, int W>
static sb200_kmers *derive_w(sb200_ctx *ctx, const sb200_kmers *kp, uint32_t B) {
    int k = (int) kp->k - 1;
    uint64_t n = kp->size * 2;
    // three padding bits above the k-mer in its last word carry the mask bit when they exist (every odd k except k = 31 mod 32);
    // otherwise the fused path keeps the bit in a byte beside the record
    const int used = 2 * (k - 32 * (W - 1));
    const int pshift = (used + 3 <= 64 && !ctx->no_mask_payload) ? used : -1;
    StagedPlan pl;
    const bool staged = !ctx->counting_passes && !ctx->atomic_partition && 2 * k >= 24 && staged_plan<W>(ctx, n, B, k, pl);
    if (!staged && (!ctx->atomic_partition || 2 * k < 24)) {
        DevBuf<uint64_t> inst(ctx, n * W);
        auto derive_kernel_ = derive_kernel<WS, W>;
        LAUNCH(ctx, derive_kernel_, div_up(kp->size, 256), 256, 0, kp->data.p, kp->size, k, pshift, inst.p);
        sb200_kmers *s = finish_set<W>(ctx, inst, n, k, B, false, false, false, pshift);
        s->instances = 0;
        return s;
    }
    SB200_REQUIRE(n < (1ull << 32), "more than 2^32-1 k-mer instances on one GPU: shard the input");
    if (staged) {
        const GroupSel gs{B, 0u, pl.p, (W == 1) ? 2 * k : 64};
        DevBuf<uint32_t> hist(ctx, (uint64_t) pl.n_groups + 1);
        hist.zero();
        const bool side_pay = pshift < 0 && !ctx->no_mask_payload;
        auto sp_count_derive_kernel_ = sp_count_derive_kernel<WS, W>;
        auto sp_scatter_derive_kernel_ = sp_scatter_derive_kernel<WS, W>;
        const size_t smem_c = (size_t) pl.n_groups * 4, smem_s = sp_stage_bytes<W>(SpCfg<W>::CAP, side_pay);
        CUDA_CHECK(cudaFuncSetAttribute(sp_count_derive_kernel_, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem_c));
        CUDA_CHECK(cudaFuncSetAttribute(sp_scatter_derive_kernel_, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem_s));
        LAUNCH(ctx, sp_count_derive_kernel_, (unsigned) ctx->num_sms, SPC_THREADS, smem_c, kp->data.p, kp->size, k, gs, 0, pl.n_groups, hist.p);
        StagedTables t;
        staged_tables<W>(ctx, pl, hist, t);
        DevBuf<uint64_t> inst(ctx, n * W), mid(ctx, n * W);
        DevBuf<uint8_t> pay, mid_pay;
        if (side_pay) { pay.alloc(ctx, n); mid_pay.alloc(ctx, n); }
        LAUNCH(ctx, sp_scatter_derive_kernel_, (unsigned) ctx->num_sms * 2, SP_THREADS, smem_s, kp->data.p, kp->size, k, pshift, gs, pl.s, pl.n_coarse, t.cur1.p,
               mid.p, mid_pay.p, SpPeers());
        staged_pass2<W>(ctx, pl.s, t.tiles, t.cur2.p, gs, last_word_mask(k), mid.p, mid_pay.p, inst.p, pay.p);
        ctx->trace_point("  candidates partitioned (staged)");
        sb200_kmers *s = finish_grouped<W>(ctx, inst, mid, n, hist.p, pl.n_groups, pl.p, k, B, false, false, false, pshift, pay.p, 0, B);
        s->instances = 0;
        return s;
    }
    const int p = choose_prefix_bits<W>(n, B, k);
    const uint32_t n_groups = (uint32_t) ((uint64_t) B << p);
    const GroupSel gs{B, 0u, p, (W == 1) ? 2 * k : 64};
    DevBuf<uint32_t> hist(ctx, (uint64_t) n_groups + 1), cursor(ctx, (uint64_t) n_groups + 1);
    hist.zero();
    auto count_pass = partition_derive_kernel<WS, W, false>;
    auto write_pass = partition_derive_kernel<WS, W, true>;
    const unsigned grid = div_up(kp->size, 256);
    LAUNCH(ctx, count_pass, grid, 256, 0, kp->data.p, kp->size, k, pshift, gs, hist.p, (uint64_t *) nullptr, (uint8_t *) nullptr);
    exclusive_scan<uint32_t>(ctx, hist.p, (uint64_t) n_groups + 1, nullptr);
    CUDA_CHECK(cudaMemcpyAsync(cursor.p, hist.p, ((size_t) n_groups + 1) * 4, cudaMemcpyDeviceToDevice, ctx->stream));
    DevBuf<uint64_t> inst(ctx, n * W), other(ctx, n * W);
    DevBuf<uint8_t> pay;
    if (pshift < 0 && !ctx->no_mask_payload) pay.alloc(ctx, n);
    LAUNCH(ctx, write_pass, grid, 256, 0, kp->data.p, kp->size, k, pshift, gs, cursor.p, inst.p, pay.p);
    ctx->trace_point("  candidates partitioned");
    sb200_kmers *s = finish_grouped<W>(ctx, inst, other, n, hist.p, n_groups, p, k, B, false, false, false, pshift, pay.p, 0, B);
    s->instances = 0;
    return s;
}

sb200_kmers *derive_kmers(sb200_ctx *ctx, const sb200_kmers *kp, unsigned B) {
    SB200_REQUIRE(kp->k >= 2, "source k-mers too short");
    SB200_REQUIRE(B >= 1 && B <= 65536, "num_buckets out of range [1,65536]");
    int WS = (int) kp->words, W = (int) ((kp->k - 1 + 31) / 32);
    if (WS == 1) return derive_w<1, 1>(ctx, kp, B);
    if (WS == 2 && W == 1) return derive_w<2, 1>(ctx, kp, B);
    if (WS == 2) return derive_w<2, 2>(ctx, kp, B);
    if (WS == 3 && W == 2) return derive_w<3, 2>(ctx, kp, B);
    if (WS == 3) return derive_w<3, 3>(ctx, kp, B);
    if (WS == 4 && W == 3) return derive_w<4, 3>(ctx, kp, B);
    return derive_w<4, 4>(ctx, kp, B);
}

// ---- sharded counting: the same two steps with the hash shuffle exposed ---------------------------------------------------
// Several GPUs share one k-mer space by giving GPU g the buckets [g*B/G, (g+1)*B/G) of the reference's own bucket
// function: (1) every GPU turns its reads (or its (k+1)-mers) into records and groups them by owner — one stable counting
// pass whose digit is the owner id; (2) the host exchanges the groups (NCCL all-to-all over NVLink, host/distributed.py);
// (3) every GPU sorts/deduplicates/counts what it received.  All copies of a k-mer meet at its owner, so multiplicities,
// per-bucket order and bucket boundaries are those of the single-GPU run, shard by shard.
template<int W>
static sb200_records *extract_records_w(sb200_ctx *ctx, const sb200_reads *rd, int K, int canonical_only, int add_rc) {
    int mode = canonical_only ? (add_rc ? MODE_CANON_RC : MODE_CANON_FWD) : (add_rc ? MODE_ALL_RC : MODE_ALL_FWD);
    uint32_t mult = (mode == MODE_ALL_RC) ? 2 : 1;
    DevBuf<uint64_t> off(ctx, rd->n_reads + 1);
    DevBuf<uint64_t> total_dev(ctx, 1);
    LAUNCH(ctx, window_count_kernel, div_up(rd->n_reads ? rd->n_reads : 1, 256), 256, 0, rd->len.p, rd->n_reads, (uint32_t) K, mult, off.p);
    exclusive_scan<uint64_t>(ctx, off.p, rd->n_reads, total_dev.p);
    uint64_t n = 0;
    ctx->fetch(&n, total_dev.p, 8);
    CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    sb200_records *r = new sb200_records();
    r->ctx = ctx; r->k = (unsigned) K; r->words = W; r->n = n;
    r->double_palindromes = mode == MODE_CANON_RC; r->marker = mode == MODE_CANON_FWD;
    r->data.alloc(ctx, n * W);
    if (n) {
        unsigned grid = (unsigned) std::min<uint64_t>((rd->n_reads + 7) / 8, (uint64_t) ctx->num_sms * 32);
        LAUNCH(ctx, extract_reads_kernel<W>, grid, 256, 0, rd->words.p, rd->word_off.p, rd->len.p, rd->n_reads, K, mode, off.p, r->data.p);
    }
    CUDA_CHECK(cudaStreamSynchronize(ctx->stream));   // blocking like every entry point: the caller hands the records to a collective on another stream
    return r;
}

// ---- extraction straight into the owner groups (canonical modes: one record per window) ----------------------------------
// extract + partition_records read and wrote every instance twice more (histogram pass, scatter pass) just to group them by owner.
// Here the extraction itself does it: a first pass counts, per read and owner, the windows that belong to that owner (the same warp
// per read, nothing written but G counters per read); after one scan of the [owner][read] matrix a second pass recomputes the windows
// and stores every record at its owner's cursor, ranked inside the warp with one ballot per owner.  Instances cross HBM once.
constexpr int MAX_FUSED_OWNERS = 32;

template<int W>
__device__ __forceinline__ uint32_t window_owner(const uint64_t *seq, uint32_t nw, uint32_t p, int K, int mode, uint32_t B, uint32_t G, uint64_t *y) {
    uint64_t x[W];
    kmer_window<W>(seq, nw, p, K, x);
    const bool minimal = kmer_canonical<W>(x, K, y);
    uint32_t b;
    if (mode == MODE_CANON_FWD && !minimal) {
#pragma unroll
        for (int j = 0; j < W; ++j) y[j] = ~0ULL;   // marker: belongs to the last bucket (rs_bucket)
        b = B - 1;
    } else {
        b = kmer_bucket<W>(y, B);
    }
    return (uint32_t) (((uint64_t) b * G) / B);
}

template<int W, bool WRITE>
__global__ void __launch_bounds__(256) extract_owner_kernel(const uint64_t *__restrict__ words, const uint64_t *__restrict__ word_off,
                                                           const uint32_t *__restrict__ len, uint64_t n_reads, int K, int mode, uint32_t B, uint32_t G,
                                                           uint32_t *__restrict__ cnt /* [G][n_reads]: counts (pass 1) / exclusive scan (pass 2) */,
                                                           uint64_t *__restrict__ out) {
    const int lane = threadIdx.x & 31;
    const uint32_t lt = (1u << lane) - 1u;
    const uint64_t warps_total = (uint64_t) gridDim.x * (blockDim.x >> 5);
    for (uint64_t r = (uint64_t) blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); r < n_reads; r += warps_total) {
        const uint32_t l = len[r];
        uint32_t acc = 0;   // lane g: windows of owner g so far (pass 1) / owner g's cursor (pass 2)
        if (WRITE && (uint32_t) lane < G) acc = cnt[(uint64_t) lane * n_reads + r];
        if (l >= (uint32_t) K) {
            const uint32_t nwin = l - K + 1;
            const uint64_t *seq = words + word_off[r];
            const uint32_t nw = (l + 31) >> 5;
            for (uint32_t p0 = 0; p0 < nwin; p0 += 32) {
                const uint32_t p = p0 + lane;
                const bool ok = p < nwin;
                uint64_t y[W];
                uint32_t o = 0xFFFFFFFFu;
                if (ok) o = window_owner<W>(seq, nw, p, K, mode, B, G, y);
                uint32_t pos = 0;
                for (uint32_t g = 0; g < G; ++g) {
                    const uint32_t m = __ballot_sync(0xffffffffu, o == g);
                    if (WRITE) {
                        const uint32_t base = __shfl_sync(0xffffffffu, acc, (int) g);
                        if (o == g) pos = base + (uint32_t) __popc(m & lt);
                    }
                    if ((uint32_t) lane == g) acc += (uint32_t) __popc(m);
                }
                if (WRITE && ok) store_rec<W>(out, pos, y);
            }
        }
        if (!WRITE && (uint32_t) lane < G) cnt[(uint64_t) lane * n_reads + r] = acc;
    }
}

__global__ void owner_totals_kernel(const uint32_t *__restrict__ off, uint64_t per_owner, uint32_t G, const uint32_t *__restrict__ total,
                                    uint32_t *__restrict__ counts) {
    const uint32_t g = threadIdx.x;
    if (g >= G) return;
    const uint32_t end = (g + 1 < G) ? off[(uint64_t) (g + 1) * per_owner] : *total;
    counts[g] = end - off[(uint64_t) g * per_owner];
}

template<int W>
static sb200_records *extract_partitioned_w(sb200_ctx *ctx, const sb200_reads *rd, int K, int mode, uint32_t B, uint32_t G, uint64_t *counts_out) {
    const uint64_t nr = rd->n_reads;
    sb200_records *r = new sb200_records();
    r->ctx = ctx; r->k = (unsigned) K; r->words = W; r->n = 0;
    r->double_palindromes = mode == MODE_CANON_RC; r->marker = mode == MODE_CANON_FWD;
    for (uint32_t g = 0; g < G; ++g) counts_out[g] = 0;
    if (nr == 0) return r;
    DevBuf<uint32_t> cnt(ctx, (uint64_t) G * nr + 1);
    DevBuf<uint32_t> total_dev(ctx, 1), counts_dev(ctx, MAX_FUSED_OWNERS);
    const unsigned grid = (unsigned) std::min<uint64_t>((nr + 7) / 8, (uint64_t) ctx->num_sms * 32);
    auto count_pass = extract_owner_kernel<W, false>;
    auto write_pass = extract_owner_kernel<W, true>;
    LAUNCH(ctx, count_pass, grid, 256, 0, rd->words.p, rd->word_off.p, rd->len.p, nr, K, mode, B, G, cnt.p, (uint64_t *) nullptr);
    exclusive_scan<uint32_t>(ctx, cnt.p, (uint64_t) G * nr, total_dev.p);
    LAUNCH(ctx, owner_totals_kernel, 1, MAX_FUSED_OWNERS, 0, cnt.p, nr, G, total_dev.p, counts_dev.p);
    uint32_t h[MAX_FUSED_OWNERS + 1];
    ctx->fetch(h, counts_dev.p, (size_t) G * 4);
    uint64_t n = 0;
    for (uint32_t g = 0; g < G; ++g) { counts_out[g] = h[g]; n += h[g]; }
    SB200_REQUIRE(n < (1ull << 32), "more than 2^32-1 k-mer instances on one GPU: shard the input");
    r->n = n;
    r->data.alloc(ctx, n * W);
    if (n) LAUNCH(ctx, write_pass, grid, 256, 0, rd->words.p, rd->word_off.p, rd->len.p, nr, K, mode, B, G, cnt.p, r->data.p);
    CUDA_CHECK(cudaStreamSynchronize(ctx->stream));   // `cnt` goes back to the allocator and the caller exchanges the records on another stream
    return r;
}

// Records of the reads grouped by owner, counts_out[g] = records of owner g.  Canonical modes take the fused kernels above; the
// modes that keep both strands extract first and partition afterwards.
sb200_records *extract_records(sb200_ctx *ctx, const sb200_reads *rd, unsigned K, int canonical_only, int add_rc);
void partition_records(sb200_ctx *ctx, sb200_records *r, unsigned B, unsigned n_parts, uint64_t *counts_out);
sb200_records *extract_records_partitioned(sb200_ctx *ctx, const sb200_reads *rd, unsigned K, int canonical_only, int add_rc, unsigned B, unsigned G,
                                           uint64_t *counts_out) {
    SB200_REQUIRE(K >= 1 && K <= 128, "K out of range [1,128]");
    SB200_REQUIRE(G >= 1 && G <= 256, "number of owners out of range [1,256]");
    SB200_REQUIRE(B >= 1 && B <= 65536 && B % G == 0, "num_buckets must be a multiple of the number of owners");
    if (!canonical_only || G > (unsigned) MAX_FUSED_OWNERS || ctx->no_fused_partition) {
        sb200_records *r = extract_records(ctx, rd, K, canonical_only, add_rc);
        partition_records(ctx, r, B, G, counts_out);
        return r;
    }
    const int mode = add_rc ? MODE_CANON_RC : MODE_CANON_FWD;
    switch ((K + 31) / 32) {
        case 1: return extract_partitioned_w<1>(ctx, rd, (int) K, mode, B, G, counts_out);
        case 2: return extract_partitioned_w<2>(ctx, rd, (int) K, mode, B, G, counts_out);
        case 3: return extract_partitioned_w<3>(ctx, rd, (int) K, mode, B, G, counts_out);
        default: return extract_partitioned_w<4>(ctx, rd, (int) K, mode, B, G, counts_out);
    }
}

sb200_records *extract_records(sb200_ctx *ctx, const sb200_reads *rd, unsigned K, int canonical_only, int add_rc) {
    SB200_REQUIRE(K >= 1 && K <= 128, "K out of range [1,128]");
    switch ((K + 31) / 32) {
        case 1: return extract_records_w<1>(ctx, rd, (int) K, canonical_only, add_rc);
        case 2: return extract_records_w<2>(ctx, rd, (int) K, canonical_only, add_rc);
        case 3: return extract_records_w<3>(ctx, rd, (int) K, canonical_only, add_rc);
        default: return extract_records_w<4>(ctx, rd, (int) K, canonical_only, add_rc);
    }
}

template<int WS, int W>
static sb200_records *derive_records_w(sb200_ctx *ctx, const sb200_kmers *kp) {
    sb200_records *r = new sb200_records();
    r->ctx = ctx; r->k = kp->k - 1; r->words = W; r->n = kp->size * 2;
    r->data.alloc(ctx, r->n * W);
    auto derive_kernel_ = derive_kernel<WS, W>;
    const int used = 2 * ((int) r->k - 32 * (W - 1));
    const int pshift = (used + 3 <= 64 && !ctx->no_mask_payload) ? used : -1;
    r->mask_payload = pshift >= 0;
    if (kp->size) LAUNCH(ctx, derive_kernel_, div_up(kp->size, 256), 256, 0, kp->data.p, kp->size, (int) r->k, pshift, r->data.p);
    CUDA_CHECK(cudaStreamSynchronize(ctx->stream));   // blocking like every entry point
    return r;
}

sb200_records *derive_records(sb200_ctx *ctx, const sb200_kmers *kp) {
    SB200_REQUIRE(kp->k >= 2, "source k-mers too short");
    int WS = (int) kp->words, W = (int) ((kp->k - 1 + 31) / 32);
    if (WS == 1) return derive_records_w<1, 1>(ctx, kp);
    if (WS == 2 && W == 1) return derive_records_w<2, 1>(ctx, kp);
    if (WS == 2) return derive_records_w<2, 2>(ctx, kp);
    if (WS == 3 && W == 2) return derive_records_w<3, 2>(ctx, kp);
    if (WS == 3) return derive_records_w<3, 3>(ctx, kp);
    if (WS == 4 && W == 3) return derive_records_w<4, 3>(ctx, kp);
    return derive_records_w<4, 4>(ctx, kp);
}

// first position of every owner's group = the scanned histogram matrix at (owner, tile 0)
__global__ void owner_starts_kernel(const uint32_t *__restrict__ offsets, uint32_t num_tiles, uint32_t n_parts, uint32_t *__restrict__ starts) {
    uint32_t o = blockIdx.x * blockDim.x + threadIdx.x;
    if (o < n_parts) starts[o] = offsets[(uint64_t) o * num_tiles];
}

// One stable counting pass whose digit is the owner id; the per-owner counts fall out of the pass's own scanned histogram.
template<int W>
static void partition_records_w(sb200_ctx *ctx, sb200_records *r, uint32_t B, uint32_t n_parts, uint64_t *counts_out) {
    DigitSel sel{-2, (int) n_parts, B, r->marker ? 1 : 0, 0, 0, last_word_mask((int) r->k)};   // the owner hash ignores a payload
    for (uint32_t p = 0; p < n_parts; ++p) counts_out[p] = 0;
    if (r->n == 0) return;
    const uint64_t n = r->n;
    SB200_REQUIRE(n < (1ull << 32), "more than 2^32-1 k-mer instances on one GPU: shard the input");
    constexpr int ITEMS = RsItems<W>::value;
    const uint32_t num_tiles = div_up(n, RS_THREADS * ITEMS);
    DevBuf<uint32_t> hist(ctx, (uint64_t) RS_BINS * num_tiles);
    DevBuf<uint64_t> scratch(ctx, n * W);
    DevBuf<uint32_t> starts(ctx, RS_BINS);
    LAUNCH(ctx, rs_hist_kernel<W>, num_tiles, RS_THREADS, 0, r->data.p, n, sel, hist.p, num_tiles);
    exclusive_scan<uint32_t>(ctx, hist.p, (uint64_t) RS_BINS * num_tiles, nullptr);
    LAUNCH(ctx, owner_starts_kernel, 1, RS_BINS, 0, hist.p, num_tiles, n_parts, starts.p);
    LAUNCH(ctx, rs_scatter_kernel<W>, num_tiles, RS_THREADS, 0, r->data.p, scratch.p, n, sel, hist.p, num_tiles);
    std::swap(r->data, scratch);   // keep the buffer that holds the grouped records
    std::vector<uint32_t> h(n_parts);
    ctx->fetch(h.data(), starts.p, (size_t) n_parts * 4);
    for (uint32_t p = 0; p < n_parts; ++p) counts_out[p] = (p + 1 < n_parts ? (uint64_t) h[p + 1] : n) - h[p];
}

void partition_records(sb200_ctx *ctx, sb200_records *r, unsigned B, unsigned n_parts, uint64_t *counts_out) {
    SB200_REQUIRE(n_parts >= 1 && n_parts <= 256, "number of owners out of range [1,256]");
    SB200_REQUIRE(B >= 1 && B <= 65536 && B % n_parts == 0, "num_buckets must be a multiple of the number of owners");
    switch (r->words) {
        case 1: partition_records_w<1>(ctx, r, B, n_parts, counts_out); break;
        case 2: partition_records_w<2>(ctx, r, B, n_parts, counts_out); break;
        case 3: partition_records_w<3>(ctx, r, B, n_parts, counts_out); break;
        default: partition_records_w<4>(ctx, r, B, n_parts, counts_out); break;
    }
}

sb200_kmers *count_records(sb200_ctx *ctx, sb200_records *r, unsigned B, int want_counts, unsigned first_bucket, unsigned n_owned) {
    SB200_REQUIRE(B >= 1 && B <= 65536, "num_buckets out of range [1,65536]");
    SB200_REQUIRE((uint64_t) first_bucket + n_owned <= B, "owned bucket range exceeds num_buckets");
    sb200_kmers *s;
    SB200_REQUIRE(!(r->mask_payload && want_counts), "records with a mask payload cannot be counted");
    if (r->n == 0) {   // an owner that received nothing (small input, many GPUs): an empty shard, not an error — the caller decides on the global total
        s = new sb200_kmers();
        s->ctx = ctx; s->k = r->k; s->words = r->words; s->num_buckets = B; s->size = 0;
        s->data.alloc(ctx, 0);
        if (want_counts) s->counts.alloc(ctx, 0);
        if (r->mask_payload || r->pay.p) s->masks_file.alloc(ctx, 4);
        s->bucket_starts.alloc(ctx, (uint64_t) B + 1);
        s->bucket_starts.zero();
        s->bucket_starts_host.assign((size_t) B + 1, 0);
        CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
        return s;
    }
    const int pshift = r->mask_payload ? 2 * ((int) r->k - 32 * ((int) r->words - 1)) : -1;
    switch (r->words) {
        case 1: s = finish_set<1>(ctx, r->data, r->n, (int) r->k, B, want_counts != 0, r->double_palindromes, r->marker, pshift, first_bucket, n_owned); break;
        case 2: s = finish_set<2>(ctx, r->data, r->n, (int) r->k, B, want_counts != 0, r->double_palindromes, r->marker, pshift, first_bucket, n_owned); break;
        case 3: s = finish_set<3>(ctx, r->data, r->n, (int) r->k, B, want_counts != 0, r->double_palindromes, r->marker, pshift, first_bucket, n_owned); break;
        default: s = finish_set<4>(ctx, r->data, r->n, (int) r->k, B, want_counts != 0, r->double_palindromes, r->marker, pshift, first_bucket, n_owned); break;
    }
    if (!want_counts) s->instances = 0;
    r->n = 0;
    return s;
}

// ---- hash-sharded path on the staged kernels -------------------------------------------------------------------------------------------
// Sender: the producer (reads / own (k+1)-mers) runs the count pass over COARSE bins of the global group key only (G x n_co <= 1024 bins:
// owner-major, so an owner's records are one contiguous range) and pass 1; records, run sizes (and payload bytes) travel.  Receiver: a
// count pass over what arrived gives its fine-group starts, pass 2 takes the tiles of every (source, coarse bin) run straight from the
// receive buffer, then the group kernel as on one GPU.  The owner's two counting passes and the sender's partition pass of round 1 are gone.

template<int W>
static bool shard_plan_w(sb200_ctx *ctx, uint64_t n_owner_est, uint32_t B, uint32_t G, int K, ShardPlan &pl) {
    if (ctx->counting_passes || ctx->atomic_partition || 2 * K < 24 || B % G) return false;
    const uint32_t n_owned = B / G;
    if (n_owned > SP_MAX_GROUPS) return false;
    pl.p = choose_prefix_bits<W>(n_owner_est, n_owned, K, ctx->group_chunk ? 0 : 7168);
    while (pl.p > 0 && ((uint64_t) n_owned << pl.p) > SP_MAX_GROUPS) --pl.p;
    pl.n_go = (uint32_t) ((uint64_t) n_owned << pl.p);
    if (n_owner_est / std::max<uint64_t>(pl.n_go, 1) > 10752 && !ctx->group_chunk) return false;   // (see staged_plan)
    // coarse bins: all owners' bins together are the local bins of a pass-1 tile
    const uint32_t max_total = std::min<uint32_t>(SP_MAX_BINS, std::max<uint32_t>(SpCfg<W>::CAP / 8, G));
    // (a pass-2 tile orders at most SP_MAX_BINS fine groups: beyond that the coarse bins stay finer than a pass-1 tile would like —
    //  shorter runs — rather than giving the staged path up; 3- and 4-word records on 8 GPUs are that case)
    pl.s = 0;
    while (pl.s < pl.p && (2u << pl.s) <= (uint32_t) SP_MAX_BINS && (uint64_t) (pl.n_go >> pl.s) * G > max_total) ++pl.s;
    if ((uint64_t) (pl.n_go >> pl.s) * G > SP_MAX_BINS || (1u << pl.s) > (uint32_t) SP_MAX_BINS) return false;
    pl.n_co = pl.n_go >> pl.s;   // s <= p: n_go = n_owned << p is a multiple of 2^s
    return pl.n_co >= 1;
}

bool shard_plan(sb200_ctx *ctx, unsigned W, uint64_t n_owner_est, unsigned B, unsigned G, unsigned K, ShardPlan *pl) {
    switch (W) {
        case 1: return shard_plan_w<1>(ctx, n_owner_est, B, G, (int) K, *pl);
        case 2: return shard_plan_w<2>(ctx, n_owner_est, B, G, (int) K, *pl);
        case 3: return shard_plan_w<3>(ctx, n_owner_est, B, G, (int) K, *pl);
        default: return shard_plan_w<4>(ctx, n_owner_est, B, G, (int) K, *pl);
    }
}

__global__ void sp_owner_totals_kernel(const uint32_t *__restrict__ cstart, uint32_t n_co, uint32_t G, uint32_t *__restrict__ out /* G + 1 */) {
    const uint32_t g = threadIdx.x;
    if (g <= G) out[g] = cstart[(uint64_t) g * n_co];
}

// after the coarse count: starts, cursors, per-owner totals on the host; counts_keep = the counts themselves (they travel with the records)
static uint64_t shard_send_tables(sb200_ctx *ctx, uint32_t G, uint32_t n_co, DevBuf<uint32_t> &hist, DevBuf<uint32_t> &counts_keep, DevBuf<uint32_t> &cur1,
                                  uint64_t *owner_counts) {
    const uint64_t nb = (uint64_t) G * n_co;
    counts_keep.alloc(ctx, nb);
    CUDA_CHECK(cudaMemcpyAsync(counts_keep.p, hist.p, nb * 4, cudaMemcpyDeviceToDevice, ctx->stream));
    exclusive_scan<uint32_t>(ctx, hist.p, nb + 1, nullptr);
    cur1.alloc(ctx, nb);
    CUDA_CHECK(cudaMemcpyAsync(cur1.p, hist.p, nb * 4, cudaMemcpyDeviceToDevice, ctx->stream));
    DevBuf<uint32_t> tot(ctx, (uint64_t) G + 1);
    LAUNCH(ctx, sp_owner_totals_kernel, 1, 128, 0, hist.p, n_co, G, tot.p);
    std::vector<uint32_t> h((size_t) G + 1);
    ctx->fetch(h.data(), tot.p, ((size_t) G + 1) * 4);
    for (uint32_t g = 0; g < G; ++g) owner_counts[g] = h[g + 1] - h[g];
    return h[G];
}

template<int W>
static sb200_records *shard_send_reads_w(sb200_ctx *ctx, const sb200_reads *rd, int K, uint32_t B, uint32_t G, const ShardPlan &pl, uint64_t *owner_counts) {
    const GroupSel gs{B, 0u, pl.p, (W == 1) ? 2 * K : 64};
    const uint32_t n_bins = G * pl.n_co;
    std::unique_ptr<sb200_records> r(new sb200_records());
    r->ctx = ctx; r->k = (unsigned) K; r->words = W; r->n = 0; r->double_palindromes = true;
    DevBuf<uint32_t> hist(ctx, (uint64_t) n_bins + 1), cur1;
    hist.zero();
    auto sp_count_reads_kernel_ = sp_count_reads_kernel<W>;
    auto sp_scatter_reads_kernel_ = sp_scatter_reads_kernel<W>;
    const size_t smem_c = (size_t) n_bins * 4, smem_s = sp_stage_bytes<W>(SpCfg<W>::CAP, false);
    CUDA_CHECK(cudaFuncSetAttribute(sp_count_reads_kernel_, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) std::max<size_t>(smem_c, 1024)));
    CUDA_CHECK(cudaFuncSetAttribute(sp_scatter_reads_kernel_, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem_s));
    // (the coarse histogram is small: two count CTAs per SM)
    if (rd->n_reads)
        LAUNCH(ctx, sp_count_reads_kernel_, (unsigned) ctx->num_sms * 2, SPC_THREADS, smem_c, rd->words.p, rd->word_off.p, rd->len.p, rd->n_reads, K, (int) PART_CANON,
               gs, pl.s, n_bins, hist.p);
    const uint64_t n = shard_send_tables(ctx, G, pl.n_co, hist, r->coarse_counts, cur1, owner_counts);
    SB200_REQUIRE(n < (1ull << 32), "more than 2^32-1 k-mer instances on one GPU: use more GPUs");
    r->n = n;
    r->data.alloc(ctx, n * W);
    if (n) {
        const uint32_t max_nwin = rd->max_len >= (uint32_t) K ? rd->max_len - (uint32_t) K + 1 : 128u;
        const uint32_t chunk_cap = 32u * std::min<uint32_t>(PART_RUN, (max_nwin + 31u) / 32u);
        const int rounds = std::max<int>(1, SpCfg<W>::CAP / (int) (SP_WARPS * chunk_cap));
        LAUNCH(ctx, sp_scatter_reads_kernel_, (unsigned) ctx->num_sms * 2, SP_THREADS, smem_s, rd->words.p, rd->word_off.p, rd->len.p, rd->n_reads, K,
               (int) PART_CANON, gs, pl.s, n_bins, rounds, cur1.p, r->data.p, SpPeers());
    }
    CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    return r.release();
}

sb200_records *shard_send_reads(sb200_ctx *ctx, const sb200_reads *rd, unsigned K, unsigned B, unsigned G, const ShardPlan *pl, uint64_t *owner_counts) {
    switch ((K + 31) / 32) {
        case 1: return shard_send_reads_w<1>(ctx, rd, (int) K, B, G, *pl, owner_counts);
        case 2: return shard_send_reads_w<2>(ctx, rd, (int) K, B, G, *pl, owner_counts);
        case 3: return shard_send_reads_w<3>(ctx, rd, (int) K, B, G, *pl, owner_counts);
        default: return shard_send_reads_w<4>(ctx, rd, (int) K, B, G, *pl, owner_counts);
    }
}

// ---- the same with the exchange inside pass 1: the runs go straight into the owners' receive buffers (SpPeers, staged_partition.cuh) ----
__global__ void sp_owner_relative_kernel(const uint32_t *__restrict__ start /* exclusive scan over [G][n_co], + total */, uint32_t n_co, uint32_t G,
                                         uint32_t *__restrict__ cur1, uint32_t *__restrict__ tot /* G + 1 */) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < G * n_co) cur1[i] = start[i] - start[(i / n_co) * n_co];
    if (i <= G) tot[i] = start[(uint64_t) i * n_co];
}

struct SpPeerCounts { uint32_t *dst[64]; };
__global__ void sp_peer_counts_kernel(const uint32_t *__restrict__ raw, uint32_t n_co, const __grid_constant__ SpPeerCounts pc) {
    uint32_t *d = pc.dst[blockIdx.x];
    for (uint32_t b = threadIdx.x; b < n_co; b += blockDim.x) d[b] = raw[(uint64_t) blockIdx.x * n_co + b];
}

// after the coarse count: per-owner totals on the host, cursors relative to the owner's run
static void shard_peer_tables(sb200_ctx *ctx, uint32_t G, uint32_t n_co, ShardPeerSend *st, uint64_t *owner_counts) {
    const uint64_t nb = (uint64_t) G * n_co;
    DevBuf<uint32_t> start(ctx, nb + 1), tot(ctx, (uint64_t) G + 1);
    CUDA_CHECK(cudaMemcpyAsync(start.p, st->raw.p, (nb + 1) * 4, cudaMemcpyDeviceToDevice, ctx->stream));
    exclusive_scan<uint32_t>(ctx, start.p, nb + 1, nullptr);
    st->cur1.alloc(ctx, nb);
    LAUNCH(ctx, sp_owner_relative_kernel, div_up(nb + 1, 256), 256, 0, start.p, n_co, G, st->cur1.p, tot.p);
    std::vector<uint32_t> h((size_t) G + 1);
    ctx->fetch(h.data(), tot.p, ((size_t) G + 1) * 4);
    CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    for (uint32_t g = 0; g < G; ++g) owner_counts[g] = h[g + 1] - h[g];
}

template<int W>
static ShardPeerSend *shard_peer_count_reads_w(sb200_ctx *ctx, const sb200_reads *rd, int K, uint32_t B, uint32_t G, const ShardPlan &pl, uint64_t *owner_counts) {
    const GroupSel gs{B, 0u, pl.p, (W == 1) ? 2 * K : 64};
    const uint32_t n_bins = G * pl.n_co;
    std::unique_ptr<ShardPeerSend> st(new ShardPeerSend());
    st->k = (unsigned) K; st->words = W; st->double_palindromes = true; st->pl = pl;
    st->raw.alloc(ctx, (uint64_t) n_bins + 1);
    st->raw.zero();
    auto sp_count_reads_kernel_ = sp_count_reads_kernel<W>;
    const size_t smem_c = (size_t) n_bins * 4;
    CUDA_CHECK(cudaFuncSetAttribute(sp_count_reads_kernel_, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) std::max<size_t>(smem_c, 1024)));
    if (rd->n_reads)
        LAUNCH(ctx, sp_count_reads_kernel_, (unsigned) ctx->num_sms * 2, SPC_THREADS, smem_c, rd->words.p, rd->word_off.p, rd->len.p, rd->n_reads, K, (int) PART_CANON,
               gs, pl.s, n_bins, st->raw.p);
    shard_peer_tables(ctx, G, pl.n_co, st.get(), owner_counts);
    return st.release();
}

template<int W>
static void shard_peer_scatter_reads_w(sb200_ctx *ctx, const sb200_reads *rd, uint32_t B, uint32_t G, ShardPeerSend *st, const SpPeers &peers) {
    const int K = (int) st->k;
    const GroupSel gs{B, 0u, st->pl.p, (W == 1) ? 2 * K : 64};
    auto sp_scatter_reads_peer_kernel_ = sp_scatter_reads_kernel<W, true>;
    const size_t smem_s = sp_stage_bytes<W>(SpCfg<W>::CAP, false);
    CUDA_CHECK(cudaFuncSetAttribute(sp_scatter_reads_peer_kernel_, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem_s));
    const uint32_t max_nwin = rd->max_len >= (uint32_t) K ? rd->max_len - (uint32_t) K + 1 : 128u;
    const uint32_t chunk_cap = 32u * std::min<uint32_t>(PART_RUN, (max_nwin + 31u) / 32u);
    const int rounds = std::max<int>(1, SpCfg<W>::CAP / (int) (SP_WARPS * chunk_cap));
    if (rd->n_reads)
        LAUNCH(ctx, sp_scatter_reads_peer_kernel_, (unsigned) ctx->num_sms * 2, SP_THREADS, smem_s, rd->words.p, rd->word_off.p, rd->len.p, rd->n_reads, K,
               (int) PART_CANON, gs, st->pl.s, G * st->pl.n_co, rounds, st->cur1.p, (uint64_t *) nullptr, peers);
}

template<int WS, int W>
static ShardPeerSend *shard_peer_count_derive_w(sb200_ctx *ctx, const sb200_kmers *kp, uint32_t B, uint32_t G, const ShardPlan &pl, uint64_t *owner_counts) {
    const int k = (int) kp->k - 1;
    const GroupSel gs{B, 0u, pl.p, (W == 1) ? 2 * k : 64};
    const uint32_t n_bins = G * pl.n_co;
    const int used = 2 * (k - 32 * (W - 1));
    std::unique_ptr<ShardPeerSend> st(new ShardPeerSend());
    st->k = (unsigned) k; st->words = W; st->pl = pl;
    st->pshift = (used + 3 <= 64 && !ctx->no_mask_payload) ? used : -1;
    st->side_pay = st->pshift < 0 && !ctx->no_mask_payload;
    st->mask_payload = st->pshift >= 0;
    st->raw.alloc(ctx, (uint64_t) n_bins + 1);
    st->raw.zero();
    auto sp_count_derive_kernel_ = sp_count_derive_kernel<WS, W>;
    const size_t smem_c = (size_t) n_bins * 4;
    CUDA_CHECK(cudaFuncSetAttribute(sp_count_derive_kernel_, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) std::max<size_t>(smem_c, 1024)));
    if (kp->size) LAUNCH(ctx, sp_count_derive_kernel_, (unsigned) ctx->num_sms * 2, SPC_THREADS, smem_c, kp->data.p, kp->size, k, gs, pl.s, n_bins, st->raw.p);
    shard_peer_tables(ctx, G, pl.n_co, st.get(), owner_counts);
    return st.release();
}

template<int WS, int W>
static void shard_peer_scatter_derive_w(sb200_ctx *ctx, const sb200_kmers *kp, uint32_t B, uint32_t G, ShardPeerSend *st, const SpPeers &peers) {
    const int k = (int) st->k;
    const GroupSel gs{B, 0u, st->pl.p, (W == 1) ? 2 * k : 64};
    auto sp_scatter_derive_peer_kernel_ = sp_scatter_derive_kernel<WS, W, true>;
    const size_t smem_s = sp_stage_bytes<W>(SpCfg<W>::CAP, st->side_pay);
    CUDA_CHECK(cudaFuncSetAttribute(sp_scatter_derive_peer_kernel_, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem_s));
    // (out_pay only tells the kernel whether the stage carries payload bytes; the peers' pointers are what it stores through)
    if (kp->size)
        LAUNCH(ctx, sp_scatter_derive_peer_kernel_, (unsigned) ctx->num_sms * 2, SP_THREADS, smem_s, kp->data.p, kp->size, k, st->pshift, gs, st->pl.s,
               G * st->pl.n_co, st->cur1.p, (uint64_t *) nullptr, st->side_pay ? peers.pay[0] : (uint8_t *) nullptr, peers);
}

ShardPeerSend *shard_peer_count_reads(sb200_ctx *ctx, const sb200_reads *rd, unsigned K, unsigned B, unsigned G, const ShardPlan *pl, uint64_t *owner_counts) {
    switch ((K + 31) / 32) {
        case 1: return shard_peer_count_reads_w<1>(ctx, rd, (int) K, B, G, *pl, owner_counts);
        case 2: return shard_peer_count_reads_w<2>(ctx, rd, (int) K, B, G, *pl, owner_counts);
        case 3: return shard_peer_count_reads_w<3>(ctx, rd, (int) K, B, G, *pl, owner_counts);
        default: return shard_peer_count_reads_w<4>(ctx, rd, (int) K, B, G, *pl, owner_counts);
    }
}

void shard_peer_scatter_reads(sb200_ctx *ctx, const sb200_reads *rd, unsigned B, unsigned G, ShardPeerSend *st, const SpPeers *peers) {
    switch (st->words) {
        case 1: shard_peer_scatter_reads_w<1>(ctx, rd, B, G, st, *peers); break;
        case 2: shard_peer_scatter_reads_w<2>(ctx, rd, B, G, st, *peers); break;
        case 3: shard_peer_scatter_reads_w<3>(ctx, rd, B, G, st, *peers); break;
        default: shard_peer_scatter_reads_w<4>(ctx, rd, B, G, st, *peers); break;
    }
}

ShardPeerSend *shard_peer_count_derive(sb200_ctx *ctx, const sb200_kmers *kp, unsigned B, unsigned G, const ShardPlan *pl, uint64_t *owner_counts) {
    int WS = (int) kp->words, W = (int) ((kp->k - 1 + 31) / 32);
    if (WS == 1) return shard_peer_count_derive_w<1, 1>(ctx, kp, B, G, *pl, owner_counts);
    if (WS == 2 && W == 1) return shard_peer_count_derive_w<2, 1>(ctx, kp, B, G, *pl, owner_counts);
    if (WS == 2) return shard_peer_count_derive_w<2, 2>(ctx, kp, B, G, *pl, owner_counts);
    if (WS == 3 && W == 2) return shard_peer_count_derive_w<3, 2>(ctx, kp, B, G, *pl, owner_counts);
    if (WS == 3) return shard_peer_count_derive_w<3, 3>(ctx, kp, B, G, *pl, owner_counts);
    if (WS == 4 && W == 3) return shard_peer_count_derive_w<4, 3>(ctx, kp, B, G, *pl, owner_counts);
    return shard_peer_count_derive_w<4, 4>(ctx, kp, B, G, *pl, owner_counts);
}

void shard_peer_scatter_derive(sb200_ctx *ctx, const sb200_kmers *kp, unsigned B, unsigned G, ShardPeerSend *st, const SpPeers *peers) {
    int WS = (int) kp->words, W = (int) st->words;
    if (WS == 1) return shard_peer_scatter_derive_w<1, 1>(ctx, kp, B, G, st, *peers);
    if (WS == 2 && W == 1) return shard_peer_scatter_derive_w<2, 1>(ctx, kp, B, G, st, *peers);
    if (WS == 2) return shard_peer_scatter_derive_w<2, 2>(ctx, kp, B, G, st, *peers);
    if (WS == 3 && W == 2) return shard_peer_scatter_derive_w<3, 2>(ctx, kp, B, G, st, *peers);
    if (WS == 3) return shard_peer_scatter_derive_w<3, 3>(ctx, kp, B, G, st, *peers);
    if (WS == 4 && W == 3) return shard_peer_scatter_derive_w<4, 3>(ctx, kp, B, G, st, *peers);
    return shard_peer_scatter_derive_w<4, 4>(ctx, kp, B, G, st, *peers);
}

// the coarse counts of every owner's part, stored into the heads of the owners' buffers
void shard_peer_send_counts(sb200_ctx *ctx, unsigned G, const ShardPeerSend *st, uint32_t *const *dst) {
    SpPeerCounts pc;
    for (unsigned g = 0; g < 64; ++g) pc.dst[g] = g < G ? dst[g] : nullptr;
    LAUNCH(ctx, sp_peer_counts_kernel, G, 128, 0, st->raw.p, st->pl.n_co, pc);
}

template<int WS, int W>
static sb200_records *shard_send_derive_w(sb200_ctx *ctx, const sb200_kmers *kp, uint32_t B, uint32_t G, const ShardPlan &pl, uint64_t *owner_counts) {
    const int k = (int) kp->k - 1;
    const GroupSel gs{B, 0u, pl.p, (W == 1) ? 2 * k : 64};
    const uint32_t n_bins = G * pl.n_co;
    const int used = 2 * (k - 32 * (W - 1));
    const int pshift = (used + 3 <= 64 && !ctx->no_mask_payload) ? used : -1;
    const bool side_pay = pshift < 0 && !ctx->no_mask_payload;
    std::unique_ptr<sb200_records> r(new sb200_records());
    r->ctx = ctx; r->k = (unsigned) k; r->words = W; r->n = 0; r->mask_payload = pshift >= 0;
    DevBuf<uint32_t> hist(ctx, (uint64_t) n_bins + 1), cur1;
    hist.zero();
    auto sp_count_derive_kernel_ = sp_count_derive_kernel<WS, W>;
    auto sp_scatter_derive_kernel_ = sp_scatter_derive_kernel<WS, W>;
    const size_t smem_c = (size_t) n_bins * 4, smem_s = sp_stage_bytes<W>(SpCfg<W>::CAP, side_pay);
    CUDA_CHECK(cudaFuncSetAttribute(sp_count_derive_kernel_, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) std::max<size_t>(smem_c, 1024)));
    CUDA_CHECK(cudaFuncSetAttribute(sp_scatter_derive_kernel_, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem_s));
    if (kp->size) LAUNCH(ctx, sp_count_derive_kernel_, (unsigned) ctx->num_sms * 2, SPC_THREADS, smem_c, kp->data.p, kp->size, k, gs, pl.s, n_bins, hist.p);
    const uint64_t n = shard_send_tables(ctx, G, pl.n_co, hist, r->coarse_counts, cur1, owner_counts);
    SB200_REQUIRE(n < (1ull << 32), "more than 2^32-1 k-mer instances on one GPU: use more GPUs");
    r->n = n;
    r->data.alloc(ctx, n * W);
    if (side_pay) r->pay.alloc(ctx, n + 1);
    if (n) LAUNCH(ctx, sp_scatter_derive_kernel_, (unsigned) ctx->num_sms * 2, SP_THREADS, smem_s, kp->data.p, kp->size, k, pshift, gs, pl.s, n_bins, cur1.p,
                  r->data.p, r->pay.p, SpPeers());
    CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    return r.release();
}

sb200_records *shard_send_derive(sb200_ctx *ctx, const sb200_kmers *kp, unsigned B, unsigned G, const ShardPlan *pl, uint64_t *owner_counts) {
    int WS = (int) kp->words, W = (int) ((kp->k - 1 + 31) / 32);
    if (WS == 1) return shard_send_derive_w<1, 1>(ctx, kp, B, G, *pl, owner_counts);
    if (WS == 2 && W == 1) return shard_send_derive_w<2, 1>(ctx, kp, B, G, *pl, owner_counts);
    if (WS == 2) return shard_send_derive_w<2, 2>(ctx, kp, B, G, *pl, owner_counts);
    if (WS == 3 && W == 2) return shard_send_derive_w<3, 2>(ctx, kp, B, G, *pl, owner_counts);
    if (WS == 3) return shard_send_derive_w<3, 3>(ctx, kp, B, G, *pl, owner_counts);
    if (WS == 4 && W == 3) return shard_send_derive_w<4, 3>(ctx, kp, B, G, *pl, owner_counts);
    return shard_send_derive_w<4, 4>(ctx, kp, B, G, *pl, owner_counts);
}

// Receiver: `got` holds the runs of all sources (run_start: G + 1 record positions), got->coarse_counts the [G x n_co] run sizes that came
// with them, got->pay the payload bytes if the senders keep them beside the records.  Consumes `got`.
template<int W>
static sb200_kmers *shard_receive_w(sb200_ctx *ctx, sb200_records *got, const uint64_t *run_start, uint32_t G, const ShardPlan &pl, uint32_t B,
                                    uint32_t first_bucket, uint32_t n_owned, bool want_counts) {
    const uint64_t n = got->n;
    const int K = (int) got->k;
    const uint64_t lw_keep = last_word_mask(K);
    const GroupSel gs{B, first_bucket, pl.p, (W == 1) ? 2 * K : 64};
    const int pshift = got->mask_payload ? 2 * (K - 32 * (W - 1)) : -1;
    // fine-group histogram of what arrived
    DevBuf<uint32_t> hist(ctx, (uint64_t) pl.n_go + 1), cur2(ctx, (uint64_t) pl.n_go + 1);
    hist.zero();
    auto sp_count_records_kernel_ = sp_count_records_kernel<W>;
    const size_t smem_c = (size_t) pl.n_go * 4;
    CUDA_CHECK(cudaFuncSetAttribute(sp_count_records_kernel_, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) std::max<size_t>(smem_c, 1024)));
    LAUNCH(ctx, sp_count_records_kernel_, (unsigned) ctx->num_sms, SPC_THREADS, smem_c, got->data.p, n, gs, lw_keep, pl.n_go, hist.p);
    exclusive_scan<uint32_t>(ctx, hist.p, (uint64_t) pl.n_go + 1, nullptr);
    CUDA_CHECK(cudaMemcpyAsync(cur2.p, hist.p, ((size_t) pl.n_go + 1) * 4, cudaMemcpyDeviceToDevice, ctx->stream));
    // segments (source, coarse bin) -> tiles
    const uint32_t n_seg = G * pl.n_co;
    DevBuf<uint32_t> seg(ctx, 3 * (uint64_t) n_seg);
    DevBuf<uint64_t> rs(ctx, (uint64_t) G + 1);
    CUDA_CHECK(cudaMemcpyAsync(rs.p, run_start, ((size_t) G + 1) * 8, cudaMemcpyHostToDevice, ctx->stream));
    LAUNCH(ctx, sp_recv_segments_kernel, G, 32, 0, got->coarse_counts.p, rs.p, G, pl.n_co, pl.s, seg.p, seg.p + n_seg, seg.p + 2 * (uint64_t) n_seg);
    StagedTiles tl;
    staged_tiles<W>(ctx, seg.p, seg.p + n_seg, seg.p + 2 * (uint64_t) n_seg, n_seg, n, tl);
    DevBuf<uint64_t> inst(ctx, n * W);
    DevBuf<uint8_t> pay;
    if (got->pay.p) pay.alloc(ctx, n + 1);
    staged_pass2<W>(ctx, pl.s, tl, cur2.p, gs, lw_keep, got->data.p, got->pay.p, inst.p, pay.p);
    CUDA_CHECK(cudaStreamSynchronize(ctx->stream));   // run_start (pageable) has been consumed
    sb200_kmers *s = finish_grouped<W>(ctx, inst, got->data, n, hist.p, pl.n_go, pl.p, K, B, want_counts, got->double_palindromes, false, pshift, pay.p,
                                       first_bucket, n_owned);
    if (!want_counts) s->instances = 0;
    got->n = 0;
    return s;
}

sb200_kmers *count_records(sb200_ctx *ctx, sb200_records *r, unsigned B, int want_counts, unsigned first_bucket, unsigned n_owned);
sb200_kmers *shard_receive(sb200_ctx *ctx, sb200_records *got, const uint64_t *run_start, unsigned G, const ShardPlan *pl, unsigned B, unsigned first_bucket,
                           unsigned n_owned, int want_counts) {
    if (got->n == 0) return count_records(ctx, got, B, want_counts, first_bucket, n_owned);   // the empty shard
    switch (got->words) {
        case 1: return shard_receive_w<1>(ctx, got, run_start, G, *pl, B, first_bucket, n_owned, want_counts != 0);
        case 2: return shard_receive_w<2>(ctx, got, run_start, G, *pl, B, first_bucket, n_owned, want_counts != 0);
        case 3: return shard_receive_w<3>(ctx, got, run_start, G, *pl, B, first_bucket, n_owned, want_counts != 0);
        default: return shard_receive_w<4>(ctx, got, run_start, G, *pl, B, first_bucket, n_owned, want_counts != 0);
    }
}

}  // namespace sb200
