// radix_sort.cuh — stable LSD radix sort of fixed-width k-mer records (W uint64 words, AoS) into the reference's
// file order: hash bucket first (KMerSegmentPolicy), then word-wise from word 0 (array_less).
//
// This replaces the reference's per-thread x per-bucket staging cells + libcxx::sort + loser-tree merge
// (C/utils/kmer_mph/kmer_splitter.hpp:111-167, kmer_index_builder.hpp:281-365) with counting passes over
// 8-bit digits: one histogram kernel, one scan of the (digit x tile) matrix and one stable scatter kernel per pass.
// Only the significant 2K bits are visited; the bucket id is recomputed from the record in the last pass(es)
// instead of being carried as payload, so records stay W words wide end to end.
#pragma once
#include "common.cuh"
#include "kmer_ops.cuh"
#include "scan.cuh"

namespace sb200 {

constexpr int RS_THREADS = 256;
constexpr int RS_WARPS = RS_THREADS / 32;
constexpr int RS_BINS = 256;

// records per thread; a tile (256 threads x ITEMS records) is staged in <= 32 KB of shared memory
#ifndef RS_ITEMS_W2
#define RS_ITEMS_W2 8
#endif
#ifndef RS_MIN_BLOCKS
#define RS_MIN_BLOCKS 4   // caps the scatter kernel at 64 registers (4 CTAs per SM): 96 registers / 2 CTAs cost 25 %, 85 / 3 another 3 % (profiles/r1b, r1c)
#endif
template<int W> struct RsItems { static constexpr int value = (W == 1) ? 16 : (W == 2) ? RS_ITEMS_W2 : 4; };

// digit selector: word >= 0 -> byte `shift/8` of that word; word == -1 -> byte of the bucket id; word == -2 -> owner of the
// bucket; word == -3 -> byte `shift/8` of the composite group key  bucket << p | top p bits of word 0  (segsort.cuh)
struct DigitSel {
    int word;
    int shift;
    uint32_t num_buckets;
    int marker;   // 1: an all-ones record is the "filtered out" marker and belongs to the last bucket (sorts last)
    int p;        // composite only: prefix bits of the value
    int top;      // composite only: significant bits of word 0 (64, or 2K for one-word records)
    uint64_t lw_keep = ~0ULL;   // bits of the LAST word that belong to the k-mer (the padding above may carry a payload, count.cu)
    uint32_t bucket_base = 0;   // composite only: first bucket this GPU owns (sharded path: every record lies in an owned bucket)
};

template<int W>
__device__ __forceinline__ uint32_t rs_bucket(const uint64_t *r, const DigitSel &d) {
    uint64_t c[W];
#pragma unroll
    for (int j = 0; j < W; ++j) c[j] = r[j];
    c[W - 1] &= d.lw_keep;
    uint32_t b = kmer_bucket<W>(c, d.num_buckets);
    if (d.marker) {
        bool m = true;
#pragma unroll
        for (int j = 0; j < W; ++j) m &= (r[j] == ~0ULL);
        if (m) b = d.num_buckets - 1;
    }
    return b;
}

template<int W>
__device__ __forceinline__ uint32_t rs_composite(const uint64_t *r, const DigitSel &d) {
    const uint32_t pre = d.p ? (uint32_t) ((r[0] >> (d.top - d.p)) & ((1ULL << d.p) - 1ULL)) : 0u;
    return ((rs_bucket<W>(r, d) - d.bucket_base) << d.p) | pre;
}

template<int W>
__device__ __forceinline__ uint32_t rs_digit(const uint64_t *r, const DigitSel &d) {
    if (d.word >= 0) return (uint32_t) (r[d.word] >> d.shift) & 0xFFu;
    if (d.word == -3) {
        // a digit that lies entirely inside the value prefix needs no bucket hash (the low pass of every two-pass grouping)
        if (d.shift + 8 <= d.p) return (uint32_t) (r[0] >> (d.top - d.p + d.shift)) & 0xFFu;
        return (rs_composite<W>(r, d) >> d.shift) & 0xFFu;
    }
    uint32_t b = rs_bucket<W>(r, d);
    if (d.word == -2) return (uint32_t) (((uint64_t) b * (uint32_t) d.shift) / d.num_buckets);   // owner of the bucket; shift = #owners
    return (b >> d.shift) & 0xFFu;
}

template<int W>
__device__ __forceinline__ void load_rec(const uint64_t *__restrict__ base, uint64_t idx, uint64_t *r) {
    if (W == 2) {
        ulonglong2 v = *reinterpret_cast<const ulonglong2 *>(base + idx * 2);
        r[0] = v.x; r[1] = v.y;
    } else if (W == 4) {
        const ulonglong2 *p = reinterpret_cast<const ulonglong2 *>(base + idx * 4);
        ulonglong2 a = p[0], b = p[1];
        r[0] = a.x; r[1] = a.y; r[2] = b.x; r[3] = b.y;
    } else {
#pragma unroll
        for (int j = 0; j < W; ++j) r[j] = base[idx * W + j];
    }
}

template<int W>
__device__ __forceinline__ void store_rec(uint64_t *__restrict__ base, uint64_t idx, const uint64_t *r) {
    if (W == 2) {
        *reinterpret_cast<ulonglong2 *>(base + idx * 2) = make_ulonglong2(r[0], r[1]);
    } else if (W == 4) {
        ulonglong2 *p = reinterpret_cast<ulonglong2 *>(base + idx * 4);
        p[0] = make_ulonglong2(r[0], r[1]);
        p[1] = make_ulonglong2(r[2], r[3]);
    } else {
#pragma unroll
        for (int j = 0; j < W; ++j) base[idx * W + j] = r[j];
    }
}

// tile histogram: hist[bin * num_tiles + tile]
template<int W>
__global__ void __launch_bounds__(RS_THREADS) rs_hist_kernel(const uint64_t *__restrict__ in, uint64_t n, DigitSel sel,
                                                            uint32_t *__restrict__ hist, uint32_t num_tiles) {
    constexpr int ITEMS = RsItems<W>::value;
    __shared__ uint32_t sh[RS_BINS];
    sh[threadIdx.x] = 0;
    __syncthreads();
    uint64_t base = (uint64_t) blockIdx.x * (RS_THREADS * ITEMS);
#pragma unroll 4
    for (int i = 0; i < ITEMS; ++i) {
        uint64_t idx = base + (uint64_t) i * RS_THREADS + threadIdx.x;
        if (idx < n) {
            uint32_t d;
            if (sel.word >= 0) {
                d = (uint32_t) (in[idx * W + sel.word] >> sel.shift) & 0xFFu;
            } else {
                uint64_t r[W];
                load_rec<W>(in, idx, r);
                d = rs_digit<W>(r, sel);
            }
            atomicAdd(&sh[d], 1u);
        }
    }
    __syncthreads();
    hist[(uint64_t) threadIdx.x * num_tiles + blockIdx.x] = sh[threadIdx.x];
}

// stable scatter, direct version (every thread stores its records straight to their bins; kept for reference — the
// staged version below replaced it): offsets[bin * num_tiles + tile] = exclusive scan of hist
template<int W>
__global__ void __launch_bounds__(RS_THREADS) rs_scatter_direct_kernel(const uint64_t *__restrict__ in, uint64_t *__restrict__ out, uint64_t n,
                                                               DigitSel sel, const uint32_t *__restrict__ offsets,
                                                               uint32_t num_tiles) {
    constexpr int ITEMS = RsItems<W>::value;
    __shared__ uint32_t cnt[RS_WARPS][RS_BINS];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < RS_WARPS * RS_BINS; i += RS_THREADS) (&cnt[0][0])[i] = 0;
    __syncthreads();

    // Only (digit, in-warp offset) is kept per record — 16 registers instead of the 16 records themselves — so that six
    // CTAs fit on an SM; the record is read again when it is moved (the tile was just read: an L1/L2 hit, not HBM).
    uint32_t info[ITEMS];   // bits 0-7 digit, bits 8-31 offset among the warp's records with the same digit
    const uint64_t warp_base = (uint64_t) blockIdx.x * (RS_THREADS * ITEMS) + (uint64_t) warp * (32 * ITEMS);
    const uint32_t lt_mask = (1u << lane) - 1u;
    uint32_t dg[ITEMS];
#pragma unroll
    for (int i = 0; i < ITEMS; ++i) {   // all loads of the tile first (memory-level parallelism), ranking afterwards
        uint64_t idx = warp_base + (uint64_t) i * 32 + lane;
        dg[i] = 0;
        if (idx < n) {
            if (sel.word >= 0) {
                dg[i] = (uint32_t) (in[idx * W + sel.word] >> sel.shift) & 0xFFu;
            } else {
                uint64_t r[W];
                load_rec<W>(in, idx, r);
                dg[i] = rs_digit<W>(r, sel);
            }
        }
    }
#pragma unroll
    for (int i = 0; i < ITEMS; ++i) {
        uint64_t idx = warp_base + (uint64_t) i * 32 + lane;
        bool ok = idx < n;
        uint32_t d = dg[i];
        uint32_t valid = __ballot_sync(0xffffffffu, ok);
        uint32_t peers = __match_any_sync(0xffffffffu, d) & valid;
        uint32_t old = ok ? cnt[warp][d] : 0;
        __syncwarp();
        if (ok && (peers & lt_mask) == 0) cnt[warp][d] = old + __popc(peers);
        __syncwarp();
        info[i] = d | ((old + __popc(peers & lt_mask)) << 8);
    }
    __syncthreads();
    {   // one thread per bin: turn per-warp counts into per-warp global bases
        uint32_t bin = threadIdx.x;
        uint32_t base = offsets[(uint64_t) bin * num_tiles + blockIdx.x];
#pragma unroll
        for (int w = 0; w < RS_WARPS; ++w) {
            uint32_t c = cnt[w][bin];
            cnt[w][bin] = base;
            base += c;
        }
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < ITEMS; ++i) {
        uint64_t idx = warp_base + (uint64_t) i * 32 + lane;
        if (idx < n) {
            uint64_t r[W];
            load_rec<W>(in, idx, r);
            store_rec<W>(out, (uint64_t) cnt[warp][info[i] & 0xFFu] + (info[i] >> 8), r);
        }
    }
}

// Sorts recs (n x W words) into (bucket, array_less) order.  `a` holds the input; `b` is scratch of the same size.
// Returns the buffer (a or b) that holds the result.
// Stable scatter with shared-memory staging.  A tile's records are ranked per warp with match.any (stable: warp order,
// then round, then lane), placed into shared memory in digit order, and only then written out: consecutive threads store
// consecutive records of the same bin, so a bin's run inside the tile leaves as one coalesced burst instead of one
// 16-byte store per record.  offsets[bin * num_tiles + tile] = exclusive scan of the histogram matrix.
template<int W>
__global__ void __launch_bounds__(RS_THREADS, RS_MIN_BLOCKS) rs_scatter_kernel(const uint64_t *__restrict__ in, uint64_t *__restrict__ out, uint64_t n,
                                                               DigitSel sel, const uint32_t *__restrict__ offsets,
                                                               uint32_t num_tiles) {
    constexpr int ITEMS = RsItems<W>::value;
    constexpr int TILE = RS_THREADS * ITEMS;
    __shared__ __align__(16) uint64_t srec[TILE * W];
    __shared__ uint8_t sdig[TILE];
    __shared__ uint32_t cnt[RS_WARPS][RS_BINS];
    __shared__ uint32_t delta[RS_BINS];     // global position - staged position, per bin (mod 2^32)
    __shared__ uint32_t s_scan[RS_THREADS / 32 + 1];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < RS_WARPS * RS_BINS; i += RS_THREADS) (&cnt[0][0])[i] = 0;
    __syncthreads();

    const uint64_t tile_base = (uint64_t) blockIdx.x * TILE;
    const uint64_t warp_base = tile_base + (uint64_t) warp * (32 * ITEMS);
    const uint32_t lt_mask = (1u << lane) - 1u;
    uint64_t rec[ITEMS][W];
    uint32_t info[ITEMS];   // bits 0-7 digit, bits 8-31 offset among the warp's records with the same digit
#pragma unroll
    for (int i = 0; i < ITEMS; ++i) {
        uint64_t idx = warp_base + (uint64_t) i * 32 + lane;
#pragma unroll
        for (int j = 0; j < W; ++j) rec[i][j] = 0;
        if (idx < n) load_rec<W>(in, idx, rec[i]);
    }
#pragma unroll
    for (int i = 0; i < ITEMS; ++i) {
        uint64_t idx = warp_base + (uint64_t) i * 32 + lane;
        bool ok = idx < n;
        uint32_t d = ok ? rs_digit<W>(rec[i], sel) : 0u;
        uint32_t valid = __ballot_sync(0xffffffffu, ok);
        uint32_t peers = __match_any_sync(0xffffffffu, d) & valid;
        uint32_t old = ok ? cnt[warp][d] : 0;
        __syncwarp();
        if (ok && (peers & lt_mask) == 0) cnt[warp][d] = old + __popc(peers);
        __syncwarp();
        info[i] = d | ((old + __popc(peers & lt_mask)) << 8);
    }
    __syncthreads();
    {   // one thread per bin: tile-local start of the bin (block scan over the 256 bin totals), per-warp bases, global delta
        const uint32_t bin = threadIdx.x;
        uint32_t tot = 0;
#pragma unroll
        for (int w = 0; w < RS_WARPS; ++w) tot += cnt[w][bin];
        uint32_t total_all;
        uint32_t start = block_exclusive_scan<uint32_t, RS_THREADS>(tot, &total_all, s_scan);
        delta[bin] = offsets[(uint64_t) bin * num_tiles + blockIdx.x] - start;
        uint32_t base = start;
#pragma unroll
        for (int w = 0; w < RS_WARPS; ++w) {
            uint32_t c = cnt[w][bin];
            cnt[w][bin] = base;
            base += c;
        }
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < ITEMS; ++i) {
        uint64_t idx = warp_base + (uint64_t) i * 32 + lane;
        if (idx < n) {
            const uint32_t d = info[i] & 0xFFu;
            const uint32_t pos = cnt[warp][d] + (info[i] >> 8);
#pragma unroll
            for (int j = 0; j < W; ++j) srec[(size_t) pos * W + j] = rec[i][j];
            sdig[pos] = (uint8_t) d;
        }
    }
    __syncthreads();
    const uint32_t count = (uint32_t) ((n - tile_base) < (uint64_t) TILE ? (n - tile_base) : (uint64_t) TILE);
    for (uint32_t q = threadIdx.x; q < count; q += RS_THREADS) {
        uint64_t r[W];
#pragma unroll
        for (int j = 0; j < W; ++j) r[j] = srec[(size_t) q * W + j];
        store_rec<W>(out, (uint64_t) (uint32_t) (q + delta[sdig[q]]), r);
    }
}

inline void append_bucket_passes(std::vector<DigitSel> &passes, uint32_t num_buckets, bool marker) {
    if (num_buckets > 1) {
        int bbits = 0;
        while ((1ull << bbits) < num_buckets) ++bbits;
        for (int s = 0; s < bbits; s += 8) passes.push_back(DigitSel{-1, s, num_buckets, marker ? 1 : 0});
    }
}

// every significant byte of the record, least significant first, then the bucket id: the complete file order
inline std::vector<DigitSel> full_passes(int W, int K, uint32_t num_buckets, bool marker) {
    std::vector<DigitSel> passes;
    for (int j = W - 1; j >= 0; --j) {
        int bits = (j == W - 1) ? (2 * K - 64 * (W - 1)) : 64;
        for (int s = 0; s < bits; s += 8) passes.push_back(DigitSel{j, s, 0, 0});
    }
    append_bucket_passes(passes, num_buckets, marker);
    return passes;
}

// the composite group key (segsort.cuh), least significant byte first: `bits` = bucket bits + p
inline std::vector<DigitSel> composite_passes(int bits, int p, int top, uint32_t num_buckets, bool marker, uint64_t lw_keep, uint32_t bucket_base = 0) {
    std::vector<DigitSel> passes;
    for (int s = 0; s < bits; s += 8) passes.push_back(DigitSel{-3, s, num_buckets, marker ? 1 : 0, p, top, lw_keep, bucket_base});
    return passes;
}

template<int W>
uint64_t *radix_sort_passes(sb200_ctx *ctx, uint64_t *a, uint64_t *b, uint64_t n, const std::vector<DigitSel> &passes) {
    SB200_REQUIRE(n < (1ull << 32), "more than 2^32-1 k-mer instances on one GPU: shard the input");
    if (n <= 1) return a;
    constexpr int ITEMS = RsItems<W>::value;
    uint32_t num_tiles = div_up(n, RS_THREADS * ITEMS);
    DevBuf<uint32_t> hist(ctx, (uint64_t) RS_BINS * num_tiles);
    uint64_t *src = a, *dst = b;
    for (const DigitSel &sel : passes) {
        LAUNCH(ctx, rs_hist_kernel<W>, num_tiles, RS_THREADS, 0, src, n, sel, hist.p, num_tiles);
        exclusive_scan<uint32_t>(ctx, hist.p, (uint64_t) RS_BINS * num_tiles, nullptr);
        LAUNCH(ctx, rs_scatter_kernel<W>, num_tiles, RS_THREADS, 0, src, dst, n, sel, hist.p, num_tiles);
        std::swap(src, dst);
    }
    return src;
}

template<int W>
uint64_t *radix_sort_records(sb200_ctx *ctx, uint64_t *a, uint64_t *b, uint64_t n, int K, uint32_t num_buckets, bool marker = false) {
    return radix_sort_passes<W>(ctx, a, b, n, full_passes(W, K, num_buckets, marker));
}

}  // namespace sb200
