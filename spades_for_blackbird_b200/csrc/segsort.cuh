// segsort.cuh — the last sorting digit, deduplication and counting in ONE shared-memory pass over groups of records.
//
// A full LSD sort of W x 64-bit records (what round-1's first version did) moves every record fifteen times through
// HBM.  Instead the records are grouped by a short COMPOSITE key with one or two stable counting passes (radix_sort.cuh):
//     composite(r) = bucket(r) << p | top p bits of the record's sort order            (bucket = KMerSegmentPolicy)
// with p chosen so that a group (all records with the same composite key) is a few thousand records and fits in the
// 227 KB of shared memory of one SM.  Groups are found by binary search over the grouped array (group_bounds_kernel), and
// one CTA per group then
//   (0) loads the group, sorts it by the NEXT 8 bits of the sort order with a shared-memory counting sort — the radix pass
//       that used to be a third trip through HBM plus a pass that marked segment boundaries — which leaves "segments" (runs
//       of equal 8-bit digit: a few dozen records, mostly copies of one genomic k-mer);
//   (1) hands every segment to ONE WARP, which deduplicates it with match.any on the record words (equal records elect
//       their lowest lane, which keeps the value and the number of copies), keeps the distinct values of the segment in
//       registers (up to 64: two per lane), merges further tiles of a long segment into that list, ranks the values against
//       each other with shuffles and writes them back in order — no shared-memory traffic, no block barrier, no quadratic
//       pass over copies (profiles/r1b_seg_chunk_source.md has the history of the block-wide version this replaced);
//   (2) scans the per-segment unique counts and packs the unique records (+ counts) at the front of the group's own range in
//       the output buffer; seg_compact_kernel closes the gaps after a scan of the group totals.
// A group larger than the shared-memory capacity (the multiplicity of genomic k-mers makes group sizes clumpy) is processed
// in ROUNDS over disjoint digit ranges, re-reading the group from L2 for each round.  Only a single digit bin that exceeds
// the capacity (a k-mer repeated thousands of times, low-complexity sequence) or an absurdly large group raises the fail
// flag, and the host then redoes the set with the generic LSD path.
#pragma once
#include "common.cuh"
#include "kmer_ops.cuh"
#include "radix_sort.cuh"
#include "scan.cuh"

namespace sb200 {

template<int W> struct SegCfg {
#ifndef SEG_CAP_W2
#define SEG_CAP_W2 6144
#define SEG_THREADS_W2 512
#endif
    static constexpr int CAP = (W <= 2) ? SEG_CAP_W2 : 3072;    // records one round of a CTA holds in shared memory (2 CTAs per SM)
    static constexpr int THREADS = (W <= 2) ? SEG_THREADS_W2 : 512;
    static constexpr int TARGET = (W <= 2) ? 7168 : 3584;       // average group size aimed for when choosing p
    static constexpr int MAX_DISTINCT = 64;                     // distinct values one segment may hold (two per lane)
};

template<int W>
__device__ __forceinline__ uint32_t group_composite(const uint64_t *r, const DigitSel &d) { return rs_composite<W>(r, d); }

struct ChunkRange {
    uint32_t s, e;      // records [s, e) of the group (s == e: empty)
};

// One thread per group: first record whose composite key is >= g (binary search; the array is grouped by composite key).
template<int W>
__global__ void __launch_bounds__(128) group_bounds_kernel(const uint64_t *__restrict__ recs, uint64_t n, DigitSel gk, uint32_t n_groups,
                                                          uint32_t *__restrict__ starts /* n_groups + 1 */) {
    uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g > n_groups) return;
    uint64_t lo = 0, hi = n;   // first index with composite >= g
    if (g == 0) hi = 0;
    if (g == n_groups) lo = n;
    while (lo < hi) {
        uint64_t mid = (lo + hi) >> 1;
        uint64_t r[W];
        load_rec<W>(recs, mid, r);
        if (rs_composite<W>(r, gk) < g) lo = mid + 1;
        else hi = mid;
    }
    starts[g] = (uint32_t) lo;
}

__global__ void group_ranges_kernel(const uint32_t *__restrict__ starts, uint32_t n_groups, ChunkRange *__restrict__ ranges) {
    uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n_groups) return;
    ChunkRange cr;
    cr.s = starts[g]; cr.e = starts[g + 1];
    ranges[g] = cr;
}

// Shared memory of group_chunk_kernel, in bytes
template<int W>
constexpr size_t seg_chunk_smem() {
    return (size_t) SegCfg<W>::CAP * W * 8      // the round's records sorted by digit; every segment then holds its unique records in order
           + (size_t) SegCfg<W>::CAP * 2         // multiplicity of the unique record at that position
           + 257 * 4                             // bin starts (exclusive scan of the digit histogram)
           + 256 * 4                             // per-bin cursors of the current round, later output offsets of the segments
           + 256 * 4                             // unique records per segment
           + 258 * 2;                            // first digit of every round
}

// a < b in record order (word-wise from word 0) without branches: 1 or 0
template<int W>
__device__ __forceinline__ uint32_t rec_less_bf(const uint64_t *a, const uint64_t *b) {
    bool lt = a[W - 1] < b[W - 1];
#pragma unroll
    for (int j = W - 2; j >= 0; --j) lt = (a[j] < b[j]) | ((a[j] == b[j]) & lt);
    return lt ? 1u : 0u;
}

// the 32 key bits right below bit `shift` of word 0 (fewer for very short k-mers): monotone in the record order inside a segment
__device__ __forceinline__ uint32_t seg_tag(uint64_t w0, int shift) {
    return shift >= 32 ? (uint32_t) (w0 >> (shift - 32)) : (uint32_t) (w0 << (32 - shift));
}

// Moves the unique records every group packed at the front of its own range to their final, contiguous place.
// One CTA per group; tile_off = exclusive scan of the per-group unique counts.
// MODE: 0 records only, 1 + multiplicities (uint32; double_pal_K = K doubles those of self-reverse-complement records), 2 + OR-ed
// mask payload (out_cnt points to bytes)
template<int W, int MODE>
__global__ void __launch_bounds__(256) seg_compact_kernel(const uint64_t *__restrict__ tmp, const uint32_t *__restrict__ tmp_cnt,
                                                         const ChunkRange *__restrict__ ranges, const uint32_t *__restrict__ tile_off,
                                                         uint32_t n_chunks, uint32_t total, uint64_t *__restrict__ out,
                                                         uint32_t *__restrict__ out_cnt, int double_pal_K = 0) {
    const uint32_t b = blockIdx.x;
    const uint32_t o = tile_off[b];
    const uint32_t cnt = (b + 1 < n_chunks ? tile_off[b + 1] : total) - o;
    const uint64_t src = ranges[b].s;
    for (uint32_t i = threadIdx.x; i < cnt; i += 256) {
        uint64_t r[W];
        load_rec<W>(tmp, src + i, r);
        store_rec<W>(out, (uint64_t) o + i, r);
        if (MODE == 1) {
            uint32_t c = tmp_cnt[src + i];
            if (double_pal_K) {   // a self-reverse-complement record is seen by both of the reference's streams: count it twice (count.cu)
                uint64_t rc[W];
                kmer_rc<W>(r, double_pal_K, rc);
                if (kmer_eq<W>(r, rc)) c *= 2;
            }
            out_cnt[o + i] = c;
        }
        if (MODE == 2) reinterpret_cast<uint8_t *>(out_cnt)[o + i] = (uint8_t) tmp_cnt[src + i];
    }
}

// dshift: the shared-memory digit is bits [dshift, dshift + 8) of word 0; the tag is the 32 bits below dshift.
constexpr uint32_t GROUP_MAX_ROUNDS_FACTOR = 32;   // groups beyond 32 x CAP records are left to the LSD fallback

// MODE 2: every record carries a 3-bit payload at bit `pshift` of its last word (the InOutMask bit it contributes, count.cu
// derive_kernel); equal records OR their bits, `lw_keep` strips the payload before any comparison.
template<int W, int MODE>
__global__ void __launch_bounds__(SegCfg<W>::THREADS) group_chunk_kernel(const uint64_t *__restrict__ recs, const ChunkRange *__restrict__ ranges,
                                                                        uint32_t *__restrict__ group_unique, uint32_t *__restrict__ fail_flag,
                                                                        uint64_t *__restrict__ out, uint32_t *__restrict__ out_cnt, int dshift,
                                                                        uint64_t lw_keep, int pshift) {
    constexpr int CAP = SegCfg<W>::CAP, THREADS = SegCfg<W>::THREADS;
    constexpr int NWARPS = THREADS / 32;
    static_assert(THREADS >= 256, "one thread per digit bin");
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint64_t *skey = reinterpret_cast<uint64_t *>(smem_raw);
    uint16_t *scnt = reinterpret_cast<uint16_t *>(skey + (size_t) CAP * W);
    uint32_t *bstart = reinterpret_cast<uint32_t *>(scnt + CAP);   // 257 entries
    uint32_t *cursor = bstart + 257;           // 256 entries
    uint32_t *ucnt = cursor + 256;             // 256 entries
    uint16_t *rdig = reinterpret_cast<uint16_t *>(ucnt + 256);   // 258 entries
    __shared__ uint32_t s_rounds, s_bad, s_next;
    __shared__ uint32_t s_scan[THREADS / 32 + 1];

    const uint32_t b = blockIdx.x;
    const ChunkRange cr = ranges[b];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t s = cr.s, gcnt = cr.e - cr.s;
    if (gcnt == 0) {
        if (threadIdx.x == 0) group_unique[b] = 0;
        return;
    }
    if (gcnt > GROUP_MAX_ROUNDS_FACTOR * (uint32_t) CAP) {
        if (threadIdx.x == 0) { group_unique[b] = 0; atomicExch(fail_flag, 1u); }
        return;
    }

    // ---- level 0a: digit histogram of the group, bin starts, split of the digit range into rounds of <= CAP records --------
    for (uint32_t i = threadIdx.x; i < 257; i += THREADS) bstart[i] = 0;
    __syncthreads();
    for (uint32_t q = threadIdx.x; q < gcnt; q += THREADS) {
        const uint64_t w0 = recs[((uint64_t) s + q) * W];
        atomicAdd(&bstart[(uint32_t) (w0 >> dshift) & 0xFFu], 1u);
    }
    __syncthreads();
    {
        const uint32_t c = threadIdx.x < 256 ? bstart[threadIdx.x] : 0u;
        uint32_t tot;
        const uint32_t ex = block_exclusive_scan<uint32_t, THREADS>(c, &tot, s_scan);
        if (threadIdx.x < 256) bstart[threadIdx.x] = ex;
        if (threadIdx.x == 0) bstart[256] = tot;
    }
    __syncthreads();
    if (threadIdx.x == 0 && gcnt <= (uint32_t) CAP) {   // one round
        rdig[0] = 0; rdig[1] = 256;
        s_rounds = 1;
        s_bad = 0;
    } else if (threadIdx.x == 0) {
        uint32_t nr = 0, first = 0, bad = 0;
        rdig[0] = 0;
        for (uint32_t d = 0; d < 256; ++d) {
            const uint32_t c = bstart[d + 1] - bstart[d];
            if (c > (uint32_t) CAP) bad = 1;
            if (bstart[d + 1] - bstart[first] > (uint32_t) CAP) { rdig[++nr] = (uint16_t) d; first = d; }
        }
        rdig[++nr] = 256;
        s_rounds = nr;
        s_bad = bad;
    }
    __syncthreads();
    if (s_bad) {
        if (threadIdx.x == 0) { group_unique[b] = 0; atomicExch(fail_flag, 1u); }
        return;
    }
    const uint32_t n_rounds = s_rounds;
    uint32_t emitted = 0;   // unique records written by the previous rounds (block-uniform)

    for (uint32_t round = 0; round < n_rounds; ++round) {
        const uint32_t d_lo = rdig[round], d_hi = rdig[round + 1];
        const uint32_t r_base = bstart[d_lo];
        const uint32_t cnt = bstart[d_hi] - r_base;
        __syncthreads();   // the previous round is done with every array
        if (cnt == 0) continue;
        // ---- level 0b: counting sort of this round's records into shared memory --------------------------------------------
        if (threadIdx.x < 256) { cursor[threadIdx.x] = 0; ucnt[threadIdx.x] = 0; }
        if (threadIdx.x == 0) s_next = d_lo;
        __syncthreads();
        for (uint32_t q0 = 0; q0 < gcnt; q0 += THREADS * 4) {   // four independent loads in flight per thread
            uint64_t in[4][W];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const uint32_t q = q0 + i * THREADS + threadIdx.x;
#pragma unroll
                for (int j = 0; j < W; ++j) in[i][j] = 0;
                if (q < gcnt) load_rec<W>(recs, (uint64_t) s + q, in[i]);
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const uint32_t q = q0 + i * THREADS + threadIdx.x;
                const uint32_t d = (uint32_t) (in[i][0] >> dshift) & 0xFFu;
                if (q < gcnt && d >= d_lo && d < d_hi) {
                    const uint32_t pos = bstart[d] - r_base + atomicAdd(&cursor[d], 1u);
#pragma unroll
                    for (int j = 0; j < W; ++j) skey[(size_t) pos * W + j] = in[i][j];
                }
            }
        }
        __syncthreads();

        // ---- level 1: one warp per segment: distinct values with multiplicities, in order, written back in place ------------
        // (segments are handed out dynamically: their cost varies with the number of distinct values)
        while (true) {
            uint32_t d = 0;
            if (lane == 0) d = atomicAdd(&s_next, 1u);
            d = __shfl_sync(0xffffffffu, d, 0);
            if (d >= d_hi) break;
            const uint32_t bs = bstart[d] - r_base, m = bstart[d + 1] - bstart[d];
            if (m == 0) continue;
            uint64_t keyA[W], keyB[W];   // distinct values of the segment: up to two per lane; VA / VB = lanes whose slot is taken
            uint32_t cA = 0, cB = 0, VA = 0, VB = 0;
            bool overflow = false;
#pragma unroll
            for (int j = 0; j < W; ++j) { keyA[j] = 0; keyB[j] = 0; }
            for (uint32_t t0 = 0; t0 < m; t0 += 32) {
                const uint32_t q = t0 + lane;
                const bool ok = q < m;
                uint64_t r[W];
#pragma unroll
                for (int j = 0; j < W; ++j) r[j] = ok ? skey[(size_t) (bs + q) * W + j] : 0ULL;
                const uint32_t pay = (MODE == 2) ? 1u << ((uint32_t) (r[W - 1] >> pshift) & 7u) : 0u;
                if (MODE == 2) r[W - 1] &= lw_keep;
                uint32_t peers = __ballot_sync(0xffffffffu, ok);
#pragma unroll
                for (int j = 0; j < W; ++j) peers &= __match_any_sync(0xffffffffu, r[j]);
                const bool is_rep = ok && ((uint32_t) lane == (uint32_t) (__ffs((int) peers) - 1));
                uint32_t mult = (uint32_t) __popc(peers);
                if (MODE == 2 && ok) mult = __reduce_or_sync(peers, pay);   // every lane of a peer group is ok and calls with the same mask
                uint32_t R = __ballot_sync(0xffffffffu, is_rep);
                if (t0 == 0) {   // first tile: its representatives are the list, each in its own lane
#pragma unroll
                    for (int j = 0; j < W; ++j) keyA[j] = r[j];
                    cA = mult;
                    VA = R;
                } else {         // later tiles of a long segment: merge every representative into the list
                    while (R) {
                        const int jl = __ffs((int) R) - 1;
                        R &= R - 1u;
                        uint64_t kj[W];
#pragma unroll
                        for (int j = 0; j < W; ++j) kj[j] = __shfl_sync(0xffffffffu, r[j], jl);
                        const uint32_t cj = __shfl_sync(0xffffffffu, mult, jl);
                        const uint32_t hitA = __ballot_sync(0xffffffffu, ((VA >> lane) & 1u) && kmer_eq<W>(keyA, kj));
                        const uint32_t hitB = VB ? __ballot_sync(0xffffffffu, ((VB >> lane) & 1u) && kmer_eq<W>(keyB, kj)) : 0u;
                        if (hitA) {
                            if (lane == __ffs((int) hitA) - 1) cA = (MODE == 2) ? (cA | cj) : (cA + cj);
                        } else if (hitB) {
                            if (lane == __ffs((int) hitB) - 1) cB = (MODE == 2) ? (cB | cj) : (cB + cj);
                        } else if (~VA) {
                            const int f = __ffs((int) ~VA) - 1;
                            if (lane == f) {
#pragma unroll
                                for (int j = 0; j < W; ++j) keyA[j] = kj[j];
                                cA = cj;
                            }
                            VA |= 1u << f;
                        } else if (~VB) {
                            const int f = __ffs((int) ~VB) - 1;
                            if (lane == f) {
#pragma unroll
                                for (int j = 0; j < W; ++j) keyB[j] = kj[j];
                                cB = cj;
                            }
                            VB |= 1u << f;
                        } else {
                            overflow = true;
                        }
                    }
                }
            }
            if (overflow) {   // more than 64 distinct values in one segment: leave the set to the LSD fallback
                if (lane == 0) atomicExch(fail_flag, 1u);
                continue;
            }
            // rank every entry among the entries (all distinct): number of smaller values.  Fast path (no second slot, and the 32 key
            // bits right below the digit — monotone in the record order inside a segment — tell all entries apart, which fails only
            // when two k-mers of one bucket agree on the whole of word 0's upper part): one 32-bit shuffle + compare per entry.
            uint32_t lessA = 0, lessB = 0;
            const uint32_t tagA = seg_tag(keyA[0], dshift);
            bool by_tag = (VB == 0u);
            if (by_tag) {
                const uint32_t same = __match_any_sync(0xffffffffu, tagA) & VA;
                by_tag = !__any_sync(0xffffffffu, ((VA >> lane) & 1u) && same != (1u << lane));
            }
            if (by_tag) {
                for (uint32_t E = VA; E; E &= E - 1u) {
                    const uint32_t te = __shfl_sync(0xffffffffu, tagA, __ffs((int) E) - 1);
                    lessA += (te < tagA) ? 1u : 0u;
                }
            } else {
            for (uint32_t E = VA; E; E &= E - 1u) {
                const int e = __ffs((int) E) - 1;
                uint64_t ke[W];
#pragma unroll
                for (int j = 0; j < W; ++j) ke[j] = __shfl_sync(0xffffffffu, keyA[j], e);
                lessA += rec_less_bf<W>(ke, keyA);
                if (VB) lessB += rec_less_bf<W>(ke, keyB);
            }
            for (uint32_t E = VB; E; E &= E - 1u) {
                const int e = __ffs((int) E) - 1;
                uint64_t ke[W];
#pragma unroll
                for (int j = 0; j < W; ++j) ke[j] = __shfl_sync(0xffffffffu, keyB[j], e);
                lessA += rec_less_bf<W>(ke, keyA);
                lessB += rec_less_bf<W>(ke, keyB);
            }
            }
            __syncwarp();
            if ((VA >> lane) & 1u) {
#pragma unroll
                for (int j = 0; j < W; ++j) skey[(size_t) (bs + lessA) * W + j] = keyA[j];
                scnt[bs + lessA] = (uint16_t) cA;
            }
            if ((VB >> lane) & 1u) {
#pragma unroll
                for (int j = 0; j < W; ++j) skey[(size_t) (bs + lessB) * W + j] = keyB[j];
                scnt[bs + lessB] = (uint16_t) cB;
            }
            if (lane == 0) ucnt[d] = (uint32_t) (__popc(VA) + __popc(VB));
        }
        __syncthreads();

        // ---- level 2: output offsets of the segments, unique records out ---------------------------------------------------
        uint32_t total_u;
        {
            const uint32_t c = (threadIdx.x >= d_lo && threadIdx.x < d_hi) ? ucnt[threadIdx.x] : 0u;
            const uint32_t ex = block_exclusive_scan<uint32_t, THREADS>(c, &total_u, s_scan);
            if (threadIdx.x < 256) cursor[threadIdx.x] = ex;
        }
        __syncthreads();
        const unsigned long long base = (unsigned long long) s + emitted;   // front of the group's own range, after the earlier rounds
        for (uint32_t d = d_lo + warp; d < d_hi; d += NWARPS) {
            const uint32_t u = ucnt[d];
            const uint32_t bs = bstart[d] - r_base;
            const unsigned long long dst = base + cursor[d];
            for (uint32_t i = lane; i < u; i += 32) {
                uint64_t r[W];
#pragma unroll
                for (int j = 0; j < W; ++j) r[j] = skey[(size_t) (bs + i) * W + j];
                store_rec<W>(out, dst + i, r);
                if (MODE != 0) out_cnt[dst + i] = scnt[bs + i];
            }
        }
        emitted += total_u;
    }   // rounds
    if (threadIdx.x == 0) group_unique[b] = emitted;
}

}  // namespace sb200
