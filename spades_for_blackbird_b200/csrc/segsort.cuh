// segsort.cuh — sort + deduplicate + count in ONE pass over records that are already grouped by a short key prefix.
//
// After three stable counting passes (two bytes of the most significant record word, then the hash bucket) the
// instance array is ordered by (bucket, 16-bit value prefix).  A "segment" is a maximal run of records sharing that
// prefix: a few dozen records for the BASELINE workloads.  Finishing the order with twelve more full-array LSD passes
// (what round-1's first version did) moves every record 12 more times through HBM; instead each CTA takes a run of
// whole segments (<= CAP records) and
//   (1) deduplicates every warp tile of 32 consecutive records with match.any on the record words — equal records elect
//       their lowest lane, which keeps the value and the number of copies; the bulk of the input (copies of the same
//       genomic k-mer) disappears here at a cost of W match instructions per 32 records;
//   (2) ranks the survivors (kept in order in shared memory, hence still grouped by segment) inside their segment by
//       direct comparison of a 32-bit tag (the key bits right below the prefix), falling back to the full key on equal
//       tags; equal survivors from different tiles merge their counts;
//   (3) compacts "first of its value" flags in sorted position with a block scan and obtains its global offset from a
//       decoupled look-back over per-CTA unique counts, so the array is read once and only unique records are written.
// Segments too long for shared memory (highly repeated k-mers, low-complexity sequence) make their CTA "dirty"; dirty
// ranges are sorted by the generic LSD path first and the CTA merely copies the result.
#pragma once
#include "common.cuh"
#include "kmer_ops.cuh"
#include "radix_sort.cuh"
#include "scan.cuh"

namespace sb200 {

template<int W> struct SegCfg {
    static constexpr int CAP = (W == 1) ? 8192 : (W == 2) ? 4096 : 2048;   // records per CTA (64 KB of keys)
    static constexpr int C = CAP / 2;        // a CTA owns the segments that START in its C-record window
    static constexpr int MAXSEG = CAP / 2;   // longest segment the shared-memory path accepts
    static constexpr int THREADS = 512;
};

struct PrefixKey {   // what defines a segment
    int shift;           // prefix = (word0 >> shift) & 0xFFFF
    uint32_t num_buckets;
    int marker;
};

template<int W>
__device__ __forceinline__ uint64_t seg_key(const uint64_t *r, const PrefixKey &pk) {
    uint32_t b = kmer_bucket<W>(r, pk.num_buckets);
    if (pk.marker) {
        bool m = true;
#pragma unroll
        for (int j = 0; j < W; ++j) m &= (r[j] == ~0ULL);
        if (m) b = pk.num_buckets - 1;
    }
    return ((uint64_t) b << 16) | ((r[0] >> pk.shift) & 0xFFFFu);
}

// head bit i = record i starts a segment.  One warp writes one 32-bit word.
template<int W>
__global__ void __launch_bounds__(256) seg_heads_kernel(const uint64_t *__restrict__ recs, uint64_t n, PrefixKey pk, uint32_t *__restrict__ hb) {
    uint64_t i = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
    int lane = threadIdx.x & 31;
    uint64_t key = ~0ULL;
    if (i < n) {
        uint64_t r[W];
        load_rec<W>(recs, i, r);
        key = seg_key<W>(r, pk);
    }
    uint64_t prev = __shfl_up_sync(0xffffffffu, key, 1);
    if (lane == 0) {
        prev = ~0ULL;
        if (i > 0 && i < n) {
            uint64_t r[W];
            load_rec<W>(recs, i - 1, r);
            prev = seg_key<W>(r, pk);
        }
    }
    bool head = (i < n) && (i == 0 || key != prev);
    uint32_t word = __ballot_sync(0xffffffffu, head);
    if (lane == 0 && (i >> 5) < ((n + 31) >> 5)) hb[i >> 5] = word;
}

// first head position >= from (searching bits up to `limit`, exclusive); returns limit if none
__device__ __forceinline__ uint64_t next_head(const uint32_t *__restrict__ hb, uint64_t from, uint64_t limit) {
    if (from >= limit) return limit;
    uint64_t w = from >> 5;
    uint32_t bits = hb[w] & (0xFFFFFFFFu << (from & 31));
    uint64_t last_w = (limit - 1) >> 5;
    while (true) {
        if (bits) {
            uint64_t p = (w << 5) + (__ffs((int) bits) - 1);
            return p < limit ? p : limit;
        }
        if (w == last_w) return limit;
        ++w;
        bits = hb[w];
    }
}

struct ChunkRange {
    uint32_t s, e;      // records [s, e) = all segments that start inside the CTA's window (s == e: owns nothing)
    uint32_t dirty;     // 1: some owned segment is longer than MAXSEG -> handled by the LSD fallback
    uint32_t side_off;  // dirty only: where the pre-sorted unique records sit in the side buffer
    uint32_t side_cnt;  // dirty only: how many
};

// one thread per CTA window: owned range and whether every owned segment fits
template<int W>
__global__ void seg_ranges_kernel(const uint32_t *__restrict__ hb, uint64_t n, uint32_t n_chunks, ChunkRange *__restrict__ ranges,
                                  uint32_t *__restrict__ dirty_list, uint32_t *__restrict__ n_dirty) {
    constexpr int C = SegCfg<W>::C, MAXSEG = SegCfg<W>::MAXSEG;
    uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= n_chunks) return;
    uint64_t lo = (uint64_t) b * C, hi = lo + C < n ? lo + C : n;
    ChunkRange cr;
    cr.dirty = 0; cr.side_off = 0; cr.side_cnt = 0;
    uint64_t s = next_head(hb, lo, hi);
    if (s >= hi) {   // no segment starts here
        cr.s = cr.e = (uint32_t) hi;
        ranges[b] = cr;
        return;
    }
    // walk the owned heads, checking segment lengths
    bool dirty = false;
    uint64_t cur = s, e = hi;
    while (true) {
        uint64_t limit = cur + 1 + MAXSEG < n ? cur + 1 + MAXSEG : n;
        uint64_t nx = next_head(hb, cur + 1, limit);
        if (nx == limit && limit < n) {   // no head within MAXSEG records: oversize segment
            dirty = true;
            nx = next_head(hb, limit, n);   // its true end (unbounded scan; rare)
        } else if (nx == limit && limit == n) {
            if (n - cur > (uint64_t) MAXSEG) dirty = true;
            nx = n;
        }
        if (nx >= hi) { e = nx; break; }
        cur = nx;
    }
    cr.s = (uint32_t) s; cr.e = (uint32_t) e; cr.dirty = dirty ? 1u : 0u;
    ranges[b] = cr;
    if (dirty) dirty_list[atomicAdd(n_dirty, 1u)] = b;
}

// decoupled look-back status word: bits 62-63 flag (0 empty, 1 aggregate, 2 inclusive prefix), bits 0-61 value
__device__ __forceinline__ unsigned long long lb_pack(unsigned long long flag, unsigned long long v) { return (flag << 62) | v; }

// Shared memory of seg_chunk_kernel, in bytes
template<int W>
constexpr size_t seg_chunk_smem() {
    return (size_t) SegCfg<W>::CAP * W * 8      // keys of the warp-tile survivors
           + (size_t) SegCfg<W>::CAP * 4         // their 32-bit tags
           + (size_t) SegCfg<W>::CAP * 2         // owner[s]: 1 + survivor index whose value is first at sorted position s (0 = none)
           + (size_t) SegCfg<W>::CAP * 2         // rcnt[p]: total multiplicity of survivor p's value (<= MAXSEG)
           + (size_t) SegCfg<W>::CAP             // bit 7: survivor starts a segment; bits 0-5: copies merged in its tile (<= 32)
           + (size_t) (SegCfg<W>::CAP / 32 + 1) * 4;   // survivors per warp tile, then its exclusive scan
}

// the 32 key bits right below the 16-bit prefix (fewer for very short k-mers): monotone in the record order inside a segment
__device__ __forceinline__ uint32_t seg_tag(uint64_t w0, int shift) {
    return shift >= 32 ? (uint32_t) (w0 >> (shift - 32)) : (uint32_t) (w0 << (32 - shift));
}

// Moves the unique records every tile packed at the front of its own range to their final, contiguous place.
// One CTA per tile; tile_off = exclusive scan of the per-tile unique counts.
template<int W, bool COUNTS>
__global__ void __launch_bounds__(256) seg_compact_kernel(const uint64_t *__restrict__ tmp, const uint32_t *__restrict__ tmp_cnt,
                                                         const ChunkRange *__restrict__ ranges, const uint32_t *__restrict__ tile_off,
                                                         uint32_t n_chunks, uint32_t total, uint64_t *__restrict__ out,
                                                         uint32_t *__restrict__ out_cnt) {
    const uint32_t b = blockIdx.x;
    const uint32_t o = tile_off[b];
    const uint32_t cnt = (b + 1 < n_chunks ? tile_off[b + 1] : total) - o;
    const uint64_t src = ranges[b].s;
    for (uint32_t i = threadIdx.x; i < cnt; i += 256) {
        uint64_t r[W];
        load_rec<W>(tmp, src + i, r);
        store_rec<W>(out, (uint64_t) o + i, r);
        if (COUNTS) out_cnt[o + i] = tmp_cnt[src + i];
    }
}

template<int W, bool COUNTS, bool USE_LOOKBACK = false>
__global__ void __launch_bounds__(SegCfg<W>::THREADS) seg_chunk_kernel(const uint64_t *__restrict__ recs, uint64_t n,
                                                                      const uint32_t *__restrict__ hb, const ChunkRange *__restrict__ ranges,
                                                                      const uint64_t *__restrict__ side_recs, const uint32_t *__restrict__ side_cnts,
                                                                      unsigned long long *__restrict__ status, uint32_t *__restrict__ tile_counter,
                                                                      uint64_t *__restrict__ out, uint32_t *__restrict__ out_cnt,
                                                                      unsigned long long *__restrict__ total_out, uint32_t n_chunks, int shift) {
    constexpr int CAP = SegCfg<W>::CAP, THREADS = SegCfg<W>::THREADS;
    constexpr int NT = CAP / 32;               // warp tiles per CTA
    constexpr int TPW = NT / (THREADS / 32);   // warp tiles per warp
    constexpr uint32_t PER = CAP / THREADS;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint64_t *pkey = reinterpret_cast<uint64_t *>(smem_raw);
    uint32_t *ptag = reinterpret_cast<uint32_t *>(pkey + (size_t) CAP * W);
    uint16_t *owner = reinterpret_cast<uint16_t *>(ptag + CAP);
    uint16_t *rcnt = owner + CAP;
    uint8_t *pinfo = reinterpret_cast<uint8_t *>(rcnt + CAP);
    uint32_t *tileoff = reinterpret_cast<uint32_t *>(pinfo + CAP);
    __shared__ uint32_t s_tile;
    __shared__ uint32_t s_np;
    __shared__ unsigned long long s_base;
    __shared__ uint32_t s_scan[THREADS / 32 + 1];

    // USE_LOOKBACK: CTAs take tiles in launch order and chain their unique counts with a decoupled look-back (single pass,
    // final positions written directly).  Otherwise (default) every CTA packs its unique records at the front of its own
    // input range and records how many; seg_compact_kernel moves them to their final place after a scan of the counts —
    // that copy replaces the right-sizing copy the host did anyway and costs less than the look-back wait (30 % of the
    // stall samples of the single-pass version: fifteen warps idle while one chases predecessors through L2).
    if (USE_LOOKBACK) {
        if (threadIdx.x == 0) s_tile = atomicAdd(tile_counter, 1u);
        __syncthreads();
    }
    const uint32_t b = USE_LOOKBACK ? s_tile : blockIdx.x;
    if (b >= n_chunks) return;
    const ChunkRange cr = ranges[b];
    const uint32_t s = cr.s, cnt = cr.dirty ? 0u : cr.e - cr.s;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;

    // ---- level 1: warp-tile deduplication, records straight from global memory into registers ---------------------------
    uint64_t rec[TPW][W];
    uint32_t mult[TPW];
    uint32_t rep_bits = 0, head_bits = 0;   // bit t: my record of tile t is a tile survivor / starts a segment
#pragma unroll
    for (int t = 0; t < TPW; ++t) {
        const uint32_t p = ((uint32_t) t * (THREADS / 32) + warp) * 32 + lane;   // tiles round-robin over the warps: a typical range fills only half of CAP
#pragma unroll
        for (int j = 0; j < W; ++j) rec[t][j] = 0;
        if (p < cnt) {
            load_rec<W>(recs, (uint64_t) s + p, rec[t]);
            const uint64_t g = (uint64_t) s + p;
            if ((__ldg(hb + (g >> 5)) >> (g & 31)) & 1u) head_bits |= 1u << t;
        }
    }
#pragma unroll
    for (int t = 0; t < TPW; ++t) {
        const uint32_t tile = (uint32_t) t * (THREADS / 32) + warp;
        const uint32_t p = tile * 32 + lane;
        const bool ok = p < cnt;
        uint32_t peers = 0xFFFFFFFFu;
        if (tile * 32 < cnt) {   // warp-uniform
#pragma unroll
            for (int j = 0; j < W; ++j) peers &= __match_any_sync(0xffffffffu, rec[t][j]);
            peers &= __ballot_sync(0xffffffffu, ok);
        }
        const bool is_rep = ok && ((uint32_t) lane == (uint32_t) (__ffs((int) peers) - 1));
        mult[t] = (uint32_t) __popc(peers);
        if (is_rep) rep_bits |= 1u << t;
        const uint32_t m = __ballot_sync(0xffffffffu, is_rep);
        if (lane == 0) tileoff[tile] = (uint32_t) __popc(m);
    }
    __syncthreads();
    {   // exclusive scan of the NT (<= 256) tile counts by the whole block
        static_assert(NT <= THREADS, "one thread per warp tile");
        const uint32_t c = threadIdx.x < NT ? tileoff[threadIdx.x] : 0u;
        uint32_t total_np;
        const uint32_t ex = block_exclusive_scan<uint32_t, THREADS>(c, &total_np, s_scan);
        if (threadIdx.x < NT) tileoff[threadIdx.x] = ex;
        if (threadIdx.x == 0) s_np = total_np;
    }
    __syncthreads();
    const uint32_t np = s_np;
#pragma unroll
    for (int t = 0; t < TPW; ++t) {
        const uint32_t tile = (uint32_t) t * (THREADS / 32) + warp;
        const bool is_rep = (rep_bits >> t) & 1u;
        const uint32_t m = __ballot_sync(0xffffffffu, is_rep);
        if (is_rep) {
            const uint32_t o = tileoff[tile] + (uint32_t) __popc(m & ((1u << lane) - 1u));
#pragma unroll
            for (int j = 0; j < W; ++j) pkey[(size_t) o * W + j] = rec[t][j];
            ptag[o] = seg_tag(rec[t][0], shift);
            pinfo[o] = (uint8_t) ((((head_bits >> t) & 1u) << 7) | (mult[t] - 1u));   // copies-1 fits 5 bits
        }
    }
    for (uint32_t i = threadIdx.x; i < np; i += THREADS) owner[i] = 0;
    __syncthreads();

    // ---- level 2: rank the survivors inside their segment ------------------------------------------------------------------
    // One warp per run of whole segments, one lane per survivor: every lane walks the segment once and all lanes read the
    // same tag per step (a shared-memory broadcast), so the trip count is uniform across the warp.  (The first version gave
    // each THREAD a survivor and let it scan its own segment: lanes of a warp sat in segments of different lengths and 29 %
    // of the kernel's stall samples were the barrier after the loop, profiles/r1b_seg_chunk_source.md.)
    {
        constexpr uint32_t NW = THREADS / 32;
        auto next_head = [&](uint32_t from) -> uint32_t {   // warp-uniform: first survivor >= from that starts a segment, else np
            for (uint32_t bb = from; bb < np; bb += 32) {
                const uint32_t q = bb + lane;
                const uint32_t hm = __ballot_sync(0xffffffffu, q < np && (pinfo[q] & 0x80u));
                if (hm) return bb + (uint32_t) __ffs((int) hm) - 1u;
            }
            return np;
        };
        // survivors [np*w/NW, np*(w+1)/NW) snapped forward to segment heads: the ranges tile [0, np) exactly
        uint32_t sb = next_head((uint32_t) ((uint64_t) np * warp / NW));
        const uint32_t send = (warp == (int) NW - 1) ? np : next_head((uint32_t) ((uint64_t) np * (warp + 1) / NW));
        while (sb < send) {
            const uint32_t se = next_head(sb + 1);
            for (uint32_t cb = sb; cb < se; cb += 32) {
                const uint32_t p = cb + lane;
                const bool act = p < se;
                const uint32_t mytag = act ? ptag[p] : 0u;
                uint64_t me[W];
#pragma unroll
                for (int j = 0; j < W; ++j) me[j] = act ? pkey[(size_t) p * W + j] : 0ULL;
                uint32_t less = 0, eq_before = 0, total = act ? (uint32_t) (pinfo[p] & 0x3Fu) + 1u : 0u;
                // Branch-free over blocks of 32 candidates: count the tags below mine and collect the positions whose tag
                // equals mine (all lanes read the same tag: a broadcast; the loads of a block are independent).  Only those
                // positions — copies of my value that survived in other tiles, or a true 48-bit-prefix tie — need the full key.
                for (uint32_t qb = sb; qb < se; qb += 32) {
                    const uint32_t nq = se - qb < 32u ? se - qb : 32u;
                    uint32_t eqm = 0;
                    uint32_t j = 0;
                    for (; j + 4 <= nq; j += 4) {
                        const uint32_t t0 = ptag[qb + j], t1 = ptag[qb + j + 1], t2 = ptag[qb + j + 2], t3 = ptag[qb + j + 3];
                        less += (uint32_t) (t0 < mytag) + (uint32_t) (t1 < mytag) + (uint32_t) (t2 < mytag) + (uint32_t) (t3 < mytag);
                        eqm |= ((uint32_t) (t0 == mytag) | ((uint32_t) (t1 == mytag) << 1) | ((uint32_t) (t2 == mytag) << 2) |
                                ((uint32_t) (t3 == mytag) << 3)) << j;
                    }
                    for (; j < nq; ++j) {
                        const uint32_t t = ptag[qb + j];
                        less += (uint32_t) (t < mytag);
                        eqm |= (uint32_t) (t == mytag) << j;
                    }
                    if (!act) eqm = 0;
                    else if (p - qb < 32u) eqm &= ~(1u << (p - qb));   // myself
                    while (eqm) {
                        const uint32_t q = qb + (uint32_t) __ffs((int) eqm) - 1u;
                        eqm &= eqm - 1u;
                        uint64_t o[W];
#pragma unroll
                        for (int jj = 0; jj < W; ++jj) o[jj] = pkey[(size_t) q * W + jj];
                        if (kmer_eq<W>(o, me)) {
                            eq_before += (q < p) ? 1u : 0u;
                            total += (uint32_t) (pinfo[q] & 0x3Fu) + 1u;
                        } else if (rec_less<W>(o, me)) {
                            ++less;
                        }
                    }
                }
                if (act && eq_before == 0) {
                    owner[sb + less] = (uint16_t) (p + 1u);
                    rcnt[p] = (uint16_t) total;
                }
            }
            sb = se;
        }
    }
    __syncthreads();
    // compact the occupied sorted positions (blocked: thread t owns positions [t*PER, (t+1)*PER))
    uint32_t own[PER];
    uint32_t local_sum = 0;
#pragma unroll
    for (uint32_t i = 0; i < PER; ++i) {
        const uint32_t sp = threadIdx.x * PER + i;
        own[i] = (sp < np) ? owner[sp] : 0u;
        local_sum += own[i] ? 1u : 0u;
    }
    uint32_t total_u;
    uint32_t pre = block_exclusive_scan<uint32_t, THREADS>(local_sum, &total_u, s_scan);
    const uint32_t my_total = cr.dirty ? cr.side_cnt : total_u;

    // ---- decoupled look-back for this CTA's global offset ------------------------------------------------------------------
    // One warp inspects 32 predecessors per step (a single thread walking them one L2 round trip at a time was 45 % of
    // this kernel's stall samples): sum the aggregates down to the nearest CTA that already knows its inclusive prefix.
    if (!USE_LOOKBACK) {
        if (threadIdx.x == 0) {
            reinterpret_cast<uint32_t *>(status)[b] = my_total;   // per-tile unique count (status doubles as the count array)
            s_base = cr.s;                                        // front of my own input range
        }
    } else if (warp == 0) {
        unsigned long long excl = 0;
        if (b > 0) {
            if (lane == 0) atomicExch(&status[b], lb_pack(1, my_total));
            long long pos = (long long) b - 1;
            while (true) {
                const long long p = pos - lane;
                unsigned long long sv = lb_pack(2, 0);   // before the first CTA: inclusive prefix 0
                if (p >= 0) sv = *reinterpret_cast<volatile unsigned long long *>(status + p);
                const unsigned long long flag = sv >> 62;
                const uint32_t empty = __ballot_sync(0xffffffffu, flag == 0);
                const uint32_t incl = __ballot_sync(0xffffffffu, flag == 2);
                const uint32_t need = incl ? (0xFFFFFFFFu >> (31 - (__ffs((int) incl) - 1))) : 0xFFFFFFFFu;   // lanes up to the first inclusive
                if (empty & need) continue;   // a predecessor we depend on has not published yet: read again
                unsigned long long v = ((need >> lane) & 1u) ? (sv & ((1ULL << 62) - 1)) : 0ULL;
#pragma unroll
                for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
                excl += v;
                if (incl) break;
                pos -= 32;
            }
        }
        if (lane == 0) {
            atomicExch(&status[b], lb_pack(2, excl + my_total));
            s_base = excl;
            if (b == n_chunks - 1) *total_out = excl + my_total;
        }
    }
    __syncthreads();
    const unsigned long long base = s_base;

    if (cr.dirty) {
        for (uint32_t i = threadIdx.x; i < cr.side_cnt; i += THREADS) {
            uint64_t r[W];
            load_rec<W>(side_recs, (uint64_t) cr.side_off + i, r);
            store_rec<W>(out, base + i, r);
            if (COUNTS) out_cnt[base + i] = side_cnts[cr.side_off + i];
        }
        return;
    }
#pragma unroll
    for (uint32_t i = 0; i < PER; ++i) {
        if (!own[i]) continue;
        const uint32_t p = own[i] - 1u;
        uint64_t r[W];
#pragma unroll
        for (int j = 0; j < W; ++j) r[j] = pkey[(size_t) p * W + j];
        const unsigned long long dst = base + pre;
        store_rec<W>(out, dst, r);
        if (COUNTS) out_cnt[dst] = rcnt[p];
        ++pre;
    }
}

}  // namespace sb200
