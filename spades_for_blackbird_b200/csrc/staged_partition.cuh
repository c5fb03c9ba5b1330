// staged_partition.cuh — producer-fused grouping of k-mer instances in two staged passes.
//
// Round 1 materialised every instance (extract / derive kernel: one write), then grouped the array with two counting passes, each a
// histogram kernel (one read) and a scatter kernel (one read + one write): the instances crossed HBM eight times before the group
// kernel saw them.  A one-pass scatter through L2 atomics (partition.cuh) moves them only twice but issues one L2 request per RECORD,
// and the L2's request rate — not its bandwidth — bounds it (profiles/r2a_ubench.log: 291 M scattered 16-byte stores + atomics = 7.4 ms).
// Here the producer is the first pass and every trip to L2 is a burst:
//   count   the producer (windows of reads / candidates of (k+1)-mers) runs once without storing anything: every CTA keeps the histogram
//           of the FINE group key  g = bucket << p | top p value bits  in shared memory (n_groups <= 49152 counters) and adds it to the
//           global histogram at the end.  One scan gives the group starts — the table the group kernel needs — and, read at a stride of
//           2^s groups, the starts of the COARSE bins  c = g >> s.
//   pass 1  the producer runs again; a CTA collects a tile of records in shared memory, orders the tile by coarse bin (shared-memory
//           histogram, scan, ranks), claims a run per non-empty bin with ONE global atomic and writes the run as a coalesced burst.
//   pass 2  tiles of one coarse bin are loaded with bulk asynchronous copies (cp.async.bulk global -> shared, mbarrier completion,
//           double buffered), ordered by the fine group inside the bin (2^s local bins) and written out the same way.
// The instances cross HBM four times (written by pass 1, read + written by pass 2, read by the group kernel); no pass reads a histogram
// matrix, and no pass makes a per-record request to L2.  Record order inside a group is arbitrary (the group kernel hashes).
// Reference stages replaced: KMerSortingSplitter's per-thread x per-bucket cells and DumpBuffers (C/utils/kmer_mph/kmer_splitter.hpp:
// 73-167), DeBruijnReadKMerSplitter / DeBruijnKMerKMerSplitter producers (kmer_splitters.hpp:25-59,159-204).
#pragma once
#include "common.cuh"
#include "kmer_ops.cuh"
#include "partition.cuh"
#include "radix_sort.cuh"
#include "scan.cuh"

namespace sb200 {

constexpr int SP_THREADS = 512;
constexpr int SP_WARPS = SP_THREADS / 32;
constexpr int SP_MAX_BINS = 1024;             // local bins of a tile (coarse bins in pass 1, fine groups of one coarse bin in pass 2)
constexpr uint32_t SP_MAX_GROUPS = 49152;     // fine groups the count kernels keep in shared memory (192 KB of counters)
constexpr int SPC_THREADS = 1024;             // count kernels: one CTA per SM (the histogram fills the shared memory), 32 warps
template<int W> struct SpCfg {
    static constexpr int CAP = (W <= 2) ? 4608 : 2304;   // records of a pass-1 tile: 72 KB of records + ordering arrays, two CTAs per SM
    static constexpr int CAP2 = (W <= 2) ? 2048 : 1024;  // records of a pass-2 tile (two bulk-copy buffers of 32 KB), two CTAs per SM
};

// ---- the shared-memory stage of one tile -----------------------------------------------------------------------------------------
struct SpStage {
    uint64_t *rec;       // CAP x W words, arrival order
    uint16_t *bin;       // local bin of every record
    uint16_t *src;       // src[d] = arrival slot of the record that leaves at position d of the bin-ordered tile
    uint8_t *pay;        // mask-bit payload beside the record (or nullptr)
    uint32_t *bstart;    // SP_MAX_BINS + 1: first position of every bin in the ordered tile
    uint32_t *cursor;    // SP_MAX_BINS: records of the bin placed so far (= its size after placement)
    uint32_t *gdelta;    // SP_MAX_BINS: global position of the bin's run minus its position in the tile
};

template<int W>
constexpr size_t sp_stage_bytes(int cap, bool with_pay) {
    return (size_t) cap * W * 8 + (size_t) cap * 2 * 2 + (with_pay ? (size_t) cap : 0) + (size_t) (3 * SP_MAX_BINS + 1) * 4 + 64;
}

template<int W>
__device__ __forceinline__ SpStage sp_carve(unsigned char *raw, int cap, bool with_pay) {
    SpStage s;
    s.rec = reinterpret_cast<uint64_t *>(raw);
    unsigned char *p = raw + (size_t) cap * W * 8;
    s.bin = reinterpret_cast<uint16_t *>(p); p += (size_t) cap * 2;
    s.src = reinterpret_cast<uint16_t *>(p); p += (size_t) cap * 2;
    s.pay = with_pay ? p : nullptr;
    if (with_pay) p += (size_t) cap;
    p = reinterpret_cast<unsigned char *>(((uintptr_t) p + 15) & ~(uintptr_t) 15);
    s.bstart = reinterpret_cast<uint32_t *>(p);
    s.cursor = s.bstart + SP_MAX_BINS + 1;
    s.gdelta = s.cursor + SP_MAX_BINS;
    return s;
}

// Before a tile is filled: clears the bin counts (all threads; a barrier must follow before the first sp_put)
__device__ __forceinline__ void sp_stage_reset(const SpStage &st, uint32_t n_bins) {
    for (uint32_t b = threadIdx.x; b <= n_bins; b += SP_THREADS) st.bstart[b] = 0;
}

// (SpPeers — the owners' receive buffers the PEER instances store through — is declared in kmer_set.cuh)
// Orders the n staged records by local bin and writes every bin's run to out[] at a position claimed from gcursor[first_bin + bin].
// All threads of the CTA call it after a barrier that follows the last sp_put; the stage may be reset and refilled when it returns.
template<int W, bool PEER = false>
__device__ __forceinline__ void sp_flush(const SpStage &st, uint32_t n, uint32_t n_bins, uint32_t *__restrict__ gcursor, uint32_t first_bin,
                                         uint64_t *__restrict__ out, uint8_t *__restrict__ out_pay, uint32_t *s_scan,
                                         const SpPeers *peers = nullptr, uint8_t *s_own = nullptr) {
    // (the bin counts were accumulated in st.bstart while the tile was filled: sp_put / sp_stage_reset)
    {   // exclusive scan of up to 1024 bin counts: two per thread
        const uint32_t b0 = 2u * threadIdx.x, b1 = b0 + 1u;
        const uint32_t c0 = b0 < n_bins ? st.bstart[b0] : 0u, c1 = b1 < n_bins ? st.bstart[b1] : 0u;
        uint32_t total;
        const uint32_t ex = block_exclusive_scan<uint32_t, SP_THREADS>(c0 + c1, &total, s_scan);
        if (b0 < n_bins) { st.bstart[b0] = ex; st.cursor[b0] = 0; }
        if (b1 < n_bins) { st.bstart[b1] = ex + c0; st.cursor[b1] = 0; }
        if (threadIdx.x == 0) st.bstart[n_bins] = total;
    }
    __syncthreads();
    for (uint32_t q = threadIdx.x; q < n; q += SP_THREADS) {
        const uint32_t b = st.bin[q];
        st.src[st.bstart[b] + atomicAdd(&st.cursor[b], 1u)] = (uint16_t) q;
    }
    __syncthreads();
    for (uint32_t b = threadIdx.x; b < n_bins; b += SP_THREADS) {   // one global atomic per non-empty bin of the tile
        const uint32_t c = st.cursor[b];
        if (c) {
            st.gdelta[b] = atomicAdd(&gcursor[first_bin + b], c) - st.bstart[b];
            if (PEER) s_own[b] = (uint8_t) ((first_bin + b) / peers->n_co);
        }
    }
    __syncthreads();
    for (uint32_t d = threadIdx.x; d < n; d += SP_THREADS) {        // consecutive d of one bin -> consecutive records in HBM
        const uint32_t q = st.src[d];
        const uint32_t g = st.gdelta[st.bin[q]] + d;
        uint64_t r[W];
        if (W == 2) {
            const ulonglong2 v = *reinterpret_cast<const ulonglong2 *>(st.rec + (size_t) q * 2);
            r[0] = v.x; r[1] = v.y;
        } else if (W == 4) {
            const ulonglong2 *p = reinterpret_cast<const ulonglong2 *>(st.rec + (size_t) q * 4);
            const ulonglong2 a = p[0], c = p[1];
            r[0] = a.x; r[1] = a.y; r[2] = c.x; r[3] = c.y;
        } else {
#pragma unroll
            for (int j = 0; j < W; ++j) r[j] = st.rec[(size_t) q * W + j];
        }
        if (PEER) {
            const uint32_t o = s_own[st.bin[q]];
            store_rec<W>(peers->rec[o], g, r);
            if (st.pay) peers->pay[o][g] = st.pay[q];
        } else {
            store_rec<W>(out, g, r);
            if (out_pay) out_pay[g] = st.pay[q];
        }
    }
    __syncthreads();
}

template<int W>
__device__ __forceinline__ void sp_put(const SpStage &st, uint32_t slot, const uint64_t *r, uint32_t bin) {
    if (W == 2) {
        *reinterpret_cast<ulonglong2 *>(st.rec + (size_t) slot * 2) = make_ulonglong2(r[0], r[1]);
    } else if (W == 4) {
        ulonglong2 *p = reinterpret_cast<ulonglong2 *>(st.rec + (size_t) slot * 4);
        p[0] = make_ulonglong2(r[0], r[1]);
        p[1] = make_ulonglong2(r[2], r[3]);
    } else {
#pragma unroll
        for (int j = 0; j < W; ++j) st.rec[(size_t) slot * W + j] = r[j];
    }
    st.bin[slot] = (uint16_t) bin;
    atomicAdd(&st.bstart[bin], 1u);   // the tile's bin histogram grows with the tile
}

// ---- reads: one warp per chunk of <= 128 windows of a read, every lane rolls through a run of consecutive windows (partition.cuh) ----
// State of a warp's walk over its reads: read index, first window of the next chunk, and the length / word offset of the read it is
// at and of the NEXT read of the warp (loaded one read ahead: a warp meets 650 reads, and three dependent loads per read — length, offset,
// words — were what the count pass waited on at 32 warps per SM).
struct SpReadCursor {
    uint64_t rd;
    uint32_t c0;
    uint32_t len_cur, len_next;
    uint64_t off_cur, off_next;
    bool primed;
};

__device__ __forceinline__ void sp_cursor_init(SpReadCursor &cur, uint64_t first_read) {
    cur.rd = first_read; cur.c0 = 0; cur.len_cur = cur.len_next = 0; cur.off_cur = cur.off_next = 0; cur.primed = false;
}

// Forms the records of the warp's next chunk.  Returns the mask of the valid records of this LANE (<= PART_RUN); rec / gid hold them.
template<int W>
__device__ __forceinline__ uint32_t sp_read_chunk(const uint64_t *__restrict__ words, const uint64_t *__restrict__ word_off, const uint32_t *__restrict__ len,
                                                  uint64_t n_reads, uint64_t warps_total, int K, int mode, const GroupSel &gs, SpReadCursor &cur,
                                                  uint64_t rec[PART_RUN][W], uint32_t gid[PART_RUN], bool &exhausted) {
    const int lane = threadIdx.x & 31;
    const uint64_t lw_mask = last_word_mask(K);
    uint32_t valid = 0;
    if (!cur.primed) {
        cur.primed = true;
        if (cur.rd < n_reads) { cur.len_cur = __ldg(len + cur.rd); cur.off_cur = __ldg(word_off + cur.rd); }
        if (cur.rd + warps_total < n_reads) { cur.len_next = __ldg(len + cur.rd + warps_total); cur.off_next = __ldg(word_off + cur.rd + warps_total); }
    }
    // skip reads that are too short / finished
    while (cur.rd < n_reads) {
        if (cur.len_cur >= (uint32_t) K && cur.c0 < cur.len_cur - (uint32_t) K + 1) break;
        cur.rd += warps_total;
        cur.c0 = 0;
        cur.len_cur = cur.len_next; cur.off_cur = cur.off_next;
        if (cur.rd + warps_total < n_reads) {
            cur.len_next = __ldg(len + cur.rd + warps_total); cur.off_next = __ldg(word_off + cur.rd + warps_total);
        }
    }
    // the words of the read AFTER this one are on their way while this one is processed (its offset arrived one read ago)
    if (cur.c0 == 0 && cur.rd + warps_total < n_reads && lane < 2)
        asm volatile("prefetch.global.L2 [%0];" ::"l"(words + cur.off_next + (uint64_t) lane * 16));
    if (cur.rd >= n_reads) { exhausted = true; return 0; }
    const uint32_t l = cur.len_cur;
    const uint32_t nwin = l - (uint32_t) K + 1, nw = (l + 31) >> 5;
    const uint64_t *seq = words + cur.off_cur;
    const uint32_t left = nwin - cur.c0;
    const uint32_t run = left >= 32u * PART_RUN ? (uint32_t) PART_RUN : (left + 31u) >> 5;
    const uint32_t p0 = cur.c0 + (uint32_t) lane * run;
    if (p0 < nwin) {
        uint64_t x[W], r[W];
        kmer_window<W>(seq, nw, p0, K, x);
        uint64_t nxt = 0;
        if (run > 1 && p0 + (uint32_t) K < l) kmer_window<1>(seq, nw, p0 + (uint32_t) K, 32, &nxt);
        kmer_rc<W>(x, K, r);
#pragma unroll
        for (int i = 0; i < PART_RUN; ++i) {
            if ((uint32_t) i < run && p0 + (uint32_t) i < nwin) {
                if (i > 0) kmer_roll<W>(x, r, K, (uint32_t) (nxt >> (2 * (i - 1))) & 3u, lw_mask);
                const bool minimal = kmer_ge_num<W>(r, x);
                const bool take_x = mode == PART_FWD || (mode != PART_REV && minimal);
                if (mode != PART_MINIMAL_ONLY || minimal) {
#pragma unroll
                    for (int j = 0; j < W; ++j) rec[i][j] = take_x ? x[j] : r[j];
                    gid[i] = group_of<W>(rec[i], gs);
                    valid |= 1u << i;
                }
            }
        }
    }
    cur.c0 += 32u * run;
    return valid;
}

// count: fine-group histogram in shared memory, flushed to the global one at the end
template<int W>
__global__ void __launch_bounds__(SPC_THREADS) sp_count_reads_kernel(const uint64_t *__restrict__ words, const uint64_t *__restrict__ word_off,
                                                                   const uint32_t *__restrict__ len, uint64_t n_reads, int K, int mode, GroupSel gs,
                                                                   int shift /* counts bins g >> shift: 0 = fine groups */, uint32_t n_groups,
                                                                   uint32_t *__restrict__ hist) {
    extern __shared__ __align__(16) unsigned char sp_smem[];
    uint32_t *sh = reinterpret_cast<uint32_t *>(sp_smem);
    for (uint32_t i = threadIdx.x; i < n_groups; i += SPC_THREADS) sh[i] = 0;
    __syncthreads();
    const uint64_t warps_total = (uint64_t) gridDim.x * (SPC_THREADS / 32);
    SpReadCursor cur;
    sp_cursor_init(cur, (uint64_t) blockIdx.x * (SPC_THREADS / 32) + (threadIdx.x >> 5));
    bool exhausted = false;
    while (!exhausted) {
        uint64_t rec[PART_RUN][W];
        uint32_t gid[PART_RUN];
        const uint32_t valid = sp_read_chunk<W>(words, word_off, len, n_reads, warps_total, K, mode, gs, cur, rec, gid, exhausted);
#pragma unroll
        for (int i = 0; i < PART_RUN; ++i)
            if ((valid >> i) & 1u) atomicAdd(&sh[gid[i] >> shift], 1u);
    }
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < n_groups; i += SPC_THREADS) {
        const uint32_t c = sh[i];
        if (c) atomicAdd(&hist[i], c);
    }
}

// pass 1: tiles of <= rounds x 16 chunks, ordered by coarse bin c = g >> s, runs written at the bins' cursors
template<int W, bool PEER = false>
__global__ void __launch_bounds__(SP_THREADS, 2) sp_scatter_reads_kernel(const uint64_t *__restrict__ words, const uint64_t *__restrict__ word_off,
                                                                     const uint32_t *__restrict__ len, uint64_t n_reads, int K, int mode, GroupSel gs,
                                                                     int s, uint32_t n_coarse, int rounds, uint32_t *__restrict__ gcursor,
                                                                     uint64_t *__restrict__ out, const __grid_constant__ SpPeers peers) {
    extern __shared__ __align__(16) unsigned char sp_smem[];
    __shared__ uint32_t s_fill, s_scan[SP_THREADS / 32 + 1];
    __shared__ uint8_t s_own[PEER ? SP_MAX_BINS : 4];
    const SpStage st = sp_carve<W>(sp_smem, SpCfg<W>::CAP, false);
    const int lane = threadIdx.x & 31;
    const uint64_t warps_total = (uint64_t) gridDim.x * SP_WARPS;
    SpReadCursor cur;
    sp_cursor_init(cur, (uint64_t) blockIdx.x * SP_WARPS + (threadIdx.x >> 5));
    bool exhausted = false;
    while (true) {
        if (threadIdx.x == 0) s_fill = 0;
        sp_stage_reset(st, n_coarse);
        __syncthreads();
        for (int round = 0; round < rounds; ++round) {
            if (exhausted) continue;
            uint64_t rec[PART_RUN][W];
            uint32_t gid[PART_RUN];
            const uint32_t valid = sp_read_chunk<W>(words, word_off, len, n_reads, warps_total, K, mode, gs, cur, rec, gid, exhausted);
            // the warp reserves room for its valid records with one shared-memory atomic
            const uint32_t mine = (uint32_t) __popc(valid);
            uint32_t incl = mine;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint32_t o = __shfl_up_sync(0xffffffffu, incl, d);
                if (lane >= d) incl += o;
            }
            uint32_t base = 0;
            if (lane == 31 && incl) base = atomicAdd(&s_fill, incl);
            base = __shfl_sync(0xffffffffu, base, 31);
            uint32_t slot = base + incl - mine;
#pragma unroll
            for (int i = 0; i < PART_RUN; ++i)
                if ((valid >> i) & 1u) sp_put<W>(st, slot++, rec[i], gid[i] >> s);
        }
        __syncthreads();
        const uint32_t n = s_fill;
        if (n) sp_flush<W, PEER>(st, n, n_coarse, gcursor, 0u, out, nullptr, s_scan, &peers, s_own);
        if (__syncthreads_and(exhausted ? 1 : 0)) break;
    }
}

// ---- k-mer candidates of (k+1)-mers (partition.cuh derive_candidates) -----------------------------------------------------------------
template<int WS, int W>
__global__ void __launch_bounds__(SPC_THREADS) sp_count_derive_kernel(const uint64_t *__restrict__ kp, uint64_t n, int k, GroupSel gs, int shift,
                                                                    uint32_t n_groups, uint32_t *__restrict__ hist) {
    extern __shared__ __align__(16) unsigned char sp_smem[];
    uint32_t *sh = reinterpret_cast<uint32_t *>(sp_smem);
    for (uint32_t i = threadIdx.x; i < n_groups; i += SPC_THREADS) sh[i] = 0;
    __syncthreads();
    for (uint64_t i = (uint64_t) blockIdx.x * SPC_THREADS + threadIdx.x; i < n; i += (uint64_t) gridDim.x * SPC_THREADS) {
        uint64_t x[WS], a[2][W];
        uint32_t bit[2];
        load_rec<WS>(kp, i, x);
        derive_candidates<WS, W>(x, k, a, bit);
        atomicAdd(&sh[group_of<W>(a[0], gs) >> shift], 1u);
        atomicAdd(&sh[group_of<W>(a[1], gs) >> shift], 1u);
    }
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < n_groups; i += SPC_THREADS) {
        const uint32_t c = sh[i];
        if (c) atomicAdd(&hist[i], c);
    }
}

template<int WS, int W, bool PEER = false>
__global__ void __launch_bounds__(SP_THREADS, 2) sp_scatter_derive_kernel(const uint64_t *__restrict__ kp, uint64_t n, int k, int pshift, GroupSel gs, int s,
                                                                      uint32_t n_coarse, uint32_t *__restrict__ gcursor, uint64_t *__restrict__ out,
                                                                      uint8_t *__restrict__ out_pay, const __grid_constant__ SpPeers peers) {
    extern __shared__ __align__(16) unsigned char sp_smem[];
    __shared__ uint32_t s_scan[SP_THREADS / 32 + 1];
    __shared__ uint8_t s_own[PEER ? SP_MAX_BINS : 4];
    constexpr int CAP = SpCfg<W>::CAP;
    constexpr uint64_t SRC_PER_TILE = CAP / 2;
    const SpStage st = sp_carve<W>(sp_smem, CAP, out_pay != nullptr);
    for (uint64_t t0 = (uint64_t) blockIdx.x * SRC_PER_TILE; t0 < n; t0 += (uint64_t) gridDim.x * SRC_PER_TILE) {
        const uint32_t cnt = (uint32_t) ((n - t0) < SRC_PER_TILE ? (n - t0) : SRC_PER_TILE);
        sp_stage_reset(st, n_coarse);
        __syncthreads();
        for (uint32_t j = threadIdx.x; j < cnt; j += SP_THREADS) {
            uint64_t x[WS], a[2][W];
            uint32_t bit[2];
            load_rec<WS>(kp, t0 + j, x);
            derive_candidates<WS, W>(x, k, a, bit);
            const uint32_t g0 = group_of<W>(a[0], gs), g1 = group_of<W>(a[1], gs);
            if (pshift >= 0) {
                a[0][W - 1] |= (uint64_t) bit[0] << pshift;
                a[1][W - 1] |= (uint64_t) bit[1] << pshift;
            } else if (st.pay) {
                st.pay[2 * j] = (uint8_t) bit[0];
                st.pay[2 * j + 1] = (uint8_t) bit[1];
            }
            sp_put<W>(st, 2 * j, a[0], g0 >> s);
            sp_put<W>(st, 2 * j + 1, a[1], g1 >> s);
        }
        __syncthreads();
        sp_flush<W, PEER>(st, 2 * cnt, n_coarse, gcursor, 0u, out, out_pay, s_scan, &peers, s_own);
    }
}

// ---- pass 2: tiles of ONE coarse bin, loaded with bulk asynchronous copies, ordered by the fine group ------------------------------------
__device__ __forceinline__ uint32_t sp_smem_addr(const void *p) { return (uint32_t) __cvta_generic_to_shared(p); }
__device__ __forceinline__ void sp_mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(sp_smem_addr(bar)), "r"(count));
}
__device__ __forceinline__ void sp_mbar_expect(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(sp_smem_addr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void sp_bulk_g2s(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(sp_smem_addr(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(sp_smem_addr(bar)) : "memory");
}
__device__ __forceinline__ void sp_mbar_wait(uint64_t *bar, uint32_t phase) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(sp_smem_addr(bar)), "r"(phase) : "memory");
}

// Pass 2 works on SEGMENTS of the pass-1 output: runs of records that share one coarse bin (one GPU: the coarse bins themselves;
// hash-sharded: the part of a coarse bin that arrived from one source GPU).  A tile = up to CAP2 consecutive records of one segment.
struct SpTile {
    uint32_t begin, cnt;   // records [begin, begin + cnt) of the pass-1 output
    uint32_t gbase;        // first fine group of the segment's coarse bin (= local bin 0 of the tile)
    uint32_t pad;
};

// one GPU: segment c = coarse bin c
__global__ void sp_coarse_segments_kernel(const uint32_t *__restrict__ fine_start, int s, uint32_t n_coarse, uint32_t n_groups,
                                          uint32_t *__restrict__ seg_begin, uint32_t *__restrict__ seg_cnt, uint32_t *__restrict__ seg_gbase) {
    const uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n_coarse) return;
    const uint32_t lo = fine_start[(uint64_t) c << s];
    const uint64_t hi_g = ((uint64_t) c + 1) << s;
    const uint32_t hi = fine_start[hi_g < n_groups ? hi_g : n_groups];
    seg_begin[c] = lo; seg_cnt[c] = hi - lo; seg_gbase[c] = c << s;
}

// hash-sharded receiver: segment (src, c) = the records of coarse bin c that arrived from rank src.  cnt[src * n_co + c] came with the
// records; run_start[src] = first record of rank src's run in the receive buffer.  One warp per source scans its row.
__global__ void sp_recv_segments_kernel(const uint32_t *__restrict__ cnt, const uint64_t *__restrict__ run_start, uint32_t n_src, uint32_t n_co, int s,
                                        uint32_t *__restrict__ seg_begin, uint32_t *__restrict__ seg_cnt, uint32_t *__restrict__ seg_gbase) {
    const uint32_t src = blockIdx.x;
    const int lane = threadIdx.x;   // 32 threads
    if (src >= n_src) return;
    uint32_t run = (uint32_t) run_start[src];
    for (uint32_t c0 = 0; c0 < n_co; c0 += 32) {
        const uint32_t c = c0 + lane;
        const uint32_t v = c < n_co ? cnt[(uint64_t) src * n_co + c] : 0u;
        uint32_t incl = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t o = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += o;
        }
        if (c < n_co) {
            const uint64_t i = (uint64_t) src * n_co + c;
            seg_begin[i] = run + incl - v; seg_cnt[i] = v; seg_gbase[i] = c << s;
        }
        run += __shfl_sync(0xffffffffu, incl, 31);
    }
}

__global__ void sp_seg_tiles_kernel(const uint32_t *__restrict__ seg_cnt, uint32_t n_seg, uint32_t cap2, uint32_t *__restrict__ n_tiles) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i <= n_seg) n_tiles[i] = i < n_seg ? (seg_cnt[i] + cap2 - 1) / cap2 : 0u;
}

// one thread per TILE: its segment by binary search over the scanned tile counts
__global__ void sp_tile_fill_kernel(const uint32_t *__restrict__ seg_begin, const uint32_t *__restrict__ seg_cnt, const uint32_t *__restrict__ seg_gbase,
                                    const uint32_t *__restrict__ tile_off, uint32_t n_seg, uint32_t cap2, SpTile *__restrict__ tiles) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= tile_off[n_seg]) return;
    uint32_t lo = 0, hi = n_seg;   // last segment with tile_off[seg] <= t (segments without tiles share their successor's offset)
    while (hi - lo > 1) {
        const uint32_t mid = (lo + hi) >> 1;
        if (tile_off[mid] <= t) lo = mid; else hi = mid;
    }
    const uint32_t o = (t - tile_off[lo]) * cap2, c = seg_cnt[lo];
    tiles[t] = SpTile{seg_begin[lo] + o, (c - o) < cap2 ? (c - o) : cap2, seg_gbase[lo], 0u};
}

template<int W>
__global__ void __launch_bounds__(SP_THREADS, 2) sp_scatter_fine_kernel(const uint64_t *__restrict__ mid, const uint8_t *__restrict__ mid_pay,
                                                                       const SpTile *__restrict__ tiles, const uint32_t *__restrict__ n_tiles_dev,
                                                                       GroupSel gs, uint64_t lw_keep, int s, uint32_t *__restrict__ gcursor,
                                                                       uint64_t *__restrict__ out, uint8_t *__restrict__ out_pay) {
    extern __shared__ __align__(128) unsigned char sp_smem[];
    __shared__ uint32_t s_scan[SP_THREADS / 32 + 1];
    __shared__ __align__(8) uint64_t s_bar[2];
    __shared__ SpTile s_tile[2];
    constexpr int CAP2 = SpCfg<W>::CAP2;
    const bool with_pay = mid_pay != nullptr;
    // two record buffers (bulk-copy destinations), then ONE set of ordering arrays: a tile is ordered and written out before the next
    unsigned char *buf[2] = {sp_smem, sp_smem + (size_t) CAP2 * W * 8};
    unsigned char *aux = sp_smem + 2 * (size_t) CAP2 * W * 8;
    SpStage st;
    st.bin = reinterpret_cast<uint16_t *>(aux); aux += (size_t) CAP2 * 2;
    st.src = reinterpret_cast<uint16_t *>(aux); aux += (size_t) CAP2 * 2;
    uint8_t *paybuf[2] = {aux, aux + CAP2};
    aux += 2 * (size_t) CAP2;
    st.bstart = reinterpret_cast<uint32_t *>(((uintptr_t) aux + 15) & ~(uintptr_t) 15);
    st.cursor = st.bstart + SP_MAX_BINS + 1;
    st.gdelta = st.cursor + SP_MAX_BINS;
    const uint32_t n_local = 1u << s;
    const uint32_t n_tiles = *n_tiles_dev;

    auto issue = [&](const SpTile &d, int b) {   // thread 0: start the copy of a tile into buffer b
        s_tile[b] = d;
        // bulk copies need 16-byte aligned addresses and sizes: record arrays are 256-byte aligned and W * 8 = 16 or 32 (W = 2, 4) keeps
        // every record boundary aligned; odd W and the payload bytes take plain loads (below)
        if (W % 2 == 0) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // the buffer was last read through the generic proxy
            sp_mbar_expect(&s_bar[b], d.cnt * (uint32_t) (W * 8));
            sp_bulk_g2s(buf[b], mid + (uint64_t) d.begin * W, d.cnt * (uint32_t) (W * 8), &s_bar[b]);
        }
    };

    if (threadIdx.x == 0) {
        sp_mbar_init(&s_bar[0], 1);
        sp_mbar_init(&s_bar[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    uint32_t t = blockIdx.x;
    SpTile next{0u, 0u, 0u, 0u};   // thread 0: descriptor of the tile after the one in flight, loaded one iteration ahead
    if (threadIdx.x == 0 && t < n_tiles) {
        issue(tiles[t], 0);
        if (t + gridDim.x < n_tiles) next = tiles[t + gridDim.x];
    }
    uint32_t phase[2] = {0, 0};
    int b = 0;
    for (; t < n_tiles; t += gridDim.x, b ^= 1) {
        __syncthreads();   // the previous iteration's flush is done with the ordering arrays; s_tile[b] is visible
        const uint32_t tn = t + gridDim.x;
        if (threadIdx.x == 0 && tn < n_tiles) {   // prefetch the next tile while this one is ordered
            issue(next, b ^ 1);
            if (tn + gridDim.x < n_tiles) next = tiles[tn + gridDim.x];
        }
        const uint32_t begin = s_tile[b].begin, cnt = s_tile[b].cnt, gbase = s_tile[b].gbase;
        st.rec = reinterpret_cast<uint64_t *>(buf[b]);
        st.pay = with_pay ? paybuf[b] : nullptr;
        if (W % 2 == 0) {
            sp_mbar_wait(&s_bar[b], phase[b]);
            phase[b] ^= 1;
        } else {
            for (uint32_t q = threadIdx.x; q < cnt * (uint32_t) W; q += SP_THREADS) st.rec[q] = mid[(uint64_t) begin * W + q];
        }
        if (with_pay)
            for (uint32_t q = threadIdx.x; q < cnt; q += SP_THREADS) st.pay[q] = mid_pay[begin + q];
        sp_stage_reset(st, n_local);
        __syncthreads();
        for (uint32_t q = threadIdx.x; q < cnt; q += SP_THREADS) {
            uint64_t r[W];
#pragma unroll
            for (int j = 0; j < W; ++j) r[j] = st.rec[(size_t) q * W + j];
            r[W - 1] &= lw_keep;   // the group key ignores a mask-bit payload in the padding
            const uint32_t bin = group_of<W>(r, gs) - gbase;
            st.bin[q] = (uint16_t) bin;
            atomicAdd(&st.bstart[bin], 1u);
        }
        __syncthreads();
        sp_flush<W>(st, cnt, n_local, gcursor, gbase, out, out_pay, s_scan);
    }
}

// count of the fine groups of records that already exist (the receiving side of the hash-sharded path): shared-memory histogram
template<int W>
__global__ void __launch_bounds__(SPC_THREADS) sp_count_records_kernel(const uint64_t *__restrict__ recs, uint64_t n, GroupSel gs, uint64_t lw_keep,
                                                                      uint32_t n_groups, uint32_t *__restrict__ hist) {
    extern __shared__ __align__(16) unsigned char sp_smem[];
    uint32_t *sh = reinterpret_cast<uint32_t *>(sp_smem);
    for (uint32_t i = threadIdx.x; i < n_groups; i += SPC_THREADS) sh[i] = 0;
    __syncthreads();
    for (uint64_t i = (uint64_t) blockIdx.x * SPC_THREADS + threadIdx.x; i < n; i += (uint64_t) gridDim.x * SPC_THREADS) {
        uint64_t r[W];
        load_rec<W>(recs, i, r);
        r[W - 1] &= lw_keep;
        atomicAdd(&sh[group_of<W>(r, gs)], 1u);
    }
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < n_groups; i += SPC_THREADS) {
        const uint32_t c = sh[i];
        if (c) atomicAdd(&hist[i], c);
    }
}

}  // namespace sb200
