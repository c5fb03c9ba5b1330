// partition.cuh — k-mer instances straight from their source into their GROUP, in one trip through HBM.
//
// The reference splits k-mers into hash buckets through per-thread x per-bucket staging cells and temp files
// (C/utils/kmer_mph/kmer_splitter.hpp:73-167, kmer_splitters.hpp:25-59,159-204).  Round 1 materialised every instance (extract /
// derive kernel), then moved the 4.7 GB array through two counting passes (histogram read, scatter read + write, twice) before the
// group kernel saw it: the instances crossed HBM eight times.  Here the producer itself is the partitioner:
//   pass 1 (count)  recomputes nothing it has to keep: every window / candidate is formed in registers, canonicalised, hashed to its
//                   group  g = bucket << p | top p value bits  (the composite key of segsort.cuh) and counted with one fire-and-forget
//                   L2 atomic (RED) on hist[g] — n_groups counters, 160 KB at config 2, L2-resident;
//   scan            group starts = exclusive scan of the histogram (also the group table the group kernel needs: no binary searches);
//   pass 2 (write)  forms the same records again, claims a slot with atomicAdd(cursor[g]) and stores the record there.
// Records of one group arrive in no particular order — the group kernel (grouphash.cuh) deduplicates by hashing and orders only the
// distinct records, so the order inside a group never mattered.  The scattered 16-byte stores land in n_groups write frontiers (one
// 128-byte line each, 5 MB in all), which the 126 MB L2 completes before it evicts them: DRAM sees whole lines.
// The instances cross HBM twice (written once, read once by the group kernel) instead of eight times.
//
// Reads: one warp per read, every lane ROLLS through a run of consecutive windows (forward window and reverse complement advance by
// one shift each; the full 2-bit-group reversal is paid once per lane, not once per window).  RtSeq::operator<<= / IsMinimal
// (C/sequence/rtseq.hpp:450-467,407-415), KMerSegmentPolicy (C/utils/kmer_mph/kmer_buckets.hpp:28-41).
#pragma once
#include "common.cuh"
#include "kmer_ops.cuh"
#include "radix_sort.cuh"

namespace sb200 {

// Group key of a record: g = (bucket - bucket_base) << p | top p bits of word 0's significant part
struct GroupSel {
    uint32_t num_buckets;
    uint32_t bucket_base;
    int p;
    int top;            // significant bits of word 0 (64, or 2K for one-word records)
};

template<int W>
__device__ __forceinline__ uint32_t group_of(const uint64_t *r, const GroupSel &gs) {
    const uint32_t pre = gs.p ? (uint32_t) ((r[0] >> (gs.top - gs.p)) & ((1ULL << gs.p) - 1ULL)) : 0u;
    return ((kmer_bucket<W>(r, gs.num_buckets) - gs.bucket_base) << gs.p) | pre;
}

constexpr int PART_RUN = 4;   // consecutive windows one lane rolls through per chunk of a read (a chunk = 128 windows)

// what a window emits: the canonical form (gbuilder / spades-core: fwd+RC streams with the IsMinimal filter fold into one record per
// window), the window itself if it is minimal (forward stream only), the window, or its reverse complement (spades-kmercount keeps
// every k-mer of both strands: one launch of PART_FWD and one of PART_REV over the same histogram / cursors)
enum PartMode { PART_CANON = 0, PART_MINIMAL_ONLY = 1, PART_FWD = 2, PART_REV = 3 };

// WRITE = false: hist[g] += 1 per emitted record.  WRITE = true: hist holds the cursors (a copy of the scanned histogram); every
// record is stored at its group's next free slot.
template<int W, bool WRITE>
__global__ void __launch_bounds__(256) partition_reads_kernel(const uint64_t *__restrict__ words, const uint64_t *__restrict__ word_off,
                                                             const uint32_t *__restrict__ len, uint64_t n_reads, int K, int mode, GroupSel gs,
                                                             uint32_t *__restrict__ hist, uint64_t *__restrict__ out) {
    const int lane = threadIdx.x & 31;
    const uint64_t warps_total = (uint64_t) gridDim.x * (blockDim.x >> 5);
    const uint64_t lw_mask = last_word_mask(K);
    for (uint64_t rd = (uint64_t) blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); rd < n_reads; rd += warps_total) {
        const uint32_t l = len[rd];
        if (l < (uint32_t) K) continue;
        const uint32_t nwin = l - (uint32_t) K + 1;
        const uint64_t *seq = words + word_off[rd];
        const uint32_t nw = (l + 31) >> 5;
        // a read of up to 128 windows is one chunk with ceil(nwin / 32) windows per lane; longer sequences (contigs fed back as
        // reads) go through in chunks of 128 windows
        const uint32_t run = nwin >= 32u * PART_RUN ? (uint32_t) PART_RUN : (nwin + 31u) >> 5;
        for (uint32_t c0 = 0; c0 < nwin; c0 += 32u * run) {
            const uint32_t p0 = c0 + (uint32_t) lane * run;
            uint64_t rec[PART_RUN][W];
            uint32_t gid[PART_RUN];
            uint32_t valid = 0;
            if (p0 < nwin) {
                uint64_t x[W], r[W];
                kmer_window<W>(seq, nw, p0, K, x);
                kmer_rc<W>(x, K, r);
                // the bases that enter the window while this lane rolls: positions p0 + K ... (at most PART_RUN - 1 of them)
                uint64_t nxt = 0;
                if (run > 1 && p0 + (uint32_t) K < l) kmer_window<1>(seq, nw, p0 + (uint32_t) K, 32, &nxt);
#pragma unroll
                for (int i = 0; i < PART_RUN; ++i) {
                    if ((uint32_t) i < run && p0 + (uint32_t) i < nwin) {
                        if (i > 0) {
                            const uint32_t c = (uint32_t) (nxt >> (2 * (i - 1))) & 3u;
                            kmer_roll<W>(x, r, K, c, lw_mask);
                        }
                        const bool minimal = kmer_ge_num<W>(r, x);   // RtSeq::IsMinimal, see kmer_canonical
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
            if (!WRITE) {
#pragma unroll
                for (int i = 0; i < PART_RUN; ++i)
                    if ((valid >> i) & 1u) atomicAdd(&hist[gid[i]], 1u);
            } else {
                uint32_t pos[PART_RUN];
#pragma unroll
                for (int i = 0; i < PART_RUN; ++i)
                    if ((valid >> i) & 1u) pos[i] = atomicAdd(&hist[gid[i]], 1u);
#pragma unroll
                for (int i = 0; i < PART_RUN; ++i)
                    if ((valid >> i) & 1u) store_rec<W>(out, pos[i], rec[i]);
            }
        }
    }
}

// ---- k-mer candidates of stored (k+1)-mers (DeBruijnKMerKMerSplitter::FillBufferFromKMers, kmer_splitters.hpp:159-176) -------------
// Every (k+1)-mer x contributes canon(x[0..k)) and canon(x[1..k]).  rc(x[0..k)) is the suffix of rc(x) and rc(x[1..k]) its prefix, so
// one reverse complement of the (k+1)-mer serves both candidates.  Each candidate carries the InOutMask bit the (k+1)-mer contributes
// to it (kmer_extension_index_builder.hpp:44-59, kmer_extension_index.hpp:92-106: AddOutgoing(next) for the prefix k-mer,
// AddIncoming(previous) for the suffix k-mer, mirrored when the k-mer is stored as its reverse complement): in the three padding bits
// above the k-mer (pshift >= 0), or — when the k-mer fills its last word (k = 31 mod 32) — in a byte array beside the records (`pay`).
template<int WS, int W, bool WRITE>
__global__ void __launch_bounds__(256) partition_derive_kernel(const uint64_t *__restrict__ kp, uint64_t n, int k, int pshift, GroupSel gs,
                                                              uint32_t *__restrict__ hist, uint64_t *__restrict__ out, uint8_t *__restrict__ pay) {
    const uint64_t i = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint64_t x[WS], a[2][W];
    uint32_t bit[2];
    load_rec<WS>(kp, i, x);
    derive_candidates<WS, W>(x, k, a, bit);
    const uint32_t g0 = group_of<W>(a[0], gs), g1 = group_of<W>(a[1], gs);
    if (!WRITE) {
        atomicAdd(&hist[g0], 1u);
        atomicAdd(&hist[g1], 1u);
    } else {
        const uint32_t p0 = atomicAdd(&hist[g0], 1u), p1 = atomicAdd(&hist[g1], 1u);
        if (pshift >= 0) {
            a[0][W - 1] |= (uint64_t) bit[0] << pshift;
            a[1][W - 1] |= (uint64_t) bit[1] << pshift;
        } else if (pay) {
            pay[p0] = (uint8_t) bit[0];
            pay[p1] = (uint8_t) bit[1];
        }
        store_rec<W>(out, p0, a[0]);
        store_rec<W>(out, p1, a[1]);
    }
}

}  // namespace sb200
