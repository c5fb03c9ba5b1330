// grouphash.cuh — deduplication, counting and the final order of one GROUP of records, by hashing instead of sorting copies.
//
// After the counting passes (radix_sort.cuh) all records with the same composite key (bucket, top p value bits) are contiguous:
// a group of a few thousand records (segsort.cuh explains the grouping).  At sequencing coverage most of them are COPIES: config 2
// has 7.1 K instances but only 1.8 K distinct (k+1)-mers per group.  group_chunk_kernel (segsort.cuh) still moved every copy through
// a shared-memory counting sort and a warp-wide match; this kernel touches a copy exactly once:
//   (1) the group is streamed from HBM (coalesced, never staged): every record probes a shared-memory hash table whose slot words hold
//       the INDEX of the first record seen with that value next to its multiplicity (one native CAS claims the slot and leaves the first
//       count; the input array is immutable, so a claimed slot is immediately comparable — no publish step, no spinning); a copy adds 1
//       to the same word (or ORs its mask payload into it);
//   (2) the occupied slots are compacted to a list of distinct records (index + count);
//   (3) only the distinct records are ordered: counting sort over 1024 bins of the 32 key bits right below the group prefix (the
//       "tag", monotone in the record order inside a group), rank inside a bin by tag (full record compare on the rare tag tie);
//   (4) records (+ counts / masks) leave in order to the front of the group's own range of the output buffer, as group_chunk_kernel
//       does; seg_compact_kernel then closes the gaps.
// Shared memory holds 2 x UMAX slot words and per-distinct bookkeeping only (48 KB; two CTAs of 512 threads per SM), so the size of a
// group is not limited by it, and neither is the multiplicity of a k-mer (it is at most the group size, and groups of 65535+ records
// take the 64-bit slots).  A group with more than UMAX distinct records is redone in R = 2, 4, ... 256 rounds over disjoint tag ranges;
// beyond that the fail flag sends the set to the LSD path (radix_sort.cuh).
// Groups of 65535+ records need 64-bit slot words: the 32-bit instance skips them and raises ctrl[1], the host then launches the
// 64-bit instance for those groups only.
// The phases are separate __noinline__ functions on purpose.  What is known (profiles/r2s_gh_test_variants.log, r2t_pytest_inline_variant.log):
// built with the phases inlined (-DGH_INLINE_PHASES) the kernel passes the stand-alone harness (tools/gh_test.cu: W = 2, counting mode) and
// 142 of the GPU tests, but the plain-dedup instance (MODE 0: k-mers without mask payload, tests/test_gpu_sharded.py
// [SB200_NO_MASK_PAYLOAD]) dies with "misaligned address"; as separate functions every instance passes the whole suite.  Whether the
// inlined form exposes a flaw of this code or of nvcc 12.9's code generation is open: compute-sanitizer (racecheck / memcheck) is closed
// on the pool this was developed on.  The inlined form is 2 % faster (7.04 against 7.19 ms per step at config 2), not worth the doubt.
#pragma once
#include "common.cuh"
#include "kmer_ops.cuh"
#include "radix_sort.cuh"
#include "scan.cuh"
#include "segsort.cuh"

namespace sb200 {

#ifndef GH_UMAX
#define GH_UMAX 4096
#endif
// GH_INLINE_PHASES (tools/gh_test.cu only): the two phases inlined into the kernel body, the variant that was miscompiled (see above)
#ifdef GH_INLINE_PHASES
#define GH_PHASE __forceinline__
#else
#define GH_PHASE __noinline__
#endif
#ifndef GH_THREADS
#define GH_THREADS 512
#endif
#ifndef GH_MIN_BLOCKS
#define GH_MIN_BLOCKS 2   // 2 x 512 threads beat 4 x 256 (7.1 against 7.9 ms): the per-group fixed work is shared by twice the threads
#endif

struct HashCfg {
    static constexpr int THREADS = GH_THREADS;
    static constexpr int UMAX = GH_UMAX;          // distinct records one round may hold
    static constexpr int TS = 2 * UMAX;           // table slots (load <= 0.5)
    static constexpr int LOG_TS = (TS == 4096) ? 12 : (TS == 8192) ? 13 : (TS == 16384) ? 14 : -1;
    static constexpr int BINS = 1024;
    static_assert(LOG_TS > 0, "GH_UMAX must be 2048, 4096 or 8192");
    static_assert(BINS % THREADS == 0 || THREADS % BINS == 0, "bins per thread");
};

// Slot word: (index of the first record with that value + 1) in the upper half, multiplicity (or OR-ed payload) in the lower half;
// 0 = empty.  One native CAS claims the slot AND leaves the first count; a copy adds 1 to the same word.
//   IdxT = uint16_t: 32-bit slots, groups below 65535 records (a multiplicity is at most the group size, so the count cannot carry)
//   IdxT = uint32_t: 64-bit slots, any group
template<typename IdxT> struct GhSlot;
template<> struct GhSlot<uint16_t> { using type = uint32_t; static constexpr int HB = 16; };
template<> struct GhSlot<uint32_t> { using type = unsigned long long; static constexpr int HB = 32; };

template<typename IdxT>
constexpr size_t gh_region_a() {
    // first life: the slot words; second life: tags + bin starts + bin cursors + binned entry ids
    size_t a1 = (size_t) HashCfg::TS * sizeof(typename GhSlot<IdxT>::type);
    size_t a2 = (size_t) HashCfg::UMAX * 4 + (size_t) (HashCfg::BINS + 1) * 4 + (size_t) HashCfg::BINS * 4 + (size_t) HashCfg::UMAX * 2;
    size_t a = a1 > a2 ? a1 : a2;
    return (a + 15) & ~(size_t) 15;
}

template<typename IdxT>
constexpr size_t group_hash_smem() {
    // region B: multiplicity + first index of every distinct record
    return gh_region_a<IdxT>() + (size_t) HashCfg::UMAX * sizeof(IdxT) * 2;
}

template<int W>
__device__ __forceinline__ void gh_load(const uint64_t *base, uint64_t idx, uint64_t *r) { load_rec<W>(base, idx, r); }

template<int W>
__device__ __forceinline__ uint32_t gh_slot(const uint64_t *k) {
    uint64_t h = k[0] * 0x9E3779B97F4A7C15ULL;
#pragma unroll
    for (int j = 1; j < W; ++j) h = (h ^ k[j]) * 0xC2B2AE3D27D4EB4FULL + (h >> 31);
    h ^= h >> 29;
    h *= 0xD6E8FEB86659FD93ULL;
    return (uint32_t) (h >> (64 - HashCfg::LOG_TS));
}

// Where record q of the group lives.  One GPU: the group is a contiguous range (recs points at its first record).  Hash-sharded
// path: the group arrives in n_src runs, one per source GPU (the senders group by the same global key, shard.cu), and is never
// made contiguous: q is a position in the logical concatenation of the parts, base[] their logical starts, off[] their physical ones.
constexpr int GH_MAX_SRC = 32;
template<bool SEG>
struct GhAddr {
    const uint64_t *recs;
    const uint8_t *pay;      // mask-bit payload beside the records (same indexing), or nullptr
    const uint32_t *base;    // SEG: shared memory, n_src + 1 logical starts
    const uint32_t *off;     // SEG: shared memory, n_src physical starts (record index into recs)
    uint32_t n_src;
    __device__ __forceinline__ uint64_t phys(uint32_t q) const {
        if (!SEG) return q;
        uint32_t s = 0;
        while (s + 1 < n_src && q >= base[s + 1]) ++s;
        return (uint64_t) off[s] + (q - base[s]);
    }
};

// The shared memory of one CTA (see group_hash_smem): every phase gets the same view.
template<typename IdxT>
struct GhSmem {
    typename GhSlot<IdxT>::type *slot;   // first life of region A
    uint32_t *ltag;       // second life of region A: tag of every distinct record
    uint32_t *binstart;   //   BINS + 1
    uint32_t *cursor;     //   BINS
    uint16_t *binned;     //   entry ids grouped by bin
    IdxT *lc;             // region B: multiplicity of every distinct record
    IdxT *lq;             //           index of its first occurrence in the group
};

template<typename IdxT>
__device__ __forceinline__ GhSmem<IdxT> gh_views(unsigned char *raw) {
    constexpr int UMAX = HashCfg::UMAX, BINS = HashCfg::BINS;
    GhSmem<IdxT> v;
    v.slot = reinterpret_cast<typename GhSlot<IdxT>::type *>(raw);
    v.ltag = reinterpret_cast<uint32_t *>(raw);
    v.binstart = v.ltag + UMAX;
    v.cursor = v.binstart + BINS + 1;
    v.binned = reinterpret_cast<uint16_t *>(v.cursor + BINS);
    v.lc = reinterpret_cast<IdxT *>(raw + gh_region_a<IdxT>());
    v.lq = v.lc + UMAX;
    return v;
}

// Phase 1+2: stream the records of the group whose tag falls into the round's range through the hash table, then compact the occupied
// slots to the list (lq, lc).  Returns the number of distinct records, or ~0u when the table got crowded (more than UMAX distinct).
template<int W, int MODE, typename IdxT, bool SEG>
__device__ GH_PHASE uint32_t gh_dedup(unsigned char *raw, const GhAddr<SEG> ga, uint32_t gcnt, int shift2,
                                          uint64_t lw_keep, int pshift, int lgR, uint32_t round, uint32_t *s_U, uint32_t *s_overflow) {
    const uint64_t *__restrict__ g = ga.recs;
    const uint8_t *__restrict__ gpay = ga.pay;
    using SlotT = typename GhSlot<IdxT>::type;
    constexpr int THREADS = HashCfg::THREADS, UMAX = HashCfg::UMAX, TS = HashCfg::TS, HB = GhSlot<IdxT>::HB;
    constexpr SlotT LOW = ((SlotT) 1 << HB) - 1;
    constexpr int MAX_PROBES = 512;
    const GhSmem<IdxT> sm = gh_views<IdxT>(raw);
    const int lane = threadIdx.x & 31;
    for (uint32_t i = threadIdx.x; i < (uint32_t) TS; i += THREADS) sm.slot[i] = 0;
    if (threadIdx.x == 0) { *s_U = 0; *s_overflow = 0; }
    __syncthreads();
    for (uint32_t q0 = 0; q0 < gcnt; q0 += THREADS * 4) {
        uint64_t in[4][W];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const uint32_t q = q0 + i * THREADS + threadIdx.x;
#pragma unroll
            for (int j = 0; j < W; ++j) in[i][j] = 0;
            if (q < gcnt) gh_load<W>(g, ga.phys(q), in[i]);
        }
        // Four records per thread move through the probe rounds together, so that their look-ups of the first occurrences (an L2 hit of
        // several hundred cycles each: the copy of a k-mer finds its slot taken and has to see the record that took it) overlap.  A
        // record whose slot holds a different value steps to the next slot and takes part in the next round.
        uint32_t h[4], val[4];
        bool todo[4];
        bool any = false;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const uint32_t q = q0 + i * THREADS + threadIdx.x;
            if (MODE == 2) {   // the mask bit this candidate carries: in its padding bits, or in the byte array beside the records
                val[i] = gpay ? ((q < gcnt) ? 1u << (gpay[ga.phys(q)] & 7u) : 0u) : 1u << ((uint32_t) (in[i][W - 1] >> pshift) & 7u);
                in[i][W - 1] &= lw_keep;
            } else {
                val[i] = (MODE == 1) ? 1u : 0u;
            }
            todo[i] = q < gcnt;
            if (lgR) todo[i] = todo[i] && (seg_tag(in[i][0], shift2) >> (32 - lgR)) == round;
            h[i] = gh_slot<W>(in[i]);
            any |= todo[i];
        }
        for (int probes = 0; any && probes < MAX_PROBES; ++probes) {
            SlotT v[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const uint32_t q = q0 + i * THREADS + threadIdx.x;
                v[i] = 0;
                if (todo[i]) {
                    v[i] = sm.slot[h[i]];
                    if (v[i] == 0) v[i] = atomicCAS(&sm.slot[h[i]], (SlotT) 0, ((SlotT) (q + 1) << HB) | val[i]);
                    if (v[i] == 0) todo[i] = false;   // claimed: the CAS left index and first count
                }
            }
            uint64_t o[4][W];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
#pragma unroll
                for (int j = 0; j < W; ++j) o[i][j] = 0;
                if (todo[i]) gh_load<W>(g, ga.phys((uint32_t) (v[i] >> HB) - 1u), o[i]);
            }
            any = false;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                if (!todo[i]) continue;
                if (MODE == 2) o[i][W - 1] &= lw_keep;
                if (kmer_eq<W>(o[i], in[i])) {
                    if (MODE == 2) atomicOr(&sm.slot[h[i]], (SlotT) val[i]);
                    else if (MODE == 1) atomicAdd(&sm.slot[h[i]], (SlotT) 1);   // cannot carry: a count is below the group size
                    todo[i] = false;
                } else {
                    h[i] = (h[i] + 1) & (TS - 1);
                    any = true;
                }
            }
        }
        if (any) atomicExch(s_overflow, 1u);   // crowded table: far more than UMAX distinct records
    }
    __syncthreads();
    for (uint32_t i0 = 0; i0 < (uint32_t) TS; i0 += THREADS) {
        const SlotT v = sm.slot[i0 + threadIdx.x];
        const uint32_t occ = __ballot_sync(0xffffffffu, v != 0);
        uint32_t base = 0;
        if (lane == 0 && occ) base = atomicAdd(s_U, (uint32_t) __popc(occ));
        base = __shfl_sync(0xffffffffu, base, 0);
        if (v != 0) {
            const uint32_t e = base + (uint32_t) __popc(occ & ((1u << lane) - 1u));
            if (e < (uint32_t) UMAX) { sm.lq[e] = (IdxT) ((v >> HB) - 1); sm.lc[e] = (IdxT) (v & LOW); }
        }
    }
    __syncthreads();
    const uint32_t U = *s_U;
    const uint32_t ov = *s_overflow;
    __syncthreads();   // everybody has read the two words before the next round resets them
    return (ov || U > (uint32_t) UMAX) ? ~0u : U;
}

// Phase 3+4: order the U distinct records (bins over the tag, rank inside the bin) and write them (+ counts) to out[obase ...).
template<int W, int MODE, typename IdxT, bool SEG>
__device__ GH_PHASE void gh_emit(unsigned char *raw, const GhAddr<SEG> ga, uint32_t U, int shift2, uint64_t lw_keep, int lgR,
                                     uint64_t *__restrict__ out, uint32_t *__restrict__ out_cnt, unsigned long long obase, uint32_t *s_wtot) {
    const uint64_t *__restrict__ g = ga.recs;
    constexpr int THREADS = HashCfg::THREADS, BINS = HashCfg::BINS;
    const GhSmem<IdxT> sm = gh_views<IdxT>(raw);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (uint32_t i = threadIdx.x; i < (uint32_t) BINS + 1; i += THREADS) sm.binstart[i] = 0;
    __syncthreads();
    const int bshift = 22 - lgR;   // bin = the 10 tag bits below the lgR bits that select the round
    for (uint32_t e0 = 0; e0 < U; e0 += THREADS * 4) {   // four first-occurrence look-ups in flight per thread
        uint64_t k[4][W];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const uint32_t e = e0 + i * THREADS + threadIdx.x;
#pragma unroll
            for (int j = 0; j < W; ++j) k[i][j] = 0;
            if (e < U) gh_load<W>(g, ga.phys(sm.lq[e]), k[i]);
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const uint32_t e = e0 + i * THREADS + threadIdx.x;
            if (e < U) {
                const uint32_t t = seg_tag(k[i][0], shift2);
                sm.ltag[e] = t;
                atomicAdd(&sm.binstart[(t >> bshift) & (BINS - 1)], 1u);
            }
        }
    }
    __syncthreads();
    {
        constexpr int PER = BINS / THREADS;
        uint32_t c[PER], sum = 0;
#pragma unroll
        for (int i = 0; i < PER; ++i) { c[i] = sm.binstart[threadIdx.x * PER + i]; sum += c[i]; }
        const uint32_t inc = warp_inclusive_scan(sum);
        if (lane == 31) s_wtot[warp] = inc;
        __syncthreads();
        uint32_t ex = inc - sum, tot = 0;
#pragma unroll
        for (int w = 0; w < THREADS / 32; ++w) {
            const uint32_t x = s_wtot[w];
            if (w < warp) ex += x;
            tot += x;
        }
#pragma unroll
        for (int i = 0; i < PER; ++i) { sm.binstart[threadIdx.x * PER + i] = ex; sm.cursor[threadIdx.x * PER + i] = 0; ex += c[i]; }
        if (threadIdx.x == 0) sm.binstart[BINS] = tot;
    }
    __syncthreads();
    for (uint32_t e = threadIdx.x; e < U; e += THREADS) {
        const uint32_t d = (sm.ltag[e] >> bshift) & (BINS - 1);
        sm.binned[sm.binstart[d] + atomicAdd(&sm.cursor[d], 1u)] = (uint16_t) e;
    }
    __syncthreads();
    for (uint32_t e0 = 0; e0 < U; e0 += THREADS * 4) {
        uint64_t me[4][W];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const uint32_t e = e0 + i * THREADS + threadIdx.x;
#pragma unroll
            for (int j = 0; j < W; ++j) me[i][j] = 0;
            if (e < U) gh_load<W>(g, ga.phys(sm.lq[e]), me[i]);
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const uint32_t e = e0 + i * THREADS + threadIdx.x;
            if (e >= U) continue;
            const uint32_t t = sm.ltag[e];
            const uint32_t d = (t >> bshift) & (BINS - 1);
            if (MODE == 2) me[i][W - 1] &= lw_keep;
            uint32_t less = 0;
            const uint32_t i0 = sm.binstart[d], i1 = sm.binstart[d + 1];
            for (uint32_t x = i0; x < i1; ++x) {
                const uint32_t o = sm.binned[x];
                const uint32_t to = sm.ltag[o];
                if (to < t) {
                    ++less;
                } else if (to == t && o != e) {   // tag tie: two distinct records agree on the upper part of word 0
                    uint64_t other[W];
                    gh_load<W>(g, ga.phys(sm.lq[o]), other);
                    if (MODE == 2) other[W - 1] &= lw_keep;
                    less += rec_less_bf<W>(other, me[i]);
                }
            }
            const unsigned long long pos = obase + i0 + less;
            store_rec<W>(out, pos, me[i]);
            if (MODE != 0) out_cnt[pos] = sm.lc[e];
        }
    }
    __syncthreads();   // the next round reuses every array
}

// Parts of the groups of a hash-sharded set: part (s, g) = cnt[s * n_groups + g] records starting at record off[s * n_groups + g]
struct GroupParts {
    const uint32_t *cnt;
    const uint32_t *off;
    uint32_t n_src;
    uint32_t n_groups;
};

// shift2: the group prefix ends at bit `shift2` of word 0; the tag is the 32 bits below it.
// MODE 0 records only, 1 + multiplicities, 2 + OR of the 3-bit mask payload at bit pshift of the last word (count.cu derive_kernel).
// SEG: the group's records lie in parts.n_src runs (GhAddr); ranges[b] then only places the OUTPUT (unique records of group b go to
// out[ranges[b].s ...), ranges[b].e - ranges[b].s = records of the group).
template<int W, int MODE, typename IdxT, bool SEG>
__global__ void __launch_bounds__(HashCfg::THREADS, GH_MIN_BLOCKS)
group_hash_kernel(const uint64_t *__restrict__ recs, const ChunkRange *__restrict__ ranges, uint32_t *__restrict__ group_unique,
                  uint32_t *__restrict__ ctrl, uint64_t *__restrict__ out, uint32_t *__restrict__ out_cnt, int shift2, uint64_t lw_keep, int pshift,
                  const uint8_t *__restrict__ pay /* MODE 2: payload bytes beside the records, or nullptr (payload in the padding bits) */,
                  GroupParts parts) {
    constexpr bool SMALL = sizeof(IdxT) == 2;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ uint32_t s_U, s_overflow;
    __shared__ uint32_t s_wtot[HashCfg::THREADS / 32];
    __shared__ uint32_t s_base[GH_MAX_SRC + 1], s_off[GH_MAX_SRC];

    const uint32_t b = blockIdx.x;
    const ChunkRange cr = ranges[b];
    const uint32_t s = cr.s, gcnt = cr.e - cr.s;
    if (SMALL ? (gcnt >= 65535u) : (gcnt < 65535u)) {   // the other instance's groups
        if (SMALL && threadIdx.x == 0) { group_unique[b] = 0; atomicExch(&ctrl[1], 1u); }
        return;
    }
    if (gcnt == 0) {
        if (threadIdx.x == 0) group_unique[b] = 0;
        return;
    }
    GhAddr<SEG> ga;
    if (SEG) {
        if (threadIdx.x == 0) {
            uint32_t acc = 0;
            for (uint32_t q = 0; q < parts.n_src; ++q) {
                s_base[q] = acc;
                s_off[q] = parts.off[(uint64_t) q * parts.n_groups + b];
                acc += parts.cnt[(uint64_t) q * parts.n_groups + b];
            }
            s_base[parts.n_src] = acc;
        }
        __syncthreads();
        ga.recs = recs; ga.pay = (MODE == 2) ? pay : nullptr; ga.base = s_base; ga.off = s_off; ga.n_src = parts.n_src;
    } else {
        ga.recs = recs + (uint64_t) s * W; ga.pay = (MODE == 2 && pay) ? pay + s : nullptr; ga.base = nullptr; ga.off = nullptr; ga.n_src = 1;
    }

    uint32_t emitted = 0;
    int lgR = 0;
    while (true) {                 // R = 1 << lgR rounds over disjoint tag ranges
        emitted = 0;
        bool bad = false;
        const uint32_t R = 1u << lgR;
        for (uint32_t round = 0; round < R; ++round) {
            const uint32_t U = gh_dedup<W, MODE, IdxT, SEG>(smem_raw, ga, gcnt, shift2, lw_keep, pshift, lgR, round, &s_U, &s_overflow);
            if (U == ~0u) { bad = true; break; }   // block-uniform
            gh_emit<W, MODE, IdxT, SEG>(smem_raw, ga, U, shift2, lw_keep, lgR, out, out_cnt, (unsigned long long) s + emitted, s_wtot);
            emitted += U;
        }
        if (!bad) break;
        if (lgR == 8) {   // a single 8-bit digit still holds more than UMAX distinct records
            if (threadIdx.x == 0) { group_unique[b] = 0; atomicExch(&ctrl[0], 1u); }
            return;
        }
        ++lgR;
    }
    if (threadIdx.x == 0) group_unique[b] = emitted;
}

}  // namespace sb200
