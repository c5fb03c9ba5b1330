// mphf.cuh — device-side lookup in the segmented BooPHF (KMerIndex::seq_idx, C/utils/kmer_mph/kmer_index.hpp:85-90;
// boomphf::mphf::lookup / getLevel / bitVector::rank, E/boomphf/BooPHF.h:465-487,609-623,303-314).
#pragma once
#include "kmer_ops.cuh"

namespace sb200 {

constexpr int MPHF_LEVELS = 25;   // boomphf default nb_levels; level 24 is the exact map (never reached in practice)

struct MphfDev {
    const uint64_t *domain;          // [B*25] bits per level
    const uint64_t *word_off;        // [B*25] first word of the level's bit-vector
    const uint64_t *rank_off;        // [B*25] first rank sample of the level
    const uint64_t *segment_starts;  // [B+1]
    const uint64_t *bits;
    const uint64_t *ranks;
    const uint32_t *pc_scan;         // set bits before every word of `bits` (whole-table builds on one GPU), or nullptr
    const uint64_t *wblk;            // walk blocks (below), or nullptr
    uint32_t num_buckets;
};

// Walk blocks: what a walk step needs from the index AND the extension masks in ONE 128-byte line — the unit DRAM serves a random
// access in anyway.  Line b holds words 3b..3b+2 of the index's bit string (bytes 0-23), the number of set bits before them
// (bytes 24-27; bit 31: the line places more than WB_MASKS keys, their masks stay in the mask array) and the mask bytes of the keys
// those 192 bits place, in bit order (bytes 28-127) — MPHF indices follow the bit order, so they are a contiguous run of the mask array.
// A step of a lookup walk then costs one DRAM access instead of three (bit-vector word, rank, mask byte): what bounds the walks of a
// shard, whose index (all GPUs' buckets) is far larger than L2.
constexpr uint32_t WB_WORDS = 3;
constexpr uint32_t WB_MASKS = 100;
constexpr uint32_t WB_LINE_WORDS = 16;

// XorshiftHashFunctors::next (BooPHF.h:94-100)
DEVINL uint64_t xs_next(uint64_t &s0, uint64_t &s1) {
    uint64_t a = s0;
    const uint64_t b = s1;
    s0 = b;
    a ^= a << 23;
    s1 = a ^ b ^ (a >> 17) ^ (b >> 26);
    return s1 + b;
}

#ifdef __CUDACC__
// Returns the global index (segment start + rank) or ~0 if the key falls through all bit levels.
template<int W>
__device__ __forceinline__ uint64_t mphf_lookup(const MphfDev &m, const uint64_t *rec) {
    uint32_t b = kmer_bucket<W>(rec, m.num_buckets);
    uint64_t s0, s1;
    xxh3_128<W>(rec, s0, s1);   // {high64, low64}: level 0 uses high, level 1 low (kmer_index.hpp:38-39)
    const uint64_t *dom = m.domain + (uint64_t) b * MPHF_LEVELS;
    for (int l = 0; l < MPHF_LEVELS - 1; ++l) {
        uint64_t h = (l == 0) ? s0 : (l == 1) ? s1 : xs_next(s0, s1);
        uint64_t d = __ldg(dom + l);
        if (d == 0) return ~0ULL;   // empty bucket: the reference's mphf is "not built"
        uint64_t pos = __umul64hi(h, d);
        const uint64_t *bv = m.bits + __ldg(m.word_off + (uint64_t) b * MPHF_LEVELS + l);
        uint64_t word = __ldg(bv + (pos >> 6));
        if ((word >> (pos & 63)) & 1ULL) {
            if (m.pc_scan) {   // the buckets' bit-vectors follow each other and a bucket sets as many bits as it has keys: global rank = index
                const uint64_t gw = __ldg(m.word_off + (uint64_t) b * MPHF_LEVELS + l) + (pos >> 6);
                return (uint64_t) __ldg(m.pc_scan + gw) + __popcll(word & ((1ULL << (pos & 63)) - 1ULL));
            }
            uint64_t r = __ldg(m.ranks + __ldg(m.rank_off + (uint64_t) b * MPHF_LEVELS + l) + (pos >> 9));
            uint64_t w0 = (pos >> 9) << 3, w1 = pos >> 6;
            for (uint64_t w = w0; w < w1; ++w) r += __popcll(__ldg(bv + w));
            r += __popcll(word & ((1ULL << (pos & 63)) - 1ULL));
            return __ldg(m.segment_starts + b) + r;
        }
    }
    return ~0ULL;
}

// Extension mask of a stored (canonical) key through the walk blocks; 0 if the key falls through all bit levels.
template<int W>
__device__ __forceinline__ uint32_t mphf_lookup_mask(const MphfDev &m, const uint8_t *__restrict__ masks, const uint64_t *rec) {
    const uint32_t b = kmer_bucket<W>(rec, m.num_buckets);
    uint64_t s0, s1;
    xxh3_128<W>(rec, s0, s1);
    const uint64_t *dom = m.domain + (uint64_t) b * MPHF_LEVELS;
    const uint64_t *wo = m.word_off + (uint64_t) b * MPHF_LEVELS;
    for (int l = 0; l < MPHF_LEVELS - 1; ++l) {
        const uint64_t h = (l == 0) ? s0 : (l == 1) ? s1 : xs_next(s0, s1);
        const uint64_t d = __ldg(dom + l);
        if (d == 0) return 0u;
        const uint64_t pos = __umul64hi(h, d);
        const uint32_t gw = (uint32_t) (__ldg(wo + l) + (pos >> 6));
        const uint32_t line = gw / WB_WORDS, j = gw - line * WB_WORDS;
        const uint64_t *lp = m.wblk + (uint64_t) line * WB_LINE_WORDS;
        const uint64_t word = __ldg(lp + j);
        if ((word >> (pos & 63)) & 1ULL) {
            uint32_t r = (uint32_t) __popcll(word & ((1ULL << (pos & 63)) - 1ULL));
            if (j > 0) r += (uint32_t) __popcll(__ldg(lp));
            if (j > 1) r += (uint32_t) __popcll(__ldg(lp + 1));
            const uint32_t hdr = __ldg(reinterpret_cast<const uint32_t *>(lp) + 2 * WB_WORDS);
            if (hdr >> 31) return __ldg(masks + (hdr & 0x7FFFFFFFu) + r);
            return __ldg(reinterpret_cast<const uint8_t *>(lp) + 8 * WB_WORDS + 4 + r);
        }
    }
    return 0u;
}

// Index of a k-mer in reading orientation: canonicalise, look up (InvertableKeyWithHash::CountIdx,
// C/utils/ph_map/key_with_hash.hpp:119-127).  *minimal tells whether x itself is the stored form.
template<int W>
__device__ __forceinline__ uint64_t mphf_lookup_oriented(const MphfDev &m, const uint64_t *x, int k, bool *minimal) {
    uint64_t c[W];
    *minimal = kmer_canonical<W>(x, k, c);
    return mphf_lookup<W>(m, c);
}
#endif

}  // namespace sb200
