// kmer_ops.cuh — device-side k-mer arithmetic and hashing, word-parallel.
//
// Record layout (reference: C/sequence/rtseq.hpp:34-131): a K-mer is W = ceil(K/32) little-endian uint64 words,
// base i (A,C,G,T = 0..3) at bits 2(i%32) of word i/32, padding bits zero.  Record order is word-wise unsigned
// lexicographic from word 0 (C/adt/array_vector.hpp:247-256), canonical form is the base-wise lexicographic minimum
// of {x, rc(x)} starting at base 0 (rtseq.hpp:407-415).  Everything here works on whole words: reverse complement
// is a 2-bit-group bit reversal (__brevll + pair swap) of the words in reverse order followed by a funnel shift, and
// the canonical test compares group-reversed words, so cost is O(W), not O(K).
//
// Hashes: XXH3 (xxHash 0.8.0, E/xxh/xxhash.h) specialised for the only input lengths that occur on this path
// (8, 16, 24, 32 bytes), seed 0, default secret: 64-bit for the bucket policy (C/utils/kmer_mph/kmer_buckets.hpp:28-41),
// 128-bit for the BooPHF levels (C/utils/kmer_mph/kmer_index.hpp:36-40).
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define DEVINL __host__ __device__ __forceinline__
#else
#define DEVINL inline
#endif

// Intrinsic wrappers: the same source compiles for the host (unit tests of the arithmetic against the oracle,
// tests/test_host_primitives.py) and for sm_100a.
DEVINL uint64_t mulhi64(uint64_t a, uint64_t b) {
#ifdef __CUDA_ARCH__
    return __umul64hi(a, b);
#else
    return (uint64_t) (((unsigned __int128) a * b) >> 64);
#endif
}
DEVINL uint64_t brev64(uint64_t v) {
#ifdef __CUDA_ARCH__
    return __brevll(v);
#else
    v = ((v >> 1) & 0x5555555555555555ULL) | ((v & 0x5555555555555555ULL) << 1);
    v = ((v >> 2) & 0x3333333333333333ULL) | ((v & 0x3333333333333333ULL) << 2);
    v = ((v >> 4) & 0x0F0F0F0F0F0F0F0FULL) | ((v & 0x0F0F0F0F0F0F0F0FULL) << 4);
    return __builtin_bswap64(v);
#endif
}
DEVINL uint64_t ldg64(const uint64_t *p) {
#ifdef __CUDA_ARCH__
    return __ldg(p);
#else
    return *p;
#endif
}

template<int W>
struct Rec {
    uint64_t w[W];
};

// ---- XXH3 secret words (XXH_readLE64(kSecret + off) for the offsets used below; xxhash.h:2513-2526) -------------------
#define XS_0   0xbe4ba423396cfeb8ULL
#define XS_8   0x1cad21f72c81017cULL
#define XS_16  0xdb979083e96dd4deULL
#define XS_24  0x1f67b3b7a4a44072ULL
#define XS_32  0x78e5c0cc4ee679cbULL
#define XS_40  0x2172ffcc7dd05a82ULL
#define XS_48  0x8e2443f7744608b8ULL
#define XS_56  0x4c263a81e69035e0ULL
#define XP64_1 0x9E3779B185EBCA87ULL
#define XP64_2 0xC2B2AE3D27D4EB4FULL
#define XP64_4 0x85EBCA77C2B2AE63ULL

DEVINL uint64_t xxh_mul128_fold64(uint64_t a, uint64_t b) { return (a * b) ^ mulhi64(a, b); }
DEVINL uint64_t xxh3_avalanche(uint64_t h) {
    h ^= h >> 37;
    h *= 0x165667919E3779F9ULL;
    h ^= h >> 32;
    return h;
}
DEVINL uint64_t rotl64(uint64_t x, int r) { return (x << r) | (x >> (64 - r)); }
DEVINL uint64_t bswap64(uint64_t x) {
#ifdef __CUDA_ARCH__
    uint32_t lo = (uint32_t) x, hi = (uint32_t) (x >> 32);
    return ((uint64_t) __byte_perm(lo, 0, 0x0123) << 32) | (uint64_t) __byte_perm(hi, 0, 0x0123);
#else
    return __builtin_bswap64(x);
#endif
}

// XXH3_64bits_withSeed(data, 8*W, 0)
template<int W>
DEVINL uint64_t xxh3_64(const uint64_t *w) {
    if (W == 1) {   // XXH3_len_4to8_64b, xxhash.h:2781-2794 (len == 8 dispatches here, :2727-2728)
        uint64_t in64 = (w[0] >> 32) | (w[0] << 32);
        uint64_t h = in64 ^ (XS_8 ^ XS_16);
        h ^= rotl64(h, 49) ^ rotl64(h, 24);   // XXH3_rrmxmx
        h *= 0x9FB21C651E98DF25ULL;
        h ^= (h >> 35) + 8;
        h *= 0x9FB21C651E98DF25ULL;
        return h ^ (h >> 28);
    } else if (W == 2) {   // XXH3_len_9to16_64b, xxhash.h:2797-2811
        uint64_t lo = w[0] ^ (XS_24 ^ XS_32);
        uint64_t hi = w[1] ^ (XS_40 ^ XS_48);
        uint64_t acc = 16 + bswap64(lo) + hi + xxh_mul128_fold64(lo, hi);
        return xxh3_avalanche(acc);
    } else {   // XXH3_len_17to128_64b with len <= 32, xxhash.h:2884-2909
        uint64_t acc = (uint64_t) (8 * W) * XP64_1;
        acc += xxh_mul128_fold64(w[0] ^ XS_0, w[1] ^ XS_8);
        acc += xxh_mul128_fold64(w[W - 2] ^ XS_16, w[W - 1] ^ XS_24);
        return xxh3_avalanche(acc);
    }
}

// XXH3_128bits(data, 8*W): returns {high64, low64} — the order KMerIndex::hash_function128 hands to BooPHF.
template<int W>
DEVINL void xxh3_128(const uint64_t *w, uint64_t &hi_out, uint64_t &lo_out) {
    if (W == 1) {   // XXH3_len_4to8_128b, xxhash.h:4271-4296
        uint64_t keyed = w[0] ^ (XS_16 ^ XS_24);
        uint64_t mul = XP64_1 + (8ull << 2);
        uint64_t lo = keyed * mul, hi = mulhi64(keyed, mul);
        hi += lo << 1;
        lo ^= hi >> 3;
        lo ^= lo >> 35;
        lo *= 0x9FB21C651E98DF25ULL;
        lo ^= lo >> 28;
        hi_out = xxh3_avalanche(hi);
        lo_out = lo;
    } else if (W == 2) {   // XXH3_len_9to16_128b, xxhash.h:4298-4331
        uint64_t bitflipl = XS_32 ^ XS_40, bitfliph = XS_48 ^ XS_56;
        uint64_t ilo = w[0], ihi = w[1];
        uint64_t a = ilo ^ ihi ^ bitflipl;
        uint64_t mlo = a * XP64_1, mhi = mulhi64(a, XP64_1);
        mlo += (uint64_t) 15 << 54;
        ihi ^= bitfliph;
        mhi += ihi + (uint64_t) (uint32_t) ihi * (uint64_t) (0x85EBCA77U - 1);
        mlo ^= bswap64(mhi);
        uint64_t hlo = mlo * XP64_2, hhi = mulhi64(mlo, XP64_2);
        hhi += mhi * XP64_2;
        lo_out = xxh3_avalanche(hlo);
        hi_out = xxh3_avalanche(hhi);
    } else {   // XXH3_len_17to128_128b with len <= 32: a single XXH128_mix32B, xxhash.h:4388-4432
        uint64_t len = 8 * W;
        uint64_t alo = len * XP64_1, ahi = 0;
        alo += xxh_mul128_fold64(w[0] ^ XS_0, w[1] ^ XS_8);
        alo ^= w[W - 2] + w[W - 1];
        ahi += xxh_mul128_fold64(w[W - 2] ^ XS_16, w[W - 1] ^ XS_24);
        ahi ^= w[0] + w[1];
        uint64_t hlo = alo + ahi;
        uint64_t hhi = alo * XP64_1 + ahi * XP64_4 + len * XP64_2;
        lo_out = xxh3_avalanche(hlo);
        hi_out = (uint64_t) 0 - xxh3_avalanche(hhi);
    }
}

// KMerSegmentPolicy (kmer_buckets.hpp:28-41): bucket = mulhi64(XXH3_64(record), num_buckets)
template<int W>
DEVINL uint32_t kmer_bucket(const uint64_t *w, uint32_t num_buckets) {
    if (num_buckets == 1) return 0;
    return (uint32_t) mulhi64(xxh3_64<W>(w), (uint64_t) num_buckets);
}

// ---- 2-bit group helpers -------------------------------------------------------------------------------------------
// reverse the order of the 32 bases of a word (base 0 <-> base 31)
DEVINL uint64_t rev2(uint64_t v) {
    v = brev64(v);
    return ((v >> 1) & 0x5555555555555555ULL) | ((v & 0x5555555555555555ULL) << 1);
}

// mask of the significant bits of the last word of a K-mer
DEVINL uint64_t last_word_mask(int K) {
    int r = K & 31;
    return r ? ((1ULL << (2 * r)) - 1) : ~0ULL;
}

// reverse complement (rtseq.hpp:79-115 computes the same value nucleotide-block-wise)
template<int W>
DEVINL void kmer_rc(const uint64_t *x, int K, uint64_t *out) {
    uint64_t y[W];
#pragma unroll
    for (int j = 0; j < W; ++j) y[j] = ~rev2(x[W - 1 - j]);   // complement = bitwise not of the 2-bit code
    int s = 2 * (32 * W - K);                                 // 0 <= s < 64
#pragma unroll
    for (int j = 0; j < W; ++j) {
        uint64_t lo = y[j] >> s;
        uint64_t hi = (j + 1 < W && s) ? (y[j + 1] << (64 - s)) : 0;
        out[j] = lo | hi;
    }
    out[W - 1] &= last_word_mask(K);
}

// x <= rc(x) base-wise from base 0  (RtSeq::IsMinimal)
template<int W>
DEVINL bool kmer_le_lex(const uint64_t *x, const uint64_t *r) {
#pragma unroll
    for (int j = 0; j < W; ++j) {
        uint64_t a = rev2(x[j]), b = rev2(r[j]);
        if (a != b) return a < b;
    }
    return true;
}

template<int W>
DEVINL bool kmer_eq(const uint64_t *a, const uint64_t *b) {
    bool e = true;
#pragma unroll
    for (int j = 0; j < W; ++j) e &= (a[j] == b[j]);
    return e;
}

// record order: word-wise from word 0 (array_less)
template<int W>
DEVINL bool rec_less(const uint64_t *a, const uint64_t *b) {
#pragma unroll
    for (int j = 0; j < W; ++j) {
        if (a[j] != b[j]) return a[j] < b[j];
    }
    return false;
}

// canonical form; returns true if x itself is minimal (out = x), false if out = rc(x).
// RtSeq::IsMinimal (rtseq.hpp:407-415) compares x and rc(x) base by base from base 0, i.e. it orders the packed integers of the
// REVERSED sequences.  reverse(rc(x)) is the complement of x and reverse(x) the complement of rc(x), and complementing (bitwise not
// inside the K-mer's bits) reverses the integer order, so
//     x <=lex rc(x)   <=>   ~rc(x) <= ~x   <=>   rc(x) >= x   as multi-word integers (word W-1 most significant):
// one word-wise compare of values that are there anyway, instead of four more 2-bit-group reversals (kmer_le_lex above is kept as
// the literal form; tests/test_host_primitives.py checks both against the oracle).
template<int W>
DEVINL bool kmer_ge_num(const uint64_t *a, const uint64_t *b) {
    bool ge = true;   // equal so far, scanning from the least significant word: a more significant word overrides
#pragma unroll
    for (int j = 0; j < W; ++j) ge = (a[j] > b[j]) | ((a[j] == b[j]) & ge);
    return ge;
}

template<int W>
DEVINL bool kmer_canonical(const uint64_t *x, int K, uint64_t *out) {
    uint64_t r[W];
    kmer_rc<W>(x, K, r);
    bool minimal = kmer_ge_num<W>(r, x);
#pragma unroll
    for (int j = 0; j < W; ++j) out[j] = minimal ? x[j] : r[j];
    return minimal;
}

// K-window starting at base `pos` of a packed sequence of `nw` words (word-aligned at seq[0])
template<int W>
DEVINL void kmer_window(const uint64_t *__restrict__ seq, uint32_t nw, uint32_t pos, int K, uint64_t *out) {
    uint32_t q = pos >> 5;
    int s = 2 * (pos & 31);
    uint64_t cur = ldg64(seq + q);
#pragma unroll
    for (int j = 0; j < W; ++j) {
        uint64_t nxt = (q + j + 1 < nw) ? ldg64(seq + q + j + 1) : 0;
        out[j] = s ? ((cur >> s) | (nxt << (64 - s))) : cur;
        cur = nxt;
    }
    out[W - 1] &= last_word_mask(K);
}

// K-window of a record held in registers: `src` has WS words, window starts at base pos (0 or 1 on this path)
template<int WS, int W>
DEVINL void kmer_subwindow(const uint64_t *src, uint32_t pos, int K, uint64_t *out) {
    int s = 2 * (int) pos;   // pos < 32
#pragma unroll
    for (int j = 0; j < W; ++j) {
        uint64_t cur = src[j];
        uint64_t nxt = (j + 1 < WS) ? src[j + 1] : 0;
        out[j] = s ? ((cur >> s) | (nxt << (64 - s))) : cur;
    }
    out[W - 1] &= last_word_mask(K);
}

DEVINL uint32_t kmer_base(const uint64_t *x, int i) { return (uint32_t) (x[i >> 5] >> (2 * (i & 31))) & 3u; }

// x << c : drop base 0, append c at position K-1 (rtseq.hpp:450-467)
template<int W>
DEVINL void kmer_shl(const uint64_t *x, int K, uint32_t c, uint64_t *out) {
#pragma unroll
    for (int j = 0; j < W; ++j) {
        uint64_t nxt = (j + 1 < W) ? x[j + 1] : 0;
        out[j] = (x[j] >> 2) | (nxt << 62);
    }
    out[W - 1] |= (uint64_t) c << (2 * ((K - 1) & 31));   // base K-1 always lives in the last word (constant index: registers)
}

// base i of a W-word k-mer for a run-time i, without dynamic register-array indexing (which would spill x to local memory)
template<int W>
DEVINL uint32_t kmer_base_w(const uint64_t *x, int i) {
    uint64_t w = x[0];
#pragma unroll
    for (int j = 1; j < W; ++j)
        if ((i >> 5) == j) w = x[j];
    return (uint32_t) (w >> (2 * (i & 31))) & 3u;
}

// The window moves one base to the right and takes base c in: x <<= c (rtseq.hpp:450-467) and, for its reverse complement, 3 - c
// enters at base 0 while the last base leaves.  Both in place; lw_mask = last_word_mask(K).
template<int W>
DEVINL void kmer_roll(uint64_t *x, uint64_t *r, int K, uint32_t c, uint64_t lw_mask) {
#pragma unroll
    for (int j = 0; j < W; ++j) x[j] = (x[j] >> 2) | ((j + 1 < W) ? (x[j + 1] << 62) : 0ULL);
    x[W - 1] |= (uint64_t) c << (2 * ((K - 1) & 31));
#pragma unroll
    for (int j = W - 1; j > 0; --j) r[j] = (r[j] << 2) | (r[j - 1] >> 62);
    r[0] = (r[0] << 2) | (uint64_t) (3u - c);
    r[W - 1] &= lw_mask;
}

// The two canonical k-mer candidates of a (k+1)-mer x (DeBruijnKMerKMerSplitter::FillBufferFromKMers, kmer_splitters.hpp:159-176)
// and the InOutMask bit x contributes to each (kmer_extension_index_builder.hpp:44-59, kmer_extension_index.hpp:92-106:
// AddOutgoing(next nucleotide) for the prefix k-mer, AddIncoming(previous nucleotide) for the suffix k-mer, mirrored when the k-mer
// is stored as its reverse complement).  rc(x[0..k)) is the suffix of rc(x) and rc(x[1..k]) its prefix, so ONE reverse complement of
// the (k+1)-mer serves both candidates.
template<int WS, int W>
DEVINL void derive_candidates(const uint64_t *x, int k, uint64_t a[2][W], uint32_t bit[2]) {
    uint64_t rx[WS];
    kmer_rc<WS>(x, k + 1, rx);
    const uint32_t pnucl = kmer_base(x, 0), nnucl = kmer_base_w<WS>(x, k);
    uint64_t f[W], b[W];
    kmer_subwindow<WS, W>(x, 0, k, f);
    kmer_subwindow<WS, W>(rx, 1, k, b);
    bool minimal = kmer_ge_num<W>(b, f);
#pragma unroll
    for (int j = 0; j < W; ++j) a[0][j] = minimal ? f[j] : b[j];
    bit[0] = minimal ? nnucl : 7u - nnucl;
    kmer_subwindow<WS, W>(x, 1, k, f);
    kmer_subwindow<WS, W>(rx, 0, k, b);
    minimal = kmer_ge_num<W>(b, f);
#pragma unroll
    for (int j = 0; j < W; ++j) a[1][j] = minimal ? f[j] : b[j];
    bit[1] = minimal ? pnucl + 4u : 3u - pnucl;
}

// InOutMask::conjugate: bit-reverse the byte (kmer_extension_index.hpp:87)
DEVINL uint32_t mask_conj(uint32_t m) {
#ifdef __CUDA_ARCH__
    return __brev(m) >> 24;
#else
    return (uint32_t) (brev64((uint64_t) m) >> 56);
#endif
}
