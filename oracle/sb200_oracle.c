/* oracle/sb200_oracle.c — TEST INFRASTRUCTURE, NOT PRODUCT CODE.  See sb200_oracle.h.
 *
 * A deliberately plain, single-threaded restatement: nucleotide-at-a-time k-mer arithmetic, qsort, sequential
 * BooPHF.  Nothing here is shared with the CUDA path (spades_for_blackbird_b200/csrc); the point is an independent
 * second implementation that has been pinned byte-for-byte to the unmodified reference (oracle/_ref/ref_driver).
 *
 * Citations: C/ = /root/reference/assembler/src/common/, E/ = /root/reference/assembler/ext/include/.
 */
#include "sb200_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

#define MAXW 4

/* ================================================================================================================
 * XXH3 (xxHash 0.8.0 as vendored in E/xxh/xxhash.h), seed 0, default secret, inputs of 8/16/24/32 bytes only.
 * ============================================================================================================== */
static const uint8_t kSecret[192] = { /* xxhash.h:2513-2526 */
    0xb8, 0xfe, 0x6c, 0x39, 0x23, 0xa4, 0x4b, 0xbe, 0x7c, 0x01, 0x81, 0x2c, 0xf7, 0x21, 0xad, 0x1c,
    0xde, 0xd4, 0x6d, 0xe9, 0x83, 0x90, 0x97, 0xdb, 0x72, 0x40, 0xa4, 0xa4, 0xb7, 0xb3, 0x67, 0x1f,
    0xcb, 0x79, 0xe6, 0x4e, 0xcc, 0xc0, 0xe5, 0x78, 0x82, 0x5a, 0xd0, 0x7d, 0xcc, 0xff, 0x72, 0x21,
    0xb8, 0x08, 0x46, 0x74, 0xf7, 0x43, 0x24, 0x8e, 0xe0, 0x35, 0x90, 0xe6, 0x81, 0x3a, 0x26, 0x4c,
    0x3c, 0x28, 0x52, 0xbb, 0x91, 0xc3, 0x00, 0xcb, 0x88, 0xd0, 0x65, 0x8b, 0x1b, 0x53, 0x2e, 0xa3,
    0x71, 0x64, 0x48, 0x97, 0xa2, 0x0d, 0xf9, 0x4e, 0x38, 0x19, 0xef, 0x46, 0xa9, 0xde, 0xac, 0xd8,
    0xa8, 0xfa, 0x76, 0x3f, 0xe3, 0x9c, 0x34, 0x3f, 0xf9, 0xdc, 0xbb, 0xc7, 0xc7, 0x0b, 0x4f, 0x1d,
    0x8a, 0x51, 0xe0, 0x4b, 0xcd, 0xb4, 0x59, 0x31, 0xc8, 0x9f, 0x7e, 0xc9, 0xd9, 0x78, 0x73, 0x64,
    0xea, 0xc5, 0xac, 0x83, 0x34, 0xd3, 0xeb, 0xc3, 0xc5, 0x81, 0xa0, 0xff, 0xfa, 0x13, 0x63, 0xeb,
    0x17, 0x0d, 0xdd, 0x51, 0xb7, 0xf0, 0xda, 0x49, 0xd3, 0x16, 0x55, 0x26, 0x29, 0xd4, 0x68, 0x9e,
    0x2b, 0x16, 0xbe, 0x58, 0x7d, 0x47, 0xa1, 0xfc, 0x8f, 0xf8, 0xb8, 0xd1, 0x7a, 0xd0, 0x31, 0xce,
    0x45, 0xcb, 0x3a, 0x8f, 0x95, 0x16, 0x04, 0x28, 0xaf, 0xd7, 0xfb, 0xca, 0xbb, 0x4b, 0x40, 0x7e,
};
#define P64_1 0x9E3779B185EBCA87ULL
#define P64_2 0xC2B2AE3D27D4EB4FULL
#define P64_4 0x85EBCA77C2B2AE63ULL
#define P32_2 0x85EBCA77U

static uint64_t sec64(unsigned off) {   /* XXH_readLE64(secret + off) */
    uint64_t v = 0;
    for (int i = 7; i >= 0; --i) v = (v << 8) | kSecret[off + i];
    return v;
}
static uint64_t rotl64(uint64_t x, int r) { return (x << r) | (x >> (64 - r)); }
static uint64_t swap64(uint64_t x) { return __builtin_bswap64(x); }
static uint64_t mul128_fold64(uint64_t a, uint64_t b) {   /* xxhash.h:2684-2688 */
    __uint128_t p = (__uint128_t) a * b;
    return (uint64_t) p ^ (uint64_t) (p >> 64);
}
static uint64_t xxh3_avalanche(uint64_t h) {              /* xxhash.h:2700-2706 */
    h ^= h >> 37;
    h *= 0x165667919E3779F9ULL;
    h ^= h >> 32;
    return h;
}
static uint64_t xxh3_rrmxmx(uint64_t h, uint64_t len) {   /* xxhash.h:2713-2721 */
    h ^= rotl64(h, 49) ^ rotl64(h, 24);
    h *= 0x9FB21C651E98DF25ULL;
    h ^= (h >> 35) + len;
    h *= 0x9FB21C651E98DF25ULL;
    return h ^ (h >> 28);
}
static uint64_t mix16(const uint64_t *in, unsigned soff) { /* XXH3_mix16B, xxhash.h:2871-2879, seed 0 */
    return mul128_fold64(in[0] ^ sec64(soff), in[1] ^ sec64(soff + 8));
}

uint64_t ora_xxh3_64(const uint64_t *w, unsigned n) {
    if (n == 1) {   /* len 8 -> XXH3_len_4to8_64b (dispatch xxhash.h:2727-2728) */
        uint32_t in1 = (uint32_t) w[0], in2 = (uint32_t) (w[0] >> 32);
        uint64_t bitflip = sec64(8) ^ sec64(16);
        uint64_t in64 = (uint64_t) in2 + ((uint64_t) in1 << 32);
        return xxh3_rrmxmx(in64 ^ bitflip, 8);
    }
    if (n == 2) {   /* len 16 -> XXH3_len_9to16_64b */
        uint64_t lo = w[0] ^ (sec64(24) ^ sec64(32));
        uint64_t hi = w[1] ^ (sec64(40) ^ sec64(48));
        uint64_t acc = 16 + swap64(lo) + hi + mul128_fold64(lo, hi);
        return xxh3_avalanche(acc);
    }
    /* len 24 / 32 -> XXH3_len_17to128_64b with len <= 32 */
    uint64_t len = 8ull * n;
    uint64_t acc = len * P64_1;
    acc += mix16(w, 0);
    acc += mix16(w + n - 2, 16);
    return xxh3_avalanche(acc);
}

void ora_xxh3_128(const uint64_t *w, unsigned n, uint64_t *hi_out, uint64_t *lo_out) {
    if (n == 1) {   /* XXH3_len_4to8_128b with len 8, xxhash.h:4271-4296 */
        uint32_t ilo = (uint32_t) w[0], ihi = (uint32_t) (w[0] >> 32);
        uint64_t in64 = (uint64_t) ilo + ((uint64_t) ihi << 32);
        uint64_t keyed = in64 ^ (sec64(16) ^ sec64(24));
        __uint128_t m = (__uint128_t) keyed * (P64_1 + (8ull << 2));
        uint64_t lo = (uint64_t) m, hi = (uint64_t) (m >> 64);
        hi += lo << 1;
        lo ^= hi >> 3;
        lo ^= lo >> 35;
        lo *= 0x9FB21C651E98DF25ULL;
        lo ^= lo >> 28;
        hi = xxh3_avalanche(hi);
        *hi_out = hi; *lo_out = lo;
        return;
    }
    if (n == 2) {   /* XXH3_len_9to16_128b with len 16, xxhash.h:4298-4331 */
        uint64_t bitflipl = sec64(32) ^ sec64(40);
        uint64_t bitfliph = sec64(48) ^ sec64(56);
        uint64_t ilo = w[0], ihi = w[1];
        __uint128_t m = (__uint128_t) (ilo ^ ihi ^ bitflipl) * P64_1;
        uint64_t mlo = (uint64_t) m, mhi = (uint64_t) (m >> 64);
        mlo += (uint64_t) (16 - 1) << 54;
        ihi ^= bitfliph;
        mhi += ihi + (uint64_t) (uint32_t) ihi * (uint64_t) (P32_2 - 1);
        mlo ^= swap64(mhi);
        __uint128_t h = (__uint128_t) mlo * P64_2;
        uint64_t hlo = (uint64_t) h, hhi = (uint64_t) (h >> 64);
        hhi += mhi * P64_2;
        *lo_out = xxh3_avalanche(hlo);
        *hi_out = xxh3_avalanche(hhi);
        return;
    }
    /* XXH3_len_17to128_128b with 16 < len <= 32, xxhash.h:4404-4432: one XXH128_mix32B(acc, in, in+len-16, secret) */
    uint64_t len = 8ull * n;
    uint64_t alo = len * P64_1, ahi = 0;
    const uint64_t *in1 = w, *in2 = w + n - 2;
    alo += mix16(in1, 0);
    alo ^= in2[0] + in2[1];
    ahi += mix16(in2, 16);
    ahi ^= in1[0] + in1[1];
    uint64_t hlo = alo + ahi;
    uint64_t hhi = alo * P64_1 + ahi * P64_4 + len * P64_2;
    *lo_out = xxh3_avalanche(hlo);
    *hi_out = (uint64_t) 0 - xxh3_avalanche(hhi);
}

/* ================================================================================================================
 * k-mer arithmetic, one nucleotide at a time (rtseq.hpp layout: nucl i at bits 2(i%32) of word i/32)
 * ============================================================================================================== */
static unsigned nwords_for(unsigned k) { return (k + 31) / 32; }
static unsigned getn(const uint64_t *x, unsigned i) { return (unsigned) (x[i >> 5] >> (2 * (i & 31))) & 3u; }
static void setn(uint64_t *x, unsigned i, unsigned c) {
    x[i >> 5] = (x[i >> 5] & ~(3ull << (2 * (i & 31)))) | ((uint64_t) c << (2 * (i & 31)));
}

void ora_kmer_rc(const uint64_t *x, unsigned k, uint64_t *out) {
    uint64_t tmp[MAXW] = {0, 0, 0, 0};
    for (unsigned i = 0; i < k; ++i) setn(tmp, i, 3u - getn(x, k - 1 - i));
    memcpy(out, tmp, nwords_for(k) * 8);
}

int ora_kmer_is_minimal(const uint64_t *x, unsigned k) {   /* rtseq.hpp:407-415 */
    for (unsigned i = 0; (i << 1) + 1 <= k; ++i) {
        unsigned front = getn(x, i), end = 3u - getn(x, k - 1 - i);
        if (front != end) return front < end;
    }
    return 1;
}

unsigned ora_bucket(const uint64_t *x, unsigned nwords, unsigned num_buckets) {   /* kmer_buckets.hpp:28-41 */
    if (num_buckets == 1) return 0;
    return (unsigned) (((__uint128_t) ora_xxh3_64(x, nwords) * num_buckets) >> 64);
}

/* window of length k starting at base `pos` of a packed sequence (word-aligned at seq[0]) */
static void window(const uint64_t *seq, uint64_t pos, unsigned k, uint64_t *out) {
    memset(out, 0, MAXW * 8);
    for (unsigned i = 0; i < k; ++i) {
        uint64_t p = pos + i;
        setn(out, i, (unsigned) (seq[p >> 5] >> (2 * (p & 31))) & 3u);
    }
}

/* ================================================================================================================
 * read packing
 * ============================================================================================================== */
static int is_nucl(char c) {   /* C/sequence/nucl.hpp:45-62 (numeric 0..3 never occurs in text input) */
    switch (c) { case 'a': case 'A': case 'c': case 'C': case 'g': case 'G': case 't': case 'T': return 1; default: return 0; }
}
static unsigned dignucl(char c) {   /* nucl.hpp:120-130 */
    if (c >= 'a' && c <= 't') c = (char) (c - 'a' + 'A');
    return c <= 'C' ? (c == 'A' ? 0u : 1u) : (c == 'G' ? 2u : 3u);
}

uint64_t ora_pack_reads(const char *ascii, const uint64_t *off, uint64_t n_reads,
                        uint64_t *words_out, uint64_t *word_off_out, uint32_t *len_out) {
    uint64_t wpos = 0;
    for (uint64_t r = 0; r < n_reads; ++r) {
        const char *s = ascii + off[r];
        uint64_t sz = off[r + 1] - off[r];
        /* longest_valid_wrapper.hpp:15-42 */
        uint64_t best_len = 0, best_pos = 0, pos = (uint64_t) -1;
        for (uint64_t i = 0; i <= sz; ++i) {
            if (i < sz && is_nucl(s[i])) {
                if (pos == (uint64_t) -1) pos = i;
            } else {
                if (pos != (uint64_t) -1 && i - pos > best_len) { best_len = i - pos; best_pos = pos; }
                pos = (uint64_t) -1;
            }
        }
        uint64_t nw = (best_len + 31) / 32;
        if (words_out) {
            for (uint64_t w = 0; w < nw; ++w) words_out[wpos + w] = 0;
            for (uint64_t i = 0; i < best_len; ++i)
                words_out[wpos + (i >> 5)] |= (uint64_t) dignucl(s[best_pos + i]) << (2 * (i & 31));
            word_off_out[r] = wpos;
            len_out[r] = (uint32_t) best_len;
        }
        wpos += nw;
    }
    if (words_out) word_off_out[n_reads] = wpos;
    return wpos;
}

/* ================================================================================================================
 * k-mer sets: (bucket, array_less) order == KMerDiskStorage file order
 * ============================================================================================================== */
struct ora_kmers {
    unsigned k, words, num_buckets;
    uint64_t size;
    uint64_t *data;
    uint64_t *bucket_starts;
    uint32_t *counts;
};

typedef struct { uint64_t w[MAXW]; uint32_t bucket; } rec_t;

static int rec_cmp(const void *pa, const void *pb) {   /* bucket, then C/adt/array_vector.hpp:247-256 */
    const rec_t *a = (const rec_t *) pa, *b = (const rec_t *) pb;
    if (a->bucket != b->bucket) return a->bucket < b->bucket ? -1 : 1;
    for (int i = 0; i < MAXW; ++i)
        if (a->w[i] != b->w[i]) return a->w[i] < b->w[i] ? -1 : 1;
    return 0;
}

typedef struct { rec_t *v; uint64_t n, cap; } recvec;
static void rv_push(recvec *rv, const uint64_t *w, unsigned nw, unsigned B) {
    if (rv->n == rv->cap) {
        rv->cap = rv->cap ? rv->cap * 2 : 1024;
        rv->v = (rec_t *) realloc(rv->v, rv->cap * sizeof(rec_t));
    }
    rec_t *r = &rv->v[rv->n++];
    memset(r, 0, sizeof(*r));
    memcpy(r->w, w, nw * 8);
    r->bucket = ora_bucket(w, nw, B);
}

static ora_kmers *finish_set(recvec *rv, unsigned K, unsigned B, int want_counts) {
    if (rv->n == 0) { free(rv->v); return NULL; }
    qsort(rv->v, rv->n, sizeof(rec_t), rec_cmp);
    unsigned W = nwords_for(K);
    ora_kmers *s = (ora_kmers *) calloc(1, sizeof(*s));
    s->k = K; s->words = W; s->num_buckets = B;
    uint64_t u = 0;
    for (uint64_t i = 0; i < rv->n; ++i)
        if (i == 0 || rec_cmp(&rv->v[i], &rv->v[i - 1]) != 0) ++u;
    s->size = u;
    s->data = (uint64_t *) malloc((u ? u : 1) * W * 8);
    s->counts = want_counts ? (uint32_t *) calloc(u ? u : 1, 4) : NULL;
    s->bucket_starts = (uint64_t *) calloc(B + 1, 8);
    uint64_t j = 0;
    for (uint64_t i = 0; i < rv->n; ++i) {
        if (i == 0 || rec_cmp(&rv->v[i], &rv->v[i - 1]) != 0) {
            memcpy(s->data + j * W, rv->v[i].w, W * 8);
            s->bucket_starts[rv->v[i].bucket + 1]++;
            ++j;
        }
        if (want_counts) s->counts[j - 1]++;
    }
    for (unsigned b = 0; b < B; ++b) s->bucket_starts[b + 1] += s->bucket_starts[b];
    free(rv->v);
    return s;
}

/* every K-window of seq[0..len) that passes the filter (kmer_splitters.hpp:25-41) */
static void push_windows(recvec *rv, const uint64_t *seq, uint32_t len, unsigned K, int canonical_only, unsigned B) {
    if (len < K) return;
    uint64_t x[MAXW];
    unsigned W = nwords_for(K);
    for (uint32_t p = 0; p + K <= len; ++p) {
        window(seq, p, K, x);
        if (canonical_only && !ora_kmer_is_minimal(x, K)) continue;
        rv_push(rv, x, W, B);
    }
}

ora_kmers *ora_count_reads(const uint64_t *words, const uint64_t *word_off, const uint32_t *len, uint64_t n_reads,
                           unsigned K, int canonical_only, int add_rc, unsigned B) {
    recvec rv = {0, 0, 0};
    uint64_t *rcbuf = NULL; uint64_t rccap = 0;
    for (uint64_t r = 0; r < n_reads; ++r) {
        const uint64_t *seq = words + word_off[r];
        push_windows(&rv, seq, len[r], K, canonical_only, B);
        if (add_rc && len[r] >= K) {   /* RCWrap: the read's reverse complement as a second read */
            uint64_t nw = ((uint64_t) len[r] + 31) / 32;
            if (nw > rccap) { rccap = nw * 2; rcbuf = (uint64_t *) realloc(rcbuf, rccap * 8); }
            memset(rcbuf, 0, nw * 8);
            for (uint32_t i = 0; i < len[r]; ++i) {
                uint32_t src = len[r] - 1 - i;
                unsigned c = 3u - ((unsigned) (seq[src >> 5] >> (2 * (src & 31))) & 3u);
                rcbuf[i >> 5] |= (uint64_t) c << (2 * (i & 31));
            }
            push_windows(&rv, rcbuf, len[r], K, canonical_only, B);
        }
    }
    free(rcbuf);
    return finish_set(&rv, K, B, 1);
}

ora_kmers *ora_derive_kmers(const ora_kmers *kp, unsigned B) {   /* kmer_splitters.hpp:159-176 */
    recvec rv = {0, 0, 0};
    unsigned K1 = kp->k, K = K1 - 1;
    uint64_t rc[MAXW];
    for (uint64_t i = 0; i < kp->size; ++i) {
        const uint64_t *x = kp->data + i * kp->words;
        uint64_t xx[MAXW] = {0, 0, 0, 0};
        memcpy(xx, x, kp->words * 8);
        push_windows(&rv, xx, K1, K, 1, B);
        memset(rc, 0, sizeof rc);
        ora_kmer_rc(xx, K1, rc);
        push_windows(&rv, rc, K1, K, 1, B);
    }
    return finish_set(&rv, K, B, 0);
}

unsigned        ora_kmers_k(const ora_kmers *s) { return s->k; }
unsigned        ora_kmers_words(const ora_kmers *s) { return s->words; }
unsigned        ora_kmers_num_buckets(const ora_kmers *s) { return s->num_buckets; }
uint64_t        ora_kmers_size(const ora_kmers *s) { return s->size; }
const uint64_t *ora_kmers_data(const ora_kmers *s) { return s->data; }
const uint64_t *ora_kmers_bucket_starts(const ora_kmers *s) { return s->bucket_starts; }
const uint32_t *ora_kmers_counts(const ora_kmers *s) { return s->counts; }
void ora_kmers_free(ora_kmers *s) {
    if (!s) return;
    free(s->data); free(s->bucket_starts); free(s->counts); free(s);
}

/* ================================================================================================================
 * BooPHF, gamma = 4, 25 levels, one per bucket (kmer_index_builder.hpp:397-410)
 * ============================================================================================================== */
#define NLEVELS 25
typedef struct {
    uint64_t nelem;
    uint64_t lastbitsetrank;
    uint64_t domain[NLEVELS];
    uint64_t nchar[NLEVELS];        /* 1 + domain/64, BooPHF.h:144-148 */
    uint64_t *bits[NLEVELS];
    uint64_t nranks[NLEVELS];
    uint64_t *ranks[NLEVELS];
    /* exact map of the last level (BooPHF.h:652-671): insertion order == file order in this single-thread port */
    uint64_t nfinal;
    uint64_t *final_hi, *final_lo, *final_idx;
} boophf;

struct ora_mphf {
    unsigned num_buckets, words;
    boophf *seg;
    uint64_t *segment_starts;   /* num_buckets + 1, with the reference's quirk (see ora_mphf_build) */
    uint64_t total;
};

static uint64_t fastrange64(uint64_t w, uint64_t p) { return (uint64_t) (((__uint128_t) w * p) >> 64); }

/* XorshiftHashFunctors::next, BooPHF.h:94-100 */
static uint64_t xs_next(uint64_t s[2]) {
    uint64_t s1 = s[0];
    const uint64_t s0 = s[1];
    s[0] = s0;
    s1 ^= s1 << 23;
    return (s[1] = (s1 ^ s0 ^ (s1 >> 17) ^ (s0 >> 26))) + s0;
}
static uint64_t iterate_hash(uint64_t s[2], unsigned level) {   /* BooPHF.h:599-606 */
    if (level == 0) return s[0];
    if (level == 1) return s[1];
    return xs_next(s);
}
static int bit_get(const uint64_t *b, uint64_t pos) { return (int) ((b[pos >> 6] >> (pos & 63)) & 1); }
static int bit_test_and_set(uint64_t *b, uint64_t pos) {
    int old = bit_get(b, pos);
    b[pos >> 6] |= 1ull << (pos & 63);
    return old;
}
/* getLevel, BooPHF.h:609-623 */
static uint64_t get_level(const boophf *m, uint64_t hi, uint64_t lo, unsigned *res_level, unsigned maxlevel) {
    uint64_t s[2] = {hi, lo};   /* kmer_index.hpp:38-39: { high64, low64 } */
    unsigned level;
    for (level = 0; level < NLEVELS - 1 && level < maxlevel; ++level) {
        uint64_t h = iterate_hash(s, level);
        if (m->bits[level] && bit_get(m->bits[level], fastrange64(h, m->domain[level]))) {
            *res_level = level;
            return h;
        }
    }
    *res_level = level;
    return iterate_hash(s, level);
}

static void boophf_build(boophf *m, const uint64_t *keys, uint64_t n, unsigned W) {
    memset(m, 0, sizeof(*m));
    m->nelem = n;
    /* init/setup, BooPHF.h:409-421,575-588 */
    double gamma = 4.0;
    uint64_t hash_domain = (uint64_t) ceil((double) n * gamma);
    double p = 1.0 - pow(((gamma * (double) n - 1) / (gamma * (double) n)), (double) n - 1);
    for (unsigned i = 0; i < NLEVELS; ++i) {
        m->domain[i] = (((uint64_t) ((double) hash_domain * pow(p, (double) i)) + 63) / 64) * 64;
        if (m->domain[i] == 0) m->domain[i] = 64;
    }
    if (n == 0) return;   /* build() returns early: default-constructed bit-vectors (BooPHF.h:425-426) */
    uint64_t *hh = (uint64_t *) malloc(n * 16);
    for (uint64_t i = 0; i < n; ++i) ora_xxh3_128(keys + i * W, W, &hh[2 * i], &hh[2 * i + 1]);
    uint64_t offset = 0;
    for (unsigned lv = 0; lv < NLEVELS; ++lv) {
        m->nchar[lv] = 1 + m->domain[lv] / 64;
        m->bits[lv] = (uint64_t *) calloc(m->nchar[lv], 8);
        uint64_t *coll = (uint64_t *) calloc(m->nchar[lv], 8);
        for (uint64_t i = 0; i < n; ++i) {   /* processHash, BooPHF.h:634-675 */
            unsigned level;
            uint64_t h = get_level(m, hh[2 * i], hh[2 * i + 1], &level, lv);
            if (level != lv) continue;
            if (lv == NLEVELS - 1) {
                uint64_t idx = m->nfinal;   /* __sync_fetch_and_add(&_final_hashidx, 1) */
                int dup = 0;
                for (uint64_t f = 0; f < m->nfinal; ++f)
                    if (m->final_hi[f] == hh[2 * i] && m->final_lo[f] == hh[2 * i + 1]) { m->final_idx[f] = (uint64_t) -1; dup = 1; }
                m->final_hi = (uint64_t *) realloc(m->final_hi, (m->nfinal + 1) * 8);
                m->final_lo = (uint64_t *) realloc(m->final_lo, (m->nfinal + 1) * 8);
                m->final_idx = (uint64_t *) realloc(m->final_idx, (m->nfinal + 1) * 8);
                m->final_hi[m->nfinal] = hh[2 * i]; m->final_lo[m->nfinal] = hh[2 * i + 1];
                m->final_idx[m->nfinal] = dup ? (uint64_t) -1 : idx;
                m->nfinal++;
            } else {
                uint64_t pos = fastrange64(h, m->domain[lv]);
                if (bit_test_and_set(m->bits[lv], pos)) bit_test_and_set(coll, pos);
            }
        }
        for (uint64_t w = 0; w < m->domain[lv] / 64; ++w) m->bits[lv][w] &= ~coll[w];   /* clearCollisions :219-228 */
        free(coll);
        /* build_ranks, BooPHF.h:289-301 */
        m->ranks[lv] = (uint64_t *) malloc((m->nchar[lv] / 8 + 2) * 8);
        uint64_t cur = offset, nr = 0;
        for (uint64_t w = 0; w < m->nchar[lv]; ++w) {
            if (((w * 64) % 512) == 0) m->ranks[lv][nr++] = cur;
            cur += (uint64_t) __builtin_popcountll(m->bits[lv][w]);
        }
        m->nranks[lv] = nr;
        offset = cur;
    }
    m->lastbitsetrank = offset;
    free(hh);
}

static uint64_t boophf_lookup(const boophf *m, uint64_t hi, uint64_t lo) {   /* BooPHF.h:465-487,303-314 */
    if (m->nelem == 0) return (uint64_t) -1;   /* !_built */
    unsigned level;
    uint64_t h = get_level(m, hi, lo, &level, NLEVELS);
    if (level == NLEVELS - 1) {
        for (uint64_t f = 0; f < m->nfinal; ++f)
            if (m->final_hi[f] == hi && m->final_lo[f] == lo)
                return m->final_idx[f] != (uint64_t) -1 ? m->final_idx[f] + m->lastbitsetrank : (uint64_t) -1;
        return (uint64_t) -1;
    }
    uint64_t pos = fastrange64(h, m->domain[level]);
    uint64_t word_idx = pos / 64, block = pos / 512;
    uint64_t r = m->ranks[level][block];
    for (uint64_t w = block * 512 / 64; w < word_idx; ++w) r += (uint64_t) __builtin_popcountll(m->bits[level][w]);
    r += (uint64_t) __builtin_popcountll(m->bits[level][word_idx] & ((1ull << (pos % 64)) - 1));
    return r;
}

ora_mphf *ora_mphf_build(const ora_kmers *s) {
    ora_mphf *m = (ora_mphf *) calloc(1, sizeof(*m));
    unsigned B = s->num_buckets;
    m->num_buckets = B; m->words = s->words; m->total = s->size;
    m->seg = (boophf *) calloc(B, sizeof(boophf));
    m->segment_starts = (uint64_t *) calloc(B + 1, 8);
    for (unsigned b = 0; b < B; ++b) {
        uint64_t n = s->bucket_starts[b + 1] - s->bucket_starts[b];
        m->segment_starts[b + 1] = n;
        boophf_build(&m->seg[b], s->data + s->bucket_starts[b] * s->words, n, s->words);
    }
    /* kmer_index_builder.hpp:427-428: the loop stops at i < segments, so the LAST entry stays a bucket size */
    for (unsigned i = 1; i < B; ++i) m->segment_starts[i] += m->segment_starts[i - 1];
    return m;
}

uint64_t ora_mphf_lookup(const ora_mphf *m, const uint64_t *rec) {   /* kmer_index.hpp:85-90 */
    unsigned b = ora_bucket(rec, m->words, m->num_buckets);
    uint64_t hi, lo;
    ora_xxh3_128(rec, m->words, &hi, &lo);
    uint64_t idx = boophf_lookup(&m->seg[b], hi, lo);
    return idx == (uint64_t) -1 ? idx : m->segment_starts[b] + idx;
}

uint64_t ora_mphf_final_level_keys(const ora_mphf *m) {
    uint64_t t = 0;
    for (unsigned b = 0; b < m->num_buckets; ++b) t += m->seg[b].nfinal;
    return t;
}

static void put(uint8_t **p, const void *src, size_t n, uint64_t *total) {
    if (*p) { memcpy(*p, src, n); *p += n; }
    *total += n;
}

uint64_t ora_mphf_serialize(const ora_mphf *m, uint8_t *out) {   /* kmer_index.hpp:99-105; BooPHF.h:514-532,316-323 */
    uint64_t total = 0;
    uint8_t *p = out;
    uint64_t nseg = m->num_buckets;
    put(&p, &nseg, 8, &total);
    for (unsigned b = 0; b < m->num_buckets; ++b) {
        const boophf *h = &m->seg[b];
        double gamma = 4.0;
        int nb_levels = NLEVELS;
        put(&p, &gamma, 8, &total);
        put(&p, &nb_levels, 4, &total);
        put(&p, &h->lastbitsetrank, 8, &total);
        put(&p, &h->nelem, 8, &total);
        for (unsigned lv = 0; lv < NLEVELS; ++lv) {
            uint64_t size = h->bits[lv] ? h->domain[lv] : 0;   /* bitVector::_size */
            uint64_t nchar = h->bits[lv] ? h->nchar[lv] : 0;
            put(&p, &size, 8, &total);
            put(&p, &nchar, 8, &total);
            if (h->bits[lv]) put(&p, h->bits[lv], nchar * 8, &total);
            uint64_t nr = h->nranks[lv];
            put(&p, &nr, 8, &total);
            if (nr) put(&p, h->ranks[lv], nr * 8, &total);
        }
        uint64_t nf = h->nfinal;
        put(&p, &nf, 8, &total);
        for (uint64_t f = 0; f < nf; ++f) {
            put(&p, &h->final_hi[f], 8, &total);
            put(&p, &h->final_lo[f], 8, &total);
            put(&p, &h->final_idx[f], 8, &total);
        }
    }
    put(&p, m->segment_starts, (m->num_buckets + 1) * 8, &total);
    return total;
}

void ora_mphf_free(ora_mphf *m) {
    if (!m) return;
    for (unsigned b = 0; b < m->num_buckets; ++b) {
        for (unsigned lv = 0; lv < NLEVELS; ++lv) { free(m->seg[b].bits[lv]); free(m->seg[b].ranks[lv]); }
        free(m->seg[b].final_hi); free(m->seg[b].final_lo); free(m->seg[b].final_idx);
    }
    free(m->seg); free(m->segment_starts); free(m);
}

/* ================================================================================================================
 * KeyWithHash (C/utils/ph_map/key_with_hash.hpp:108-207): a k-mer in reading orientation + idx of its canonical form
 * ============================================================================================================== */
typedef struct { uint64_t w[MAXW]; uint64_t idx; int minimal; } kwh_t;

static kwh_t make_kwh(const ora_mphf *m, const uint64_t *x, unsigned k) {
    kwh_t r;
    memset(&r, 0, sizeof r);
    memcpy(r.w, x, nwords_for(k) * 8);
    r.minimal = ora_kmer_is_minimal(r.w, k);
    if (r.minimal) r.idx = ora_mphf_lookup(m, r.w);
    else { uint64_t rc[MAXW] = {0, 0, 0, 0}; ora_kmer_rc(r.w, k, rc); r.idx = ora_mphf_lookup(m, rc); }
    return r;
}
static uint8_t invert_byte(uint8_t b) {   /* InOutMask::conjugate, kmer_extension_index.hpp:87 */
    uint8_t r = 0;
    for (int i = 0; i < 8; ++i) if (b & (1 << i)) r |= (uint8_t) (1 << (7 - i));
    return r;
}
static uint8_t get_value(const uint8_t *data, const kwh_t *k) {   /* storing_traits.hpp:45-52 */
    return k->minimal ? data[k->idx] : invert_byte(data[k->idx]);
}
static kwh_t kwh_shl(const ora_mphf *m, const kwh_t *a, unsigned k, unsigned c) {   /* kwh << c */
    uint64_t x[MAXW] = {0, 0, 0, 0};
    for (unsigned i = 0; i + 1 < k; ++i) setn(x, i, getn(a->w, i + 1));
    setn(x, k - 1, c);
    return make_kwh(m, x, k);
}
static kwh_t kwh_rc(const ora_mphf *m, const kwh_t *a, unsigned k) {   /* !kwh */
    uint64_t x[MAXW] = {0, 0, 0, 0};
    ora_kmer_rc(a->w, k, x);
    return make_kwh(m, x, k);
}
static int kwh_eq(const kwh_t *a, const kwh_t *b) { return memcmp(a->w, b->w, sizeof a->w) == 0; }
static const int UNIQUE[16] = {0, 1, 1, 0, 1, 0, 0, 0, 1, 0, 0, 0, 0, 0, 0, 0};
static const int NEXT[16] = {-1, 0, 1, -1, 2, -1, -1, -1, 3, -1, -1, -1, -1, -1, -1, -1};
static const int COUNT[16] = {0, 1, 1, 2, 1, 2, 2, 3, 1, 2, 2, 3, 2, 3, 3, 4};
static int inv_position(int nucl, int as_is) { return as_is ? nucl : 7 - nucl; }

void ora_fill_masks(const ora_kmers *kp, const ora_mphf *m, uint8_t *data) {   /* kmer_extension_index_builder.hpp:44-59 */
    unsigned K1 = kp->k, k = K1 - 1;
    memset(data, 0, m->total);
    for (uint64_t i = 0; i < kp->size; ++i) {
        uint64_t x[MAXW] = {0, 0, 0, 0}, pre[MAXW], suf[MAXW];
        memcpy(x, kp->data + i * kp->words, kp->words * 8);
        unsigned pnucl = getn(x, 0), nnucl = getn(x, K1 - 1);
        window(x, 0, k, pre);
        window(x, 1, k, suf);
        kwh_t a = make_kwh(m, pre, k), b = make_kwh(m, suf, k);
        data[a.idx] |= (uint8_t) (1u << inv_position((int) nnucl, a.minimal));       /* AddOutgoing :92-98 */
        data[b.idx] |= (uint8_t) (1u << inv_position((int) pnucl + 4, b.minimal));   /* AddIncoming :100-106 */
    }
}

/* ================================================================================================================
 * early tip clipper, processed in final_kmers order by one thread
 * ============================================================================================================== */
typedef struct { kwh_t *v; uint64_t n, cap; } kwhvec;
static void kv_push(kwhvec *kv, const kwh_t *k) {
    if (kv->n == kv->cap) { kv->cap = kv->cap ? kv->cap * 2 : 64; kv->v = (kwh_t *) realloc(kv->v, kv->cap * sizeof(kwh_t)); }
    kv->v[kv->n++] = *k;
}

static void find_forward(const ora_mphf *m, const uint8_t *data, unsigned k, kwh_t kh, kwhvec *tip, uint64_t bound) {
    /* early_simplification.hpp:110-121 */
    for (;;) {
        uint8_t mk = get_value(data, &kh);
        if (!(tip->n < bound && UNIQUE[mk >> 4] && UNIQUE[mk & 15])) break;
        kv_push(tip, &kh);
        kh = kwh_shl(m, &kh, k, (unsigned) NEXT[mk & 15]);
    }
    kv_push(tip, &kh);
    uint8_t mk = get_value(data, &kh);
    if (!UNIQUE[mk >> 4] || (mk & 15) != 0) tip->n = 0;
}

uint64_t ora_tipclip(const ora_kmers *s, const ora_mphf *m, uint8_t *data, uint64_t bound) {
    unsigned k = s->k;
    uint64_t removed_total = 0;
    kwhvec tips[4] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
    kwhvec tipped = {0, 0, 0};
    for (uint64_t i = 0; i < s->size; ++i) {
        kwh_t fw = make_kwh(m, s->data + i * s->words, k);
        kwh_t rc = kwh_rc(m, &fw, k);
        for (int o = 0; o < 2; ++o) {   /* early_simplification.hpp:62-73 */
            kwh_t kh = o ? rc : fw;
            uint8_t mask = get_value(data, &kh);
            if (COUNT[mask & 15] < 2) continue;
            /* RemoveForward :141-154 */
            uint64_t max = 0;
            for (unsigned c = 0; c < 4; ++c) {
                tips[c].n = 0;
                if (mask & (1u << c)) {
                    kwh_t khc = kwh_shl(m, &kh, k, c);
                    find_forward(m, data, k, khc, &tips[c], bound);
                    uint64_t len = tips[c].n == 0 ? (uint64_t) -1 : tips[c].n;
                    if (len > max) max = len;
                }
            }
            uint64_t removed = 0;   /* RemoveTips :131-139 */
            for (unsigned c = 0; c < 4; ++c)
                if (tips[c].n < max) {
                    for (uint64_t t = 0; t < tips[c].n; ++t) data[tips[c].v[t].idx] = 0;
                    removed += tips[c].n;
                }
            removed_total += removed;
            if (removed) kv_push(&tipped, &kh);
        }
    }
    for (uint64_t t = 0; t < tipped.n; ++t) {   /* RemoveInconsistentForwardLinks :20-35 */
        kwh_t kh = tipped.v[t];
        uint8_t mask = get_value(data, &kh);
        for (unsigned c = 0; c < 4; ++c) {
            if (!(mask & (1u << c))) continue;
            kwh_t nx = kwh_shl(m, &kh, k, c);
            uint8_t nm = get_value(data, &nx);
            if (!(nm & (1u << (4 + getn(kh.w, 0)))))
                data[kh.idx] &= (uint8_t) ~(1u << inv_position((int) c, kh.minimal));   /* DeleteOutgoing */
        }
    }
    for (int c = 0; c < 4; ++c) free(tips[c].v);
    free(tipped.v);
    return removed_total;
}

/* ================================================================================================================
 * unbranching paths and perfect loops
 * ============================================================================================================== */
struct ora_seqs { uint64_t count, n_loops; uint64_t *off; char *chars; uint64_t cap_seqs, cap_chars; };

typedef struct { uint8_t *v; uint64_t n, cap; } nucvec;   /* one code 0..3 per byte */
static void nv_push(nucvec *nv, unsigned c) {
    if (nv->n == nv->cap) { nv->cap = nv->cap ? nv->cap * 2 : 256; nv->v = (uint8_t *) realloc(nv->v, nv->cap); }
    nv->v[nv->n++] = (uint8_t) c;
}
static int nuc_less_rc(const uint8_t *s, uint64_t n) {   /* s < !s, sequence.hpp:222-230 */
    for (uint64_t i = 0; i < n; ++i) {
        unsigned a = s[i], b = 3u - s[n - 1 - i];
        if (a != b) return a < b;
    }
    return 0;
}
static void emit(ora_seqs *out, const uint8_t *s, uint64_t n, int rc) {
    if (out->count + 1 >= out->cap_seqs) {
        out->cap_seqs = out->cap_seqs ? out->cap_seqs * 2 : 64;
        out->off = (uint64_t *) realloc(out->off, (out->cap_seqs + 1) * 8);
    }
    uint64_t base = out->off[out->count];
    if (base + n + 1 > out->cap_chars) {
        out->cap_chars = (base + n + 1) * 2;
        out->chars = (char *) realloc(out->chars, out->cap_chars);
    }
    for (uint64_t i = 0; i < n; ++i)
        out->chars[base + i] = "ACGT"[rc ? 3u - s[n - 1 - i] : s[i]];
    out->off[++out->count] = base + n;
}
/* CleanCondensed(sequence) + CleanCondensed(!sequence): both strands share the mask entry (:288-304) */
static void clean_condensed(const ora_mphf *m, uint8_t *data, unsigned k, const uint8_t *s, uint64_t n) {
    uint64_t x[MAXW];
    for (uint64_t p = 0; p + k <= n; ++p) {
        memset(x, 0, sizeof x);
        for (unsigned i = 0; i < k; ++i) setn(x, i, s[p + i]);
        kwh_t kh = make_kwh(m, x, k);
        data[kh.idx] = 0;
    }
}
/* ConstructSequenceWithEdge, :236-245 */
static void construct_sequence(const ora_mphf *m, const uint8_t *data, unsigned k, kwh_t start, kwh_t end, nucvec *nv) {
    nv->n = 0;
    for (unsigned i = 0; i < k; ++i) nv_push(nv, getn(start.w, i));
    nv_push(nv, getn(end.w, k - 1));
    kwh_t istart = start, iend = end;
    for (;;) {
        uint8_t mask = get_value(data, &end);   /* StepRightIfPossible :226-234 */
        if (!(UNIQUE[mask & 15] && UNIQUE[mask >> 4])) break;
        kwh_t nx = kwh_shl(m, &end, k, (unsigned) NEXT[mask & 15]);
        start = end; end = nx;
        if (kwh_eq(&start, &istart) && kwh_eq(&end, &iend)) break;
        nv_push(nv, getn(end.w, k - 1));
    }
}

ora_seqs *ora_unitigs(const ora_kmers *s, const ora_mphf *m, uint8_t *data, int with_loops) {
    unsigned k = s->k;
    ora_seqs *out = (ora_seqs *) calloc(1, sizeof(*out));
    out->off = (uint64_t *) calloc(65, 8); out->cap_seqs = 64;
    nucvec nv = {0, 0, 0};
    /* CalculateSequences :267-286 over the whole file (chunks are contiguous and concatenated in order) */
    for (uint64_t i = 0; i < s->size; ++i) {
        kwh_t kh = make_kwh(m, s->data + i * s->words, k);
        uint8_t ext = get_value(data, &kh);
        if (UNIQUE[ext & 15] && UNIQUE[ext >> 4]) continue;   /* not a junction */
        kwh_t inv = kwh_rc(m, &kh, k);
        for (int o = 0; o < 2; ++o) {   /* AddStartDeEdges :212-224 */
            kwh_t v = o ? inv : kh;
            if (o && v.minimal) continue;
            uint8_t mask = get_value(data, &v);
            for (unsigned c = 0; c < 4; ++c) {
                if (!(mask & (1u << c))) continue;
                kwh_t e = kwh_shl(m, &v, k, c);
                construct_sequence(m, data, k, v, e, &nv);
                if (nuc_less_rc(nv.v, nv.n)) continue;
                emit(out, nv.v, nv.n, 0);
            }
        }
    }
    if (with_loops) {
        uint64_t n_paths = out->count;
        /* ExtractUnbranchingPathsAndLoops :377-384: isolate everything on the extracted paths first */
        nucvec tmp = {0, 0, 0};
        for (uint64_t q = 0; q < n_paths; ++q) {
            uint64_t n = out->off[q + 1] - out->off[q];
            tmp.n = 0;
            for (uint64_t i = 0; i < n; ++i) {
                char ch = out->chars[out->off[q] + i];
                nv_push(&tmp, ch == 'A' ? 0u : ch == 'C' ? 1u : ch == 'G' ? 2u : 3u);
            }
            clean_condensed(m, data, k, tmp.v, tmp.n);
        }
        /* CollectLoops :308-344 — candidates are decided on the post-cleaning masks, then walked in file order */
        uint8_t *cand = (uint8_t *) calloc(s->size ? s->size : 1, 1);
        for (uint64_t i = 0; i < s->size; ++i) {
            kwh_t kh = make_kwh(m, s->data + i * s->words, k);
            uint8_t ext = get_value(data, &kh);
            cand[i] = (uint8_t) (UNIQUE[ext & 15] && UNIQUE[ext >> 4]);
        }
        for (uint64_t i = 0; i < s->size; ++i) {
            if (!cand[i]) continue;
            kwh_t kh = make_kwh(m, s->data + i * s->words, k);
            uint8_t ext = get_value(data, &kh);
            if (!(UNIQUE[ext & 15] && UNIQUE[ext >> 4])) continue;
            /* ConstructLoopFromVertex :248-265 */
            kwh_t nx = kwh_shl(m, &kh, k, (unsigned) NEXT[ext & 15]);
            construct_sequence(m, data, k, kh, nx, &nv);
            uint64_t n = nv.n;
            int64_t split = -1;
            for (uint64_t p = 0; p + k + 1 <= n; ++p) {
                int self_rc = 1;
                for (unsigned j = 0; j < k + 1; ++j)
                    if (nv.v[p + j] != 3u - nv.v[p + k - j]) { self_rc = 0; break; }
                if (self_rc) { split = (int64_t) p; break; }
            }
            if (split < 0) {
                emit(out, nv.v, n, nuc_less_rc(nv.v, n));
                out->n_loops++;
                clean_condensed(m, data, k, nv.v, n);
            } else {   /* SplitLoop :248-253 */
                uint64_t pos = (uint64_t) split;
                emit(out, nv.v + pos, k + 1, nuc_less_rc(nv.v + pos, k + 1));
                clean_condensed(m, data, k, nv.v + pos, k + 1);
                tmp.n = 0;
                for (uint64_t j = pos + 1; j < n - k; ++j) nv_push(&tmp, nv.v[j]);
                for (uint64_t j = 0; j < pos + k; ++j) nv_push(&tmp, nv.v[j]);
                emit(out, tmp.v, tmp.n, nuc_less_rc(tmp.v, tmp.n));
                clean_condensed(m, data, k, tmp.v, tmp.n);
                out->n_loops += 2;
            }
        }
        free(cand);
        free(tmp.v);
    }
    free(nv.v);
    return out;
}

uint64_t        ora_seqs_count(const ora_seqs *s) { return s->count; }
uint64_t        ora_seqs_n_loops(const ora_seqs *s) { return s->n_loops; }
const uint64_t *ora_seqs_offsets(const ora_seqs *s) { return s->off; }
const char     *ora_seqs_chars(const ora_seqs *s) { return s->chars; }
void ora_seqs_free(ora_seqs *s) { if (!s) return; free(s->off); free(s->chars); free(s); }
