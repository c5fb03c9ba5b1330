// tests/host_prims.cpp — TEST INFRASTRUCTURE: compiles spades_for_blackbird_b200/csrc/kmer_ops.cuh for the HOST so that the
// word-parallel k-mer arithmetic and the XXH3 specialisations the kernels use can be checked against the oracle without
// a GPU (tests/test_host_primitives.py).  Not part of the product library.
#include <cstring>
#include "../spades_for_blackbird_b200/csrc/kmer_ops.cuh"
#include "../spades_for_blackbird_b200/csrc/mphf.cuh"

template<int W> static void ops(const uint64_t *x, int K, uint64_t *rc, int *minimal, uint64_t *h64, uint64_t *h128, uint64_t *shl,
                                uint32_t c, uint32_t *bucket, uint32_t B) {
    kmer_rc<W>(x, K, rc);
    uint64_t can[W];
    *minimal = kmer_canonical<W>(x, K, can) ? 1 : 0;
    *h64 = xxh3_64<W>(x);
    xxh3_128<W>(x, h128[0], h128[1]);
    kmer_shl<W>(x, K, c, shl);
    *bucket = kmer_bucket<W>(x, B);
}
extern "C" void hp_ops(int W, const uint64_t *x, int K, uint64_t *rc, int *minimal, uint64_t *h64, uint64_t *h128, uint64_t *shl,
                       uint32_t c, uint32_t *bucket, uint32_t B) {
    switch (W) {
        case 1: ops<1>(x, K, rc, minimal, h64, h128, shl, c, bucket, B); break;
        case 2: ops<2>(x, K, rc, minimal, h64, h128, shl, c, bucket, B); break;
        case 3: ops<3>(x, K, rc, minimal, h64, h128, shl, c, bucket, B); break;
        default: ops<4>(x, K, rc, minimal, h64, h128, shl, c, bucket, B); break;
    }
}
extern "C" void hp_window(int W, const uint64_t *seq, uint32_t nw, uint32_t pos, int K, uint64_t *out) {
    switch (W) {
        case 1: kmer_window<1>(seq, nw, pos, K, out); break;
        case 2: kmer_window<2>(seq, nw, pos, K, out); break;
        case 3: kmer_window<3>(seq, nw, pos, K, out); break;
        default: kmer_window<4>(seq, nw, pos, K, out); break;
    }
}
extern "C" void hp_subwindow(int WS, int W, const uint64_t *src, uint32_t pos, int K, uint64_t *out) {
    if (WS == 1) kmer_subwindow<1, 1>(src, pos, K, out);
    else if (WS == 2 && W == 1) kmer_subwindow<2, 1>(src, pos, K, out);
    else if (WS == 2) kmer_subwindow<2, 2>(src, pos, K, out);
    else if (WS == 3 && W == 2) kmer_subwindow<3, 2>(src, pos, K, out);
    else if (WS == 3) kmer_subwindow<3, 3>(src, pos, K, out);
    else if (WS == 4 && W == 3) kmer_subwindow<4, 3>(src, pos, K, out);
    else kmer_subwindow<4, 4>(src, pos, K, out);
}
extern "C" uint64_t hp_xs_next(uint64_t *s) { return sb200::xs_next(s[0], s[1]); }
extern "C" uint32_t hp_mask_conj(uint32_t m) { return mask_conj(m); }

// rolling window + reverse complement (partition.cuh): start at window `pos`, roll `steps` times; outputs the last x and r
template<int W> static void roll(const uint64_t *seq, uint32_t nw, uint32_t pos, int K, int steps, uint64_t *xo, uint64_t *ro) {
    uint64_t x[W], r[W];
    kmer_window<W>(seq, nw, pos, K, x);
    kmer_rc<W>(x, K, r);
    for (int i = 0; i < steps; ++i) {
        uint32_t q = pos + K + i;
        uint32_t c = (uint32_t) (seq[q >> 5] >> (2 * (q & 31))) & 3u;
        kmer_roll<W>(x, r, K, c, last_word_mask(K));
    }
    for (int j = 0; j < W; ++j) { xo[j] = x[j]; ro[j] = r[j]; }
}
extern "C" void hp_roll(int W, const uint64_t *seq, uint32_t nw, uint32_t pos, int K, int steps, uint64_t *xo, uint64_t *ro) {
    switch (W) {
        case 1: roll<1>(seq, nw, pos, K, steps, xo, ro); break;
        case 2: roll<2>(seq, nw, pos, K, steps, xo, ro); break;
        case 3: roll<3>(seq, nw, pos, K, steps, xo, ro); break;
        default: roll<4>(seq, nw, pos, K, steps, xo, ro); break;
    }
}
template<int WS, int W> static void cand(const uint64_t *x, int k, uint64_t *out, uint32_t *bit) {
    uint64_t a[2][W];
    derive_candidates<WS, W>(x, k, a, bit);
    for (int c = 0; c < 2; ++c)
        for (int j = 0; j < W; ++j) out[c * 4 + j] = a[c][j];
}
extern "C" void hp_candidates(int WS, int W, const uint64_t *x, int k, uint64_t *out /* 2 x 4 words */, uint32_t *bit /* 2 */) {
    if (WS == 1) cand<1, 1>(x, k, out, bit);
    else if (WS == 2 && W == 1) cand<2, 1>(x, k, out, bit);
    else if (WS == 2) cand<2, 2>(x, k, out, bit);
    else if (WS == 3 && W == 2) cand<3, 2>(x, k, out, bit);
    else if (WS == 3) cand<3, 3>(x, k, out, bit);
    else if (WS == 4 && W == 3) cand<4, 3>(x, k, out, bit);
    else cand<4, 4>(x, k, out, bit);
}
