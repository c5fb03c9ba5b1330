"""Word-parallel k-mer arithmetic + XXH3 specialisations of the CUDA path (csrc/kmer_ops.cuh, compiled for the host)
against the nucleotide-at-a-time oracle.  Bit-exact."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import oracle_lib as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def hp():
    so = os.path.join(ROOT, "tests", "_host_prims.so")
    src = os.path.join(ROOT, "tests", "host_prims.cpp")
    hdr = os.path.join(ROOT, "spades_for_blackbird_b200", "csrc", "kmer_ops.cuh")
    if not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        subprocess.check_call(["/usr/bin/g++", "-O1", "-std=c++17", "-shared", "-fPIC", "-x", "c++", "-Wno-unknown-pragmas",
                               "-o", so, src])
    return C.CDLL(so)


def rand_kmer(rng, K):
    W = (K + 31) // 32
    codes = rng.integers(0, 4, size=K)
    w = np.zeros(4, dtype=np.uint64)
    for i, c in enumerate(codes):
        w[i // 32] |= np.uint64(int(c)) << np.uint64(2 * (i % 32))
    return w, W


def p64(a):
    return a.ctypes.data_as(O.u64p)


@pytest.mark.parametrize("K", [1, 5, 21, 22, 31, 32, 33, 55, 56, 63, 64, 65, 77, 78, 95, 96, 97, 127, 128])
def test_kmer_ops_match_oracle(hp, K):
    L = O.lib()
    rng = np.random.default_rng(K)
    for trial in range(200):
        x, W = rand_kmer(rng, K)
        if trial % 10 == 0 and K % 2 == 0:   # force a self-reverse-complement k-mer
            half = [int((int(x[i // 32]) >> (2 * (i % 32))) & 3) for i in range(K // 2)]
            full = half + [3 - c for c in reversed(half)]
            x[:] = 0
            for i, c in enumerate(full):
                x[i // 32] |= np.uint64(c) << np.uint64(2 * (i % 32))
        if trial % 10 in (1, 2, 3) and K >= 2:   # near-palindrome: x and rc(x) agree up to one base, anywhere (every word decides IsMinimal once)
            half = [int((int(x[i // 32]) >> (2 * (i % 32))) & 3) for i in range((K + 1) // 2)]
            full = (half + [3 - c for c in reversed(half[:K // 2])])[:K]
            j = int(rng.integers(0, K))
            full[j] = (full[j] + 1 + int(rng.integers(0, 3))) % 4
            x[:] = 0
            for i, c in enumerate(full):
                x[i // 32] |= np.uint64(c) << np.uint64(2 * (i % 32))
        rc = np.zeros(4, dtype=np.uint64); shl = np.zeros(4, dtype=np.uint64); h128 = np.zeros(2, dtype=np.uint64)
        minimal = C.c_int(); h64 = C.c_uint64(); bucket = C.c_uint32()
        c = int(rng.integers(0, 4)); B = int(rng.integers(1, 2000))
        hp.hp_ops(W, p64(x), K, p64(rc), C.byref(minimal), C.byref(h64), p64(h128), p64(shl), c, C.byref(bucket), B)
        orc = np.zeros(4, dtype=np.uint64)
        L.ora_kmer_rc(p64(x), K, p64(orc))
        assert np.array_equal(rc[:W], orc[:W])
        assert minimal.value == L.ora_kmer_is_minimal(p64(x), K)
        assert h64.value == L.ora_xxh3_64(p64(x), W)
        hi = C.c_uint64(); lo = C.c_uint64()
        L.ora_xxh3_128(p64(x), W, C.byref(hi), C.byref(lo))
        assert (int(h128[0]), int(h128[1])) == (hi.value, lo.value)
        assert bucket.value == L.ora_bucket(p64(x), W, B)
        # x << c
        codes = [int((int(x[i // 32]) >> (2 * (i % 32))) & 3) for i in range(K)]
        exp = codes[1:] + [c]
        got = [int((int(shl[i // 32]) >> (2 * (i % 32))) & 3) for i in range(K)]
        assert got == exp
        for j in range(W, 4):
            assert rc[j] == 0 or j >= W


@pytest.mark.parametrize("K", [5, 22, 32, 33, 56, 64, 78, 96, 128])
def test_window_extraction(hp, K):
    rng = np.random.default_rng(1000 + K)
    W = (K + 31) // 32
    for L_ in (K, K + 1, K + 31, K + 32, K + 33, 150, 250):
        if L_ < K:
            continue
        codes = rng.integers(0, 4, size=L_)
        nw = (L_ + 31) // 32
        seq = np.zeros(nw, dtype=np.uint64)
        for i, c in enumerate(codes):
            seq[i // 32] |= np.uint64(int(c)) << np.uint64(2 * (i % 32))
        for pos in range(L_ - K + 1):
            out = np.zeros(4, dtype=np.uint64)
            hp.hp_window(W, p64(seq), nw, pos, K, p64(out))
            got = [int((int(out[i // 32]) >> (2 * (i % 32))) & 3) for i in range(K)]
            assert got == [int(c) for c in codes[pos:pos + K]]
            # padding bits are zero
            total_bits = 2 * K
            for j in range(W):
                lo = 64 * j
                if total_bits < lo + 64:
                    assert int(out[j]) >> max(total_bits - lo, 0) == 0


@pytest.mark.parametrize("K1", [6, 22, 32, 33, 34, 56, 64, 65, 78, 96, 97, 128])
def test_subwindow(hp, K1):
    rng = np.random.default_rng(2000 + K1)
    k = K1 - 1
    WS, W = (K1 + 31) // 32, (k + 31) // 32
    for _ in range(100):
        x, _w = rand_kmer(rng, K1)
        codes = [int((int(x[i // 32]) >> (2 * (i % 32))) & 3) for i in range(K1)]
        for pos in (0, 1):
            out = np.zeros(4, dtype=np.uint64)
            hp.hp_subwindow(WS, W, p64(x), pos, k, p64(out))
            got = [int((int(out[i // 32]) >> (2 * (i % 32))) & 3) for i in range(k)]
            assert got == codes[pos:pos + k]
            if 2 * k < 64 * W:
                assert int(out[W - 1]) >> (2 * k - 64 * (W - 1)) == 0


def test_mask_conj(hp):
    for m in range(256):
        assert hp.hp_mask_conj(m) == int("{:08b}".format(m)[::-1], 2)


def _pack(codes):
    w = np.zeros(max((len(codes) + 31) // 32, 1) + 1, dtype=np.uint64)
    for i, c in enumerate(codes):
        w[i // 32] |= np.uint64(int(c)) << np.uint64(2 * (i % 32))
    return w


def _unpack(w, K):
    return [int((int(w[i // 32]) >> (2 * (i % 32))) & 3) for i in range(K)]


@pytest.mark.parametrize("K", [5, 22, 32, 33, 56, 64, 65, 78, 96, 97, 127, 128])
def test_rolling_window_and_rc(hp, K):
    """partition.cuh: a lane extracts one window + its reverse complement and ROLLS both through the next bases"""
    rng = np.random.default_rng(3000 + K)
    W = (K + 31) // 32
    for L_ in (K + 3, K + 40, 150 if K < 140 else K + 9, 250):
        codes = [int(c) for c in rng.integers(0, 4, size=L_)]
        seq = _pack(codes)
        nw = (L_ + 31) // 32
        for pos in range(0, L_ - K + 1, 7):
            for steps in range(0, min(4, L_ - K - pos + 1)):
                xo = np.zeros(4, dtype=np.uint64); ro = np.zeros(4, dtype=np.uint64)
                hp.hp_roll(W, p64(seq), nw, pos, K, steps, p64(xo), p64(ro))
                want = codes[pos + steps:pos + steps + K]
                assert _unpack(xo, K) == want
                assert _unpack(ro, K) == [3 - c for c in reversed(want)]
                for j in range(W):   # padding bits stay zero
                    if 2 * K < 64 * (j + 1):
                        assert int(xo[j]) >> max(2 * K - 64 * j, 0) == 0 and int(ro[j]) >> max(2 * K - 64 * j, 0) == 0


@pytest.mark.parametrize("K1", [6, 22, 32, 33, 34, 56, 64, 65, 78, 96, 97, 128])
def test_derive_candidates(hp, K1):
    """the two canonical k-mer candidates of a (k+1)-mer and their InOutMask bits, from ONE reverse complement"""
    rng = np.random.default_rng(4000 + K1)
    k = K1 - 1
    WS, W = (K1 + 31) // 32, (k + 31) // 32
    for trial in range(150):
        x, _w = rand_kmer(rng, K1)
        codes = _unpack(x, K1)
        if trial % 5 == 0:   # k-mer candidates that are (near-)palindromes
            half = codes[:(k + 1) // 2]
            pal = (half + [3 - c for c in reversed(half[:k // 2])])[:k]
            codes = pal + [codes[-1]] if trial % 10 == 0 else [codes[0]] + pal
            x = np.zeros(4, dtype=np.uint64)
            x[:len(_pack(codes)) - 1] = _pack(codes)[:-1]
        out = np.zeros(8, dtype=np.uint64); bit = np.zeros(2, dtype=np.uint32)
        hp.hp_candidates(WS, W, p64(x), k, p64(out), bit.ctypes.data_as(C.POINTER(C.c_uint32)))
        for which, sub in ((0, codes[:k]), (1, codes[1:])):
            rc = [3 - c for c in reversed(sub)]
            minimal = sub <= rc
            assert _unpack(out[4 * which:4 * which + 4], k) == (sub if minimal else rc)
            if which == 0:
                want_bit = codes[k] if minimal else 7 - codes[k]
            else:
                want_bit = codes[0] + 4 if minimal else 3 - codes[0]
            assert int(bit[which]) == want_bit
