"""ctypes binding of oracle/libsb200_oracle.so — TEST INFRASTRUCTURE (the CPU restatement, see oracle/sb200_oracle.h).

Only tests/, bench.py's cpu_baseline leg and __graft_entry__.smoke() import this module; the product package
(spades_for_blackbird_b200) never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
_LIB = None

u64p = C.POINTER(C.c_uint64)
u32p = C.POINTER(C.c_uint32)
u8p = C.POINTER(C.c_uint8)


def build_oracle():
    """Compile oracle/libsb200_oracle.so if it is missing or stale."""
    so = os.path.join(ORACLE_DIR, "libsb200_oracle.so")
    src = os.path.join(ORACLE_DIR, "sb200_oracle.c")
    if not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", ORACLE_DIR, "oracle"], stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is not None:
        return _LIB
    L = C.CDLL(build_oracle())
    vp = C.c_void_p
    L.ora_xxh3_64.restype = C.c_uint64
    L.ora_xxh3_64.argtypes = [u64p, C.c_uint]
    L.ora_xxh3_128.restype = None
    L.ora_xxh3_128.argtypes = [u64p, C.c_uint, u64p, u64p]
    L.ora_kmer_rc.restype = None
    L.ora_kmer_rc.argtypes = [u64p, C.c_uint, u64p]
    L.ora_kmer_is_minimal.restype = C.c_int
    L.ora_kmer_is_minimal.argtypes = [u64p, C.c_uint]
    L.ora_bucket.restype = C.c_uint
    L.ora_bucket.argtypes = [u64p, C.c_uint, C.c_uint]
    L.ora_pack_reads.restype = C.c_uint64
    L.ora_pack_reads.argtypes = [C.c_char_p, u64p, C.c_uint64, u64p, u64p, u32p]
    L.ora_count_reads.restype = vp
    L.ora_count_reads.argtypes = [u64p, u64p, u32p, C.c_uint64, C.c_uint, C.c_int, C.c_int, C.c_uint]
    L.ora_derive_kmers.restype = vp
    L.ora_derive_kmers.argtypes = [vp, C.c_uint]
    for name, rt in [("ora_kmers_k", C.c_uint), ("ora_kmers_words", C.c_uint), ("ora_kmers_num_buckets", C.c_uint),
                     ("ora_kmers_size", C.c_uint64), ("ora_kmers_data", u64p), ("ora_kmers_bucket_starts", u64p),
                     ("ora_kmers_counts", u32p)]:
        getattr(L, name).restype = rt
        getattr(L, name).argtypes = [vp]
    L.ora_kmers_free.argtypes = [vp]
    L.ora_mphf_build.restype = vp
    L.ora_mphf_build.argtypes = [vp]
    L.ora_mphf_lookup.restype = C.c_uint64
    L.ora_mphf_lookup.argtypes = [vp, u64p]
    L.ora_mphf_final_level_keys.restype = C.c_uint64
    L.ora_mphf_final_level_keys.argtypes = [vp]
    L.ora_mphf_serialize.restype = C.c_uint64
    L.ora_mphf_serialize.argtypes = [vp, u8p]
    L.ora_mphf_free.argtypes = [vp]
    L.ora_fill_masks.restype = None
    L.ora_fill_masks.argtypes = [vp, vp, u8p]
    L.ora_tipclip.restype = C.c_uint64
    L.ora_tipclip.argtypes = [vp, vp, u8p, C.c_uint64]
    L.ora_unitigs.restype = vp
    L.ora_unitigs.argtypes = [vp, vp, u8p, C.c_int]
    L.ora_seqs_count.restype = C.c_uint64
    L.ora_seqs_count.argtypes = [vp]
    L.ora_seqs_n_loops.restype = C.c_uint64
    L.ora_seqs_n_loops.argtypes = [vp]
    L.ora_seqs_offsets.restype = u64p
    L.ora_seqs_offsets.argtypes = [vp]
    L.ora_seqs_chars.restype = C.POINTER(C.c_char)
    L.ora_seqs_chars.argtypes = [vp]
    L.ora_seqs_free.argtypes = [vp]
    _LIB = L
    return L


def _p(a, t):
    return a.ctypes.data_as(t)


def pack_reads(reads):
    """list[str] -> (words u64[], word_off u64[n+1], len u32[n]) in the reference's 2-bit layout (longest ACGT run)."""
    L = lib()
    enc = [r.encode() for r in reads]
    off = np.zeros(len(enc) + 1, dtype=np.uint64)
    off[1:] = np.cumsum([len(e) for e in enc], dtype=np.uint64)
    blob = b"".join(enc)
    n = len(enc)
    nw = L.ora_pack_reads(blob, _p(off, u64p), n, None, None, None)
    words = np.zeros(max(nw, 1), dtype=np.uint64)
    word_off = np.zeros(n + 1, dtype=np.uint64)
    lens = np.zeros(max(n, 1), dtype=np.uint32)
    L.ora_pack_reads(blob, _p(off, u64p), n, _p(words, u64p), _p(word_off, u64p), _p(lens, u32p))
    return words[:nw], word_off, lens[:n]


class KmerSet:
    def __init__(self, handle):
        self.h = handle
        L = lib()
        self.k = L.ora_kmers_k(handle)
        self.words = L.ora_kmers_words(handle)
        self.num_buckets = L.ora_kmers_num_buckets(handle)
        self.size = L.ora_kmers_size(handle)
        self.data = np.ctypeslib.as_array(L.ora_kmers_data(handle), shape=(self.size, self.words)).copy()
        self.bucket_starts = np.ctypeslib.as_array(L.ora_kmers_bucket_starts(handle), shape=(self.num_buckets + 1,)).copy()
        cp = L.ora_kmers_counts(handle)
        self.counts = np.ctypeslib.as_array(cp, shape=(self.size,)).copy() if cp else None

    def __del__(self):
        if self.h:
            lib().ora_kmers_free(self.h)
            self.h = None


def count_reads(words, word_off, lens, K, canonical_only, add_rc, num_buckets):
    L = lib()
    words = np.ascontiguousarray(words, dtype=np.uint64)
    if words.size == 0:
        words = np.zeros(1, dtype=np.uint64)
    word_off = np.ascontiguousarray(word_off, dtype=np.uint64)
    lens = np.ascontiguousarray(lens, dtype=np.uint32)
    n = len(word_off) - 1
    if lens.size == 0:
        lens = np.zeros(1, dtype=np.uint32)
    h = L.ora_count_reads(_p(words, u64p), _p(word_off, u64p), _p(lens, u32p), n, K, int(canonical_only), int(add_rc),
                          num_buckets)
    return KmerSet(h) if h else None


def derive_kmers(kpomers, num_buckets):
    h = lib().ora_derive_kmers(kpomers.h, num_buckets)
    return KmerSet(h) if h else None


class Mphf:
    def __init__(self, kmers):
        self.kmers = kmers
        self.h = lib().ora_mphf_build(kmers.h)

    def lookup(self, rec):
        rec = np.ascontiguousarray(rec, dtype=np.uint64)
        return lib().ora_mphf_lookup(self.h, _p(rec, u64p))

    def lookup_all(self, recs):
        return np.array([self.lookup(r) for r in recs], dtype=np.uint64)

    def final_level_keys(self):
        return lib().ora_mphf_final_level_keys(self.h)

    def serialize(self):
        L = lib()
        n = L.ora_mphf_serialize(self.h, None)
        buf = np.zeros(n, dtype=np.uint8)
        L.ora_mphf_serialize(self.h, _p(buf, u8p))
        return buf

    def __del__(self):
        if self.h:
            lib().ora_mphf_free(self.h)
            self.h = None


def fill_masks(kpomers, mphf):
    data = np.zeros(max(mphf.kmers.size, 1), dtype=np.uint8)
    lib().ora_fill_masks(kpomers.h, mphf.h, _p(data, u8p))
    return data[:mphf.kmers.size]


def tipclip(kmers, mphf, data, bound):
    """In place on `data`; returns the number of removed k-mers."""
    return lib().ora_tipclip(kmers.h, mphf.h, _p(data, u8p), bound)


def unitigs(kmers, mphf, data, with_loops=True):
    """Returns (list[str] in the reference's output order, n_loops); `data` is mutated like the reference's masks."""
    L = lib()
    h = L.ora_unitigs(kmers.h, mphf.h, _p(data, u8p), int(with_loops))
    n = L.ora_seqs_count(h)
    off = np.ctypeslib.as_array(L.ora_seqs_offsets(h), shape=(n + 1,)).copy()
    total = int(off[n])
    chars = C.string_at(L.ora_seqs_chars(h), total) if total else b""
    out = [chars[int(off[i]):int(off[i + 1])].decode() for i in range(n)]
    nl = L.ora_seqs_n_loops(h)
    L.ora_seqs_free(h)
    return out, nl


def gbuilder(reads, k, num_buckets, tip_bound=None, with_loops=True):
    """Whole path on the CPU oracle.  Returns a dict of every artefact the reference driver dumps."""
    words, word_off, lens = pack_reads(reads)
    kp = count_reads(words, word_off, lens, k + 1, True, True, num_buckets)
    if kp is None:
        return None
    km = derive_kmers(kp, num_buckets)
    mp = Mphf(km)
    masks_idx = fill_masks(kp, mp)
    clipped = 0
    if tip_bound is not None:
        clipped = tipclip(km, mp, masks_idx, tip_bound)
    idx = mp.lookup_all(km.data)
    res = dict(kpomers=kp, kmers=km, mphf=mp, idx=idx, masks_idx=masks_idx.copy(), clipped=clipped,
               index_bin=mp.serialize())
    work = masks_idx.copy()
    res["unitigs"], res["n_loops"] = unitigs(km, mp, work, with_loops)
    return res
