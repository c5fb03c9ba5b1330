"""ctypes binding of libspades_b200.so (include/sb200.h) with the reference's own vocabulary on top.

The classes mirror the C++ interfaces of the reference for this path (same names, argument meaning and failure
behaviour), so that the parity tests read like A/test/debruijn/construction_test.cpp:

    kmers::KMerDiskCounter<RtSeq>::Count            C/utils/kmer_mph/kmer_index_builder.hpp:241-267
    kmers::KMerDiskStorage<RtSeq>                   C/utils/kmer_mph/kmer_index_builder.hpp:48-191
    kmers::KMerIndexBuilder / KMerIndex             C/utils/kmer_mph/kmer_index_builder.hpp:368-453, kmer_index.hpp:25-147
    utils::DeBruijnExtensionIndex(Builder)          C/utils/extension_index/kmer_extension_index{,_builder}.hpp
    debruijn_graph::EarlyTipClipperProcessor        C/assembly_graph/construction/early_simplification.hpp:37-160
    debruijn_graph::UnbranchingPathExtractor        C/assembly_graph/construction/debruijn_graph_constructor.hpp:182-388

There is no CPU path here: importing works anywhere, but creating a Context without the built library or without an
sm_100 GPU raises (the reference FATAL_ERRORs; we raise Sb200Error with the library's message).
"""
import ctypes as C
import os

import numpy as np

_PKG = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB_PATH = os.path.join(_PKG, "libspades_b200.so")
_lib = None

u64p = C.POINTER(C.c_uint64)
u32p = C.POINTER(C.c_uint32)
u8p = C.POINTER(C.c_uint8)
vp = C.c_void_p


class Sb200Error(RuntimeError):
    pass


class ConstructParams(C.Structure):
    _fields_ = [("k", C.c_uint), ("num_buckets", C.c_uint), ("tip_clip", C.c_int), ("tip_length_bound", C.c_uint64),
                ("with_loops", C.c_int), ("fetch_kmers", C.c_int)]


class GraphView(C.Structure):
    _fields_ = [("n_kpomers", C.c_uint64), ("n_kmers", C.c_uint64), ("n_unitigs", C.c_uint64), ("n_loops", C.c_uint64),
                ("unitig_bases", C.c_uint64), ("n_unitig_words", C.c_uint64), ("kpomer_instances", C.c_uint64),
                ("clipped", C.c_uint64),
                ("kpomers", u64p), ("kpomer_counts", u32p), ("kpomer_bucket_starts", u64p),
                ("kmers", u64p), ("kmer_bucket_starts", u64p),
                ("masks", u8p), ("index_bytes", u8p), ("index_size", C.c_uint64),
                ("unitig_words", u64p), ("unitig_word_off", u64p), ("unitig_len", u32p),
                ("h2d_bytes", C.c_uint64), ("d2h_bytes", C.c_uint64)]


class ShardInfo(C.Structure):
    _fields_ = [("rank", C.c_int), ("size", C.c_int), ("whole_set_fallback", C.c_int), ("gathered", C.c_int),
                ("total_kpomers", C.c_uint64), ("total_kmers", C.c_uint64), ("total_instances", C.c_uint64), ("total_unitigs", C.c_uint64),
                ("total_unitig_bases", C.c_uint64), ("n_loops", C.c_uint64), ("clipped", C.c_uint64), ("bytes_sent", C.c_uint64),
                ("exchange_ms", C.c_double), ("stage_ms", C.c_double * 8), ("record_bytes", C.c_uint64)]


STAGE_NAMES = ("count_kpomers", "count_kmers", "mphf", "masks", "tipclip", "unitigs", "gather", "total")

EXPORTS = {   # symbol -> (restype, argtypes); tests check that the library exports every one of them
    "sb200_create": (C.c_int, [C.c_int, C.POINTER(vp)]),
    "sb200_destroy": (None, [vp]),
    "sb200_last_error": (C.c_char_p, [vp]),
    "sb200_synchronize": (C.c_int, [vp]),
    "sb200_stream": (vp, [vp]),
    "sb200_kernel_launches": (C.c_uint64, [vp, C.c_int]),
    "sb200_profile": (C.c_int, [vp, C.c_int]),
    "sb200_profile_report": (C.c_int, [vp, C.c_char_p, C.c_uint64, u64p]),
    "sb200_reads_upload": (C.c_int, [vp, u64p, u64p, u32p, C.c_uint64, C.POINTER(vp)]),
    "sb200_reads_wrap_device": (C.c_int, [vp, vp, vp, vp, C.c_uint64, C.c_uint64, C.POINTER(vp)]),
    "sb200_reads_free": (None, [vp]),
    "sb200_count": (C.c_int, [vp, vp, C.c_uint, C.c_int, C.c_int, C.c_uint, C.POINTER(vp)]),
    "sb200_derive_kmers": (C.c_int, [vp, vp, C.c_uint, C.POINTER(vp)]),
    "sb200_kmers_k": (C.c_uint, [vp]),
    "sb200_kmers_words": (C.c_uint, [vp]),
    "sb200_kmers_num_buckets": (C.c_uint, [vp]),
    "sb200_kmers_size": (C.c_uint64, [vp]),
    "sb200_kmers_instances": (C.c_uint64, [vp]),
    "sb200_kmers_bucket_starts": (C.c_int, [vp, u64p]),
    "sb200_kmers_download": (C.c_int, [vp, C.c_uint64, C.c_uint64, u64p]),
    "sb200_kmers_counts_download": (C.c_int, [vp, C.c_uint64, C.c_uint64, u32p]),
    "sb200_kmers_device_records": (vp, [vp]),
    "sb200_kmers_device_counts": (vp, [vp]),
    "sb200_kmers_free": (None, [vp]),
    "sb200_mphf_build": (C.c_int, [vp, vp, C.POINTER(vp)]),
    "sb200_mphf_size": (C.c_uint64, [vp]),
    "sb200_mphf_mem_size": (C.c_uint64, [vp]),
    "sb200_mphf_lookup": (C.c_int, [vp, vp, u64p, C.c_uint64, u64p]),
    "sb200_mphf_seq_idx": (C.c_int, [vp, u64p, u64p]),
    "sb200_mphf_serialize": (C.c_int, [vp, u8p, u64p]),
    "sb200_mphf_free": (None, [vp]),
    "sb200_ext_build": (C.c_int, [vp, vp, vp, vp, C.POINTER(vp)]),
    "sb200_ext_masks_download": (C.c_int, [vp, u8p]),
    "sb200_ext_idx_download": (C.c_int, [vp, u32p]),
    "sb200_ext_free": (None, [vp]),
    "sb200_tipclip": (C.c_int, [vp, vp, vp, vp, C.c_uint64, u64p]),
    "sb200_unitigs_extract": (C.c_int, [vp, vp, vp, vp, C.c_int, C.POINTER(vp)]),
    "sb200_unitigs_count": (C.c_uint64, [vp]),
    "sb200_unitigs_loops": (C.c_uint64, [vp]),
    "sb200_unitigs_total_bases": (C.c_uint64, [vp]),
    "sb200_unitigs_total_words": (C.c_uint64, [vp]),
    "sb200_unitigs_download": (C.c_int, [vp, u64p, u64p, u32p]),
    "sb200_unitigs_free": (None, [vp]),
    "sb200_records_extract": (C.c_int, [vp, vp, C.c_uint, C.c_int, C.c_int, C.POINTER(vp)]),
    "sb200_records_derive": (C.c_int, [vp, vp, C.POINTER(vp)]),
    "sb200_records_partition": (C.c_int, [vp, vp, C.c_uint, C.c_uint, u64p]),
    "sb200_records_extract_partitioned": (C.c_int, [vp, vp, C.c_uint, C.c_int, C.c_int, C.c_uint, C.c_uint, u64p, C.POINTER(vp)]),
    "sb200_records_alloc": (C.c_int, [vp, C.c_uint64, C.c_uint, C.c_int, C.POINTER(vp)]),
    "sb200_records_size": (C.c_uint64, [vp]),
    "sb200_records_words": (C.c_uint, [vp]),
    "sb200_records_k": (C.c_uint, [vp]),
    "sb200_records_flags": (C.c_int, [vp]),
    "sb200_records_device": (vp, [vp]),
    "sb200_records_free": (None, [vp]),
    "sb200_count_records": (C.c_int, [vp, vp, C.c_uint, C.c_int, C.POINTER(vp)]),
    "sb200_count_records_owned": (C.c_int, [vp, vp, C.c_uint, C.c_uint, C.c_uint, C.c_int, C.POINTER(vp)]),
    "sb200_mphf_build_sharded": (C.c_int, [vp, vp, u64p, C.POINTER(vp)]),
    "sb200_mphf_arrays": (C.c_int, [vp, C.POINTER(vp), u64p, C.POINTER(vp), u64p]),
    "sb200_mphf_complete": (C.c_int, [vp, vp]),
    "sb200_ext_masks_device": (C.c_int, [vp, C.POINTER(vp), u64p]),
    "sb200_unitigs_extract_local": (C.c_int, [vp, vp, vp, vp, u64p, C.POINTER(vp)]),
    "sb200_unitigs_device": (C.c_int, [vp, C.POINTER(vp), C.POINTER(vp), C.POINTER(vp)]),
    "sb200_coverage_map_build": (C.c_int, [vp, vp, C.POINTER(vp)]),
    "sb200_coverage_map_index": (vp, [vp]),
    "sb200_coverage_map_size": (C.c_uint64, [vp]),
    "sb200_coverage_map_values_download": (C.c_int, [vp, u32p]),
    "sb200_coverage_map_free": (None, [vp]),
    "sb200_unitigs_coverage": (C.c_int, [vp, vp, vp, C.c_uint32, u64p, u64p]),
    "sb200_comm_unique_id": (C.c_int, [u8p]),
    "sb200_comm_last_error": (C.c_char_p, []),
    "sb200_comm_create_nccl": (C.c_int, [vp, C.c_int, C.c_int, u8p, C.POINTER(vp)]),
    "sb200_comm_create_local": (C.c_int, [C.c_int, C.POINTER(vp)]),
    "sb200_comm_rank": (C.c_int, [vp]),
    "sb200_comm_size": (C.c_int, [vp]),
    "sb200_comm_free": (None, [vp]),
    "sb200_construct_sharded": (C.c_int, [vp, vp, vp, C.POINTER(ConstructParams), C.c_int, C.POINTER(vp)]),
    "sb200_shard_kpomers": (vp, [vp]),
    "sb200_shard_kmers": (vp, [vp]),
    "sb200_shard_mphf": (vp, [vp]),
    "sb200_shard_ext": (vp, [vp]),
    "sb200_shard_unitigs": (vp, [vp]),
    "sb200_shard_info": (C.c_int, [vp, C.POINTER(ShardInfo)]),
    "sb200_shard_walk_stats": (C.c_int, [vp, u64p]),
    "sb200_shard_free": (None, [vp]),
    "sb200_multi_create": (C.c_int, [C.c_int, C.POINTER(C.c_int), C.POINTER(vp)]),
    "sb200_multi_destroy": (None, [vp]),
    "sb200_multi_last_error": (C.c_char_p, [vp]),
    "sb200_multi_size": (C.c_int, [vp]),
    "sb200_multi_context": (vp, [vp, C.c_int]),
    "sb200_multi_construct": (C.c_int, [vp, u64p, u64p, u32p, C.c_uint64, C.POINTER(ConstructParams), C.POINTER(vp)]),
    "sb200_construct": (C.c_int, [vp, u64p, u64p, u32p, C.c_uint64, C.POINTER(ConstructParams), C.POINTER(vp)]),
    "sb200_graph_get": (C.c_int, [vp, C.POINTER(GraphView)]),
    "sb200_graph_free": (None, [vp]),
}


def load_library():
    """dlopen libspades_b200.so and type every export.  Raises if the library is not built (no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise Sb200Error("libspades_b200.so is not built (%s): run __graft_entry__.build() / make -C "
                         "spades_for_blackbird_b200/csrc; there is no CPU fallback" % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, (rt, at) in EXPORTS.items():
        fn = getattr(lib, name)   # AttributeError if the symbol is missing
        fn.restype = rt
        fn.argtypes = at
    _lib = lib
    return lib


def _p(a, t):
    return a.ctypes.data_as(t)


class Context:
    """One GPU: sb200_ctx."""

    def __init__(self, device=0):
        self.lib = load_library()
        h = vp()
        rc = self.lib.sb200_create(device, C.byref(h))
        if rc != 0:
            raise Sb200Error(self.lib.sb200_last_error(None).decode())
        self.h = h
        self.device = device

    def check(self, rc):
        if rc != 0:
            raise Sb200Error(self.lib.sb200_last_error(self.h).decode())

    def synchronize(self):
        self.check(self.lib.sb200_synchronize(self.h))

    def stream(self):
        return self.lib.sb200_stream(self.h)

    def kernel_launches(self, reset=False):
        return self.lib.sb200_kernel_launches(self.h, int(reset))

    def profile(self, enable=True):
        self.check(self.lib.sb200_profile(self.h, int(enable)))

    def profile_report(self):
        """[(kernel name, launches, total ms)] sorted by time, measured with CUDA events on the launching stream"""
        need = C.c_uint64()
        self.check(self.lib.sb200_profile_report(self.h, None, 0, C.byref(need)))
        buf = C.create_string_buffer(need.value + 16)
        self.check(self.lib.sb200_profile_report(self.h, buf, need.value + 16, C.byref(need)))
        rows = []
        for line in buf.value.decode().splitlines():
            name, n, ms = line.split("\t")
            rows.append((name, int(n), float(ms)))
        return rows

    def close(self):
        if self.h:
            self.lib.sb200_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class ReadStreams:
    """io::ReadStreamList<io::SingleReadSeq> already converted to the binary 2-bit layout, resident in HBM."""

    def __init__(self, ctx, words, word_off, lens):
        self.ctx = ctx
        words = np.ascontiguousarray(words, dtype=np.uint64)
        if words.size == 0:
            words = np.zeros(1, dtype=np.uint64)
        word_off = np.ascontiguousarray(word_off, dtype=np.uint64)
        lens = np.ascontiguousarray(lens, dtype=np.uint32)
        self.n_reads = len(word_off) - 1
        if lens.size == 0:
            lens = np.zeros(1, dtype=np.uint32)
        h = vp()
        ctx.check(ctx.lib.sb200_reads_upload(ctx.h, _p(words, u64p), _p(word_off, u64p), _p(lens, u32p), self.n_reads, C.byref(h)))
        self.h = h

    @classmethod
    def from_device(cls, ctx, d_words, d_word_off, d_len, n_reads, n_words):
        """Device pointers (ints), e.g. torch tensors' data_ptr(): no host round trip."""
        self = cls.__new__(cls)
        self.ctx = ctx
        self.n_reads = n_reads
        h = vp()
        ctx.check(ctx.lib.sb200_reads_wrap_device(ctx.h, d_words, d_word_off, d_len, n_reads, n_words, C.byref(h)))
        self.h = h
        return self

    def free(self):
        if self.h:
            self.ctx.lib.sb200_reads_free(self.h)
            self.h = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class KMerDiskStorage:
    """Sorted-unique k-mers per hash bucket (device resident); accessors named after kmers::KMerDiskStorage."""

    def __init__(self, ctx, handle):
        self.ctx, self.h = ctx, handle
        lib = ctx.lib
        self._k = lib.sb200_kmers_k(handle)
        self.words = lib.sb200_kmers_words(handle)
        self._nb = lib.sb200_kmers_num_buckets(handle)
        self._size = lib.sb200_kmers_size(handle)
        self.instances = lib.sb200_kmers_instances(handle)
        bs = np.zeros(self._nb + 1, dtype=np.uint64)
        lib.sb200_kmers_bucket_starts(handle, _p(bs, u64p))
        self.bucket_starts = bs

    def k(self):
        return self._k

    def num_buckets(self):
        return self._nb

    def total_kmers(self):
        return self._size

    def bucket_size(self, i):
        return int(self.bucket_starts[i + 1] - self.bucket_starts[i])

    def _download(self, first, count):
        out = np.zeros((max(count, 1), self.words), dtype=np.uint64)
        self.ctx.check(self.ctx.lib.sb200_kmers_download(self.h, first, count, _p(out, u64p)))
        return out[:count]

    def bucket(self, i):
        """records of bucket i, i.e. the contents of the reference's kmers<i> file"""
        return self._download(int(self.bucket_starts[i]), self.bucket_size(i))

    def final_kmers(self):
        """merge(): all buckets concatenated, the reference's final_kmers file"""
        return self._download(0, self._size)

    def counts(self):
        out = np.zeros(max(self._size, 1), dtype=np.uint32)
        self.ctx.check(self.ctx.lib.sb200_kmers_counts_download(self.h, 0, self._size, _p(out, u32p)))
        return out[:self._size]

    def free(self):
        if self.h and getattr(self, "owned", True):
            self.ctx.lib.sb200_kmers_free(self.h)
        self.h = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class KMerDiskCounter:
    """kmers::KMerDiskCounter<RtSeq> over a DeBruijnReadKMerSplitter (reads) or a DeBruijnKMerKMerSplitter (k-mers)."""

    def __init__(self, ctx, source, K, canonical_only=True, add_rc=True):
        self.ctx, self.source, self.K = ctx, source, K
        self.canonical_only, self.add_rc = canonical_only, add_rc

    def Count(self, num_buckets, num_threads=1):
        h = vp()
        lib = self.ctx.lib
        if isinstance(self.source, ReadStreams):
            self.ctx.check(lib.sb200_count(self.ctx.h, self.source.h, self.K, int(self.canonical_only), int(self.add_rc),
                                           num_buckets, C.byref(h)))
        else:
            assert self.K == self.source.k() - 1
            self.ctx.check(lib.sb200_derive_kmers(self.ctx.h, self.source.h, num_buckets, C.byref(h)))
        return KMerDiskStorage(self.ctx, h)

    CountAll = Count


class KMerIndex:
    """kmers::KMerIndex: per-bucket BooPHF + segment starts, built by KMerIndexBuilder::BuildIndex(index, storage)."""

    def __init__(self, ctx, storage):
        self.ctx, self.storage = ctx, storage
        h = vp()
        ctx.check(ctx.lib.sb200_mphf_build(ctx.h, storage.h, C.byref(h)))
        self.h = h

    def size(self):
        return self.ctx.lib.sb200_mphf_size(self.h)

    def mem_size(self):
        return self.ctx.lib.sb200_mphf_mem_size(self.h)

    def seq_idx(self, records):
        records = np.ascontiguousarray(records, dtype=np.uint64).reshape(-1, self.storage.words)
        out = np.zeros(max(len(records), 1), dtype=np.uint64)
        self.ctx.check(self.ctx.lib.sb200_mphf_lookup(self.ctx.h, self.h, _p(records, u64p), len(records), _p(out, u64p)))
        return out[:len(records)]

    def seq_idx_one(self, record):
        """KMerIndex::seq_idx(const Seq &): one key, looked up on the host"""
        rec = np.ascontiguousarray(record, dtype=np.uint64)
        out = C.c_uint64()
        self.ctx.check(self.ctx.lib.sb200_mphf_seq_idx(self.h, _p(rec, u64p), C.byref(out)))
        return out.value

    def serialize(self):
        n = C.c_uint64()
        self.ctx.check(self.ctx.lib.sb200_mphf_serialize(self.h, None, C.byref(n)))
        buf = np.zeros(n.value, dtype=np.uint8)
        self.ctx.check(self.ctx.lib.sb200_mphf_serialize(self.h, _p(buf, u8p), C.byref(n)))
        return buf

    def free(self):
        if self.h and getattr(self, "owned", True):
            self.ctx.lib.sb200_mphf_free(self.h)
        self.h = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class DeBruijnExtensionIndex:
    """utils::DeBruijnExtensionIndex<>: KMerIndex + final k-mers + one InOutMask byte per k-mer."""

    def __init__(self, ctx, k):
        self.ctx, self._k = ctx, k
        self.kmers = None     # KMerDiskStorage of k-mers (KeyIteratingMap's final_kmers)
        self.index = None     # KMerIndex
        self.h = None         # sb200_ext

    def k(self):
        return self._k

    def size(self):
        return self.kmers.total_kmers()

    def data(self):
        """PerfectHashMap::data_: raw masks in MPHF-index order"""
        out = np.zeros(max(self.size(), 1), dtype=np.uint8)
        self.ctx.check(self.ctx.lib.sb200_ext_masks_download(self.h, _p(out, u8p)))
        return out[:self.size()]

    def idx(self):
        """ConstructKWH(kmer).idx() for every k-mer of final_kmers"""
        out = np.zeros(max(self.size(), 1), dtype=np.uint32)
        self.ctx.check(self.ctx.lib.sb200_ext_idx_download(self.h, _p(out, u32p)))
        return out[:self.size()].astype(np.uint64)

    def free(self):
        if self.h:
            self.ctx.lib.sb200_ext_free(self.h)
            self.h = None
        if self.index:
            self.index.free()
        if self.kmers:
            self.kmers.free()

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class DeBruijnExtensionIndexBuilder:
    def BuildExtensionIndexFromStream(self, index, streams, nthreads=1, num_buckets=None):
        """kmer_extension_index_builder.hpp:62-80.  Returns the (k+1)-mer storage, like the reference."""
        ctx = index.ctx
        nb = num_buckets if num_buckets is not None else 10 * nthreads
        kpomers = KMerDiskCounter(ctx, streams, index.k() + 1, canonical_only=True, add_rc=True).Count(nb, nthreads)
        self.BuildExtensionIndexFromKPOMers(index, kpomers, nthreads)
        return kpomers

    def BuildExtensionIndexFromKPOMers(self, index, kpomers, nthreads=1):
        """kmer_extension_index_builder.hpp:82-106"""
        ctx = index.ctx
        if kpomers.k() != index.k() + 1:
            raise Sb200Error("VERIFY(kpomers.k() == index.k() + 1)")
        index.kmers = KMerDiskCounter(ctx, kpomers, index.k()).Count(kpomers.num_buckets(), nthreads)
        index.index = KMerIndex(ctx, index.kmers)
        h = vp()
        ctx.check(ctx.lib.sb200_ext_build(ctx.h, kpomers.h, index.kmers.h, index.index.h, C.byref(h)))
        index.h = h


class EarlyTipClipperProcessor:
    def __init__(self, index, length_bound):
        self.index, self.length_bound = index, length_bound

    def ClipTips(self):
        ix = self.index
        removed = C.c_uint64()
        ix.ctx.check(ix.ctx.lib.sb200_tipclip(ix.ctx.h, ix.kmers.h, ix.index.h, ix.h, self.length_bound, C.byref(removed)))
        return removed.value


def unpack_sequences(words, word_off, lens):
    """packed 2-bit sequences -> list[str]"""
    out = []
    lut = np.frombuffer(b"ACGT", dtype=np.uint8)
    shifts = (2 * np.arange(32, dtype=np.uint64))
    for i in range(len(lens)):
        w = words[int(word_off[i]):int(word_off[i + 1])]
        codes = ((w[:, None] >> shifts[None, :]) & np.uint64(3)).astype(np.uint8).reshape(-1)[:int(lens[i])]
        out.append(lut[codes].tobytes().decode())
    return out


class UnbranchingPathExtractor:
    def __init__(self, index, k):
        self.index, self.k = index, k
        self.n_loops = 0
        self.h = None   # the last result stays in HBM (CoverageHashMap.edge_coverage reads it there) until free()

    def _run(self, with_loops):
        ix = self.index
        lib, ctx = ix.ctx.lib, ix.ctx
        self.free()
        h = vp()
        ctx.check(lib.sb200_unitigs_extract(ctx.h, ix.kmers.h, ix.index.h, ix.h, int(with_loops), C.byref(h)))
        self.h = h
        self.n_loops = lib.sb200_unitigs_loops(h)
        return download_unitigs(ctx, h)

    def free(self):
        if self.h:
            self.index.ctx.lib.sb200_unitigs_free(self.h)
            self.h = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass

    def ExtractUnbranchingPaths(self, nchunks=1, packed=False):
        r = self._run(False)
        return r if packed else unpack_sequences(*r)

    def ExtractUnbranchingPathsAndLoops(self, nchunks=1, packed=False):
        r = self._run(True)
        return r if packed else unpack_sequences(*r)


class CoverageHashMap:
    """utils::PerfectHashMap<RtSeq, uint32_t> of CoverageHashMapBuilder::BuildIndex (ph_map/coverage_hash_map_builder.hpp:39-54): a KMerIndex
    over the (k+1)-mer storage + the multiplicity of every (k+1)-mer in index order; edge_coverage() is GraphCoverageFiller
    (assembly_graph/graph_support/coverage_filling.hpp:45-95) over a unitig set."""

    def __init__(self, ctx, kpomers):
        self.ctx, self.kpomers = ctx, kpomers
        h = vp()
        ctx.check(ctx.lib.sb200_coverage_map_build(ctx.h, kpomers.h, C.byref(h)))
        self.h = h
        self.index = KMerIndex.__new__(KMerIndex)
        self.index.ctx, self.index.storage, self.index.h, self.index.owned = ctx, kpomers, vp(ctx.lib.sb200_coverage_map_index(h)), False

    def data(self):
        n = self.ctx.lib.sb200_coverage_map_size(self.h)
        out = np.zeros(max(n, 1), dtype=np.uint32)
        self.ctx.check(self.ctx.lib.sb200_coverage_map_values_download(self.h, _p(out, u32p)))
        return out[:n]

    def edge_coverage(self, unitigs_handle, averaging_range=50):
        """(kc, flank) for the sequences of a sb200_unitigs handle: kc[i] = KC:i of GFA, flank[i] = (start, end) raw flanking coverage"""
        n = self.ctx.lib.sb200_unitigs_count(unitigs_handle)
        kc = np.zeros(max(n, 1), dtype=np.uint64)
        fl = np.zeros(max(2 * n, 2), dtype=np.uint64)
        self.ctx.check(self.ctx.lib.sb200_unitigs_coverage(self.ctx.h, self.h, unitigs_handle, averaging_range, _p(kc, u64p), _p(fl, u64p)))
        return kc[:n], fl[:2 * n].reshape(-1, 2)

    def free(self):
        if self.h:
            self.ctx.lib.sb200_coverage_map_free(self.h)
            self.h = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class Graph:
    """Result of sb200_construct: host-resident (pinned) buffers of the whole path."""

    def __init__(self, ctx, handle):
        self.ctx, self.h = ctx, handle
        v = GraphView()
        ctx.lib.sb200_graph_get(handle, C.byref(v))
        self.view = v

    def _arr(self, ptr, n, dtype):
        if not ptr or n == 0:
            return np.zeros(0, dtype=dtype)
        return np.ctypeslib.as_array(ptr, shape=(n,))

    def masks(self):
        return self._arr(self.view.masks, self.view.n_kmers, np.uint8)

    def index_bytes(self):
        return self._arr(self.view.index_bytes, self.view.index_size, np.uint8)

    def unitigs_packed(self):
        v = self.view
        return (self._arr(v.unitig_words, v.n_unitig_words, np.uint64), self._arr(v.unitig_word_off, v.n_unitigs + 1, np.uint64),
                self._arr(v.unitig_len, v.n_unitigs, np.uint32))

    def unitigs(self):
        return unpack_sequences(*self.unitigs_packed())

    def free(self):
        if self.h:
            self.ctx.lib.sb200_graph_free(self.h)
            self.h = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


def construct(ctx, words, word_off, lens, k, num_buckets, tip_clip=False, tip_length_bound=0, with_loops=True, fetch_kmers=False):
    """Whole path with host buffers on both sides (spades-gbuilder's body, A/projects/gbuilder/main.cpp:165-181)."""
    p = ConstructParams(k, num_buckets, int(tip_clip), tip_length_bound, int(with_loops), int(fetch_kmers))
    h = vp()
    n = len(word_off) - 1
    ctx.check(ctx.lib.sb200_construct(ctx.h, _p(words, u64p), _p(word_off, u64p), _p(lens, u32p), n, C.byref(p), C.byref(h)))
    return Graph(ctx, h)


# ---- hash-sharded path: one call per rank (csrc/shard.cu) -----------------------------------------------------------------------
def download_unitigs(ctx, h):
    """(words, word_off, len) of a sb200_unitigs handle"""
    lib = ctx.lib
    n, nw = lib.sb200_unitigs_count(h), lib.sb200_unitigs_total_words(h)
    words = np.zeros(max(nw, 1), dtype=np.uint64)
    off = np.zeros(n + 1, dtype=np.uint64)
    lens = np.zeros(max(n, 1), dtype=np.uint32)
    ctx.check(lib.sb200_unitigs_download(h, _p(words, u64p), _p(off, u64p), _p(lens, u32p)))
    return words[:nw], off, lens[:n]


class Comm:
    """sb200_comm: NCCL (one rank per GPU) or local (virtual ranks = threads of this process)."""

    def __init__(self, lib, handle):
        self.lib, self.h = lib, handle
        self.rank, self.size = lib.sb200_comm_rank(handle), lib.sb200_comm_size(handle)

    @staticmethod
    def unique_id():
        lib = load_library()
        buf = np.zeros(128, dtype=np.uint8)
        if lib.sb200_comm_unique_id(_p(buf, u8p)) != 0:
            raise Sb200Error(lib.sb200_comm_last_error().decode())
        return buf

    @classmethod
    def nccl(cls, ctx, rank, world, unique_id):
        uid = np.ascontiguousarray(unique_id, dtype=np.uint8)
        h = vp()
        ctx.check(ctx.lib.sb200_comm_create_nccl(ctx.h, rank, world, _p(uid, u8p), C.byref(h)))
        return cls(ctx.lib, h)

    @classmethod
    def local(cls, world):
        lib = load_library()
        arr = (vp * world)()
        if lib.sb200_comm_create_local(world, arr) != 0:
            raise Sb200Error(lib.sb200_comm_last_error().decode())
        return [cls(lib, vp(arr[r])) for r in range(world)]

    def free(self):
        if self.h:
            self.lib.sb200_comm_free(self.h)
            self.h = None


class Shard:
    """One rank's result of sb200_construct_sharded: its shards of both k-mer tables, the whole index + masks, its unitig slice."""

    def __init__(self, ctx, handle):
        self.ctx, self.h = ctx, handle
        lib = ctx.lib
        info = ShardInfo()
        lib.sb200_shard_info(handle, C.byref(info))
        self.info = info
        self.stage_ms = {STAGE_NAMES[i]: float(info.stage_ms[i]) for i in range(8)}
        self.kpomers = KMerDiskStorage(ctx, vp(lib.sb200_shard_kpomers(handle)))
        self.kpomers.owned = False
        self.kmers = KMerDiskStorage(ctx, vp(lib.sb200_shard_kmers(handle)))
        self.kmers.owned = False
        self.index = KMerIndex.__new__(KMerIndex)
        self.index.ctx, self.index.storage, self.index.h, self.index.owned = ctx, self.kmers, vp(lib.sb200_shard_mphf(handle)), False
        self.ext = vp(lib.sb200_shard_ext(handle))
        self.unitigs_h = vp(lib.sb200_shard_unitigs(handle))

    def masks(self):
        """PerfectHashMap::data_ of the WHOLE index (every rank holds all of it)"""
        out = np.zeros(max(self.info.total_kmers, 1), dtype=np.uint8)
        self.ctx.check(self.ctx.lib.sb200_ext_masks_download(self.ext, _p(out, u8p)))
        return out[:self.info.total_kmers]

    def idx(self):
        n = self.kmers.total_kmers()
        out = np.zeros(max(n, 1), dtype=np.uint32)
        self.ctx.check(self.ctx.lib.sb200_ext_idx_download(self.ext, _p(out, u32p)))
        return out[:n].astype(np.uint64)

    def walk_stats(self):
        out = np.zeros(6 * self.info.size, dtype=np.uint64)
        self.ctx.lib.sb200_shard_walk_stats(self.h, _p(out, u64p))
        return out.reshape(self.info.size, 6)

    def unitigs_packed(self):
        return download_unitigs(self.ctx, self.unitigs_h)

    def unitigs(self):
        return unpack_sequences(*self.unitigs_packed())

    def free(self):
        if self.h:
            self.ctx.lib.sb200_shard_free(self.h)
            self.h = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


def construct_sharded(ctx, comm, reads, k, num_buckets, tip_clip=False, tip_length_bound=0, with_loops=True, gather_to=-1):
    """This rank's part of the hash-sharded path (every rank of `comm` must call it, each with its own slice of the reads)."""
    p = ConstructParams(k, num_buckets, int(tip_clip), tip_length_bound, int(with_loops), 0)
    h = vp()
    ctx.check(ctx.lib.sb200_construct_sharded(ctx.h, comm.h, reads.h, C.byref(p), int(gather_to), C.byref(h)))
    return Shard(ctx, h)


class MultiContext:
    """sb200_multi: several GPUs (or virtual ranks on one GPU when a device id repeats) driven by the library from ONE host thread —
    SURVEY.md 8(b)'s sb200_create(n_gpus, device_ids)."""

    def __init__(self, device_ids):
        self.lib = load_library()
        ids = (C.c_int * len(device_ids))(*device_ids)
        h = vp()
        if self.lib.sb200_multi_create(len(device_ids), ids, C.byref(h)) != 0:
            raise Sb200Error(self.lib.sb200_multi_last_error(None).decode())
        self.h = h

    def construct(self, words, word_off, lens, k, num_buckets, tip_clip=False, tip_length_bound=0, with_loops=True, fetch_kmers=True):
        p = ConstructParams(k, num_buckets, int(tip_clip), tip_length_bound, int(with_loops), int(fetch_kmers))
        words = np.ascontiguousarray(words, dtype=np.uint64)
        word_off = np.ascontiguousarray(word_off, dtype=np.uint64)
        lens = np.ascontiguousarray(lens, dtype=np.uint32)
        g = vp()
        if self.lib.sb200_multi_construct(self.h, _p(words, u64p), _p(word_off, u64p), _p(lens, u32p), len(word_off) - 1, C.byref(p), C.byref(g)) != 0:
            raise Sb200Error(self.lib.sb200_multi_last_error(self.h).decode())

        class _Ctx:   # what Graph needs from a context
            lib = self.lib
        return Graph(_Ctx, g)

    def close(self):
        if self.h:
            self.lib.sb200_multi_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
