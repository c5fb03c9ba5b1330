"""Hash-sharded graph construction over several GPUs of one node: one process (rank) per GPU.

The k-mer space is sharded with the reference's own bucket function (KMerSegmentPolicy, C/utils/kmer_mph/kmer_buckets.hpp:
28-41): rank g of G owns the buckets [g*B/G, (g+1)*B/G), a contiguous range of the reference's file order, so the shards
concatenated in rank order ARE the single-GPU (= reference) result.  Exchange steps (the reference does the same shuffle
through kmers_raw<i> files, kmer_splitter.hpp:140-161):

    1. reads are split by read index; every rank extracts canonical (k+1)-mer instances and groups them by owner
    2. all-to-all #1 (instances)            -> owners sort / deduplicate / count: their shard of the (k+1)-mer storage
    3. owners derive k-mer candidates, group them by the owner of the K-MER
    4. all-to-all #2 (candidates)           -> owners deduplicate: their shard of the k-mer table (final_kmers)
    5. all-gather of bucket sizes (tiny)    -> segment starts / level geometry of the whole KMerIndex on every rank
    6. every rank builds the BooPHF levels of its own buckets inside the global layout; all-reduce(sum) of the bit-vectors
       and rank samples (disjoint supports: sum == or) -> the whole index on every rank
    7. every rank sets the mask bits of its own (k+1)-mers in a global mask array; all-reduce(sum) (distinct (k+1)-mers set
       distinct bits) -> all masks on every rank
    8. every rank walks the start edges of the junctions in its own k-mer shard (needs only masks + index): its slice of
       the unitig list, already in the reference's global order; gather to rank 0.

Compute is behind a small backend interface and communication behind a `Comm`, so the orchestration below runs unchanged
on NCCL (bench.py, one rank per B200), on gloo with a CPU backend (tests/test_distributed_cpu.py, world_size 2) and with
several virtual ranks inside one process on one GPU (tests/test_gpu_sharded.py).
"""
import ctypes as C
import threading

import numpy as np

from . import binding as B


# ---------------------------------------------------------------------------------------------------------------- comms
class TorchComm:
    """torch.distributed (nccl for CUDA tensors, gloo for CPU tensors)."""

    def __init__(self, group=None):
        import torch.distributed as dist
        self.dist, self.group = dist, group
        self.rank, self.size = dist.get_rank(group), dist.get_world_size(group)

    def all_gather_obj(self, obj):
        out = [None] * self.size
        self.dist.all_gather_object(out, obj, group=self.group)
        return out

    def all_to_all_v(self, send, send_counts, width, alloc=None):
        """send: 1-D int64 tensor holding sum(send_counts) records of `width` words, grouped by destination.  alloc(n) may
        provide the receive buffer for n records (e.g. the library's own record array: no staging copy)."""
        import torch
        counts = self.all_gather_obj([int(c) for c in send_counts])
        recv_counts = [counts[src][self.rank] for src in range(self.size)]
        n = sum(recv_counts)
        recv = alloc(n) if alloc else torch.empty(n * width, dtype=torch.int64, device=send.device)
        self.dist.all_to_all_single(recv, send, [c * width for c in recv_counts], [int(c) * width for c in send_counts],
                                    group=self.group)
        return recv, recv_counts

    def all_reduce_sum_(self, t):
        self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM, group=self.group)
        return t

    def gather_v(self, t, dst=0):
        """variable-length 1-D tensors -> list on dst (None elsewhere)"""
        import torch
        sizes = self.all_gather_obj(int(t.numel()))
        if self.rank == dst:
            out = [torch.empty(s, dtype=t.dtype, device=t.device) for s in sizes]
            reqs = [self.dist.irecv(out[src], src=src, group=self.group) for src in range(self.size) if src != dst and sizes[src]]
            out[dst].copy_(t)
            for r in reqs:
                r.wait()
            return out
        if t.numel():
            self.dist.send(t, dst=dst, group=self.group)
        return None

    def barrier(self):
        self.dist.barrier(group=self.group)


class LocalComm:
    """G virtual ranks = G threads of one process (each with its own sb200 context on the same GPU, or on the CPU)."""

    class Shared:
        def __init__(self, size):
            self.size = size
            self.barrier = threading.Barrier(size)
            self.slots = [None] * size

    def __init__(self, shared, rank):
        self.sh, self.rank, self.size = shared, rank, shared.size

    def _exchange(self, obj):
        self.sh.slots[self.rank] = obj
        self.sh.barrier.wait()
        out = list(self.sh.slots)
        self.sh.barrier.wait()
        return out

    def all_gather_obj(self, obj):
        return self._exchange(obj)

    def all_to_all_v(self, send, send_counts, width, alloc=None):
        import torch
        offs = np.concatenate([[0], np.cumsum(send_counts)]).astype(np.int64) * width
        parts = [send[int(offs[d]):int(offs[d + 1])] for d in range(self.size)]
        allparts = self._exchange(parts)
        mine = [allparts[src][self.rank] for src in range(self.size)]
        if alloc:
            recv = alloc(sum(int(p.numel()) for p in mine) // width)
            o = 0
            for p in mine:
                recv[o:o + p.numel()].copy_(p)
                o += p.numel()
            if recv.is_cuda:
                torch.cuda.synchronize(recv.device)
        else:
            recv = torch.cat([p.clone() for p in mine]) if mine else send[:0]
        self.sh.barrier.wait()   # senders keep their buffers alive until everyone has copied
        return recv, [int(p.numel()) // width for p in mine]

    def all_reduce_sum_(self, t):
        allt = self._exchange(t)
        if self.rank == 0:
            acc = allt[0].clone()
            for o in allt[1:]:
                acc += o
            self.sh.slots[0] = acc
        self.sh.barrier.wait()
        res = self.sh.slots[0]
        self.sh.barrier.wait()
        t.copy_(res)
        self.sh.barrier.wait()
        return t

    def gather_v(self, t, dst=0):
        allt = self._exchange(t.clone())
        return allt if self.rank == dst else None

    def barrier(self):
        self.sh.barrier.wait()


# ------------------------------------------------------------------------------------------------------------- backends
class _CudaView:
    """zero-copy torch view of device memory owned by libspades_b200 (via __cuda_array_interface__)"""

    def __init__(self, ptr, n, typestr):
        self.__cuda_array_interface__ = {"shape": (int(n),), "typestr": typestr, "data": (int(ptr), False), "version": 2}


def cuda_view(ptr, n, typestr, device):
    import torch
    if n == 0:
        return torch.empty(0, dtype={"<i8": torch.int64, "|u1": torch.uint8, "<i4": torch.int32}[typestr], device=device)
    return torch.as_tensor(_CudaView(ptr, n, typestr), device=device)


class GpuShardBackend:
    """One rank's compute on its B200 through the C ABI (include/sb200.h, 'hash-sharded path')."""

    def __init__(self, ctx, device):
        self.ctx, self.lib, self.device = ctx, ctx.lib, device

    def sync(self):
        import torch
        torch.cuda.synchronize(self.device)

    # -- records --------------------------------------------------------------------------------------------------------
    def _partition(self, rec, num_buckets, n_owners):
        counts = np.zeros(n_owners, dtype=np.uint64)
        self.ctx.check(self.lib.sb200_records_partition(self.ctx.h, rec, num_buckets, n_owners, counts.ctypes.data_as(B.u64p)))
        n, w = self.lib.sb200_records_size(rec), self.lib.sb200_records_words(rec)
        view = cuda_view(self.lib.sb200_records_device(rec), n * w, "<i8", self.device)
        return view, counts.astype(np.int64).tolist(), w

    def extract_partition(self, reads, K, num_buckets, n_owners):
        rec = B.vp()
        counts = np.zeros(n_owners, dtype=np.uint64)
        self.ctx.check(self.lib.sb200_records_extract_partitioned(self.ctx.h, reads.h, K, 1, 1, num_buckets, n_owners,
                                                                  counts.ctypes.data_as(B.u64p), C.byref(rec)))
        n, w = self.lib.sb200_records_size(rec), self.lib.sb200_records_words(rec)
        view = cuda_view(self.lib.sb200_records_device(rec), max(n * w, 1), "<i8", self.device)[:n * w]
        return rec, view, counts.astype(np.int64).tolist(), w

    def derive_partition(self, kpomers, num_buckets, n_owners):
        rec = B.vp()
        self.ctx.check(self.lib.sb200_records_derive(self.ctx.h, kpomers.h, C.byref(rec)))
        view, counts, w = self._partition(rec, num_buckets, n_owners)
        return rec, view, counts, w

    def free_records(self, rec):
        self.lib.sb200_records_free(rec)

    def records_flags(self, rec):
        return int(self.lib.sb200_records_flags(rec))

    def alloc_records(self, n, K, flags):
        """empty record array for n received records + its int64 view (the all-to-all writes straight into it)"""
        rec = B.vp()
        self.ctx.check(self.lib.sb200_records_alloc(self.ctx.h, n, K, int(flags), C.byref(rec)))
        w = self.lib.sb200_records_words(rec)
        return rec, cuda_view(self.lib.sb200_records_device(rec), max(n * w, 1), "<i8", self.device)[:n * w]

    def count_records(self, rec, num_buckets, want_counts, owned=None):
        """owned = (first_bucket, n_owned): the bucket range every received record lies in (this rank's share)"""
        h = B.vp()
        try:
            if owned is None:
                self.ctx.check(self.lib.sb200_count_records(self.ctx.h, rec, num_buckets, int(want_counts), C.byref(h)))
            else:
                self.ctx.check(self.lib.sb200_count_records_owned(self.ctx.h, rec, num_buckets, int(owned[0]), int(owned[1]), int(want_counts),
                                                                  C.byref(h)))
        finally:
            self.lib.sb200_records_free(rec)
        return B.KMerDiskStorage(self.ctx, h)

    def count(self, recv, n, K, num_buckets, want_counts, flags):
        """recv: int64 CUDA tensor with n records (flags = records_flags of the senders) -> this rank's shard as a KMerDiskStorage"""
        rec = B.vp()
        self.ctx.check(self.lib.sb200_records_alloc(self.ctx.h, n, K, int(flags), C.byref(rec)))
        w = self.lib.sb200_records_words(rec)
        if n:
            cuda_view(self.lib.sb200_records_device(rec), n * w, "<i8", self.device).copy_(recv)
            self.sync()
        h = B.vp()
        try:
            self.ctx.check(self.lib.sb200_count_records(self.ctx.h, rec, num_buckets, int(want_counts), C.byref(h)))
        finally:
            self.lib.sb200_records_free(rec)
        return B.KMerDiskStorage(self.ctx, h)

    # -- index, masks, unitigs ---------------------------------------------------------------------------------------------
    def mphf_build(self, kmers, global_sizes):
        gs = np.ascontiguousarray(global_sizes, dtype=np.uint64)
        h = B.vp()
        self.ctx.check(self.lib.sb200_mphf_build_sharded(self.ctx.h, kmers.h, gs.ctypes.data_as(B.u64p), C.byref(h)))
        idx = B.KMerIndex.__new__(B.KMerIndex)
        idx.ctx, idx.storage, idx.h = self.ctx, kmers, h
        return idx

    def mphf_arrays(self, index):
        bits, ranks = B.vp(), B.vp()
        nb, nr = C.c_uint64(), C.c_uint64()
        self.lib.sb200_mphf_arrays(index.h, C.byref(bits), C.byref(nb), C.byref(ranks), C.byref(nr))
        return cuda_view(bits.value, nb.value, "<i8", self.device), cuda_view(ranks.value, nr.value, "<i8", self.device)

    def mphf_complete(self, index):
        """after the all-reduce of the bit-vectors: per-word prefix popcounts for one-read ranks"""
        self.ctx.check(self.lib.sb200_mphf_complete(self.ctx.h, index.h))

    def ext_build(self, kpomers, kmers, index):
        h = B.vp()
        self.ctx.check(self.lib.sb200_ext_build(self.ctx.h, kpomers.h, kmers.h, index.h, C.byref(h)))
        return h

    def ext_masks(self, ext):
        p, n = B.vp(), C.c_uint64()
        self.lib.sb200_ext_masks_device(ext, C.byref(p), C.byref(n))
        return cuda_view(p.value, n.value, "|u1", self.device)

    def free_ext(self, ext):
        self.lib.sb200_ext_free(ext)

    def unitigs_local(self, kmers, index, ext):
        stats = np.zeros(6, dtype=np.uint64)
        h = B.vp()
        self.ctx.check(self.lib.sb200_unitigs_extract_local(self.ctx.h, kmers.h, index.h, ext, stats.ctypes.data_as(B.u64p), C.byref(h)))
        return stats, (h if h.value else None)

    def unitigs_views(self, u):
        n, nw = self.lib.sb200_unitigs_count(u), self.lib.sb200_unitigs_total_words(u)
        w, o, ln = B.vp(), B.vp(), B.vp()
        self.lib.sb200_unitigs_device(u, C.byref(w), C.byref(o), C.byref(ln))
        return (cuda_view(w.value, nw, "<i8", self.device), cuda_view(o.value, n + 1, "<i8", self.device),
                cuda_view(ln.value, n, "<i4", self.device))

    def free_unitigs(self, u):
        self.lib.sb200_unitigs_free(u)


# ------------------------------------------------------------------------------------------------------------- the path
class ShardedResult:
    def __init__(self):
        self.kpomers = self.kmers = self.index = self.ext = self.unitigs = None
        self.global_kmer_bucket_sizes = None
        self.stats = None
        self.gathered = None   # rank 0: (words, word_off, len) lists per rank


def count_shard(backend, comm, make_records, K, num_buckets, want_counts, double_palindromes):
    """steps 1-2 / 3-4: group by owner, all-to-all, sort/dedup/count the received records"""
    rec, view, counts, width = make_records()
    flags = backend.records_flags(rec)   # the same on every rank: double palindromes / marker / mask payload
    if hasattr(backend, "alloc_records"):   # receive straight into the library's record array
        holder = {}

        def alloc(n):
            holder["rec"], v = backend.alloc_records(n, K, flags)
            return v
        comm.all_to_all_v(view, counts, width, alloc=alloc)
        backend.sync()   # the collective ran on torch's / NCCL's stream; the library works on its own
        backend.free_records(rec)
        n_owned = num_buckets // comm.size   # owner g holds buckets [g B/G, (g+1) B/G): sb200_records_partition's owner function
        return backend.count_records(holder["rec"], num_buckets, want_counts, owned=(comm.rank * n_owned, n_owned))
    recv, recv_counts = comm.all_to_all_v(view, counts, width)
    backend.sync()
    backend.free_records(rec)
    return backend.count(recv, sum(recv_counts), K, num_buckets, want_counts, flags)


def construct_sharded(backend, comm, reads, k, num_buckets, gather_to=0, keep=False):
    """The whole path for this rank's slice of the reads.  Returns a ShardedResult (device-resident shard handles)."""
    G = comm.size
    if num_buckets % G:
        raise B.Sb200Error("num_buckets (%d) must be a multiple of the number of GPUs (%d)" % (num_buckets, G))
    res = ShardedResult()
    import time as _time
    marks = []

    def mark(label):   # stages are blocking on the library's stream; collectives are followed by backend.sync()
        marks.append((label, _time.perf_counter()))
    mark("start")
    res.kpomers = count_shard(backend, comm, lambda: backend.extract_partition(reads, k + 1, num_buckets, G), k + 1, num_buckets,
                              True, True)
    mark("count_kpomers (partition + all-to-all + sort)")
    res.kmers = count_shard(backend, comm, lambda: backend.derive_partition(res.kpomers, num_buckets, G), k, num_buckets,
                            False, False)
    mark("count_kmers (derive + partition + all-to-all + sort)")
    # 5. global bucket sizes: every bucket is non-empty on exactly one rank
    local_sizes = np.diff(res.kmers.bucket_starts).astype(np.int64)
    sizes = np.sum(np.stack(comm.all_gather_obj(local_sizes)), axis=0).astype(np.uint64)
    res.global_kmer_bucket_sizes = sizes
    # 6. index: own buckets inside the global layout, then sum over ranks
    res.index = backend.mphf_build(res.kmers, sizes)
    bits, ranks = backend.mphf_arrays(res.index)
    if G > 1:
        comm.all_reduce_sum_(bits)
        comm.all_reduce_sum_(ranks)
        backend.sync()
    if hasattr(backend, "mphf_complete"):
        backend.mphf_complete(res.index)   # the index is whole now (also for G = 1 through this driver: the build saw global sizes)
    mark("mphf (build + all-reduce)")
    # 7. masks
    res.ext = backend.ext_build(res.kpomers, res.kmers, res.index)
    masks = backend.ext_masks(res.ext)
    if G > 1:
        comm.all_reduce_sum_(masks)
        backend.sync()
    mark("masks (fill + all-reduce)")
    # 8. unitigs of the junctions in my shard
    stats, u = backend.unitigs_local(res.kmers, res.index, res.ext)
    mark("unitigs (local walks)")
    allstats = np.stack(comm.all_gather_obj(stats.astype(np.int64)))
    res.stats = allstats
    long_chains = int(allstats[:, 1].sum())
    loops = int(allstats[:, 0].sum()) != 2 * int(allstats[0, 5])
    if long_chains or loops:
        raise B.Sb200Error("the sharded extraction met %s: run this input through the single-GPU path (pointer jumping) "
                           "[per-rank stats (chain vertices, long chains, edges, kept, bases, non-junction k-mers): %s]" % (
                               "chains longer than the direct-walk limit" if long_chains else "perfect loops", allstats.tolist()))
    res.unitigs = u
    if gather_to is not None:
        w, o, ln = backend.unitigs_views(u)
        gw, go, gl = comm.gather_v(w, gather_to), comm.gather_v(o, gather_to), comm.gather_v(ln, gather_to)
        if comm.rank == gather_to:
            res.gathered = (gw, go, gl)
        backend.sync()
        mark("gather")
    res.stage_ms = {marks[i][0]: 1e3 * (marks[i][1] - marks[i - 1][1]) for i in range(1, len(marks))}
    return res


def unpack_gathered(gathered):
    """rank 0: concatenated unitig strings in global (= reference) order"""
    out = []
    gw, go, gl = gathered
    for w, o, ln in zip(gw, go, gl):
        w = w.cpu().numpy().view(np.uint64)
        o = o.cpu().numpy().view(np.uint64)
        ln = ln.cpu().numpy().view(np.uint32)
        out += B.unpack_sequences(w, o, ln)
    return out
