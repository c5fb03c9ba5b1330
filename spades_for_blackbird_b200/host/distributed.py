"""Launching the hash-sharded path: one rank per GPU (NCCL) or G virtual ranks on one GPU (threads).

The orchestration itself is C++ (csrc/shard.cu over csrc/comm.cu: NCCL C API / local copies) behind ONE C-ABI call per rank,
`sb200_construct_sharded` (include/sb200.h).  What is left to the host language is what a launcher does:
  * hand the 128-byte NCCL id from rank 0 to the other ranks (here: over torch.distributed, whatever backend it runs on),
  * give every rank its slice of the reads,
  * run G virtual ranks as threads against a local communicator (tests on a one-GPU box).

The k-mer space is sharded with the reference's own bucket function (KMerSegmentPolicy, C/utils/kmer_mph/kmer_buckets.hpp:28-41):
rank g of G owns the buckets [g*B/G, (g+1)*B/G), a contiguous range of the reference's file order, so the shards concatenated in
rank order ARE the single-GPU (= reference) result.

`count_shard` + `TorchComm` below are the PROTOCOL MODEL of the two record exchanges (group by owner, all-to-all of variable-sized
groups, per-owner sort / dedup / count) over plain torch tensors: tests/test_distributed_cpu.py runs it with gloo, world size 2, and
a CPU stand-in for the per-rank compute, so that ownership and exchange bookkeeping are covered where no GPU exists.
"""
import threading

import numpy as np

from . import binding as B


# ------------------------------------------------------------------------------------------------------------- launchers
def nccl_comm(ctx, rank=None, world=None):
    """NCCL communicator of the library for this rank.  torch.distributed must be initialised (any backend): it only carries
    the 128-byte id from rank 0 to the others."""
    import torch
    import torch.distributed as dist
    rank = dist.get_rank() if rank is None else rank
    world = dist.get_world_size() if world is None else world
    uid = torch.from_numpy(B.Comm.unique_id().copy()) if rank == 0 else torch.zeros(128, dtype=torch.uint8)
    if world > 1:
        if dist.get_backend() == "nccl":
            dev = torch.device("cuda", ctx.device)
            t = uid.to(dev)
            dist.broadcast(t, src=0)
            uid = t.cpu()
        else:
            dist.broadcast(uid, src=0)
    return B.Comm.nccl(ctx, rank, world, uid.numpy())


def slice_reads(words, word_off, lens, rank, world):
    """reads [n*rank/world, n*(rank+1)/world) of a packed read set, re-based"""
    n = len(lens)
    lo, hi = n * rank // world, n * (rank + 1) // world
    w0, w1 = int(word_off[lo]), int(word_off[hi])
    return words[w0:w1], word_off[lo:hi + 1] - word_off[lo], lens[lo:hi]


def run_virtual_ranks(G, work, devices=None):
    """work(rank, ctx, comm) -> result, on G threads with one context each (device devices[r], default 0) and a local
    communicator.  Returns the list of results; the first real error of any rank is raised."""
    comms = B.Comm.local(G)
    results, errors = [None] * G, [None] * G

    def run(r):
        ctx = None
        try:
            ctx = B.Context(devices[r] if devices else 0)
            results[r] = work(r, ctx, comms[r])
        except Exception as e:   # noqa: BLE001
            errors[r] = e
        finally:
            if ctx is not None:
                ctx.close()

    threads = [threading.Thread(target=run, args=(r,)) for r in range(G)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    for c in comms:
        c.free()
    real = [e for e in errors if e is not None and "another rank of the local communicator failed" not in str(e)]
    if real:
        raise real[0]
    for e in errors:
        if e is not None:
            raise e
    return results


# ---------------------------------------------------------------------------------------------- protocol model (CPU tests)
class TorchComm:
    """torch.distributed (gloo for CPU tensors)."""

    def __init__(self, group=None):
        import torch.distributed as dist
        self.dist, self.group = dist, group
        self.rank, self.size = dist.get_rank(group), dist.get_world_size(group)

    def all_gather_obj(self, obj):
        out = [None] * self.size
        self.dist.all_gather_object(out, obj, group=self.group)
        return out

    def all_to_all_v(self, send, send_counts, width):
        """send: 1-D int64 tensor holding sum(send_counts) records of `width` words, grouped by destination"""
        import torch
        counts = self.all_gather_obj([int(c) for c in send_counts])
        recv_counts = [counts[src][self.rank] for src in range(self.size)]
        recv = torch.empty(sum(recv_counts) * width, dtype=torch.int64, device=send.device)
        self.dist.all_to_all_single(recv, send, [c * width for c in recv_counts], [int(c) * width for c in send_counts], group=self.group)
        return recv, recv_counts

    def all_reduce_sum_(self, t):
        self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM, group=self.group)
        return t

    def barrier(self):
        self.dist.barrier(group=self.group)


def count_shard(backend, comm, make_records, K, num_buckets, want_counts):
    """the record exchange of steps 1-2 / 3-4 (csrc/shard.cu exchange_and_count): group by owner, all-to-all, sort / dedup /
    count what arrived.  `backend` supplies the per-rank compute."""
    rec, view, counts, width = make_records()
    flags = backend.records_flags(rec)   # the same on every rank: double palindromes / marker / mask payload
    recv, recv_counts = comm.all_to_all_v(view, counts, width)
    backend.free_records(rec)
    return backend.count(recv, sum(recv_counts), K, num_buckets, want_counts, flags)
