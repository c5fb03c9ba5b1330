"""world_size-2 gloo test of the exchange protocol of the sharded path (the model in spades_for_blackbird_b200/host/distributed.py of
what csrc/shard.cu does over NCCL): ownership by bucket range, all-to-all of variable-sized record groups, shard-wise sort / dedup /
count — with a small CPU stand-in for the per-rank compute (pure Python on the oracle's primitives; the CUDA path is tested on the
GPU in tests/test_gpu_sharded.py and, on real NCCL, tests/test_gpu_nccl.py).  Shards concatenated in rank order must equal the oracle's
single-process k-mer sets."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _pack(s):
    w = [0, 0, 0, 0]
    for i, ch in enumerate(s):
        w[i // 32] |= "ACGT".index(ch) << (2 * (i % 32))
    return w[:(len(s) + 31) // 32]


def _rc(s):
    return s[::-1].translate(str.maketrans("ACGT", "TGCA"))


class CpuShardBackend:
    """the GpuShardBackend calls used by count_shard(), on CPU tensors"""

    def __init__(self, O):
        self.O = O

    def sync(self):
        pass

    def free_records(self, rec):
        pass

    def _bucket(self, words, B):
        a = np.array(words, dtype=np.uint64)
        return self.O.lib().ora_bucket(a.ctypes.data_as(self.O.u64p), len(words), B)

    def records_flags(self, rec):
        return rec["flags"]

    def _group(self, recs, W, B, G, flags=0):
        import torch
        owners = [self._bucket(r, B) * G // B for r in recs]
        order = sorted(range(len(recs)), key=lambda i: owners[i])     # stable
        flat = [w for i in order for w in recs[i]]
        counts = [owners.count(g) for g in range(G)]
        t = torch.from_numpy(np.array(flat, dtype=np.uint64).view(np.int64).copy()) if flat else torch.empty(0, dtype=torch.int64)
        return {"flags": flags}, t, counts, W

    def extract_partition(self, reads, K, B, G):
        recs = []
        for r in reads:
            for p in range(len(r) - K + 1):
                x = r[p:p + K]
                recs.append(_pack(min(x, _rc(x))))
        return self._group(recs, (K + 31) // 32, B, G, flags=1)   # fwd+RC canonical counting: palindromes count twice

    def derive_partition(self, kp, B, G):
        K = kp.K
        recs = []
        for s in kp.strings:
            for x in (s[:K - 1], s[1:]):
                recs.append(_pack(min(x, _rc(x))))
        return self._group(recs, (K - 1 + 31) // 32, B, G)

    def count(self, recv, n, K, B, want_counts, flags):
        double_palindromes = bool(flags & 1)
        W = (K + 31) // 32
        a = recv.numpy().view(np.uint64).reshape(n, W)
        keyed = {}
        for row in a:
            key = tuple(int(v) for v in row)
            keyed[key] = keyed.get(key, 0) + 1
        items = sorted(keyed.items(), key=lambda kv: (self._bucket(list(kv[0]), B), kv[0]))

        def unpack(words):
            return "".join("ACGT"[(words[i // 32] >> (2 * (i % 32))) & 3] for i in range(K))

        class Shard:
            pass
        s = Shard()
        s.K = K
        s.records = np.array([list(kv[0]) for kv in items], dtype=np.uint64).reshape(-1, W)
        s.strings = [unpack(kv[0]) for kv in items]
        s.counts = np.array([kv[1] * (2 if double_palindromes and unpack(kv[0]) == _rc(unpack(kv[0])) else 1) for kv in items],
                            dtype=np.uint32)
        s.buckets = [self._bucket(list(kv[0]), B) for kv in items]
        return s


def _worker(rank, world, port, k, B, reads, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist
    import oracle_lib as O
    from spades_for_blackbird_b200.host import distributed as D
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    try:
        comm = D.TorchComm()
        be = CpuShardBackend(O)
        n = len(reads)
        mine = reads[n * rank // world: n * (rank + 1) // world]
        kp = D.count_shard(be, comm, lambda: be.extract_partition(mine, k + 1, B, world), k + 1, B, True)
        km = D.count_shard(be, comm, lambda: be.derive_partition(kp, B, world), k, B, False)
        # ownership: only my bucket range
        lo, hi = rank * B // world, (rank + 1) * B // world
        assert all(lo <= b < hi for b in kp.buckets) and all(lo <= b < hi for b in km.buckets)
        np.savez(os.path.join(out_dir, "rank%d.npz" % rank), kp=kp.records, kc=kp.counts, km=km.records)
        # launcher plumbing: every rank's slice of a packed read set, re-based; the union is the whole set
        words, word_off, lens = O.pack_reads(reads)
        w, o, ln = D.slice_reads(words, word_off, lens, rank, world)
        assert int(o[0]) == 0 and len(o) == len(ln) + 1 and int(o[-1]) == len(w)
        import torch
        tot = torch.tensor([len(ln), len(w)], dtype=torch.int64)
        comm.all_reduce_sum_(tot)
        assert tot.tolist() == [len(lens), len(words)]
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("k,B", [(21, 20), (33, 10)])
def test_sharded_counting_over_gloo(tmp_path, k, B):
    import torch.multiprocessing as mp
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib as O
    from spades_for_blackbird_b200.host import synth
    genome = synth.random_genome(600, k)
    reads = synth.codes_to_strings(synth.sample_pairs(genome, 40, 70, 160, 0.01, k + 1))
    reads.append("ACGTACGTACG" + "CGTACGTACGT" if k == 21 else reads[0])   # a self-reverse-complement 22-mer: counted twice
    want = O.gbuilder(reads, k, B)
    port = _free_port()
    mp.spawn(_worker, args=(2, port, k, B, reads, str(tmp_path)), nprocs=2, join=True)
    parts = [np.load(os.path.join(str(tmp_path), "rank%d.npz" % r)) for r in range(2)]
    W1, W0 = (k + 1 + 31) // 32, (k + 31) // 32
    assert np.array_equal(np.concatenate([p["kp"].reshape(-1, W1) for p in parts]), want["kpomers"].data)
    assert np.array_equal(np.concatenate([p["kc"] for p in parts]), want["kpomers"].counts)
    assert np.array_equal(np.concatenate([p["km"].reshape(-1, W0) for p in parts]), want["kmers"].data)
