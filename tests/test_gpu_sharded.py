"""The hash-sharded path with G virtual ranks (threads, one sb200 context each) on ONE GPU: every kernel and every piece of
the orchestration of spades_for_blackbird_b200/host/distributed.py except the NCCL transport itself, which is replaced by
LocalComm.  Shards concatenated in rank order must equal the reference's single-process result bit for bit."""
import threading

import numpy as np
import pytest

import oracle_lib as O
from spades_for_blackbird_b200.host import binding as B
from spades_for_blackbird_b200.host import distributed as D
from spades_for_blackbird_b200.host import synth

pytestmark = pytest.mark.gpu


def run_virtual_ranks(G, reads, k, nb):
    import torch
    dev = torch.device("cuda", 0)
    words, word_off, lens = O.pack_reads(reads)
    n = len(lens)
    shared = D.LocalComm.Shared(G)
    results, errors = [None] * G, [None] * G

    def work(r):
        try:
            ctx = B.Context(0)
            lo, hi = n * r // G, n * (r + 1) // G
            w0, w1 = int(word_off[lo]), int(word_off[hi])
            streams = B.ReadStreams(ctx, words[w0:w1], word_off[lo:hi + 1] - word_off[lo], lens[lo:hi])
            backend = D.GpuShardBackend(ctx, dev)
            res = D.construct_sharded(backend, D.LocalComm(shared, r), streams, k, nb, gather_to=0)
            out = dict(kpomers=res.kpomers.final_kmers(), counts=res.kpomers.counts(), kp_starts=res.kpomers.bucket_starts,
                       kmers=res.kmers.final_kmers(), km_starts=res.kmers.bucket_starts,
                       masks=backend.ext_masks(res.ext).cpu().numpy(), stats=res.stats)
            if r == 0:
                out["unitigs"] = D.unpack_gathered(res.gathered)
            results[r] = out
        except Exception as e:   # noqa: BLE001
            errors[r] = e
            try:
                shared.barrier.abort()
            except Exception:
                pass

    threads = [threading.Thread(target=work, args=(r,)) for r in range(G)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    for e in errors:
        if e is not None and not isinstance(e, threading.BrokenBarrierError):
            raise e
    for e in errors:
        if e is not None:
            raise e
    return results


@pytest.mark.parametrize("G", [2, 4])
@pytest.mark.parametrize("name", ["ecoli1k_k21", "multiword_k77", "multiword_k127"])
def test_sharded_equals_reference(G, name):
    from conftest import load_golden
    g = load_golden(name)
    nb = g["buckets"] if g["buckets"] % G == 0 else g["buckets"] * G
    if nb != g["buckets"]:
        pytest.skip("bucket count of the fixture is not a multiple of G")
    res = run_virtual_ranks(G, g["reads"], g["k"], nb)
    kp = np.concatenate([r["kpomers"] for r in res]).reshape(-1)
    assert np.array_equal(kp, g["kpomers"])
    assert np.array_equal(np.concatenate([r["counts"] for r in res]), g["coverage"])
    assert np.array_equal(np.concatenate([r["kmers"] for r in res]).reshape(-1), g["kmers"])
    for r in res:
        assert np.array_equal(r["masks"][:len(g["masks_idx"])], g["masks_idx"])    # every rank holds all masks
    assert res[0]["unitigs"] == g["unitigs"]
    # ownership: rank r holds exactly the buckets [r*B/G, (r+1)*B/G)
    for r, out in enumerate(res):
        sizes = np.diff(out["km_starts"])
        owned = np.zeros(nb, dtype=bool)
        owned[r * nb // G:(r + 1) * nb // G] = True
        assert (sizes[~owned] == 0).all()


def test_sharded_random_vs_oracle():
    k, nb, G = 55, 20, 2
    genome = synth.random_genome(8000, 123)
    reads = synth.codes_to_strings(synth.sample_pairs(genome, 700, 150, 350, 0.006, 124))
    want = O.gbuilder(reads, k, nb)
    res = run_virtual_ranks(G, reads, k, nb)
    assert np.array_equal(np.concatenate([r["kpomers"] for r in res]), want["kpomers"].data)
    assert np.array_equal(np.concatenate([r["kmers"] for r in res]), want["kmers"].data)
    assert np.array_equal(res[0]["masks"][:len(want["masks_idx"])], want["masks_idx"])
    assert res[0]["unitigs"] == want["unitigs"]


@pytest.mark.parametrize("env", ["SB200_NO_FUSED_PARTITION", "SB200_NO_PLACE"])
def test_sharded_alternative_paths_agree(monkeypatch, env):
    """The sharded path with one of its shortcuts switched off — extraction straight into the owner groups (then: extract, partition
    pass), one-read ranks through the rebuilt prefix popcounts (then: rank samples) — gives the same shards, masks and unitigs."""
    k, nb, G = 33, 40, 4
    genome = synth.random_genome(12000, 321)
    reads = synth.codes_to_strings(synth.sample_pairs(genome, 900, 120, 300, 0.005, 322))
    want = O.gbuilder(reads, k, nb)
    monkeypatch.setenv(env, "1")
    res = run_virtual_ranks(G, reads, k, nb)
    assert np.array_equal(np.concatenate([r["kpomers"] for r in res]), want["kpomers"].data)
    assert np.array_equal(np.concatenate([r["counts"] for r in res]), want["kpomers"].counts)
    assert np.array_equal(np.concatenate([r["kmers"] for r in res]), want["kmers"].data)
    assert np.array_equal(res[0]["masks"][:len(want["masks_idx"])], want["masks_idx"])
    assert res[0]["unitigs"] == want["unitigs"]
