"""The hash-sharded path with G virtual ranks (threads, one sb200 context each, a local communicator) on ONE GPU: every kernel and all
of the C++ orchestration of csrc/shard.cu — everything but the NCCL transport itself, which tests/test_gpu_nccl.py covers on two real
GPUs.  Shards concatenated in rank order must equal the reference's single-process result bit for bit."""
import numpy as np
import pytest

import oracle_lib as O
from spades_for_blackbird_b200.host import binding as B
from spades_for_blackbird_b200.host import distributed as D
from spades_for_blackbird_b200.host import synth

pytestmark = pytest.mark.gpu


def run_sharded(G, reads, k, nb, tip_bound=None, gather_to=0):
    words, word_off, lens = O.pack_reads(reads)

    def work(r, ctx, comm):
        w, o, ln = D.slice_reads(words, word_off, lens, r, G)
        streams = B.ReadStreams(ctx, w, o, ln)
        sh = B.construct_sharded(ctx, comm, streams, k, nb, tip_clip=tip_bound is not None, tip_length_bound=tip_bound or 0, gather_to=gather_to)
        out = dict(kpomers=sh.kpomers.final_kmers(), counts=sh.kpomers.counts(), kp_starts=sh.kpomers.bucket_starts,
                   kmers=sh.kmers.final_kmers(), km_starts=sh.kmers.bucket_starts, masks=sh.masks(), idx=sh.idx(),
                   index_bin=sh.index.serialize(), unitigs=sh.unitigs(), clipped=int(sh.info.clipped), fallback=bool(sh.info.whole_set_fallback),
                   gathered=bool(sh.info.gathered), n_unitigs=int(sh.info.total_unitigs), stage_ms=sh.stage_ms, bytes_sent=int(sh.info.bytes_sent))
        sh.free(); streams.free()
        return out

    return D.run_virtual_ranks(G, work)


def check(res, want, G, nb, gathered=True):
    """want: dict with the reference's (or the oracle's) kpomers / coverage / kmers / idx / masks_idx / index_bin / unitigs"""
    assert np.array_equal(np.concatenate([r["kpomers"] for r in res]).reshape(-1), want["kpomers"])
    assert np.array_equal(np.concatenate([r["counts"] for r in res]), want["coverage"])
    assert np.array_equal(np.concatenate([r["kmers"] for r in res]).reshape(-1), want["kmers"])
    assert np.array_equal(np.concatenate([r["idx"] for r in res]), want["idx"])          # MPHF index of every k-mer, shard by shard
    for r in res:
        assert np.array_equal(r["masks"], want["masks_idx"])                               # every rank holds all masks ...
        if want.get("index_bin") is not None:
            assert np.array_equal(r["index_bin"], want["index_bin"])                       # ... and the whole KMerIndex, byte for byte
    if gathered:
        assert res[0]["unitigs"] == list(want["unitigs"])
    else:
        assert [u for r in res for u in r["unitigs"]] == list(want["unitigs"])           # rank order = the reference's order
    for r, out in enumerate(res):   # ownership: rank r holds exactly the buckets [r*B/G, (r+1)*B/G)
        owned = np.zeros(nb, dtype=bool)
        owned[r * nb // G:(r + 1) * nb // G] = True
        assert (np.diff(out["km_starts"])[~owned] == 0).all() and (np.diff(out["kp_starts"])[~owned] == 0).all()


def golden_or_oracle(name, G):
    """the reference's fixture when its bucket count divides by G, else the oracle (pinned to the same fixtures) on the fixture's reads with
    the next multiple of G buckets — no sharded case is skipped"""
    from conftest import load_golden
    g = load_golden(name)
    tip = int(g["tip_bound"]) if g["tip_bound"] >= 0 else None
    if g["buckets"] % G == 0:
        full = (np.diff(np.concatenate([[0], np.cumsum(g["kp_bucket_sizes"])])) >= 0).all()
        want = dict(kpomers=g["kpomers"], coverage=g["coverage"], kmers=g["kmers"], idx=g["idx"], masks_idx=g["masks_idx"],
                    index_bin=g["index_bin"] if full else None, unitigs=g["unitigs"], clipped=int(g["clipped"]))
        return g["reads"], g["k"], g["buckets"], tip, want
    nb = g["buckets"] * G
    w = O.gbuilder(g["reads"], g["k"], nb, tip_bound=tip)
    want = dict(kpomers=w["kpomers"].data.reshape(-1), coverage=w["kpomers"].counts, kmers=w["kmers"].data.reshape(-1), idx=w["idx"],
                masks_idx=w["masks_idx"], index_bin=w["index_bin"], unitigs=w["unitigs"], clipped=w.get("clipped", 0))
    return g["reads"], g["k"], nb, tip, want


@pytest.mark.parametrize("G", [2, 4])
@pytest.mark.parametrize("name", ["ecoli1k_k21", "ecoli1k_k55", "multiword_k77", "multiword_k127", "loops_k21", "tipclip_k21", "tipclip_k33",
                                  "ctest_SplitThread2"])
def test_sharded_equals_reference(G, name):
    reads, k, nb, tip, want = golden_or_oracle(name, G)
    res = run_sharded(G, reads, k, nb, tip_bound=tip)
    if not (np.diff(np.concatenate([r["km_starts"][r_ * nb // G:(r_ + 1) * nb // G + 1] for r_, r in enumerate(res)])) != 0).all():
        want = dict(want, index_bin=None)   # the reference does not serialise an index with empty buckets (make_golden.py)
    check(res, want, G, nb)
    if tip is not None:
        assert all(r["clipped"] == want["clipped"] for r in res)
    if name == "loops_k21":
        assert all(r["fallback"] for r in res)    # perfect loops: rank 0 extracted the whole set


def test_sharded_random_vs_oracle_slices_left_sharded():
    k, nb, G = 55, 20, 2
    genome = synth.random_genome(8000, 123)
    reads = synth.codes_to_strings(synth.sample_pairs(genome, 700, 150, 350, 0.006, 124))
    w = O.gbuilder(reads, k, nb)
    want = dict(kpomers=w["kpomers"].data.reshape(-1), coverage=w["kpomers"].counts, kmers=w["kmers"].data.reshape(-1), idx=w["idx"],
                masks_idx=w["masks_idx"], index_bin=w["index_bin"], unitigs=w["unitigs"])
    res = run_sharded(G, reads, k, nb, gather_to=-1)
    check(res, want, G, nb, gathered=False)
    assert not any(r["fallback"] or r["gathered"] for r in res)
    assert all(len(r["unitigs"]) > 0 for r in res) and res[0]["bytes_sent"] > 0


def test_sharded_long_chains_go_to_rank0():
    """error-free reads: one 6 kbp unitig, far beyond the direct-walk limit — the whole-set extraction on rank 0 (pointer jumping)"""
    k, nb, G = 31, 20, 4
    genome = synth.random_genome(6000, 77)
    reads = synth.codes_to_strings(synth.sample_pairs(genome, 600, 120, 300, 0.0, 78))
    w = O.gbuilder(reads, k, nb)
    assert max(len(u) for u in w["unitigs"]) > 2000
    want = dict(kpomers=w["kpomers"].data.reshape(-1), coverage=w["kpomers"].counts, kmers=w["kmers"].data.reshape(-1), idx=w["idx"],
                masks_idx=w["masks_idx"], index_bin=w["index_bin"], unitigs=w["unitigs"])
    res = run_sharded(G, reads, k, nb)
    check(res, want, G, nb)
    assert all(r["fallback"] for r in res) and all(len(r["unitigs"]) == 0 for r in res[1:])


def test_sharded_empty_shards():
    """two short reads on four ranks with four buckets: most owners receive nothing, two ranks have no reads at all"""
    k, nb, G = 21, 4, 4
    reads = ["ACGTTGCATGCCGATAGCTAGCTAGGATCCA", "TTGACCGATAGCTAGCTAGGATCCATTGACA"]
    w = O.gbuilder(reads, k, nb)
    want = dict(kpomers=w["kpomers"].data.reshape(-1), coverage=w["kpomers"].counts, kmers=w["kmers"].data.reshape(-1), idx=w["idx"],
                masks_idx=w["masks_idx"], index_bin=None, unitigs=w["unitigs"])
    res = run_sharded(G, reads, k, nb)
    check(res, want, G, nb)


def test_sharded_no_kmers_fails_on_every_rank():
    with pytest.raises(B.Sb200Error, match="No kmers were extracted"):
        run_sharded(2, ["ACGT", "GGCA"], 21, 4)


@pytest.mark.parametrize("env", ["SB200_COUNTING_PASSES", "SB200_COUNTING_PASSES,SB200_NO_FUSED_PARTITION", "SB200_NO_PLACE", "SB200_NO_MASK_PAYLOAD",
                                 "SB200_NO_PEER_STORES"])
def test_sharded_alternative_paths_agree(monkeypatch, env):
    """The sharded path with one of its shortcuts switched off — the staged sender / receiver kernels (then: round 1's extraction straight
    into the owner groups + the owner's counting passes; with NO_FUSED_PARTITION: extract, partition pass), k-mer indices from the build's
    placement record and one-read ranks (then: lookups, rank samples), masks through the k-mer sort + all-gather of slices (then: lookups
    + all-reduce) — gives the same shards, masks and unitigs."""
    k, nb, G = 33, 40, 4
    genome = synth.random_genome(12000, 321)
    reads = synth.codes_to_strings(synth.sample_pairs(genome, 900, 120, 300, 0.005, 322))
    w = O.gbuilder(reads, k, nb)
    want = dict(kpomers=w["kpomers"].data.reshape(-1), coverage=w["kpomers"].counts, kmers=w["kmers"].data.reshape(-1), idx=w["idx"],
                masks_idx=w["masks_idx"], index_bin=w["index_bin"], unitigs=w["unitigs"])
    for e in env.split(","):
        monkeypatch.setenv(e.split("=")[0], e.split("=")[1] if "=" in e else "1")
    res = run_sharded(G, reads, k, nb)
    check(res, want, G, nb)


@pytest.mark.parametrize("ids", [[0, 0], [0, 0, 0, 0]])
def test_multi_context_one_host_thread(ids):
    """sb200_multi (SURVEY 8(b)'s sb200_create(n_gpus, device_ids)): the library splits host reads over its ranks, runs the sharded path
    from its own threads and returns ONE graph in the single-GPU layout.  A repeated device id = virtual ranks on this GPU."""
    from conftest import load_golden
    for name in ("ecoli1k_k55", "loops_k21", "tipclip_k33"):
        g = load_golden(name)
        nb = g["buckets"] if g["buckets"] % len(ids) == 0 else g["buckets"] * len(ids)
        tip = int(g["tip_bound"]) if g["tip_bound"] >= 0 else None
        if nb == g["buckets"]:
            want = dict(kpomers=g["kpomers"], coverage=g["coverage"], kmers=g["kmers"], masks_idx=g["masks_idx"], unitigs=g["unitigs"],
                        clipped=int(g["clipped"]))
        else:
            w = O.gbuilder(g["reads"], g["k"], nb, tip_bound=tip)
            want = dict(kpomers=w["kpomers"].data.reshape(-1), coverage=w["kpomers"].counts, kmers=w["kmers"].data.reshape(-1),
                        masks_idx=w["masks_idx"], unitigs=w["unitigs"], clipped=w.get("clipped", 0))
        words, word_off, lens = O.pack_reads(g["reads"])
        mc = B.MultiContext(ids)
        try:
            gr = mc.construct(words, word_off, lens, g["k"], nb, tip_clip=tip is not None, tip_length_bound=tip or 0, fetch_kmers=True)
            v = gr.view
            assert np.array_equal(np.ctypeslib.as_array(v.kpomers, shape=(len(want["kpomers"]),)), want["kpomers"]), name
            assert np.array_equal(np.ctypeslib.as_array(v.kpomer_counts, shape=(v.n_kpomers,)), want["coverage"]), name
            assert np.array_equal(np.ctypeslib.as_array(v.kmers, shape=(len(want["kmers"]),)), want["kmers"]), name
            assert np.array_equal(gr.masks(), want["masks_idx"]), name
            assert gr.unitigs() == list(want["unitigs"]), name
            assert v.clipped == want["clipped"]
            ks = np.ctypeslib.as_array(v.kmer_bucket_starts, shape=(nb + 1,))
            assert ks[0] == 0 and ks[-1] == v.n_kmers and (np.diff(ks.astype(np.int64)) >= 0).all()
            gr.free()
        finally:
            mc.close()
    with pytest.raises(B.Sb200Error, match="No kmers were extracted"):
        mc = B.MultiContext([0, 0])
        try:
            mc.construct(*O.pack_reads(["ACGT", "GGCA"]), 21, 4)
        finally:
            mc.close()
