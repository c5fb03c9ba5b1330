"""The CUDA path (through the C ABI, spades_for_blackbird_b200/host/binding.py) against fixtures produced by the unmodified
reference (tests/golden/*.npz) and against the CPU oracle on seeded random inputs.  Everything is integer / byte /
index work: the bar is bit-exact equality, including the reference's output ORDER of k-mers and unitigs."""
import hashlib

import numpy as np
import pytest

import oracle_lib as O
from spades_for_blackbird_b200.host import binding as B
from spades_for_blackbird_b200.host import synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    c = B.Context(0)
    yield c
    c.close()


def build_index(ctx, reads, k, nb):
    words, word_off, lens = O.pack_reads(reads)
    streams = B.ReadStreams(ctx, words, word_off, lens)
    index = B.DeBruijnExtensionIndex(ctx, k)
    kpomers = B.DeBruijnExtensionIndexBuilder().BuildExtensionIndexFromStream(index, streams, num_buckets=nb)
    return streams, index, kpomers


def test_golden_whole_path(ctx, golden):
    g = golden
    streams, index, kpomers = build_index(ctx, g["reads"], g["k"], g["buckets"])
    assert np.array_equal(kpomers.final_kmers().reshape(-1), g["kpomers"])
    assert np.array_equal(np.diff(kpomers.bucket_starts), g["kp_bucket_sizes"])
    assert np.array_equal(kpomers.counts(), g["coverage"])            # coverage_hash_map_builder.hpp:15-38
    for b in (0, g["buckets"] // 2, g["buckets"] - 1):                # kmers<i> files
        lo, hi = int(kpomers.bucket_starts[b]), int(kpomers.bucket_starts[b + 1])
        assert np.array_equal(kpomers.bucket(b).reshape(-1), g["kpomers"].reshape(-1, kpomers.words)[lo:hi].reshape(-1))
    assert np.array_equal(index.kmers.final_kmers().reshape(-1), g["kmers"])
    assert np.array_equal(index.index.seq_idx(index.kmers.final_kmers()), g["idx"])
    assert np.array_equal(index.idx(), g["idx"])
    km_recs = index.kmers.final_kmers()
    for i in range(0, len(km_recs), max(len(km_recs) // 50, 1)):            # KMerIndex::seq_idx(const Seq &): one key at a time, on the host
        assert index.index.seq_idx_one(km_recs[i]) == int(g["idx"][i])
    if (np.diff(index.kmers.bucket_starts) > 0).all():
        assert np.array_equal(index.index.serialize(), g["index_bin"])    # KMerIndex::serialize, byte for byte
    if g["tip_bound"] >= 0:
        assert B.EarlyTipClipperProcessor(index, g["tip_bound"]).ClipTips() == int(g["clipped"])
    assert np.array_equal(index.data(), g["masks_idx"])
    ex = B.UnbranchingPathExtractor(index, g["k"])
    assert ex.ExtractUnbranchingPathsAndLoops() == g["unitigs"]
    # the coverage map (a KMerIndex over the (k+1)-mers + counts in ITS index order) and the per-edge coverage
    cov = B.CoverageHashMap(ctx, kpomers)
    kp_idx = cov.index.seq_idx(kpomers.final_kmers())
    assert np.array_equal(np.sort(kp_idx), np.arange(kpomers.total_kmers(), dtype=np.uint64))       # minimal perfect over the (k+1)-mers
    assert np.array_equal(cov.data()[kp_idx.astype(np.int64)], g["coverage"])                         # PerfectHashMap<RtSeq, uint32_t>::data_
    kc, flank = cov.edge_coverage(ex.h, 50)
    import sys, os
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
    import sb200_gfa
    idx_of = {tuple(int(v) for v in rec): int(i) for rec, i in zip(index.kmers.final_kmers(), g["idx"])}
    lines = sb200_gfa.gfa_lines(g["unitigs"], g["k"], lambda s: idx_of[tuple(int(v) for v in O.pack_reads([s])[0])], kc=[int(x) for x in kc])
    assert sorted(lines) == g["gfa_cov"]                                                             # spades-gbuilder -c: DP:f / KC:i
    fl = g["flanking"][np.argsort(g["flanking"][:, 0])]
    assert np.array_equal(flank.astype(np.int64), fl[:, 1:])                                          # FlankingCoverage raw values
    cov.free(); ex.free()


def test_golden_kmercount(ctx, golden):
    g = golden
    words, word_off, lens = O.pack_reads(g["reads"])
    streams = B.ReadStreams(ctx, words, word_off, lens)
    kc = B.KMerDiskCounter(ctx, streams, g["k"], canonical_only=False, add_rc=True).CountAll(16)
    assert np.array_equal(kc.final_kmers().reshape(-1), g["kc_final"])
    assert np.array_equal(np.diff(kc.bucket_starts), g["kc_bucket_sizes"])
    if g["name"] == "ecoli1k_k21":
        assert hashlib.md5(kc.final_kmers().tobytes()).hexdigest() == "d47405a3a23aed21661194c705a67970"


def test_golden_one_shot(ctx, golden):
    g = golden
    words, word_off, lens = O.pack_reads(g["reads"])
    gr = B.construct(ctx, words, word_off, lens, g["k"], g["buckets"], tip_clip=g["tip_bound"] >= 0,
                     tip_length_bound=max(g["tip_bound"], 0), fetch_kmers=True)
    v = gr.view
    assert v.n_kmers == len(g["idx"]) and v.clipped == int(g["clipped"])
    assert np.array_equal(gr.masks(), g["masks_idx"])
    assert gr.unitigs() == g["unitigs"]
    assert np.array_equal(np.ctypeslib.as_array(v.kpomers, shape=(len(g["kpomers"]),)), g["kpomers"])
    assert np.array_equal(np.ctypeslib.as_array(v.kpomer_counts, shape=(v.n_kpomers,)), g["coverage"])
    assert np.array_equal(np.ctypeslib.as_array(v.kmers, shape=(len(g["kmers"]),)), g["kmers"])
    if (np.diff(np.ctypeslib.as_array(v.kmer_bucket_starts, shape=(g["buckets"] + 1,))) > 0).all():
        assert np.array_equal(gr.index_bytes(), g["index_bin"])


@pytest.mark.parametrize("k,read_len,seed", [(21, 100, 1), (31, 100, 2), (33, 120, 3), (55, 150, 4), (63, 150, 5),
                                              (77, 150, 6), (95, 200, 7), (97, 250, 8), (127, 250, 9)])
def test_random_reads_vs_oracle(ctx, k, read_len, seed):
    genome = synth.random_genome(6000, seed)
    codes = synth.sample_pairs(genome, 500, read_len, 2 * read_len + 50, 0.006, seed + 100)
    reads = synth.codes_to_strings(codes)
    nb = 10 * (1 + seed % 4)
    want = O.gbuilder(reads, k, nb, tip_bound=read_len - k if seed % 2 else None)
    streams, index, kpomers = build_index(ctx, reads, k, nb)
    assert np.array_equal(kpomers.final_kmers(), want["kpomers"].data)
    assert np.array_equal(kpomers.counts(), want["kpomers"].counts)
    assert np.array_equal(index.kmers.final_kmers(), want["kmers"].data)
    assert np.array_equal(index.idx(), want["idx"])
    assert np.array_equal(index.index.serialize(), want["index_bin"])
    if seed % 2:
        assert B.EarlyTipClipperProcessor(index, read_len - k).ClipTips() == want["clipped"]
    assert np.array_equal(index.data(), want["masks_idx"])
    assert B.UnbranchingPathExtractor(index, k).ExtractUnbranchingPathsAndLoops() == want["unitigs"]


def test_perfect_loops_of_many_sizes(ctx):
    """cycles without junctions, including power-of-two lengths and self-reverse-complement loops (SplitLoop)"""
    def rc(s):
        return s[::-1].translate(str.maketrans("ACGT", "TGCA"))
    k = 21
    reads = []
    for n, seed in [(64, 1), (128, 2), (100, 3), (256, 4), (333, 5), (32, 6)]:
        s = "".join("ACGT"[c] for c in synth.random_genome(n, 50 + seed))
        reads += [(s + s + s)[i:i + 60] for i in range(0, n, 7)]
    for n, seed in [(40, 7), (64, 8)]:
        a = "".join("ACGT"[c] for c in synth.random_genome(n, 70 + seed))
        s = a + rc(a)
        reads += [(s + s + s)[i:i + 70] for i in range(0, 2 * n, 5)]
    want = O.gbuilder(reads, k, 10)
    assert want["n_loops"] >= 8
    streams, index, kpomers = build_index(ctx, reads, k, 10)
    ex = B.UnbranchingPathExtractor(index, k)
    assert ex.ExtractUnbranchingPathsAndLoops() == want["unitigs"]
    assert ex.n_loops == want["n_loops"]
    assert ex.ExtractUnbranchingPaths() == O.gbuilder(reads, k, 10, with_loops=False)["unitigs"]


def test_edge_cases(ctx):
    # reads shorter than K, reads with Ns, empty input: the reference FATALs with "No kmers were extracted"
    words, word_off, lens = O.pack_reads(["ACGT", "NNNN", ""])
    streams = B.ReadStreams(ctx, words, word_off, lens)
    with pytest.raises(B.Sb200Error, match="No kmers were extracted"):
        B.KMerDiskCounter(ctx, streams, 22).Count(10)
    # exactly one window
    r = "ACGTTGCAAGGCTTAACCGGTA"
    words, word_off, lens = O.pack_reads([r])
    streams = B.ReadStreams(ctx, words, word_off, lens)
    kp = B.KMerDiskCounter(ctx, streams, 22).Count(10)
    want = O.count_reads(words, word_off, lens, 22, True, True, 10)
    assert np.array_equal(kp.final_kmers(), want.data) and np.array_equal(kp.counts(), want.counts)
    # a self-reverse-complement (k+1)-mer is counted twice per occurrence (both streams keep it)
    pal = "ACGTACGTACG" + "CGTACGTACGT"
    assert pal == pal[::-1].translate(str.maketrans("ACGT", "TGCA"))
    words, word_off, lens = O.pack_reads([pal, "G" + pal + "T"])
    streams = B.ReadStreams(ctx, words, word_off, lens)
    kp = B.KMerDiskCounter(ctx, streams, 22).Count(4)
    want = O.count_reads(words, word_off, lens, 22, True, True, 4)
    assert np.array_equal(kp.final_kmers(), want.data) and np.array_equal(kp.counts(), want.counts)
    assert 4 in kp.counts().tolist()
    # forward-only canonical mode (non RC-wrapped streams) and plain forward mode
    g = synth.random_genome(2000, 3)
    reads = synth.codes_to_strings(synth.sample_pairs(g, 100, 80, 200, 0.01, 4))
    words, word_off, lens = O.pack_reads(reads)
    streams = B.ReadStreams(ctx, words, word_off, lens)
    for canon, addrc in [(True, False), (False, False)]:
        for K in (22, 32, 64):
            got = B.KMerDiskCounter(ctx, streams, K, canonical_only=canon, add_rc=addrc).Count(7)
            want = O.count_reads(words, word_off, lens, K, canon, addrc, 7)
            assert np.array_equal(got.final_kmers(), want.data), (canon, addrc, K)
            assert np.array_equal(got.counts(), want.counts)
            assert np.array_equal(got.bucket_starts, want.bucket_starts)
    # single bucket
    got = B.KMerDiskCounter(ctx, streams, 22).Count(1)
    want = O.count_reads(words, word_off, lens, 22, True, True, 1)
    assert np.array_equal(got.final_kmers(), want.data)


def test_size_independent_properties_at_scale(ctx):
    """2 M reads-bases scale (too big for the oracle in seconds): sortedness, uniqueness, conservation of instances,
    MPHF bijectivity, mask symmetry, unitig k-mer cover."""
    k, nb = 55, 80
    words, word_off, lens = synth.isolate_config(genome_len=200_000, coverage=30.0)
    streams = B.ReadStreams(ctx, words, word_off, lens)
    index = B.DeBruijnExtensionIndex(ctx, k)
    kp = B.DeBruijnExtensionIndexBuilder().BuildExtensionIndexFromStream(index, streams, num_buckets=nb)
    rec = kp.final_kmers()
    cnt = kp.counts()
    assert int(cnt.sum()) == int((lens.astype(np.int64) - k).clip(min=0).sum())          # every window counted once
    for b in range(nb):                                                                  # sorted + unique per bucket
        r = rec[int(kp.bucket_starts[b]):int(kp.bucket_starts[b + 1])]
        key = r[:, 0].astype(object) * (1 << 64) + r[:, 1].astype(object) if len(r) < 2000 else None
        assert np.all((r[1:, 0] > r[:-1, 0]) | ((r[1:, 0] == r[:-1, 0]) & (r[1:, 1] > r[:-1, 1])))
    idx = index.idx()
    assert np.array_equal(np.sort(idx), np.arange(index.size(), dtype=np.uint64))        # minimal perfect
    masks = index.data()
    out_deg = np.array([bin(m & 15).count("1") for m in range(256)])[masks]
    in_deg = np.array([bin(m >> 4).count("1") for m in range(256)])[masks]
    # every canonical (k+1)-mer sets exactly two bits, and distinct edges set distinct bits
    assert int(out_deg.sum()) + int(in_deg.sum()) == 2 * kp.total_kmers()
    w, off, ln = B.UnbranchingPathExtractor(index, k).ExtractUnbranchingPathsAndLoops(packed=True)
    # unitigs partition the edges: each canonical (k+1)-mer lies on exactly one kept sequence (a sequence that is its own
    # reverse complement holds both strands of its edges, hence >=)
    edges = int((ln.astype(np.int64) - k).sum())
    assert kp.total_kmers() <= edges <= int(1.001 * kp.total_kmers())


def test_long_chains_and_forced_pointer_jumping(ctx, monkeypatch):
    """Chains longer than the direct-walk limit (error-free reads: one 6 kbp unitig) take the pointer-jumping path;
    SB200_FORCE_JUMP=1 sends ordinary inputs through it as well — both must equal the oracle."""
    k, nb = 31, 10
    genome = synth.random_genome(6000, 77)
    codes = synth.sample_pairs(genome, 600, 120, 300, 0.0, 78)
    reads = synth.codes_to_strings(codes)
    want = O.gbuilder(reads, k, nb)
    assert max(len(u) for u in want["unitigs"]) > 2000
    streams, index, kpomers = build_index(ctx, reads, k, nb)
    assert B.UnbranchingPathExtractor(index, k).ExtractUnbranchingPathsAndLoops() == want["unitigs"]
    monkeypatch.setenv("SB200_FORCE_JUMP", "1")
    ctx2 = B.Context(0)
    try:
        from conftest import load_golden
        for name in ("ecoli1k_k21", "multiword_k77", "tipclip_k33", "loops_k21"):
            g = load_golden(name)
            streams, index, kpomers = build_index(ctx2, g["reads"], g["k"], g["buckets"])
            if g["tip_bound"] >= 0:
                assert B.EarlyTipClipperProcessor(index, g["tip_bound"]).ClipTips() == int(g["clipped"])
            assert B.UnbranchingPathExtractor(index, g["k"]).ExtractUnbranchingPathsAndLoops() == g["unitigs"], name
            index.free(); kpomers.free(); streams.free()
    finally:
        ctx2.close()


@pytest.mark.parametrize("env", [("SB200_LINKS",), ("SB200_NO_WALK_BLOCKS",), ("SB200_WALK_CAPTURE_WORDS",), ("SB200_LINKS", "SB200_WALK_CAPTURE_WORDS")])
def test_direct_walk_variants(monkeypatch, env):
    """The direct walks resolve every step by an MPHF lookup through the walk blocks (default, and what a table shard does).
    Variants with the same unitigs in the same order: SB200_LINKS=1 follows a link table instead (one lookup per vertex up front);
    SB200_NO_WALK_BLOCKS=1 reads bit-vector, rank and mask array separately; a capture buffer of one word per start edge makes the
    emitting pass re-walk longer kept paths (lookup walks and link walks)."""
    for e in env:
        monkeypatch.setenv(e, "1")
    ctx2 = B.Context(0)
    try:
        from conftest import load_golden
        for name in ("ecoli1k_k55", "multiword_k127", "tipclip_k21", "ctest_SplitThread2"):
            g = load_golden(name)
            streams, index, kpomers = build_index(ctx2, g["reads"], g["k"], g["buckets"])
            if g["tip_bound"] >= 0:
                assert B.EarlyTipClipperProcessor(index, g["tip_bound"]).ClipTips() == int(g["clipped"])
            assert B.UnbranchingPathExtractor(index, g["k"]).ExtractUnbranchingPathsAndLoops() == g["unitigs"], name
            index.free(); kpomers.free(); streams.free()
    finally:
        ctx2.close()


@pytest.mark.parametrize("k,nb,cov", [(21, 1, 300), (33, 2, 250), (55, 1, 200)])
def test_clumpy_groups_take_several_rounds(ctx, k, nb, cov):
    """High coverage makes group sizes of the shared-memory sort clumpy (every genomic k-mer brings `cov` copies at once): groups
    larger than the shared-memory capacity are processed in several digit-range rounds.  Same sets, counts, unitigs."""
    genome = synth.random_genome(2500, 300 + k)
    n_pairs = int(2500 * cov / (2 * 100))
    codes = synth.sample_pairs(genome, n_pairs, 100, 250, 0.004, 301 + k)
    reads = synth.codes_to_strings(codes)
    want = O.gbuilder(reads, k, nb)
    streams, index, kpomers = build_index(ctx, reads, k, nb)
    assert np.array_equal(kpomers.final_kmers(), want["kpomers"].data)
    assert np.array_equal(kpomers.counts(), want["kpomers"].counts)
    assert np.array_equal(index.kmers.final_kmers(), want["kmers"].data)
    assert B.UnbranchingPathExtractor(index, k).ExtractUnbranchingPathsAndLoops() == want["unitigs"]


@pytest.mark.parametrize("copies", [12000, 70000])
def test_massively_repeated_kmer(ctx, copies):
    """One read repeated 12 000 / 70 000 times.  The hashing group kernel keeps one slot per DISTINCT record, so the multiplicity
    does not matter to it (the sorting kernel overflowed its shared memory and sent the set to the LSD path); a group of 65535+
    records (70 000 copies) goes to the 64-bit-slot instance of the kernel.  Multiplicities included."""
    genome = synth.random_genome(400, 11)
    base = synth.codes_to_strings(synth.sample_pairs(genome, 40, 100, 250, 0.0, 12))
    reads = base + [base[0]] * copies
    k, nb = 31, 10
    want = O.gbuilder(reads, k, nb)
    streams, index, kpomers = build_index(ctx, reads, k, nb)
    assert np.array_equal(kpomers.final_kmers(), want["kpomers"].data)
    assert np.array_equal(kpomers.counts(), want["kpomers"].counts)
    assert int(kpomers.counts().max()) >= copies
    assert np.array_equal(index.kmers.final_kmers(), want["kmers"].data)
    assert B.UnbranchingPathExtractor(index, k).ExtractUnbranchingPathsAndLoops() == want["unitigs"]


@pytest.mark.parametrize("k,nb,pairs,read_len", [(31, 1, 800, 100), (63, 2, 1280, 150)])
def test_groups_of_mostly_distinct_records_take_tag_rounds(ctx, k, nb, pairs, read_len):
    """Coverage ~1: almost every record of a group is distinct, so groups hold more than the 4096 distinct records one round of the
    hashing kernel keeps; they are redone in rounds over disjoint tag ranges (grouphash.cuh).  Same sets, counts, unitigs."""
    genome = synth.random_genome(300000, 900 + k)   # ~110 K instances per bucket: groups of ~7 K records, ~5 K of them distinct
    codes = synth.sample_pairs(genome, pairs, read_len, 2 * read_len + 50, 0.002, 901 + k)
    reads = synth.codes_to_strings(codes)
    want = O.gbuilder(reads, k, nb)
    streams, index, kpomers = build_index(ctx, reads, k, nb)
    assert np.array_equal(kpomers.final_kmers(), want["kpomers"].data)
    assert np.array_equal(kpomers.counts(), want["kpomers"].counts)
    assert np.array_equal(index.kmers.final_kmers(), want["kmers"].data)
    assert np.array_equal(index.data(), want["masks_idx"])
    assert B.UnbranchingPathExtractor(index, k).ExtractUnbranchingPathsAndLoops() == want["unitigs"]


def test_multi_k_on_resident_reads(ctx):
    """BASELINE configs[2]: K = 21, 33, 55, 77 back to back on the SAME device-resident packed reads (the reference re-reads its binary
    read files for every K; nothing else is shared between the iterations), each against the oracle."""
    genome = synth.random_genome(20000, 77)
    codes = synth.sample_pairs(genome, 2000, 150, 350, 0.005, 78)
    reads = synth.codes_to_strings(codes)
    words, word_off, lens = O.pack_reads(reads)
    streams = B.ReadStreams(ctx, words, word_off, lens)
    for k in (21, 33, 55, 77):
        want = O.gbuilder(reads, k, 80)
        index = B.DeBruijnExtensionIndex(ctx, k)
        kp = B.DeBruijnExtensionIndexBuilder().BuildExtensionIndexFromStream(index, streams, num_buckets=80)
        assert np.array_equal(kp.final_kmers(), want["kpomers"].data)
        assert np.array_equal(kp.counts(), want["kpomers"].counts)
        assert np.array_equal(index.kmers.final_kmers(), want["kmers"].data)
        assert np.array_equal(index.data(), want["masks_idx"])
        assert B.UnbranchingPathExtractor(index, k).ExtractUnbranchingPathsAndLoops() == want["unitigs"]
        index.free(); kp.free()


def test_low_complexity_reads_overflow_a_segment(ctx):
    """Reads that are 90 % A: thousands of distinct k-mers share their leading bits, one shared-memory segment holds more
    distinct values than a warp of the sorting kernel keeps in registers (fail flag -> LSD path); the hashing kernel bins them by the
    key bits below the shared prefix and takes tag rounds when a group holds too many."""
    rng = np.random.default_rng(5)
    reads = ["".join("ACGT"[c] for c in np.where(rng.random(70) < 0.9, 0, rng.integers(0, 4, 70))) for _ in range(4000)]
    k, nb = 21, 10
    want = O.gbuilder(reads, k, nb)
    streams, index, kpomers = build_index(ctx, reads, k, nb)
    assert np.array_equal(kpomers.final_kmers(), want["kpomers"].data)
    assert np.array_equal(kpomers.counts(), want["kpomers"].counts)
    assert np.array_equal(index.kmers.final_kmers(), want["kmers"].data)
    assert B.UnbranchingPathExtractor(index, k).ExtractUnbranchingPathsAndLoops() == want["unitigs"]


@pytest.mark.parametrize("env", ["SB200_ATOMIC_PARTITION", "SB200_COUNTING_PASSES"])
def test_alternative_grouping_paths(monkeypatch, env):
    """The default grouping is the staged producer-fused partition (staged_partition.cuh).  SB200_ATOMIC_PARTITION=1: reads / (k+1)-mers
    go straight into their groups through L2 atomics (partition.cuh); SB200_COUNTING_PASSES=1: round 1's extract / derive + two
    histogram + scatter passes — the same tables, masks and unitigs for one- to four-word records, the kmercount modes and the tip clipper."""
    monkeypatch.setenv(env, "1")
    ctx2 = B.Context(0)
    try:
        from conftest import load_golden
        for name in ("ecoli1k_k21", "ecoli1k_k55", "multiword_k77", "multiword_k127", "tipclip_k33", "loops_k21"):
            g = load_golden(name)
            streams, index, kpomers = build_index(ctx2, g["reads"], g["k"], g["buckets"])
            assert np.array_equal(kpomers.final_kmers().reshape(-1), g["kpomers"]), name
            assert np.array_equal(kpomers.counts(), g["coverage"]), name
            assert np.array_equal(index.kmers.final_kmers().reshape(-1), g["kmers"]), name
            if g["tip_bound"] >= 0:
                assert B.EarlyTipClipperProcessor(index, g["tip_bound"]).ClipTips() == int(g["clipped"])
            assert np.array_equal(index.data(), g["masks_idx"]), name
            assert B.UnbranchingPathExtractor(index, g["k"]).ExtractUnbranchingPathsAndLoops() == g["unitigs"], name
            kc = B.KMerDiskCounter(ctx2, streams, g["k"], canonical_only=False, add_rc=True).CountAll(16)
            assert np.array_equal(kc.final_kmers().reshape(-1), g["kc_final"]), name
            index.free(); kpomers.free(); streams.free(); kc.free()
    finally:
        ctx2.close()


def test_baseline_config2_full_size(monkeypatch):
    """BASELINE configs[1] at full size (4.6 Mbp genome, 2x150 at 100x, k = 55, 80 buckets: 291 M (k+1)-mer instances) — far
    beyond what the oracle does in seconds, so: size-independent properties of the tables, and the two independent
    implementations of the path must agree (hashing group kernel + masks OR-ed inside the sort + link-table walks  vs.  sorting
    group kernel + masks by MPHF lookups + lookup walks)."""
    import hashlib
    k, nb = 55, 80
    words, word_off, lens = synth.isolate_config()
    instances = int((lens.astype(np.int64) - k).clip(min=0).sum())

    def run(ctx):
        streams = B.ReadStreams(ctx, words, word_off, lens)
        index = B.DeBruijnExtensionIndex(ctx, k)
        kp = B.DeBruijnExtensionIndexBuilder().BuildExtensionIndexFromStream(index, streams, num_buckets=nb)
        w, off, ln = B.UnbranchingPathExtractor(index, k).ExtractUnbranchingPathsAndLoops(packed=True)
        return streams, index, kp, (w, off, ln)

    ctx1 = B.Context(0)
    try:
        streams, index, kp, (w, off, ln) = run(ctx1)
        assert kp.instances == instances
        cnt = kp.counts()
        assert int(cnt.astype(np.int64).sum()) == instances                                   # every window counted exactly once
        for table in (kp, index.kmers):
            rec = table.final_kmers()
            st = table.bucket_starts
            assert st[0] == 0 and st[-1] == len(rec) and np.all(np.diff(st.astype(np.int64)) > 0)
            inc = (rec[1:, 0] > rec[:-1, 0]) | ((rec[1:, 0] == rec[:-1, 0]) & (rec[1:, 1] > rec[:-1, 1]))
            inc[st[1:-1].astype(np.int64) - 1] = True                                         # order restarts at a bucket boundary
            assert inc.all()                                                                  # strictly increasing = sorted + unique
            del rec, inc
        idx = index.idx()
        assert len(idx) == index.size() and np.array_equal(np.sort(idx), np.arange(index.size(), dtype=idx.dtype))   # minimal perfect
        masks = index.data()
        pop = np.array([bin(m).count("1") for m in range(256)], dtype=np.int64)
        assert int(pop[masks].sum()) == 2 * kp.total_kmers()                                  # every (k+1)-mer sets exactly two bits
        edges = int((ln.astype(np.int64) - k).sum())
        assert kp.total_kmers() <= edges <= int(1.001 * kp.total_kmers())                     # unitigs partition the edges
        digest = [hashlib.md5(np.ascontiguousarray(a).tobytes()).hexdigest() for a in (masks, w, off, ln)]
        n_unitigs = len(ln)
        index.free(); kp.free(); streams.free()
    finally:
        ctx1.close()
    monkeypatch.setenv("SB200_NO_MASK_PAYLOAD", "1")
    monkeypatch.setenv("SB200_LINKS", "1")               # the link-table walks instead of lookups through the walk blocks
    monkeypatch.setenv("SB200_GROUP_KERNEL", "chunk")   # and the sorting group kernel instead of the hashing one
    monkeypatch.setenv("SB200_NO_PLACE", "1")            # and k-mer indices by MPHF lookups instead of the build's placement record
    monkeypatch.setenv("SB200_COUNTING_PASSES", "1")     # and extract / derive + counting passes instead of the staged producer-fused partition
    monkeypatch.setenv("SB200_MPHF_STATE_PER_KEY", "1")  # and a state record per key out of MPHF level 0 instead of level 1 re-reading the keys
    ctx2 = B.Context(0)
    try:
        streams, index, kp, (w, off, ln) = run(ctx2)
        assert len(ln) == n_unitigs
        assert [hashlib.md5(np.ascontiguousarray(a).tobytes()).hexdigest() for a in (index.data(), w, off, ln)] == digest
        index.free(); kp.free(); streams.free()
    finally:
        ctx2.close()
