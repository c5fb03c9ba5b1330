#!/usr/bin/env python
"""Stage-by-stage parity report of the CUDA path against the golden fixtures (debug aid; run on the GPU box)."""
import os
import sys
import time
import traceback

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import golden_names, load_golden  # noqa: E402
import oracle_lib as O  # noqa: E402
from spades_for_blackbird_b200.host import binding as B  # noqa: E402


def check(name, ok, extra=""):
    print("   %-28s %s %s" % (name, "OK " if ok else "FAIL", extra), flush=True)
    return ok


def first_diff(a, b):
    a = np.asarray(a).reshape(-1); b = np.asarray(b).reshape(-1)
    if a.shape != b.shape:
        return "shape %s vs %s" % (a.shape, b.shape)
    d = np.nonzero(a != b)[0]
    return "" if len(d) == 0 else "first diff at %d: %s vs %s (%d diffs)" % (d[0], a[d[0]], b[d[0]], len(d))


def run_case(ctx, g):
    ok = True
    words, word_off, lens = O.pack_reads(g["reads"])
    k, nb = g["k"], g["buckets"]
    reads = B.ReadStreams(ctx, words, word_off, lens)
    # kmercount
    kc = B.KMerDiskCounter(ctx, reads, k, canonical_only=False, add_rc=True).Count(16)
    ok &= check("kmercount final_kmers", np.array_equal(kc.final_kmers().reshape(-1), g["kc_final"]), first_diff(kc.final_kmers(), g["kc_final"]))
    ok &= check("kmercount bucket sizes", np.array_equal(np.diff(kc.bucket_starts), g["kc_bucket_sizes"]))
    kc.free()
    index = B.DeBruijnExtensionIndex(ctx, k)
    kp = B.KMerDiskCounter(ctx, reads, k + 1, True, True).Count(nb)
    ok &= check("kpomers", np.array_equal(kp.final_kmers().reshape(-1), g["kpomers"]), first_diff(kp.final_kmers(), g["kpomers"]))
    ok &= check("kpomer bucket sizes", np.array_equal(np.diff(kp.bucket_starts), g["kp_bucket_sizes"]))
    ok &= check("coverage", np.array_equal(kp.counts(), g["coverage"]), first_diff(kp.counts(), g["coverage"]))
    B.DeBruijnExtensionIndexBuilder().BuildExtensionIndexFromKPOMers(index, kp)
    ok &= check("kmers", np.array_equal(index.kmers.final_kmers().reshape(-1), g["kmers"]), first_diff(index.kmers.final_kmers(), g["kmers"]))
    idx = index.index.seq_idx(index.kmers.final_kmers())
    ok &= check("mphf idx", np.array_equal(idx, g["idx"]), first_diff(idx, g["idx"]))
    ok &= check("ext idx", np.array_equal(index.idx(), g["idx"]))
    if (np.diff(index.kmers.bucket_starts) > 0).all():
        ser = index.index.serialize()
        ok &= check("KMerIndex::serialize", np.array_equal(ser, g["index_bin"]), first_diff(ser, g["index_bin"]))
    clipped = 0
    if g["tip_bound"] >= 0:
        clipped = B.EarlyTipClipperProcessor(index, g["tip_bound"]).ClipTips()
        ok &= check("tipclip removed", clipped == int(g["clipped"]), "%d vs %d" % (clipped, int(g["clipped"])))
    ok &= check("masks", np.array_equal(index.data(), g["masks_idx"]), first_diff(index.data(), g["masks_idx"]))
    ex = B.UnbranchingPathExtractor(index, k)
    u = ex.ExtractUnbranchingPathsAndLoops()
    same = u == g["unitigs"]
    extra = ""
    if not same:
        extra = "n=%d vs %d; set-equal=%s" % (len(u), len(g["unitigs"]), sorted(u) == sorted(g["unitigs"]))
    ok &= check("unitigs (ordered)", same, extra + " loops=%d" % ex.n_loops)
    # one-shot API
    gr = B.construct(ctx, words, word_off, lens, k, nb, tip_clip=g["tip_bound"] >= 0, tip_length_bound=max(g["tip_bound"], 0))
    ok &= check("construct(): unitigs", gr.unitigs() == g["unitigs"])
    ok &= check("construct(): masks", np.array_equal(gr.masks(), g["masks_idx"]))
    gr.free()
    index.free(); kp.free(); reads.free()
    return ok


def main():
    names = sys.argv[1:] or golden_names()
    ctx = B.Context(0)
    bad = []
    for name in names:
        g = load_golden(name)
        print("== %s (k=%d, B=%d, %d reads)" % (name, g["k"], g["buckets"], len(g["reads"])), flush=True)
        t0 = time.time()
        try:
            if not run_case(ctx, g):
                bad.append(name)
        except Exception:
            traceback.print_exc()
            bad.append(name)
        print("   %.2fs" % (time.time() - t0), flush=True)
    print("FAILED: %s" % bad if bad else "ALL OK")
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
