#!/usr/bin/env python
"""Small driver for ncu: one count + derive pass over a reduced isolate workload (defaults: 1.5 Mbp genome, 100x)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from spades_for_blackbird_b200.host import binding as B, synth
glen = int(sys.argv[1]) if len(sys.argv) > 1 else 1_500_000
stages = sys.argv[2] if len(sys.argv) > 2 else "count"
k = 55
words, word_off, lens = synth.isolate_config(genome_len=glen, coverage=100.0)
ctx = B.Context(0)
streams = B.ReadStreams(ctx, words, word_off, lens)
for it in range(2):
    kp = B.KMerDiskCounter(ctx, streams, k + 1, True, True).Count(80)
    index = B.DeBruijnExtensionIndex(ctx, k)
    index.kmers = B.KMerDiskCounter(ctx, kp, k).Count(80)
    if stages == "all":
        index.index = B.KMerIndex(ctx, index.kmers)
        h = B.vp()
        ctx.check(ctx.lib.sb200_ext_build(ctx.h, kp.h, index.kmers.h, index.index.h, B.C.byref(h)))
        index.h = h
        B.UnbranchingPathExtractor(index, k).ExtractUnbranchingPathsAndLoops(packed=True)
    print("iter", it, kp.total_kmers(), index.kmers.total_kmers(), flush=True)
    index.free(); kp.free()
print("done")
