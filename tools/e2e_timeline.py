#!/usr/bin/env python
"""Stage timeline of sb200_construct (SB200_TIMELINE=1) on the bench workload: where the end-to-end time goes."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["SB200_TIMELINE"] = "1"
import argparse
import numpy as np, torch
import bench
from spades_for_blackbird_b200.host import binding as B
a = argparse.Namespace(genome_len=4_600_000, coverage=100.0, read_len=150, k=55, buckets=80)
words, word_off, lens, _ = bench.make_reads(a)
pw = torch.from_numpy(words.view(np.int64)).pin_memory(); po = torch.from_numpy(word_off.view(np.int64)).pin_memory(); pl = torch.from_numpy(lens.view(np.int32)).pin_memory()
hw, ho, hl = pw.numpy().view(np.uint64), po.numpy().view(np.uint64), pl.numpy().view(np.uint32)
ctx = B.Context(0)
for fetch in (True, True, True, False, False):
    t0 = time.perf_counter()
    g = B.construct(ctx, hw, ho, hl, a.k, a.buckets, fetch_kmers=fetch)
    t1 = time.perf_counter()
    g.free()
    print("fetch_kmers=%s: construct %.1f ms, free %.1f ms" % (fetch, (t1 - t0) * 1e3, (time.perf_counter() - t1) * 1e3), file=sys.stderr)
