"""Debug helper: count the (k+1)-mers and k-mers of a golden fixture with the group kernel under test, report the first differences."""
import sys, os
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import load_golden
import oracle_lib as O
from spades_for_blackbird_b200.host import binding as B

name = sys.argv[1] if len(sys.argv) > 1 else "ecoli1k_k21"
g = load_golden(name)
ctx = B.Context(0)
words, word_off, lens = O.pack_reads(g["reads"])
streams = B.ReadStreams(ctx, words, word_off, lens)
kp = B.KMerDiskCounter(ctx, streams, g["k"] + 1, True, True).Count(g["buckets"])
got = kp.final_kmers().reshape(-1, kp.words)
want = g["kpomers"].reshape(-1, kp.words)
print("kpomers", got.shape, want.shape, "equal", got.shape == want.shape and np.array_equal(got, want))
if got.shape == want.shape:
    bad = np.nonzero((got != want).any(axis=1))[0]
    print("first bad rows", bad[:10])
    print("counts equal", np.array_equal(kp.counts(), g["coverage"]))
else:
    gs = {tuple(r) for r in got.tolist()}; ws = {tuple(r) for r in want.tolist()}
    print("distinct got", len(gs), "missing", len(ws - gs), "extra", len(gs - ws))
km = B.KMerDiskCounter(ctx, kp, g["k"]).Count(g["buckets"])
got = km.final_kmers().reshape(-1, km.words)
want = g["kmers"].reshape(-1, km.words)
print("kmers", got.shape, want.shape, "equal", got.shape == want.shape and np.array_equal(got, want))
if got.shape != want.shape:
    gs = {tuple(r) for r in got.tolist()}; ws = {tuple(r) for r in want.tolist()}
    print("distinct got", len(gs), "missing", len(ws - gs), "extra", len(gs - ws))
    for r in list(gs - ws)[:5]:
        print("extra %x" % r[0])
else:
    bad = np.nonzero((got != want).any(axis=1))[0]
    print("first bad rows", bad[:10])
got = kp.final_kmers().reshape(-1, kp.words); want = g["kpomers"].reshape(-1, kp.words)
gs = {tuple(r) for r in got.tolist()}; ws = {tuple(r) for r in want.tolist()}
print("KP distinct got", len(gs), "missing", len(ws - gs), "extra", len(gs - ws))
bs = kp.bucket_starts
print("bucket sizes equal", np.array_equal(np.diff(bs), g["kp_bucket_sizes"]))
for r in range(min(14, len(got))):
    print(r, " ".join("%016x" % x for x in got[r]), "|", " ".join("%016x" % x for x in want[r]), kp.counts()[r], g["coverage"][r])
