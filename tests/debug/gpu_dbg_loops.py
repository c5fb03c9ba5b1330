import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import load_golden
import oracle_lib as O
from spades_for_blackbird_b200.host import binding as B
def rc(s): return s[::-1].translate(str.maketrans("ACGT", "TGCA"))
g = load_golden("loops_k21")
ctx = B.Context(0)
words, word_off, lens = O.pack_reads(g["reads"])
reads = B.ReadStreams(ctx, words, word_off, lens)
index = B.DeBruijnExtensionIndex(ctx, g["k"])
kp = B.DeBruijnExtensionIndexBuilder().BuildExtensionIndexFromStream(index, reads, num_buckets=g["buckets"])
ex = B.UnbranchingPathExtractor(index, g["k"])
u = ex.ExtractUnbranchingPathsAndLoops()
w = g["unitigs"]
print("n", len(u), len(w), "loops", ex.n_loops)
for i, (a, b) in enumerate(zip(u, w)):
    if a != b:
        print(i, len(a), len(b), "rc-equal" if a == rc(b) else "", "sameset" if sorted(a) == sorted(b) else "")
        print("  got ", a[:80], "...", a[-30:])
        print("  want", b[:80], "...", b[-30:])
        if len(a) == len(b):
            # rotation check
            d = b + b
            print("  rotation:", a[:40] in d, rc(a)[:40] in d)
