#!/usr/bin/env python
"""Regenerate tests/golden/*.npz from the UNMODIFIED reference (oracle/_ref/ref_driver, built by `make -C oracle ref`
from the sources under /root/reference).  Run in the build container only; the fixtures travel, the reference does not.

    python tests/golden/make_golden.py

Each fixture holds the input reads and every artefact the reference produced for them:
  kpomers / kp_bucket_sizes / coverage    sorted-unique canonical (k+1)-mers per hash bucket, their multiplicities
  kmers / idx / masks_idx / index_bin     final_kmers, MPHF index of each, InOutMask array in index order, KMerIndex::serialize
  unitigs / clipped                       UnbranchingPathExtractor output (reference order) / tip-clipper count
  gfa / fastg                             FastGraphFromSequencesConstructor + gfa::GFAWriter / io::FastgWriter on those unitigs, lines / records sorted
  gfa_cov / flanking                      the same GFA after FillCoverageAndFlankingFromPHM (spades-gbuilder -c), raw flanking coverage per edge
  kc_final                                spades-kmercount style final_kmers (non-canonical, 16 buckets)
"""
import os
import shutil
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from spades_for_blackbird_b200.host import synth  # noqa: E402

DRIVER = os.path.join(ROOT, "oracle", "_ref", "ref_driver")


def run_ref(reads, k, buckets, mode="gbuilder", tip_bound=None, coverage=True, threads=1):
    tmp = tempfile.mkdtemp(prefix="sb200_golden_")
    try:
        rp = os.path.join(tmp, "reads.txt")
        with open(rp, "w") as f:
            f.write("\n".join(reads) + "\n")
        out = os.path.join(tmp, "out")
        cmd = [DRIVER, "--mode", mode, "--reads", rp, "--out", out, "-k", str(k), "-t", str(threads),
               "--buckets", str(buckets), "--quiet"]
        if tip_bound is not None:
            cmd += ["--tip-bound", str(tip_bound)]
        if coverage and mode == "gbuilder":
            cmd += ["--coverage"]
        subprocess.check_call(cmd, stdout=subprocess.DEVNULL)
        res = {}
        if mode == "kmercount":
            res["kc_final"] = np.fromfile(os.path.join(out, "final_kmers"), dtype=np.uint64)
            res["kc_bucket_sizes"] = np.fromfile(os.path.join(out, "bucket_sizes.u64"), dtype=np.uint64)
            return res
        res["kpomers"] = np.concatenate(
            [np.fromfile(os.path.join(out, "kpomers.%d" % b), dtype=np.uint64) for b in range(buckets)])
        res["kp_bucket_sizes"] = np.fromfile(os.path.join(out, "kpomer_bucket_sizes.u64"), dtype=np.uint64)
        res["kmers"] = np.fromfile(os.path.join(out, "final_kmers"), dtype=np.uint64)
        res["idx"] = np.fromfile(os.path.join(out, "idx.u64"), dtype=np.uint64)
        res["masks_idx"] = np.fromfile(os.path.join(out, "masks_idx.u8"), dtype=np.uint8)
        res["index_bin"] = np.fromfile(os.path.join(out, "index.bin"), dtype=np.uint8)
        if coverage:
            res["coverage"] = np.fromfile(os.path.join(out, "coverage.u32"), dtype=np.uint32)
        with open(os.path.join(out, "unitigs.txt")) as f:
            res["unitigs"] = np.array([l.strip() for l in f if l.strip()])
        # spades-gbuilder --gfa on the same unitigs: the line ORDER depends on the reference's adjacency containers, the SET does not
        with open(os.path.join(out, "graph.gfa")) as f:
            res["gfa"] = np.array(sorted(l.rstrip("\n") for l in f if l.strip()))
        # spades-gbuilder --fastg: one FASTA record per oriented edge; compared record by record after sorting
        with open(os.path.join(out, "graph.fastg")) as f:
            res["fastg"] = np.array(sorted(">" + r.rstrip("\n") for r in f.read().split(">") if r.strip()))
        if coverage:
            # spades-gbuilder -c: the same graph after FillCoverageAndFlankingFromPHM (DP:f / KC:i filled), and the raw flanking
            # coverage of every canonical edge / its conjugate (id order)
            with open(os.path.join(out, "graph_cov.gfa")) as f:
                res["gfa_cov"] = np.array(sorted(l.rstrip("\n") for l in f if l.strip()))
            res["flanking"] = np.array([[int(x) for x in l.split()] for l in open(os.path.join(out, "flanking.txt")) if l.strip()],
                                       dtype=np.int64).reshape(-1, 3)
        clipped = 0
        for line in open(os.path.join(out, "timing.txt")):
            if line.startswith("clipped "):
                clipped = int(line.split()[1])
        res["clipped"] = np.array(clipped)
        return res
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


def rc(s):
    return s[::-1].translate(str.maketrans("ACGT", "TGCA"))


def codes_str(a):
    return "".join("ACGT"[int(c)] for c in a)


def reads_from_genome(g, read_len, n, err, seed, circular=False):
    rng = np.random.default_rng(seed)
    s = codes_str(g)
    if circular:
        s2 = s + s[:read_len]
    out = []
    for _ in range(n):
        if circular:
            p = int(rng.integers(0, len(s)))
            r = s2[p:p + read_len]
        else:
            p = int(rng.integers(0, len(s) - read_len + 1))
            r = s[p:p + read_len]
        r = list(r)
        for i in range(len(r)):
            if rng.random() < err:
                r[i] = "ACGT"[(("ACGT".index(r[i])) + int(rng.integers(1, 4))) % 4]
        r = "".join(r)
        out.append(rc(r) if rng.random() < 0.5 else r)
    return out


def cases():
    """name -> (reads, k, buckets, tip_bound)"""
    c = {}
    # BASELINE config #1: assembler/test_dataset (E. coli 1K), read as plain sequences (the .fq.gz is parsed once here).
    import gzip
    ds = "/root/reference/assembler/test_dataset"
    reads = []
    for fn in ("ecoli_1K_1.fq.gz", "ecoli_1K_2.fq.gz"):
        lines = gzip.open(os.path.join(ds, fn), "rt").read().split("\n")
        reads.append([lines[i] for i in range(1, len(lines), 4)])
    # spades-gbuilder/kmercount see the library as left file then right file (order is irrelevant to the sets)
    ecoli = reads[0] + reads[1]
    c["ecoli1k_k21"] = (ecoli, 21, 80, None)
    c["ecoli1k_k55"] = (ecoli, 55, 80, None)
    # A/test/debruijn/construction_test.cpp:32-66 (k=5, one stream -> 10 buckets)
    ct = {
        "SimpleThread": ["ACAAACCACCA"],
        "SimpleThread2": ["ACAAACCACCC", "AAACCACCCAC"],
        "SplitThread": ["ACAAACCACCA", "ACAAACAACCC"],
        "SplitThread2": ["ACAAACCACCA", "ACAAACAACCA"],
        "Buldge": ["ACAAAACACCA", "ACAAACCACCA"],
        "CondenseSimple": ["CGAAACCAC", "CGAAAACAC", "AACCACACC", "AAACACACC"],
    }
    for name, rd in ct.items():
        c["ctest_" + name] = (rd, 5, 10, None)
    # perfect loop (circular 300-mer), a self-reverse-complement loop a.rc(a), reads with Ns, a short read
    g_loop = synth.random_genome(300, 7)
    loop_reads = reads_from_genome(g_loop, 80, 60, 0.0, 8, circular=True)
    a = codes_str(synth.random_genome(120, 9))
    selfrc = a + rc(a)
    selfrc_reads = reads_from_genome(np.array(["ACGT".index(ch) for ch in selfrc], dtype=np.uint8), 90, 80, 0.0, 10,
                                     circular=True)
    g_lin = synth.random_genome(1500, 11)
    lin_reads = reads_from_genome(g_lin, 100, 200, 0.01, 12)
    lin_reads[3] = lin_reads[3][:40] + "N" + lin_reads[3][41:]
    lin_reads[5] = "NNNN" + lin_reads[5][4:60] + "NN" + lin_reads[5][62:]
    lin_reads[7] = "ACGTACGTAC"
    lin_reads[9] = "N" * 30
    lin_reads[11] = lin_reads[11].lower()
    c["loops_k21"] = (loop_reads + selfrc_reads + lin_reads, 21, 20, None)
    # multi-word records: (k, k+1) word counts (1,2) (2,2) (3,3) (3,3 full) (4,4 full)
    g_mw = synth.random_genome(3000, 21)
    mw_reads = reads_from_genome(g_mw, 250, 240, 0.004, 22)
    for k in (31, 33, 63, 77, 95, 127):
        c["multiword_k%d" % k] = (mw_reads, k, 10, None)
    # early tip clipper (spades-core: bound = RL - k, construction.cpp:300-304), single thread => sequential order
    g_tc = synth.random_genome(4000, 31)
    tc_reads = reads_from_genome(g_tc, 100, 1600, 0.01, 32)
    c["tipclip_k21"] = (tc_reads, 21, 10, 100 - 21)
    c["tipclip_k33"] = (tc_reads, 33, 10, 100 - 33)
    return c


def binary_reads_fixture():
    """binreads/: the reference's own .seq/.off files (io::BinaryWriter::ToBinary, binary_converter.cpp:50-113) for 257 reads
    of mixed lengths with Ns — pins the on-disk read format the C++ adapters parse and emit (tests/test_cpp_adapters.py)."""
    rng = np.random.default_rng(77)
    reads = []
    for i in range(257):
        n = int(rng.integers(0, 200))
        s = "".join("ACGT"[c] for c in rng.integers(0, 4, n))
        if i % 5 == 0 and n > 10:
            p = int(rng.integers(0, n))
            s = s[:p] + "N" * int(rng.integers(1, 4)) + s[p:]
        reads.append(s if s else "N")
    d = os.path.join(HERE, "binreads")
    os.makedirs(d, exist_ok=True)
    with open(os.path.join(d, "reads.txt"), "w") as f:
        f.write("\n".join(reads) + "\n")
    tmp = tempfile.mkdtemp(prefix="sb200_golden_")
    try:
        subprocess.check_call([DRIVER, "--mode", "tobinary", "--reads", os.path.join(d, "reads.txt"), "--out", tmp, "-k", "5", "-t", "1",
                               "--quiet"], stdout=subprocess.DEVNULL)
        for ext in ("seq", "off"):
            shutil.copy(os.path.join(tmp, "lib." + ext), os.path.join(d, "lib." + ext))
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    print("binreads: %d reads, lib.seq %d bytes" % (len(reads), os.path.getsize(os.path.join(d, "lib.seq"))))


def main():
    if not os.path.exists(DRIVER):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "ref"])
    binary_reads_fixture()
    if "--binreads-only" in sys.argv:
        return
    for name, (reads, k, B, tip) in cases().items():
        res = run_ref(reads, k, B, tip_bound=tip)
        res.update(run_ref(reads, k, 16, mode="kmercount"))
        res["reads"] = np.array("\n".join(reads))
        res["k"] = np.array(k)
        res["buckets"] = np.array(B)
        res["tip_bound"] = np.array(-1 if tip is None else tip)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **res)
        print("%-20s k=%-3d B=%-3d reads=%-5d kpomers=%-7d kmers=%-7d unitigs=%-5d clipped=%d" % (
            name, k, B, len(reads), len(res["kp_bucket_sizes"]) and int(res["kp_bucket_sizes"].sum()),
            len(res["idx"]), len(res["unitigs"]), int(res["clipped"])))


if __name__ == "__main__":
    main()
