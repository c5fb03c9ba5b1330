"""The C++ host side (include/sb200_adapters.hpp, driven by tools/sb200_gbuilder.cpp the way spades-gbuilder /
spades-kmercount drive the reference's classes).  CPU: it builds with plain g++ against the C ABI, round-trips the
reference's .seq/.off read format, and refuses to run without a GPU.  GPU: its files equal the reference's."""
import os
import struct
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TOOL = os.path.join(ROOT, "tools", "sb200_gbuilder")


@pytest.fixture(scope="module")
def tool():
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "spades_for_blackbird_b200", "csrc"), "tools"])
    assert os.path.exists(TOOL)
    return TOOL


def write_reads(path, reads):
    with open(path, "w") as f:
        f.write("\n".join(reads) + "\n")


def test_builds_and_fails_loudly_without_gpu(tool, tmp_path):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    write_reads(tmp_path / "r.txt", ["ACGTACGTACGTTTGACCA"])
    p = subprocess.run([tool, "--reads", str(tmp_path / "r.txt"), "--out", str(tmp_path), "-k", "5"], capture_output=True, text=True)
    assert p.returncode == 255
    assert "no CPU fallback" in p.stderr


def test_binary_read_files_layout(tool, tmp_path):
    """--write-binary emits PREFIX.seq/.off in the reference's layout (read_stream.hpp:19-36, single_read.hpp:279-299,
    sequence.hpp:399-441, binary_converter.cpp:50-113): parsed here independently, byte by byte."""
    reads = ["ACGTNNACGTACGGT", "T" * 70, "", "NNNN", "ACGT" * 40 + "N" + "A" * 3] + ["ACGTTGCA" * 5] * 230
    write_reads(tmp_path / "r.txt", [r if r else "N" for r in reads])
    subprocess.run([tool, "--reads", str(tmp_path / "r.txt"), "--write-binary", str(tmp_path / "lib"), "--out", str(tmp_path), "-k", "5"],
                   capture_output=True)
    seq = open(tmp_path / "lib.seq", "rb").read()
    off = np.frombuffer(open(tmp_path / "lib.off", "rb").read(), dtype=np.uint64)
    want = ["ACGTACGGT", "T" * 70, "", "", "ACGT" * 40] + ["ACGTTGCA" * 5] * 230     # LongestValid: first longest ACGT run
    count, max_len, total = struct.unpack_from("<QQQ", seq, 0)
    assert (count, max_len, total) == (len(want), max(map(len, want)), sum(map(len, want)))
    pos, starts = 24, []
    for i, w in enumerate(want):
        starts.append(pos)
        (n,) = struct.unpack_from("<Q", seq, pos)
        assert n == len(w)
        nw = (n + 31) // 32
        words = np.frombuffer(seq, dtype=np.uint64, count=nw, offset=pos + 8)
        got = "".join("ACGT"[(int(words[j >> 5]) >> (2 * (j & 31))) & 3] for j in range(n))
        assert got == w
        raw = (reads[i] if reads[i] else "N")
        left = raw.find(w) if w else struct.unpack_from("<H", seq, pos + 8 + 8 * nw)[0]
        assert struct.unpack_from("<HH", seq, pos + 8 + 8 * nw) == (left, len(raw) - left - n)   # what LongestValid trimmed
        pos += 8 + 8 * nw + 4
    assert pos == len(seq)
    assert list(off) == starts[::100]         # one offset per CHUNK = 100 records


@pytest.mark.gpu
def test_gbuilder_files_equal_reference(tool, tmp_path, golden):
    g = golden
    if g["buckets"] % 10:
        pytest.skip("bucket count is not 10 x threads")
    write_reads(tmp_path / "r.txt", g["reads"])
    args = [tool, "--reads", str(tmp_path / "r.txt"), "--write-binary", str(tmp_path / "lib"), "--out", str(tmp_path), "-k", str(g["k"]),
            "-t", str(g["buckets"] // 10), "--coverage", "--self-check"]
    if g["tip_bound"] >= 0:
        args += ["--tip-clip", str(g["tip_bound"])]
    out = subprocess.run(args, capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "self-check OK" in out.stdout          # KMerDiskStorage bucket iterators + KMerIndex::seq_idx(const Seq &) on the host
    assert np.array_equal(np.fromfile(tmp_path / "kpomers", dtype=np.uint64), g["kpomers"])
    assert np.array_equal(np.fromfile(tmp_path / "final_kmers", dtype=np.uint64), g["kmers"])
    assert np.array_equal(np.fromfile(tmp_path / "coverage.u32", dtype=np.uint32), g["coverage"])
    assert np.array_equal(np.fromfile(tmp_path / "masks_idx.u8", dtype=np.uint8), g["masks_idx"])
    assert open(tmp_path / "unitigs.txt").read().split() == g["unitigs"]
    # a12: link records keyed by the MPHF -> vertices -> GFA; "byte-identical after canonical ordering" (line order in the reference
    # follows its adjacency containers)
    assert sorted(open(tmp_path / "graph.gfa").read().splitlines()) == list(g["gfa"])
    assert sorted(">" + r.rstrip("\n") for r in open(tmp_path / "graph.fastg").read().split(">") if r.strip()) == list(g["fastg"])
    # spades-gbuilder -c — per-edge coverage from the (k+1)-mer multiplicities (DP:f / KC:i), flanking coverage
    assert sorted(open(tmp_path / "graph_cov.gfa").read().splitlines()) == list(g["gfa_cov"])
    fl = np.array([[int(x) for x in l.split()] for l in open(tmp_path / "flanking.txt")], dtype=np.int64).reshape(-1, 3)
    assert np.array_equal(fl[np.argsort(fl[:, 0])], g["flanking"][np.argsort(g["flanking"][:, 0])])
    # second run from the binary read files it wrote (io::BinaryFileStream path): same graph
    out2 = tmp_path / "o2"
    out2.mkdir()
    subprocess.check_call([tool, "--binary", str(tmp_path / "lib"), "--out", str(out2), "-k", str(g["k"]), "-t", str(g["buckets"] // 10)]
                          + (["--tip-clip", str(g["tip_bound"])] if g["tip_bound"] >= 0 else []))
    assert open(out2 / "unitigs.txt").read().split() == g["unitigs"]


@pytest.mark.gpu
def test_kmercount_file_equals_reference(tool, tmp_path, golden):
    g = golden
    write_reads(tmp_path / "r.txt", g["reads"])
    subprocess.check_call([tool, "--mode", "kmercount", "--reads", str(tmp_path / "r.txt"), "--out", str(tmp_path), "-k", str(g["k"])])
    assert np.array_equal(np.fromfile(tmp_path / "final_kmers", dtype=np.uint64), g["kc_final"])


def test_binary_read_files_equal_reference(tool, tmp_path):
    """tests/golden/binreads/lib.seq|.off were written by the reference's io::BinaryWriter (make_golden.py); the adapter's
    writer must reproduce them byte for byte from the same raw reads, and its reader must accept the reference's files."""
    gold = os.path.join(ROOT, "tests", "golden", "binreads")
    subprocess.run([tool, "--reads", os.path.join(gold, "reads.txt"), "--write-binary", str(tmp_path / "lib"), "--out", str(tmp_path), "-k", "5"],
                   capture_output=True)
    for ext in ("seq", "off"):
        assert open(tmp_path / ("lib." + ext), "rb").read() == open(os.path.join(gold, "lib." + ext), "rb").read()
    # reader: reference files in, same files out
    subprocess.run([tool, "--binary", os.path.join(gold, "lib"), "--write-binary", str(tmp_path / "again"), "--out", str(tmp_path), "-k", "5"],
                   capture_output=True)
    for ext in ("seq", "off"):
        assert open(tmp_path / ("again." + ext), "rb").read() == open(os.path.join(gold, "lib." + ext), "rb").read()
