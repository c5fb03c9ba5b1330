"""The hash-sharded path on REAL NCCL: torch.distributed.run with one process per GPU (2, and 4 when the box has them), fixtures of
the unmodified reference incl. perfect loops, tip clipping, four-word records and an error-free genome (chains beyond the direct-walk
limit).  Skipped on a box with a single GPU (tests/test_gpu_sharded.py covers the same orchestration with virtual ranks there)."""
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

import oracle_lib as O
from spades_for_blackbird_b200.host import synth

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
NAMES = ["ecoli1k_k55", "loops_k21", "tipclip_k21", "multiword_k127", "errorfree_k31"]


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _gpus():
    import torch
    return torch.cuda.device_count()


@pytest.mark.parametrize("G", [2, 4])
def test_sharded_over_nccl(tmp_path, G):
    if _gpus() < G:
        pytest.skip("needs %d GPUs" % G)
    from conftest import load_golden
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(G), "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), os.path.join(ROOT, "tests", "nccl_worker.py"), str(tmp_path), ",".join(NAMES)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    for name in NAMES:
        parts = [np.load(os.path.join(str(tmp_path), "%s_rank%d.npz" % (name, g))) for g in range(G)]
        nb = int(parts[0]["nb"])
        if name == "errorfree_k31":
            reads, k, tip = synth.codes_to_strings(synth.sample_pairs(synth.random_genome(6000, 77), 600, 120, 300, 0.0, 78)), 31, None
            gold = None
        else:
            gold = load_golden(name)
            reads, k, tip = gold["reads"], gold["k"], (int(gold["tip_bound"]) if gold["tip_bound"] >= 0 else None)
        if gold is not None and nb == gold["buckets"]:
            want = dict(kpomers=gold["kpomers"], coverage=gold["coverage"], kmers=gold["kmers"], idx=gold["idx"], masks_idx=gold["masks_idx"],
                        unitigs=list(gold["unitigs"]), clipped=int(gold["clipped"]))
        else:
            w = O.gbuilder(reads, k, nb, tip_bound=tip)
            want = dict(kpomers=w["kpomers"].data.reshape(-1), coverage=w["kpomers"].counts, kmers=w["kmers"].data.reshape(-1), idx=w["idx"],
                        masks_idx=w["masks_idx"], unitigs=list(w["unitigs"]), clipped=w.get("clipped", 0))
        assert np.array_equal(np.concatenate([p["kpomers"] for p in parts]).reshape(-1), want["kpomers"]), name
        assert np.array_equal(np.concatenate([p["counts"] for p in parts]), want["coverage"]), name
        assert np.array_equal(np.concatenate([p["kmers"] for p in parts]).reshape(-1), want["kmers"]), name
        assert np.array_equal(np.concatenate([p["idx"] for p in parts]), want["idx"]), name
        for p in parts:
            assert np.array_equal(p["masks"], want["masks_idx"]), name
            assert np.array_equal(p["index_bin"], parts[0]["index_bin"]), name       # the same whole index on every rank
        assert list(parts[0]["unitigs"]) == want["unitigs"], name
        if tip is not None:
            assert all(int(p["clipped"]) == want["clipped"] for p in parts), name
        if name in ("loops_k21", "errorfree_k31"):
            assert all(bool(p["fallback"]) for p in parts), name
        assert int(parts[0]["bytes_sent"]) > 0
