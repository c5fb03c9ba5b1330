"""bench.py without a GPU: the reference arm (`--impl reference` runs oracle/_ref/ref_driver — the unmodified reference — on the host
cores) prints the contract's JSON line, and the synthetic workloads of BASELINE configs 2-5 are deterministic in their seeds."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

from spades_for_blackbird_b200.host import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "oracle", "_ref", "ref_driver")


@pytest.mark.skipif(not os.path.exists(REF), reason="oracle/_ref/ref_driver not built (python -c 'import __graft_entry__ as g; g.build()')")
def test_reference_arm_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--config", "1", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == "reads_to_condensed_dbg_throughput" and line["unit"] == "Gbp/s"
    assert line["higher_is_better"] is True and line["n_gpus"] == 1 and line["steps"] == 1 and line["warmup"] == 0
    assert line["value"] > 0 and line["ms_per_step"] > 0
    cb = line["cpu_baseline"]
    assert cb["kind"] == "reference" and cb["cores"] >= 1 and cb["value"] == line["value"] and cb["sample"]
    e2e = line["e2e"]
    assert e2e["value"] == line["value"] and e2e["unit"] == line["unit"] and e2e["h2d_bytes_per_step"] == 0 and e2e["d2h_bytes_per_step"] == 0
    assert line["config"]["config"] == 1 and "workload" in line["config"]


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", LOCAL_RANK="1", WORLD_SIZE="2")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--config", "1"],
                       capture_output=True, text=True, timeout=600, cwd=ROOT, env=env)
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_workloads_are_deterministic_and_shaped():
    for cfg, w in synth.WORKLOADS.items():
        assert {"ks", "scaling", "read_len", "seed"} <= set(w), cfg
    assert synth.WORKLOADS[4]["ks"] == (127,) and synth.WORKLOADS[3]["ks"] == (21, 33, 55, 77) and synth.WORKLOADS[5]["scaling"] == "strong"
    for cfg in (2, 5):
        w = synth.WORKLOADS[cfg]
        a = synth.workload_reads(cfg, 0, 1, max_reads=2000)
        b = synth.workload_reads(cfg, 0, 1, max_reads=2000)
        assert all(np.array_equal(x, y) for x, y in zip(a, b)), cfg                 # same seed, same reads
        words, off, ln = a
        assert len(ln) == 2000 and int(ln[0]) == w["read_len"] and len(off) == 2001
        assert len(words) == 2000 * ((w["read_len"] + 31) // 32)
        tail = w["read_len"] % 32                                                    # padding bits of every read's last word are zero
        if tail:
            last = words.reshape(2000, -1)[:, -1]
            assert not np.any(last >> np.uint64(2 * tail))
    # weak scaling: another rank of a 2-GPU job samples different reads (of a genome twice as long)
    r1 = synth.workload_reads(2, 1, 2, max_reads=2000)
    assert not np.array_equal(r1[0], synth.workload_reads(2, 0, 2, max_reads=2000)[0])
    # strong scaling: the ranks' chunks of config 5 interleave to the one-rank read set
    lens, pairs = synth._metagenome(synth.WORKLOADS[5])
    assert len(lens) == 200 and lens.min() >= 1_000_000 and lens.max() <= 8_000_000
    assert abs(int(pairs.sum()) * 2 * 150 - 1_000_000_000) < 5_000_000
