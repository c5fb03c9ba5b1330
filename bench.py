#!/usr/bin/env python
"""bench.py — reads -> condensed de Bruijn graph throughput (BASELINE.json metric) on N B200s of one node.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--config 1..5] [--verify]

Workloads (--config, 1-based index into BASELINE.json `configs`; generators in host/synth.py):
    1  assembler/test_dataset (E. coli 1K, real reads, from tests/golden/ecoli1k_k21.npz), k = 21
    2  synthetic 4.6 Mbp genome, 2x150 at 100x, k = 55                      <- default: the configuration the metric is quoted on
    3  the same reads, k = 21, 33, 55, 77 back to back on the resident reads
    4  2x250 at 80x, k = 127 (four-word records)
    5  metagenome mix, 200 genomes, ~1 Gbp of 2x150 reads, k = 55 (strong scaling: the read set is split over the ranks)
A step = one pass of the whole hot path over the read set (config 3: one pass per K):
    packed reads -> canonical (k+1)-mers -> sort/dedup(+counts) -> k-mers -> BooPHF MPHF -> extension masks -> unitigs
`value`  : input bases (x number of Ks) / device time with the packed reads already resident in HBM (CUDA events on the library's stream)
`e2e`    : the same through sb200_construct() / sb200_construct_sharded() with pinned HOST buffers on both sides, as SURVEY.md 8(d) defines
           the metric: from "packed reads resident in pinned host memory" to "unitig buffers + index resident on host" (H2D of the reads, D2H
           of the KMerIndex bytes, the mask array and the packed unitigs inside the timed region); `e2e_all_tables` also brings both k-mer
           tables home (the reference's temporary files)
`roofline`: N = 1: the kernel with the largest measured share of the step — algorithmic bytes per launch / its mean launch duration,
           measured live with CUDA events around every launch of the timed steps.  N > 1: the record exchange against NVLink.
`cpu_baseline` / --impl reference: the UNMODIFIED reference (oracle/_ref/ref_driver, compiled from /root/reference) on the box's
           host cores over a bounded sample of the same read set, same bucket count.
`parity` : digests of the tables / masks / unitigs against the reference — N = 1: the reference's own run of the cpu_baseline leg on
           the same sample (and, with --verify, tests/golden/fullsize_digests.json at full size); N > 1 with --verify: rank 0 redoes
           the whole read set on one GPU and every rank's shard must match its slice.
"""
import argparse
import hashlib
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
from spades_for_blackbird_b200.host import synth  # noqa: E402

METRIC = "reads_to_condensed_dbg_throughput"
# dram__bytes_read.sum + dram__bytes_write.sum per launch (bytes) from the committed ncu --set full capture of this same command
# (profiles/); keys are the names of sb200_profile_report().
NCU_TRAFFIC = {}
try:
    NCU_TRAFFIC = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
except Exception:
    pass
UNIT = "Gbp/s"
NVLINK_PEAK_GBS = 770.0   # measured peer copy per direction per GPU on this pool (B200_PROFILING.md; 900 nominal)


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", type=int, default=2, choices=[1, 2, 3, 4, 5])
    ap.add_argument("--buckets", type=int, default=80)
    ap.add_argument("--cpu-sample-reads", type=int, default=3_000_000,
                    help="reads of the workload the reference CPU path is timed on (about 10-20 s of CPU work per pass)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--e2e-tables", action="store_true", help="N > 1: the e2e leg also brings every rank's shard of both k-mer tables home")
    ap.add_argument("--no-parity", action="store_true", help="skip the digest comparison with the reference run of the cpu_baseline leg")
    ap.add_argument("--verify", action="store_true", help="full-size digest check (see module docstring)")
    ap.add_argument("--breakdown", action="store_true", help="print the per-kernel timing table to stderr")
    return ap.parse_args()


def workload(a):
    if a.config == 1:
        return dict(name="assembler/test_dataset E. coli 1K paired reads, k=21 (BASELINE configs[0])", ks=(21,), scaling="weak")
    return synth.WORKLOADS[a.config]


def workload_config(a, n_gpus):
    w = workload(a)
    cfg = {"workload": w["name"], "config": a.config, "k": list(w["ks"]) if len(w["ks"]) > 1 else w["ks"][0], "num_buckets": a.buckets,
           "error_rate": synth.ERROR_RATE if a.config != 1 else None,
           "reads_per_gpu": "all" if n_gpus == 1 else ("1/%d of the read set" % n_gpus if w["scaling"] == "strong" else "one full read set each"),
           "l2": "inputs and every intermediate are larger than L2 (>= 115 MB of reads, GBs of k-mer instances)" if a.config != 1 else
                 "tiny input: launch-latency bound, listed for completeness"}
    for key in ("genome_len", "read_len", "coverage", "seed", "total_bases", "n_genomes"):
        if key in w:
            cfg[key] = w[key]
    return cfg


def load_reads(a, rank=0, world=1, max_reads=None):
    if a.config == 1:
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        reads = str(np.load(os.path.join(ROOT, "tests", "golden", "ecoli1k_k21.npz"))["reads"]).split("\n")
        lut = np.zeros(256, dtype=np.uint8)
        lut[ord("C")] = 1; lut[ord("G")] = 2; lut[ord("T")] = 3
        parts, offs, lens = [], [0], []
        for r in reads:
            codes = lut[np.frombuffer(r.encode(), dtype=np.uint8)][None, :]
            parts.append(synth.pack_codes(codes)[0])
            offs.append(offs[-1] + len(parts[-1]))
            lens.append(len(r))
        return np.concatenate(parts), np.array(offs, dtype=np.uint64), np.array(lens, dtype=np.uint32)
    return synth.workload_reads(a.config, rank, world, max_reads)


def md5(a):
    return hashlib.md5(np.ascontiguousarray(a).tobytes()).hexdigest()


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, device):
        self.rows, self.proc, self.device = [], None, device

    def start(self):
        q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.device), "--query-gpu=" + q, "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            pass
        sm = [float(r[0]) for r in self.rows if len(r) >= 7 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 7 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 7 for i in range(4) if r[3 + i].lower().startswith("active")})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(sm)}


class ReferenceSample:
    """A bounded sample of the workload for the reference's CPU path: the FIRST `n_reads` reads of the same read set (same genome,
    same chunk seeds as the GPU arm), written once as plain sequences for oracle/_ref/ref_driver."""

    def __init__(self, a, n_reads):
        self.a = a
        self.k = workload(a)["ks"][0] if a.config != 3 else 55
        self.driver = os.path.join(ROOT, "oracle", "_ref", "ref_driver")
        self.tmp = None
        self.n_reads = 0
        if not os.path.exists(self.driver):
            return
        self.tmp = tempfile.mkdtemp(prefix="sb200_bench_")
        self.reads_path = os.path.join(self.tmp, "reads.txt")
        with open(self.reads_path, "w") as f:
            if a.config == 1:
                reads = str(np.load(os.path.join(ROOT, "tests", "golden", "ecoli1k_k21.npz"))["reads"]).split("\n")
                f.write("\n".join(reads) + "\n")
                self.n_reads = len(reads)
            else:
                for c in synth.workload_chunks(a.config):   # chunk c of the sample IS chunk c of the workload
                    c = c[:max(n_reads - self.n_reads, 0)]
                    if len(c) == 0:
                        break
                    f.write("\n".join(synth.codes_to_strings(c)) + "\n")
                    self.n_reads += len(c)

    def available(self):
        return self.tmp is not None

    def run(self, cores, dump=False):
        """One timed pass of the unmodified reference path (KMerDiskCounter -> ExtensionIndex -> UnbranchingPathExtractor) with
        `cores` threads and the GPU arm's bucket count.  Returns (Gbp/s, seconds, bases); read parsing and the dumps are outside the
        driver's `path_total`."""
        out = os.path.join(self.tmp, "out")
        subprocess.call(["rm", "-rf", out])
        cmd = [self.driver, "--mode", "gbuilder", "--reads", self.reads_path, "--out", out, "-k", str(self.k), "-t", str(cores),
               "--buckets", str(self.a.buckets), "--quiet"]
        subprocess.check_call(cmd + (["--no-graph"] if dump else ["--no-dump"]), stdout=subprocess.DEVNULL)
        t = {}
        for line in open(os.path.join(out, "timing.txt")):
            p = line.split()
            t[p[0]] = float(p[1])
        secs = t["path_total"]
        bases = int(t["bases"])
        self.out = out
        return bases / secs / 1e9, secs, bases

    def digests(self):
        """md5 of what the reference dumped (run(dump=True)): the k-mer table, the mask array in index order, the index bytes, the
        unitigs (packed like the device output)"""
        d = {}
        h = hashlib.md5()
        for b in range(self.a.buckets):
            h.update(open(os.path.join(self.out, "kpomers.%d" % b), "rb").read())
        d["kpomers"] = h.hexdigest()
        for name, fn in (("kmers", "final_kmers"), ("masks", "masks_idx.u8"), ("index", "index.bin")):
            d[name] = hashlib.md5(open(os.path.join(self.out, fn), "rb").read()).hexdigest()
        words, word_off, lens = synth.pack_text_sequences(open(os.path.join(self.out, "unitigs.txt"), "rb").read())
        d["unitig_words"], d["unitig_len"] = md5(words), md5(lens)
        d["n_unitigs"] = int(len(lens))
        return d

    def close(self):
        if self.tmp:
            subprocess.call(["rm", "-rf", self.tmp])
            self.tmp = None


def gpu_digests(B, ctx, words, word_off, lens, k, buckets):
    """the same digests from one pass of the CUDA path over host reads"""
    g = B.construct(ctx, words, word_off, lens, k, buckets, fetch_kmers=True)
    v = g.view
    wu, ou, lu = g.unitigs_packed()
    W1, W0 = (k + 1 + 31) // 32, (k + 31) // 32
    d = {"kpomers": md5(np.ctypeslib.as_array(v.kpomers, shape=(v.n_kpomers * W1,))), "kmers": md5(np.ctypeslib.as_array(v.kmers, shape=(v.n_kmers * W0,))),
         "coverage": md5(np.ctypeslib.as_array(v.kpomer_counts, shape=(v.n_kpomers,))), "masks": md5(g.masks()), "index": md5(g.index_bytes()),
         "unitig_words": md5(wu), "unitig_len": md5(lu), "n_unitigs": int(len(lu))}
    g.free()
    return d


def reference_arm(a):
    """--impl reference: the reference's own CPU implementation of the path on this box's host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    cores = os.cpu_count() or 1
    smp = ReferenceSample(a, a.cpu_sample_reads)
    if not smp.available():
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/ref_driver is not built (no /root/reference at build time)"}))
        return 0
    vals, secs_all = [], []
    try:
        for i in range(a.warmup + a.steps):
            r = smp.run(cores)
            if i >= a.warmup:
                vals.append(r[0]); secs_all.append(r[1])
            bases = r[2]
    finally:
        smp.close()
    v = float(np.mean(vals))
    sample = "first %d reads (%.1f Mbp) of the workload; whole reference path incl. its temp-file I/O, %d buckets, k=%d, %.1f s per step, -t %d" % (
        smp.n_reads, bases / 1e6, a.buckets, smp.k, float(np.mean(secs_all)), cores)
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup,
            "ms_per_step": 1e3 * float(np.mean(secs_all)), "higher_is_better": True, "scaling": workload(a)["scaling"], "vs_baseline": None,
            "dtype": "u64", "data": "synthetic" if a.config != 1 else "assembler/test_dataset reads", "config": workload_config(a, 1),
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "reference", "sample": sample},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))
    return 0


def cpu_baseline_and_parity(a, B, ctx, want_parity):
    """rank 0: the reference on a bounded sample (timed), and — same sample, same bucket count — the digest comparison with the CUDA path"""
    cores = os.cpu_count() or 1
    smp = ReferenceSample(a, a.cpu_sample_reads)
    if not smp.available():
        return None, None
    cpu = parity = None
    try:
        r = smp.run(cores, dump=want_parity)
        cpu = {"value": r[0], "unit": UNIT, "cores": cores, "kind": "reference",
               "sample": "first %d reads (%.1f Mbp) of the workload, whole reference path in %.1f s, %d buckets, k=%d, -t %d" % (
                   smp.n_reads, r[2] / 1e6, r[1], a.buckets, smp.k, cores)}
        if want_parity:
            ref = smp.digests()
            words, word_off, lens = load_reads(a, max_reads=smp.n_reads)
            got = gpu_digests(B, ctx, words, word_off, lens, smp.k, a.buckets)
            keys = ("kpomers", "kmers", "masks", "index", "unitig_words", "unitig_len", "n_unitigs")
            parity = {"against": "oracle/_ref/ref_driver (unmodified reference) on the cpu_baseline sample: first %d reads, k=%d, %d buckets" % (
                          smp.n_reads, smp.k, a.buckets),
                      "equal": {key: bool(ref[key] == got[key]) for key in keys}, "n_unitigs": got["n_unitigs"]}
            parity["ok"] = all(parity["equal"].values())
    finally:
        smp.close()
    return cpu, parity


def verify_full_size(a, B, ctx, words, word_off, lens):
    """--verify at N = 1: digests of the FULL workload against what the reference produced for it (tests/golden/fullsize_digests.json)"""
    try:
        ref_all = json.load(open(os.path.join(ROOT, "tests", "golden", "fullsize_digests.json")))
    except Exception:
        return {"ok": None, "note": "tests/golden/fullsize_digests.json not found"}
    out = {}
    for k in workload(a)["ks"]:
        ref = ref_all.get("config%d_k%d" % (a.config, k))
        if ref is None:
            out["k%d" % k] = {"ok": None, "note": "no reference digest for this configuration"}
            continue
        got = gpu_digests(B, ctx, words, word_off, lens, k, a.buckets)
        eq = {"kpomers": got["kpomers"] == ref["kpomers_md5"], "coverage": got["coverage"] == ref["coverage_md5"], "kmers": got["kmers"] == ref["kmers_md5"],
              "masks": got["masks"] == ref["masks_idx_md5"], "index": got["index"] == ref["index_bin_md5"],
              "unitig_words": got["unitig_words"] == ref["unitig_words_md5"], "unitig_len": got["unitig_len"] == ref["unitig_len_md5"]}
        out["k%d" % k] = {"equal": {k_: bool(v_) for k_, v_ in eq.items()}, "ok": all(eq.values()), "n_unitigs": got["n_unitigs"]}
    out["against"] = "the unmodified reference on the full workload (tests/golden/fullsize_digests.json, made by tests/golden/make_fullsize_digests.py)"
    out["ok"] = all(v.get("ok") for key, v in out.items() if key.startswith("k"))
    return out


def main_sharded(a, world, rank, local_rank):
    """N > 1: one rank per GPU, k-mer space sharded by hash bucket; the orchestration is C++ (csrc/shard.cu over NCCL).  Weak-scaling
    workloads grow the genome with N (every rank brings one full read set); the metagenome is split over the ranks."""
    import torch
    import torch.distributed as dist
    from spades_for_blackbird_b200.host import binding as B
    from spades_for_blackbird_b200.host import distributed as D

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist.init_process_group("nccl", device_id=dev)
    if a.buckets % world:
        raise SystemExit("--buckets must be a multiple of the number of GPUs")
    w = workload(a)
    ks = w["ks"]
    words, word_off, lens = load_reads(a, rank, world)
    pw = torch.from_numpy(words.view(np.int64)).pin_memory()
    hw = pw.numpy().view(np.uint64)
    t = torch.tensor([int(lens.astype(np.int64).sum())], device=dev, dtype=torch.int64)
    dist.all_reduce(t)
    total_bases = int(t.item()) * len(ks)

    ctx = B.Context(local_rank)
    comm = D.nccl_comm(ctx, rank, world)
    stream = torch.cuda.ExternalStream(ctx.stream(), device=dev)
    streams = B.ReadStreams(ctx, hw, word_off, lens)
    info, stage_acc, pinned_bufs = {}, {}, {}
    xchg = {"bytes": 0, "ms": 0.0}

    def pinned(name, nbytes):   # grow-only pinned host buffers, reused by every step
        t_ = pinned_bufs.get(name)
        if t_ is None or t_.numel() < nbytes:
            t_ = torch.empty(int(nbytes * 1.05) + 64, dtype=torch.uint8).pin_memory()
            pinned_bufs[name] = t_
        return t_

    def step(e2e):
        rs = B.ReadStreams(ctx, hw, word_off, lens) if e2e else streams      # e2e: H2D of the reads inside the timed region
        d2h = 0
        for k in ks:
            # device-resident figure: unitigs gathered to rank 0 (the north-star's "final gather of unitig fragments");
            # e2e: every rank brings its own shard AND its own unitig slice home over its own PCIe link
            sh = B.construct_sharded(ctx, comm, rs, k, a.buckets, gather_to=-1 if e2e else 0)
            if not e2e:
                for k_, v_ in sh.stage_ms.items():
                    stage_acc[k_] = stage_acc.get(k_, 0.0) + v_
                xchg["bytes"] += int(sh.info.record_bytes); xchg["all"] = xchg.get("all", 0) + int(sh.info.bytes_sent); xchg["ms"] += float(sh.info.exchange_ms)
            info.update(kpomers=int(sh.info.total_kpomers), instances=int(sh.info.total_instances), kmers=int(sh.info.total_kmers),
                        unitigs=int(sh.info.total_unitigs), unitig_bases=int(sh.info.total_unitig_bases),
                        whole_set_fallback=bool(sh.info.whole_set_fallback))
            if e2e and a.e2e_tables:
                kp, km = sh.kpomers, sh.kmers
                hp = pinned("kp", kp.total_kmers() * kp.words * 8)
                hc = pinned("kc", kp.total_kmers() * 4)
                hk = pinned("km", km.total_kmers() * km.words * 8)
                ctx.check(ctx.lib.sb200_kmers_download(kp.h, 0, kp.total_kmers(), B.C.cast(hp.data_ptr(), B.u64p)))
                ctx.check(ctx.lib.sb200_kmers_counts_download(kp.h, 0, kp.total_kmers(), B.C.cast(hc.data_ptr(), B.u32p)))
                ctx.check(ctx.lib.sb200_kmers_download(km.h, 0, km.total_kmers(), B.C.cast(hk.data_ptr(), B.u64p)))
                d2h += kp.total_kmers() * (kp.words * 8 + 4) + km.total_kmers() * km.words * 8
            if e2e:
                lib = ctx.lib
                n, nw = lib.sb200_unitigs_count(sh.unitigs_h), lib.sb200_unitigs_total_words(sh.unitigs_h)
                uw, uo, ul = pinned("uw", nw * 8 + 8), pinned("uo", (n + 1) * 8), pinned("ul", n * 4 + 4)
                ctx.check(lib.sb200_unitigs_download(sh.unitigs_h, B.C.cast(uw.data_ptr(), B.u64p), B.C.cast(uo.data_ptr(), B.u64p),
                                                     B.C.cast(ul.data_ptr(), B.u32p)))
                d2h += nw * 8 + (n + 1) * 8 + n * 4
                if rank == 0:   # one copy of the extension index: the masks and the KMerIndex bytes of the whole index
                    hm = pinned("masks", int(sh.info.total_kmers))
                    ctx.check(lib.sb200_ext_masks_download(sh.ext, B.C.cast(hm.data_ptr(), B.u8p)))
                    nb_ = B.C.c_uint64()
                    ctx.check(lib.sb200_mphf_serialize(sh.index.h, None, B.C.byref(nb_)))
                    hi_ = pinned("index", nb_.value)
                    ctx.check(lib.sb200_mphf_serialize(sh.index.h, B.C.cast(hi_.data_ptr(), B.u8p), B.C.byref(nb_)))
                    d2h += int(sh.info.total_kmers) + nb_.value
            sh.free()
        if e2e:
            rs.free()
        return d2h

    def barrier():
        dist.barrier()
        torch.cuda.synchronize()

    def timed(e2e):
        for _ in range(a.warmup):
            step(e2e)
        stage_acc.clear(); xchg["bytes"] = 0; xchg["all"] = 0; xchg["ms"] = 0.0
        barrier()
        t0 = time.perf_counter()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        d2h = 0
        for _ in range(a.steps):
            d2h = step(e2e)
        e1.record(stream)
        barrier()
        ms = max(e0.elapsed_time(e1), (time.perf_counter() - t0) * 1e3 if e2e else 0.0)   # e2e: host-side work is part of the call
        t_ = torch.tensor([ms, float(d2h)], device=dev, dtype=torch.float64)
        mx = t_.clone()
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        dist.all_reduce(t_, op=dist.ReduceOp.SUM)
        return float(mx[0].item()) / a.steps, int(t_[1].item())

    sampler = ClockSampler(local_rank)
    sampler.start()
    ctx.kernel_launches(reset=True)
    ms_per_step, _ = timed(False)
    stage_ms = {k_: v_ / a.steps for k_, v_ in stage_acc.items()}
    launches = ctx.kernel_launches()
    clocks = sampler.stop()
    xb, xm = xchg["bytes"] / a.steps, xchg["ms"] / a.steps
    if a.breakdown:   # one more step with per-kernel CUDA events: rank 0's kernels against its stage times (the rest is exchange + waiting)
        ctx.profile(True)
        stage_acc.clear()
        step(False)
        torch.cuda.synchronize()
        rep = ctx.profile_report()
        ctx.profile(False)
        if rank == 0:
            print("stage_ms (profiled step): %s" % json.dumps({k_: round(v_, 3) for k_, v_ in stage_acc.items()}), file=sys.stderr)
            for name, n, t in rep:
                print("%-40s %6d launches %10.3f ms/step" % (name, n, t), file=sys.stderr)
        dist.barrier()
    ms_e2e, d2h = (None, 0) if a.no_e2e else timed(True)

    parity = None
    if a.verify:   # every rank's shard against its slice of a ONE-GPU run over the whole read set (rank 0)
        k = ks[0]
        sh = B.construct_sharded(ctx, comm, streams, k, a.buckets, gather_to=0)
        mine = {"kpomers": md5(sh.kpomers.final_kmers()), "counts": md5(sh.kpomers.counts()), "kmers": md5(sh.kmers.final_kmers())}
        allm = [None] * world
        dist.all_gather_object(allm, mine)
        if rank == 0:
            uw, uo, ul = sh.unitigs_packed()
            masks = sh.masks()
            parts = [load_reads(a, r, world) for r in range(world)]
            wa = np.concatenate([p[0] for p in parts])
            la = np.concatenate([p[2] for p in parts])
            oa = np.concatenate([[0], np.cumsum((la.astype(np.int64) + 31) // 32)]).astype(np.uint64)
            sh.free()
            rs1 = B.ReadStreams(ctx, wa, oa, la)
            index = B.DeBruijnExtensionIndex(ctx, k)
            kp = B.DeBruijnExtensionIndexBuilder().BuildExtensionIndexFromStream(index, rs1, num_buckets=a.buckets)
            w1, o1, l1 = B.UnbranchingPathExtractor(index, k).ExtractUnbranchingPathsAndLoops(packed=True)
            eq = {"masks": bool(np.array_equal(masks, index.data())), "unitig_words": bool(np.array_equal(uw, w1)), "unitig_len": bool(np.array_equal(ul, l1))}
            n_own = a.buckets // world
            for r in range(world):
                lo, hi = int(kp.bucket_starts[r * n_own]), int(kp.bucket_starts[(r + 1) * n_own])
                lo0, hi0 = int(index.kmers.bucket_starts[r * n_own]), int(index.kmers.bucket_starts[(r + 1) * n_own])
                eq["rank%d" % r] = bool(md5(kp._download(lo, hi - lo)) == allm[r]["kpomers"] and md5(index.kmers._download(lo0, hi0 - lo0)) == allm[r]["kmers"]
                                        and md5(kp.counts()[lo:hi]) == allm[r]["counts"])
            parity = {"against": "the same %d read sets through the single-GPU path on rank 0 (k=%d)" % (world, k), "equal": eq, "ok": all(eq.values())}
            index.free(); kp.free(); rs1.free()
        else:
            sh.free()
        dist.barrier()

    cpu = None
    if rank == 0 and not a.no_cpu_baseline:
        cpu, _ = cpu_baseline_and_parity(a, B, ctx, False)
    if rank == 0:
        cfg = workload_config(a, world)
        if w["scaling"] == "weak":
            cfg["workload"] += "; N>1: genome x %d, every rank brings one full read set, k-mer space sharded by hash bucket" % world
        W1 = 8 * ((ks[0] + 1 + 31) // 32)
        nv = xb / (xm * 1e-3) / 1e9 if xm > 0 else None
        roof = {"bound": "nvlink", "achieved": nv, "peak": NVLINK_PEAK_GBS, "unit": "GB/s", "frac": (nv / NVLINK_PEAK_GBS) if nv else None,
                "traffic": None,
                "kernel": "sp_scatter_reads_kernel<W, PEER> + sp_scatter_derive_kernel<WS, W, PEER>: pass 1 of the two groupings, storing every owner's runs "
                          "into the owner's receive buffer over NVLink (the record exchanges)",
                "record_bytes_leaving_this_gpu_per_step": xb, "all_bytes_leaving_this_gpu_per_step": xchg.get("all", 0) / a.steps,
                "exchange_ms_per_step": xm, "share_of_step": xm / ms_per_step if ms_per_step else None,
                "peak_source": "measured peer copy per direction per GPU (B200_PROFILING.md); 900 GB/s nominal; tools/ubench_p2p.cu: 717 GB/s by SM stores, 780 by the copy engine",
                "note": "achieved = record bytes rank 0 stored into other ranks' HBM / device time of the two kernels that produce AND send them "
                        "(producer compute included: the kernels are not pure copies); instance record = %d B" % W1}
        line = {"metric": METRIC, "value": total_bases / (ms_per_step * 1e-3) / 1e9, "unit": UNIT, "n_gpus": world, "steps": a.steps,
                "warmup": a.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": w["scaling"], "vs_baseline": None,
                "dtype": "u64", "data": "synthetic", "config": cfg, "clocks": clocks, "gpu_launches": int(launches),
                "e2e": None if ms_e2e is None else {"value": total_bases / (ms_e2e * 1e-3) / 1e9, "unit": UNIT,
                                                    "h2d_bytes_per_step": int(words.nbytes + word_off.nbytes + lens.nbytes) * world,
                                                    "d2h_bytes_per_step": int(d2h), "ms_per_step": ms_e2e,
                                                    "returns": "every rank: its slice of the packed unitigs over its own PCIe link; rank 0: the extension "
                                                               "index (KMerIndex bytes + mask array) — the metric's 'unitig buffers + index resident on host'"
                                                               + ("; every rank also its shard of both k-mer tables" if a.e2e_tables else "")},
                "roofline": roof, "cpu_baseline": cpu, "stage_ms_rank0": stage_ms, "parity": parity,
                "counts": {k_: (int(v_) if not isinstance(v_, bool) else v_) for k_, v_ in info.items()},
                "exchange": "2 x peer-store pass 1 (k-mer instances, k-mer candidates; CUDA IPC mappings over NVLink), one NCCL send/recv group for the MPHF bit-vector / rank / mask slices, gather of unitigs"}
        print(json.dumps(line))
    comm.free()
    dist.destroy_process_group()
    return 0


def main():
    a = parse_args()
    if a.impl == "reference":
        return reference_arm(a)
    if int(os.environ.get("WORLD_SIZE", "1")) > 1:
        return main_sharded(a, int(os.environ["WORLD_SIZE"]), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")))

    import torch
    from spades_for_blackbird_b200.host import binding as B

    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    w = workload(a)
    ks = w["ks"]

    words, word_off, lens = load_reads(a)
    total_bases = int(lens.astype(np.int64).sum())

    ctx = B.Context(local_rank)
    stream = torch.cuda.ExternalStream(ctx.stream(), device=dev)
    streams = B.ReadStreams(ctx, words, word_off, lens)   # resident in HBM before the timed region

    stage_s = {}

    def timed(name, fn):
        # every stage call is blocking (it ends with a stream synchronise), so host time == device time + launch gaps
        t0 = time.perf_counter()
        r = fn()
        stage_s[name] = stage_s.get(name, 0.0) + time.perf_counter() - t0
        return r

    def step():
        # the same call sequence as DeBruijnExtensionIndexBuilder::BuildExtensionIndexFromStream + UnbranchingPathExtractor, once per K
        stats = {}
        for k in ks:
            index = B.DeBruijnExtensionIndex(ctx, k)
            kp = timed("count_kpomers", lambda: B.KMerDiskCounter(ctx, streams, k + 1, True, True).Count(a.buckets))
            index.kmers = timed("count_kmers", lambda: B.KMerDiskCounter(ctx, kp, k).Count(a.buckets))
            index.index = timed("mphf", lambda: B.KMerIndex(ctx, index.kmers))
            h = B.vp()
            timed("masks", lambda: ctx.check(ctx.lib.sb200_ext_build(ctx.h, kp.h, index.kmers.h, index.index.h, B.C.byref(h))))
            index.h = h
            u = B.vp()
            timed("unitigs", lambda: ctx.check(ctx.lib.sb200_unitigs_extract(ctx.h, index.kmers.h, index.index.h, index.h, 1, B.C.byref(u))))
            stats[k] = (kp.total_kmers(), kp.instances, index.size(), ctx.lib.sb200_unitigs_count(u), ctx.lib.sb200_unitigs_total_bases(u))
            ctx.lib.sb200_unitigs_free(u)
            index.free(); kp.free()
        return stats

    def barrier():
        torch.cuda.synchronize()

    for _ in range(a.warmup):
        stats = step()
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    stage_s.clear()
    ctx.kernel_launches(reset=True)
    ctx.profile(True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record(stream)
    for _ in range(a.steps):
        stats = step()
    e1.record(stream)
    barrier()
    ms = e0.elapsed_time(e1)
    launches = ctx.kernel_launches()
    report = ctx.profile_report()
    ctx.profile(False)
    clocks = sampler.stop()
    ms_per_step = ms / a.steps
    value = total_bases * len(ks) / (ms_per_step * 1e-3) / 1e9

    # ---- roofline of the dominant kernel ------------------------------------------------------------------------------
    # achieved = ALGORITHMIC bytes of one step's launches of that kernel / their measured device time.  Algorithmic bytes follow
    # SURVEY.md section 8(d)'s single-pass model: every launch reads its input once and writes its output once (DESIGN.md lists
    # the per-kernel formulas).  W1/W0 = bytes of a (k+1)-mer / k-mer record; sums run over the Ks of the step.
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured copy bandwidth (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback (B200_PROFILING.md)"
    kernel_ms = {name: (n, t) for name, n, t in report}
    alg, path_bytes = {}, 0.0

    def add(name, v):
        alg[name] = alg.get(name, 0.0) + v

    n_inst_all = 0
    for k in ks:
        n_kp, n_inst, n_km, n_unitigs, unitig_bases = stats[k]
        n_inst_all += n_inst
        W1, W0 = 8 * ((k + 1 + 31) // 32), 8 * ((k + 31) // 32)
        n_words_out = (unitig_bases + 31 * n_unitigs) / 32.0 * 8     # packed unitig bytes (upper bound on padding)
        add("extract_reads_kernel<W>", total_bases / 4 + n_inst * W1)
        # S2 + S4: instances in, unique (+ count / + mask byte) out — the group kernel
        add("group_hash_kernel_", n_inst * W1 + n_kp * (W1 + 4) + 2 * n_kp * W0 + n_km * (W0 + 1))
        add("group_chunk_kernel_", n_inst * W1 + n_kp * (W1 + 4) + 2 * n_kp * W0 + n_km * (W0 + 1))
        add("derive_kernel_", n_kp * W1 + 2 * n_kp * W0)
        # staged producer-fused grouping (csrc/staged_partition.cuh): the reads are produced twice (count, pass 1), the instances cross
        # HBM in pass 1 (write) and pass 2 (read + write)
        add("sp_count_reads_kernel_", total_bases / 4)
        add("sp_scatter_reads_kernel_", total_bases / 4 + n_inst * W1)
        add("sp_scatter_fine_kernel_", 2 * (n_inst * W1 + 2 * n_kp * W0))
        add("sp_count_derive_kernel_", n_kp * W1)
        add("sp_scatter_derive_kernel_", n_kp * W1 + 2 * n_kp * W0)
        add("mphf_level1_kernel_", n_km * W0 + 0.22 * n_km * 24)      # keys in; a state record for the keys level 0 did not place
        add("walk_emit_captured_kernel<W>", 2 * n_words_out)          # captured nucleotides in, packed unitigs out
        add("fill_masks_kernel_", n_kp * W1 + n_km)
        add("index_of_kmers_kernel<W>", n_km * (W0 + 8 + 2))
        add("index_from_place_kernel", n_km * (4 + 8 + 2))            # placement in; idx, inv, mask byte (read + write) out
        add("links_kernel<W>", n_km * (W0 + 4 + 1 + 8))                # k-mer, idx, mask in; two link words out
        add("walk_measure_links_kernel<W>", 2 * n_km * 4)               # every link word is read once
        add("walk_emit_links_kernel<W>", n_km * 4 + n_words_out)
        add("walk_measure_kernel<W>", n_km * (W0 + 1))
        add("walk_emit_kernel<W>", n_km * 1 + n_words_out)
        add("mphf_level0_kernel<W>", n_km * W0 + n_km * 5.8 / 8)
        add("sort_records", n_inst * W1 + 2 * n_kp * W0)
        # whole path against the single-pass model of SURVEY.md 8(d): sum of the stage formulas S1..S7
        path_bytes += (total_bases / 4 + n_inst * W1) + (n_inst * W1 + n_kp * (W1 + 4)) + (n_kp * W1 + 2 * n_kp * (W0 + 1)) + \
                      (2 * n_kp * (W0 + 1) + n_km * (W0 + 1)) + (n_km * W0 + n_km * 5.8 / 8) + (n_km * (W0 + 1) + n_km) + \
                      (n_km * (W0 + 1) + (n_kp + k * n_unitigs) / 4)
    for name in ("rs_scatter_kernel<W>", "rs_hist_kernel<W>"):
        if name in kernel_ms:
            per_sort = kernel_ms[name][0] / a.steps / (2.0 * len(ks))   # launches per sort
            mult = 2 if "scatter" in name else 1                          # a counting pass reads and writes every record once
            alg[name] = per_sort * mult * alg["sort_records"]
    alg.pop("sort_records")
    roof = None
    top = next(((name, n, t) for name, n, t in report if name in alg), None)
    if top:
        name, n_l, t_l = top
        bytes_step = float(alg[name])
        achieved = bytes_step * a.steps / (t_l * 1e-3) / 1e9
        roof = {"bound": "hbm", "kernel": name, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": NCU_TRAFFIC.get(name), "peak_source": peak_src, "launches_per_step": n_l / a.steps,
                "mean_launch_ms": t_l / n_l, "share_of_step": t_l / ms,
                "algorithmic_bytes_per_launch": bytes_step * a.steps / n_l,
                "note": "top kernel by measured device time; traffic = dram read+write bytes per launch from profiles/ (ncu --set full), null if not captured"}
    path_roof = {"algorithmic_bytes_per_step": path_bytes, "achieved": path_bytes / (ms_per_step * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                 "frac": path_bytes / (ms_per_step * 1e-3) / 1e9 / peak}
    kernels_roof = [{"kernel": name, "ms_per_step": t / a.steps, "achieved_gbs": alg[name] * a.steps / (t * 1e-3) / 1e9,
                     "frac": alg[name] * a.steps / (t * 1e-3) / 1e9 / peak} for name, n, t in report if name in alg]
    breakdown = [{"kernel": name, "launches": n // a.steps if a.steps else n, "ms_per_step": t / a.steps} for name, n, t in report[:12]]
    if a.breakdown:
        for name, n, t in report:
            print("%-40s %6d launches %10.3f ms/step" % (name, n // a.steps, t / a.steps), file=sys.stderr)

    # ---- end to end through the C ABI with host buffers ---------------------------------------------------------------
    e2e = e2e_graph = None
    if not a.no_e2e:
        pw = torch.from_numpy(words.view(np.int64)).pin_memory()
        po = torch.from_numpy(word_off.view(np.int64)).pin_memory()
        pl = torch.from_numpy(lens.view(np.int32)).pin_memory()
        hw, ho, hl = pw.numpy().view(np.uint64), po.numpy().view(np.uint64), pl.numpy().view(np.uint32)

        def one(fetch_kmers):
            h2d = d2h = 0
            for k in ks:
                g = B.construct(ctx, hw, ho, hl, k, a.buckets, fetch_kmers=fetch_kmers)
                h2d += g.view.h2d_bytes; d2h += g.view.d2h_bytes
                g.free()
            return h2d, d2h

        def timed_e2e(fetch_kmers):
            h2d, d2h = one(fetch_kmers)   # warm-up (also sizes the pinned result pool)
            for _ in range(max(a.warmup - 1, 0)):
                one(fetch_kmers)
            barrier()
            t0 = time.perf_counter()
            e0.record(stream)
            for _ in range(a.steps):
                one(fetch_kmers)
            e1.record(stream)
            barrier()
            ms_e = max(e0.elapsed_time(e1), (time.perf_counter() - t0) * 1e3)   # host-side work (pinned pool, serialisation) is part of the call
            return {"value": total_bases * len(ks) / (ms_e / a.steps * 1e-3) / 1e9, "unit": UNIT, "h2d_bytes_per_step": int(h2d),
                    "d2h_bytes_per_step": int(d2h), "ms_per_step": ms_e / a.steps}

        # e2e = the metric as SURVEY.md 8(d) defines it: wall time from "packed reads resident in pinned host memory" to "unitig buffers +
        # index resident on host" — the extension index (KMerIndex bytes + mask array) and the packed unitigs come home, which is what
        # spades-gbuilder consumes after the path; the two k-mer tables are the reference's TEMPORARY files (it deletes them) and stay
        # on the device, where KMerDiskStorage::bucket() downloads them on demand
        e2e = timed_e2e(False)
        e2e["returns"] = "KMerIndex bytes + mask array (the extension index) and the packed unitigs: the metric's 'unitig buffers + index resident on host'"
        # everything the reference's builders ever materialise, both k-mer tables included (2.9 GB over PCIe at config 2)
        e2e_graph = timed_e2e(True)
        e2e_graph["returns"] = "additionally both k-mer tables: (k+1)-mers + counts, k-mers (the reference's temporary KMerDiskStorage files)"

    cpu = parity = None
    if not a.no_cpu_baseline:
        cpu, parity = cpu_baseline_and_parity(a, B, ctx, not a.no_parity)
    verify = verify_full_size(a, B, ctx, words, word_off, lens) if a.verify else None

    n_kp, n_inst, n_km, n_unitigs, unitig_bases = stats[ks[-1]]
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": 1, "steps": a.steps, "warmup": a.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": w["scaling"], "vs_baseline": None, "dtype": "u64",
            "data": "synthetic" if a.config != 1 else "assembler/test_dataset reads", "config": workload_config(a, 1), "clocks": clocks,
            "gpu_launches": int(launches),
            "e2e": e2e, "e2e_all_tables": e2e_graph, "roofline": roof, "cpu_baseline": cpu, "parity": parity, "verify_full_size": verify,
            "roofline_path": path_roof, "roofline_kernels": kernels_roof,
            "kmers_counted_per_s": n_inst_all / (stage_s["count_kpomers"] / a.steps),
            "stage_ms": {k_: 1e3 * v_ / a.steps for k_, v_ in stage_s.items()},
            "counts": {"kpomer_instances": int(n_inst), "kpomers": int(n_kp), "kmers": int(n_km), "unitigs": int(n_unitigs),
                       "unitig_bases": int(unitig_bases), "input_bases": total_bases, "k": int(ks[-1])},
            "breakdown": breakdown}
    print(json.dumps(line))
    return 0


if __name__ == "__main__":
    sys.exit(main())
