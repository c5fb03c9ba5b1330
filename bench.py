#!/usr/bin/env python
"""bench.py — reads -> condensed de Bruijn graph throughput (BASELINE.json metric) on N B200s of one node.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]

Workload (config.workload): BASELINE.json configs[1] — synthetic 4.6 Mbp genome, 2x150 paired reads at 100x
(1 533 333 pairs, 460 Mbp), 0.5 % substitutions, k = 55, 80 hash buckets (= the reference's 10 x 8 threads).
A step = one pass of the whole hot path over that read set:
    packed reads -> canonical (k+1)-mers -> sort/dedup(+counts) -> k-mers -> BooPHF MPHF -> extension masks -> unitigs
`value`  : input bases / device time with the packed reads already resident in HBM (CUDA events on the library's stream)
`e2e`    : the same through sb200_construct() with pinned HOST buffers on both sides (H2D of the reads and D2H of
           (k+1)-mers + counts, k-mers, masks, MPHF and unitigs inside the timed region)
`e2e_graph_only`: the same call without the two k-mer tables (masks, index and unitigs come home; the tables stay on the device)
`roofline`: the kernel with the largest measured share of the step — algorithmic bytes per launch / its mean launch duration,
           measured live with CUDA events around every launch of the timed steps
`cpu_baseline` / --impl reference: the UNMODIFIED reference (oracle/_ref/ref_driver, compiled from /root/reference)
           on the box's host cores over a bounded sample of the same read set.
"""
import argparse
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
    NCU_TRAFFIC = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "profiles", "ncu_traffic.json")))
except Exception:
    pass
UNIT = "Gbp/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--genome-len", type=int, default=4_600_000)
    ap.add_argument("--coverage", type=float, default=100.0)
    ap.add_argument("--read-len", type=int, default=150)
    ap.add_argument("--k", type=int, default=55)
    ap.add_argument("--buckets", type=int, default=80)
    ap.add_argument("--cpu-sample-reads", type=int, default=3_000_000,
                    help="reads of the workload the reference CPU path is timed on (about 10-20 s of CPU work per pass)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--breakdown", action="store_true", help="print the per-kernel timing table to stderr")
    return ap.parse_args()


def workload_config(a, n_gpus):
    return {"workload": "synthetic isolate %.1f Mbp genome, 2x%d reads at %.0fx, k=%d (BASELINE configs[1])" % (
                a.genome_len / 1e6, a.read_len, a.coverage, a.k),
            "k": a.k, "num_buckets": a.buckets, "genome_len": a.genome_len, "read_len": a.read_len, "coverage": a.coverage,
            "error_rate": 0.005, "seed": 42, "reads_per_gpu": "all" if n_gpus == 1 else "1/%d" % n_gpus,
            "l2": "inputs and every intermediate are larger than L2 (>= 115 MB reads, 4.7 GB of k-mer instances)"}


def make_reads(a, codes_only=False):
    n_pairs = int(a.genome_len * a.coverage / (2 * a.read_len))
    g = synth.random_genome(a.genome_len, 42)
    chunks_w = []
    chunk = 200_000
    first_codes = None
    for s in range(0, n_pairs, chunk):
        c = synth.sample_pairs(g, min(chunk, n_pairs - s), a.read_len, 350 if a.read_len <= 150 else 500, 0.005, 1042 + s)
        if first_codes is None:
            first_codes = c
        chunks_w.append(synth.pack_codes(c)[0])
    words = np.concatenate(chunks_w)
    n = 2 * n_pairs
    wpr = (a.read_len + 31) // 32
    word_off = np.arange(n + 1, dtype=np.uint64) * np.uint64(wpr)
    lens = np.full(n, a.read_len, dtype=np.uint32)
    return words, word_off, lens, first_codes


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
    same chunk seeds as make_reads), written once as plain sequences for oracle/_ref/ref_driver."""

    def __init__(self, a, n_reads):
        self.a = a
        self.driver = os.path.join(ROOT, "oracle", "_ref", "ref_driver")
        self.tmp = None
        self.n_reads = 0
        if not os.path.exists(self.driver):
            return
        n_pairs_all = int(a.genome_len * a.coverage / (2 * a.read_len))
        n_pairs = min((n_reads + 1) // 2, n_pairs_all)
        g = synth.random_genome(a.genome_len, 42)
        self.tmp = tempfile.mkdtemp(prefix="sb200_bench_")
        self.reads_path = os.path.join(self.tmp, "reads.txt")
        with open(self.reads_path, "w") as f:
            for s in range(0, n_pairs, 200_000):       # the chunking and seeds of make_reads: chunk c of the sample IS chunk c of the workload
                c = synth.sample_pairs(g, min(200_000, n_pairs_all - s), a.read_len, 350 if a.read_len <= 150 else 500, 0.005, 1042 + s)
                c = c[:2 * (n_pairs - s)]
                f.write("\n".join(synth.codes_to_strings(c)) + "\n")
                self.n_reads += len(c)

    def available(self):
        return self.tmp is not None

    def run(self, cores):
        """One timed pass of the unmodified reference path (KMerDiskCounter -> ExtensionIndex -> UnbranchingPathExtractor) with
        `cores` threads.  Returns (Gbp/s, seconds, bases); read parsing and output are outside the driver's `path_total`."""
        out = os.path.join(self.tmp, "out")
        subprocess.call(["rm", "-rf", out])
        subprocess.check_call([self.driver, "--mode", "gbuilder", "--reads", self.reads_path, "--out", out, "-k", str(self.a.k), "-t", str(cores),
                               "--quiet", "--no-dump"], stdout=subprocess.DEVNULL)
        t = {}
        for line in open(os.path.join(out, "timing.txt")):
            p = line.split()
            t[p[0]] = float(p[1])
        secs = t["path_total"]
        bases = int(t["bases"])
        return bases / secs / 1e9, secs, bases

    def close(self):
        if self.tmp:
            subprocess.call(["rm", "-rf", self.tmp])
            self.tmp = None


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
    sample = "first %d reads (%.1f Mbp) of the workload; whole reference path incl. its temp-file I/O, %.1f s per step, -t %d" % (
        smp.n_reads, bases / 1e6, float(np.mean(secs_all)), cores)
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup,
            "ms_per_step": 1e3 * float(np.mean(secs_all)), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u64", "data": "synthetic", "config": workload_config(a, 1),
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "reference", "sample": sample},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))
    return 0


def main_sharded(a, world, rank, local_rank):
    """N > 1: one rank per GPU, k-mer space sharded by hash bucket (host/distributed.py).  Weak scaling: the genome grows
    with N (4.6 Mbp x N) at fixed coverage, so every rank brings the same 460 Mbp of reads as the N = 1 workload."""
    import torch
    import torch.distributed as dist
    from spades_for_blackbird_b200.host import binding as B
    from spades_for_blackbird_b200.host import distributed as D

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist.init_process_group("nccl", device_id=dev)
    if a.buckets % world:
        raise SystemExit("--buckets must be a multiple of the number of GPUs")
    # this rank's slice of the read set: pairs sampled from the shared genome with rank-specific seeds
    g = synth.random_genome(a.genome_len * world, 42)
    n_pairs = int(a.genome_len * a.coverage / (2 * a.read_len))
    chunks = []
    for s in range(0, n_pairs, 200_000):
        c = synth.sample_pairs(g, min(200_000, n_pairs - s), a.read_len, 350 if a.read_len <= 150 else 500, 0.005,
                               1042 + s + 7_000_003 * rank)
        chunks.append(synth.pack_codes(c)[0])
    words = np.concatenate(chunks)
    n = 2 * n_pairs
    wpr = (a.read_len + 31) // 32
    word_off = np.arange(n + 1, dtype=np.uint64) * np.uint64(wpr)
    lens = np.full(n, a.read_len, dtype=np.uint32)
    pw = torch.from_numpy(words.view(np.int64)).pin_memory()
    hw = pw.numpy().view(np.uint64)
    total_bases = int(lens.astype(np.int64).sum()) * world

    ctx = B.Context(local_rank)
    comm = D.TorchComm()
    backend = D.GpuShardBackend(ctx, dev)
    streams = B.ReadStreams(ctx, hw, word_off, lens)
    info = {}
    stage_acc = {}
    pinned_bufs = {}

    def pinned(name, nbytes):   # grow-only pinned host buffers, reused by every step
        t = pinned_bufs.get(name)
        if t is None or t.numel() < nbytes:
            t = torch.empty(int(nbytes * 1.05) + 64, dtype=torch.uint8).pin_memory()
            pinned_bufs[name] = t
        return t

    def step(e2e):
        rs = B.ReadStreams(ctx, hw, word_off, lens) if e2e else streams      # e2e: H2D of the reads inside the timed region
        res = D.construct_sharded(backend, comm, rs, a.k, a.buckets, gather_to=0)
        if not e2e:
            for k_, v_ in res.stage_ms.items():
                stage_acc[k_] = stage_acc.get(k_, 0.0) + v_
        info.update(kpomers=res.kpomers.total_kmers(), instances=res.kpomers.instances, kmers=res.kmers.total_kmers(),
                    unitigs=int(res.stats[:, 3].sum()), unitig_bases=int(res.stats[:, 4].sum()))
        d2h = 0
        if e2e:   # every rank brings its shard of the tables home (pinned buffers); rank 0 also the masks and the gathered unitigs
            kp, km = res.kpomers, res.kmers
            hp = pinned("kp", kp.total_kmers() * kp.words * 8)
            hc = pinned("kc", kp.total_kmers() * 4)
            hk = pinned("km", km.total_kmers() * km.words * 8)
            ctx.check(ctx.lib.sb200_kmers_download(kp.h, 0, kp.total_kmers(), B.C.cast(hp.data_ptr(), B.u64p)))
            ctx.check(ctx.lib.sb200_kmers_counts_download(kp.h, 0, kp.total_kmers(), B.C.cast(hc.data_ptr(), B.u32p)))
            ctx.check(ctx.lib.sb200_kmers_download(km.h, 0, km.total_kmers(), B.C.cast(hk.data_ptr(), B.u64p)))
            d2h = kp.total_kmers() * (kp.words * 8 + 4) + km.total_kmers() * km.words * 8
            if rank == 0:
                m = backend.ext_masks(res.ext)
                parts = [m] + [t for lst in res.gathered for t in lst]
                for i, t in enumerate(parts):
                    nb = t.numel() * t.element_size()
                    pinned("g%d" % i, nb)[:nb].copy_(t.contiguous().view(torch.uint8).reshape(-1), non_blocking=True)
                    d2h += nb
                torch.cuda.synchronize()
        res.kpomers.free(); res.kmers.free(); res.index.free()
        backend.free_ext(res.ext); backend.free_unitigs(res.unitigs)
        if e2e:
            rs.free()
        return d2h

    def barrier():
        dist.barrier()
        torch.cuda.synchronize()

    def timed(e2e):
        for _ in range(a.warmup):
            step(e2e)
        stage_acc.clear()
        barrier()
        t0 = time.perf_counter()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        d2h = 0
        for _ in range(a.steps):
            d2h = step(e2e)
        torch.cuda.synchronize()
        e1.record()
        barrier()
        ms = max(e0.elapsed_time(e1), 0.0)
        ms = max(ms, (time.perf_counter() - t0) * 1e3 - 1.0) if ms == 0.0 else ms
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()) / a.steps, d2h

    sampler = ClockSampler(local_rank)
    sampler.start()
    ctx.kernel_launches(reset=True)
    ms_per_step, _ = timed(False)
    stage_ms = {k_: v_ / a.steps for k_, v_ in stage_acc.items()}
    launches = ctx.kernel_launches()
    clocks = sampler.stop()
    ms_e2e, d2h = (None, 0) if a.no_e2e else timed(True)
    if rank == 0:
        cfg = workload_config(a, world)
        cfg["workload"] += "; N>1: genome %.1f Mbp x %d, every rank brings 460 Mbp of reads, k-mer space sharded by hash bucket" % (
            a.genome_len / 1e6, world)
        line = {"metric": METRIC, "value": total_bases / (ms_per_step * 1e-3) / 1e9, "unit": UNIT, "n_gpus": world, "steps": a.steps,
                "warmup": a.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "u64", "data": "synthetic", "config": cfg, "clocks": clocks, "gpu_launches": int(launches),
                "e2e": None if ms_e2e is None else {"value": total_bases / (ms_e2e * 1e-3) / 1e9, "unit": UNIT,
                                                    "h2d_bytes_per_step": int(words.nbytes + word_off.nbytes + lens.nbytes),
                                                    "d2h_bytes_per_step": int(d2h), "ms_per_step": ms_e2e,
                                                    "returns": "per rank: its shard of (k+1)-mers + counts and k-mers; rank 0: masks + all unitigs"},
                "roofline": None, "cpu_baseline": None, "stage_ms_rank0": stage_ms,
                "counts": {k_: int(v_) for k_, v_ in info.items()},
                "exchange": "2 x NCCL all-to-all (k-mer instances, k-mer candidates), all-reduce of MPHF bit-vectors and masks, gather of unitigs"}
        print(json.dumps(line))
    dist.destroy_process_group()
    return 0


def main():
    a = parse_args()
    if a.impl == "reference":
        return reference_arm(a)
    if int(os.environ.get("WORLD_SIZE", "1")) > 1:
        return main_sharded(a, int(os.environ["WORLD_SIZE"]), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")))

    import torch
    import torch.distributed as dist
    from spades_for_blackbird_b200.host import binding as B

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)

    words, word_off, lens, first_codes = make_reads(a)
    if world > 1:   # weak scaling: every rank works on its own read set of the full shape (replicas of the workload)
        pass
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
        # the same call sequence as DeBruijnExtensionIndexBuilder::BuildExtensionIndexFromStream + UnbranchingPathExtractor
        index = B.DeBruijnExtensionIndex(ctx, a.k)
        kp = timed("count_kpomers", lambda: B.KMerDiskCounter(ctx, streams, a.k + 1, True, True).Count(a.buckets))
        index.kmers = timed("count_kmers", lambda: B.KMerDiskCounter(ctx, kp, a.k).Count(a.buckets))
        index.index = timed("mphf", lambda: B.KMerIndex(ctx, index.kmers))
        h = B.vp()
        timed("masks", lambda: ctx.check(ctx.lib.sb200_ext_build(ctx.h, kp.h, index.kmers.h, index.index.h, B.C.byref(h))))
        index.h = h
        u = B.vp()
        timed("unitigs", lambda: ctx.check(ctx.lib.sb200_unitigs_extract(ctx.h, index.kmers.h, index.index.h, index.h, 1, B.C.byref(u))))
        stats = (kp.total_kmers(), kp.instances, index.size(), ctx.lib.sb200_unitigs_count(u), ctx.lib.sb200_unitigs_total_bases(u))
        ctx.lib.sb200_unitigs_free(u)
        index.free(); kp.free()
        return stats

    def barrier():
        if world > 1:
            dist.barrier()
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
    if world > 1:
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    ms_per_step = ms / a.steps
    value = world * total_bases / (ms_per_step * 1e-3) / 1e9

    # ---- roofline of the dominant kernel ------------------------------------------------------------------------------
    # achieved = ALGORITHMIC bytes of one step's launches of that kernel / their measured device time.  Algorithmic bytes follow
    # SURVEY.md section 8(d)'s single-pass model: every launch reads its input once and writes its output once (DESIGN.md lists
    # the per-kernel formulas).  W1/W0 = bytes of a (k+1)-mer / k-mer record.
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured copy bandwidth (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback (B200_PROFILING.md)"
    n_kp, n_inst, n_km, n_unitigs, unitig_bases = stats
    kernel_ms = {name: (n, t) for name, n, t in report}
    W1, W0 = 8 * ((a.k + 1 + 31) // 32), 8 * ((a.k + 31) // 32)
    n_words_out = (unitig_bases + 31 * n_unitigs) / 32.0 * 8     # packed unitig bytes (upper bound on padding)
    alg = {   # algorithmic bytes per STEP of every launch of the kernel
        "extract_reads_kernel<W>": total_bases / 4 + n_inst * W1,
        # S2 + S4: instances in, unique (+ count / + mask byte) out — the shared-memory group sort
        "seg_chunk_kernel_": n_inst * W1 + n_kp * (W1 + 4) + 2 * n_kp * W0 + n_km * (W0 + 1),
        "derive_kernel_": n_kp * W1 + 2 * n_kp * W0,
        "fill_masks_kernel_": n_kp * W1 + n_km,
        "index_of_kmers_kernel<W>": n_km * (W0 + 8 + 2),
        "index_from_place_kernel": n_km * (4 + 8 + 2),            # placement in; idx, inv, mask byte (read + write) out
        "links_kernel<W>": n_km * (W0 + 4 + 1 + 8),                # k-mer, idx, mask in; two link words out
        "walk_measure_links_kernel<W>": 2 * n_km * 4,               # every link word is read once
        "walk_emit_links_kernel<W>": n_km * 4 + n_words_out,
        "walk_measure_kernel<W>": n_km * (W0 + 1),
        "walk_emit_kernel<W>": n_km * 1 + n_words_out,
        "mphf_level0_kernel<W>": n_km * W0 + n_km * 5.8 / 8,
    }
    for name in ("rs_scatter_kernel<W>", "rs_hist_kernel<W>"):
        if name in kernel_ms:
            per_sort = kernel_ms[name][0] / a.steps / 2.0        # launches per sort
            mult = 2 if "scatter" in name else 1                   # a counting pass reads and writes every record once
            alg[name] = per_sort * mult * (n_inst * W1 + 2 * n_kp * W0)
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
    # whole path against the single-pass model of SURVEY.md 8(d): sum of the stage formulas S1..S7
    path_bytes = (total_bases / 4 + n_inst * W1) + (n_inst * W1 + n_kp * (W1 + 4)) + (n_kp * W1 + 2 * n_kp * (W0 + 1)) + \
                 (2 * n_kp * (W0 + 1) + n_km * (W0 + 1)) + (n_km * W0 + n_km * 5.8 / 8) + (n_km * (W0 + 1) + n_km) + \
                 (n_km * (W0 + 1) + (n_kp + a.k * n_unitigs) / 4)
    path_roof = {"algorithmic_bytes_per_step": path_bytes, "achieved": path_bytes / (ms_per_step * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                 "frac": path_bytes / (ms_per_step * 1e-3) / 1e9 / peak}
    kernels_roof = [{"kernel": name, "ms_per_step": t / a.steps, "achieved_gbs": alg[name] * a.steps / (t * 1e-3) / 1e9,
                     "frac": alg[name] * a.steps / (t * 1e-3) / 1e9 / peak} for name, n, t in report if name in alg]
    breakdown = [{"kernel": name, "launches": n // a.steps if a.steps else n, "ms_per_step": t / a.steps} for name, n, t in report[:12]]
    if a.breakdown and rank == 0:
        for name, n, t in report:
            print("%-40s %6d launches %10.3f ms/step" % (name, n // a.steps, t / a.steps), file=sys.stderr)

    # ---- end to end through the C ABI with host buffers ---------------------------------------------------------------
    e2e = e2e_graph = None
    if not a.no_e2e:
        pw = torch.from_numpy(words.view(np.int64)).pin_memory()
        po = torch.from_numpy(word_off.view(np.int64)).pin_memory()
        pl = torch.from_numpy(lens.view(np.int32)).pin_memory()
        hw, ho, hl = pw.numpy().view(np.uint64), po.numpy().view(np.uint64), pl.numpy().view(np.uint32)
        def timed_e2e(fetch_kmers):
            g = B.construct(ctx, hw, ho, hl, a.k, a.buckets, fetch_kmers=fetch_kmers)   # warm-up (also sizes the pinned result pool)
            h2d, d2h = g.view.h2d_bytes, g.view.d2h_bytes
            g.free()
            for _ in range(max(a.warmup - 1, 0)):
                B.construct(ctx, hw, ho, hl, a.k, a.buckets, fetch_kmers=fetch_kmers).free()
            barrier()
            t0 = time.perf_counter()
            e0.record(stream)
            for _ in range(a.steps):
                B.construct(ctx, hw, ho, hl, a.k, a.buckets, fetch_kmers=fetch_kmers).free()
            e1.record(stream)
            barrier()
            ms_e = max(e0.elapsed_time(e1), (time.perf_counter() - t0) * 1e3)   # host-side work (pinned pool, serialisation) is part of the call
            if world > 1:
                t = torch.tensor([ms_e], device=dev, dtype=torch.float64)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                ms_e = float(t.item())
            return {"value": world * total_bases / (ms_e / a.steps * 1e-3) / 1e9, "unit": UNIT, "h2d_bytes_per_step": int(h2d),
                    "d2h_bytes_per_step": int(d2h), "ms_per_step": ms_e / a.steps}

        # headline: EVERYTHING the reference's builders leave behind comes home — both k-mer tables (its KMerDiskStorage files) included
        e2e = timed_e2e(True)
        e2e["returns"] = "(k+1)-mers + counts, k-mers, masks, KMerIndex bytes, packed unitigs"
        # what spades-gbuilder itself consumes after the path (sequences + index for the link records; the k-mer tables are temporary
        # files it deletes): the tables stay on the device and KMerDiskStorage::bucket() downloads on demand
        e2e_graph = timed_e2e(False)
        e2e_graph["returns"] = "masks, KMerIndex bytes, packed unitigs (k-mer tables stay device-resident, fetched on demand)"

    cpu = None
    if rank == 0 and world == 1 and not a.no_cpu_baseline:
        cores = os.cpu_count() or 1
        smp = ReferenceSample(a, a.cpu_sample_reads)
        if smp.available():
            try:
                r = smp.run(cores)
            finally:
                smp.close()
            cpu = {"value": r[0], "unit": UNIT, "cores": cores, "kind": "reference",
                   "sample": "first %d reads (%.1f Mbp) of the workload, whole reference path in %.1f s, -t %d" % (
                       smp.n_reads, r[2] / 1e6, r[1], cores)}

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
                "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64",
                "data": "synthetic", "config": workload_config(a, world), "clocks": clocks, "gpu_launches": int(launches),
                "e2e": e2e, "e2e_graph_only": e2e_graph, "roofline": roof, "cpu_baseline": cpu, "roofline_path": path_roof, "roofline_kernels": kernels_roof,
                "kmers_counted_per_s": world * n_inst / (stage_s["count_kpomers"] / a.steps),
                "stage_ms": {k_: 1e3 * v_ / a.steps for k_, v_ in stage_s.items()},
                "counts": {"kpomer_instances": int(n_inst), "kpomers": int(n_kp), "kmers": int(n_km), "unitigs": int(n_unitigs),
                           "unitig_bases": int(unitig_bases), "input_bases": total_bases},
                "breakdown": breakdown}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
