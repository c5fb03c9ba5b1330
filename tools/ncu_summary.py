#!/usr/bin/env python
"""Turn the ncu CSV exports a GPU run left in gpurun_out/ into the tracked summaries under profiles/.

    python tools/ncu_summary.py <tag> <round-name>      e.g.  python tools/ncu_summary.py r1c r1a

Writes profiles/<round>_launches.csv (the `--metrics gpu__time_duration.sum` launch list as captured),
profiles/<round>_ncu_summary.md (per-kernel totals of the launch list + the `--set full` metrics of the captured launches)
and profiles/ncu_traffic.json (dram read+write bytes per launch, keyed by the kernel names of sb200_profile_report(),
which bench.py copies into roofline.traffic)."""
import collections
import csv
import json
import os
import re
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "gpurun_out")
PROF = os.path.join(ROOT, "profiles")


def base(name):
    name = name.replace("void ", "")
    return re.split(r"[<(]", name)[0].strip().rstrip("_")


def to_f(x):
    try:
        return float(x.replace(",", ""))
    except Exception:
        return float("nan")


def main():
    tag, rnd = sys.argv[1], sys.argv[2]
    os.makedirs(PROF, exist_ok=True)
    md = ["# ncu summary, %s (B200, `bench.py --steps 1 --warmup 1|0 --no-e2e --no-cpu-baseline`, workload = BASELINE configs[1])" % rnd, ""]
    # ---- launch list -----------------------------------------------------------------------------------------------
    src = os.path.join(OUT, "launches_%s.csv" % tag)
    lines = [l for l in open(src) if l.startswith('"')]
    with open(os.path.join(PROF, "%s_launches.csv" % rnd), "w") as f:
        f.writelines(lines)
    rows = list(csv.DictReader(lines))
    tot = collections.OrderedDict()
    for r in rows:
        b = base(r["Kernel Name"])
        n, t = tot.get(b, (0, 0.0))
        tot[b] = (n + 1, t + to_f(r["Metric Value"]) / 1e6)
    total = sum(t for _, t in tot.values())
    md += ["## Launch list (`ncu --metrics gpu__time_duration.sum --clock-control none`): %d launches, %.2f ms of kernel time over 2 passes of the path"
           % (len(rows), total), "", "Per-launch times under ncu are serialised and cold-cache; the SHARE of each kernel is what is comparable with",
           "bench.py's live CUDA-event breakdown.", "", "| kernel | launches | total ms | share |", "|---|---:|---:|---:|"]
    for b, (n, t) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
        md.append("| %s | %d | %.3f | %.1f %% |" % (b, n, t, 100 * t / total))
    md.append("")
    # ---- full capture ------------------------------------------------------------------------------------------------
    raw = os.path.join(OUT, "full_%s_raw.csv" % tag)
    traffic = {}
    if os.path.exists(raw):
        rr = list(csv.reader(open(raw)))
        hdr, units = rr[0], rr[1]
        ix = {h: i for i, h in enumerate(hdr)}
        stall = [h for h in hdr if h.startswith("smsp__pcsamp_warps_issue_stalled_") and "not_issued" not in h]
        cols = [("gpu__time_duration.sum", "time"), ("dram__bytes_read.sum", "dram rd"), ("dram__bytes_write.sum", "dram wr"),
                ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram %"), ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm %"),
                ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1 %"), ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2 %"),
                ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ %"), ("launch__registers_per_thread", "regs"),
                ("smsp__inst_executed.sum", "warp insts")]
        cols = [c for c in cols if c[0] in ix]
        md += ["## `ncu --set full --clock-control none` captures (one row per captured launch)", "",
               "| kernel | grid x block | " + " | ".join(c[1] for c in cols) + " | top stall reasons (pc samples) |",
               "|---|---|" + "---:|" * len(cols) + "---|"]
        per = collections.defaultdict(list)
        for r in rr[2:]:
            name = r[ix["Kernel Name"]]
            b = base(name)
            tpl = re.search(r"<([^>]*)>", name)
            vals = []
            for c, _ in cols:
                v, u = r[ix[c]], units[ix[c]]
                vals.append("%s %s" % (v, u) if u and u not in ("%", "inst", "register/thread") else v)
            st = sorted(((to_f(r[ix[h]]), h.replace("smsp__pcsamp_warps_issue_stalled_", "")) for h in stall), reverse=True)
            ssum = sum(x for x, _ in st) or 1.0
            tops = ", ".join("%s %.0f%%" % (h, 100 * x / ssum) for x, h in st[:4])
            md.append("| %s%s | %s x %s | %s | %s |" % (b, "<%s>" % tpl.group(1) if tpl else "", r[ix["launch__grid_size"]], r[ix["launch__block_size"]],
                                                      " | ".join(vals), tops))

            def gb(col):
                v, u = to_f(r[ix[col]]), units[ix[col]]
                return v * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}.get(u, 1.0)
            per[b].append(gb("dram__bytes_read.sum") + gb("dram__bytes_write.sum"))
        md.append("")
        for b, v in per.items():
            traffic[b] = sum(v) / len(v)
    # keys as sb200_profile_report() prints them (the LAUNCH macro's kernel expression)
    names = {"seg_chunk_kernel": "seg_chunk_kernel_", "group_chunk_kernel": "seg_chunk_kernel_", "group_hash_kernel": "group_hash_kernel_", "sp_scatter_fine_kernel": "sp_scatter_fine_kernel_", "sp_scatter_reads_kernel": "sp_scatter_reads_kernel_", "sp_count_reads_kernel": "sp_count_reads_kernel_", "sp_scatter_derive_kernel": "sp_scatter_derive_kernel_", "sp_count_derive_kernel": "sp_count_derive_kernel_", "mphf_level1_kernel": "mphf_level1_kernel_",
             "links_kernel": "links_kernel<W>", "index_from_place_kernel": "index_from_place_kernel",
             "walk_measure_links_kernel": "walk_measure_links_kernel<W>", "walk_emit_links_kernel": "walk_emit_links_kernel<W>", "rs_scatter_kernel": "rs_scatter_kernel<W>", "rs_hist_kernel": "rs_hist_kernel<W>",
             "walk_measure_kernel": "walk_measure_kernel<W>", "walk_emit_kernel": "walk_emit_kernel<W>", "fill_masks_kernel": "fill_masks_kernel_",
             "extract_reads_kernel": "extract_reads_kernel<W>", "derive_kernel": "derive_kernel_", "index_of_kmers_kernel": "index_of_kmers_kernel<W>",
             "mphf_level0_kernel": "mphf_level0_kernel<W>", "seg_heads_kernel": "seg_heads_kernel<W>"}
    out = {}
    for b, v in traffic.items():
        out[names.get(b, b)] = v
        out[b] = v
    if out:
        json.dump(out, open(os.path.join(PROF, "ncu_traffic.json"), "w"), indent=1, sort_keys=True)
    extra = os.path.join(OUT, "bench_%s.json" % tag)
    if os.path.exists(extra):
        shutil.copy(extra, os.path.join(PROF, "%s_bench_line.json" % rnd))
        md += ["## bench.py line of the same build (not under ncu)", "", "See `%s_bench_line.json`." % rnd, ""]
    open(os.path.join(PROF, "%s_ncu_summary.md" % rnd), "w").write("\n".join(md) + "\n")
    print("\n".join(md[:40]))


if __name__ == "__main__":
    main()
