#!/usr/bin/env python
"""Top source lines by warp-stall samples from an `ncu --page source --csv` export (tools/ncu_source.sh).

    python tools/ncu_source_summary.py gpurun_out/src_<tag>.csv profiles/<name>.md "<title>"
"""
import collections
import csv
import sys


def main():
    src, dst, title = sys.argv[1], sys.argv[2], sys.argv[3]
    rows = list(csv.reader(open(src)))
    cur, hdr, func = None, None, ""
    agg = collections.defaultdict(lambda: [0, 0, "", collections.Counter()])
    total = 0
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            cur = r[1]
            continue
        if r[0] == "Function Name":
            func = r[1]
            continue
        if r[0] == "Line No":
            hdr = r
            continue
        if not hdr or len(r) < 8:
            continue
        d = dict(zip(hdr, r))
        try:
            ln, samp, inst = int(d["Line No"]), int(d["# Samples"] or 0), int(d["Instructions Executed"] or 0)
        except Exception:
            continue
        key = (cur.split("/")[-1], ln)
        a = agg[key]
        a[0] += samp
        a[1] += inst
        a[2] = r[1]
        for k, v in d.items():
            if k.startswith("stall_") and "Not Issued" not in k and v and v != "0":
                try:
                    a[3][k[6:]] += int(v)
                except Exception:
                    pass
        total += samp
    out = ["# %s" % title, "", "`%s`" % func, "", "%d warp-stall samples (`ncu --set full --import-source on`, source page)." % total, "",
           "| samples | warp insts | file:line | source | top stall reasons |", "|---:|---:|---|---|---|"]
    for (f, ln), (s, inst, text, st) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:30]:
        out.append("| %.1f %% | %d | %s:%d | `%s` | %s |" % (100.0 * s / max(total, 1), inst, f, ln, text.strip().replace("|", "\\|")[:100],
                                                        ", ".join("%s %d" % kv for kv in st.most_common(3))))
    open(dst, "w").write("\n".join(out) + "\n")
    print("\n".join(out[:20]))


if __name__ == "__main__":
    main()
