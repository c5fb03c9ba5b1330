#!/usr/bin/env python
"""md5 digests of what the UNMODIFIED reference (oracle/_ref/ref_driver) produces for the BASELINE workloads at FULL size, so that the
`-m gpu` tests and bench.py --verify compare the CUDA path with the reference itself at the sizes the bench is quoted on — not with
another of this repo's own code paths.  Run in the build container (needs /root/reference compiled: `make -C oracle ref`); only the
digests travel (tests/golden/fullsize_digests.json), the generator of the reads is deterministic (host/synth.py workload_reads).

    python tests/golden/make_fullsize_digests.py 2 4 [5:2000000]      # config[:max_reads]
"""
import hashlib
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from spades_for_blackbird_b200.host import synth  # noqa: E402

DRIVER = os.path.join(ROOT, "oracle", "_ref", "ref_driver")
OUT = os.path.join(HERE, "fullsize_digests.json")


def md5_files(paths):
    h = hashlib.md5()
    for p in paths:
        with open(p, "rb") as f:
            while True:
                b = f.read(1 << 24)
                if not b:
                    break
                h.update(b)
    return h.hexdigest()


def md5_arr(a):
    return hashlib.md5(np.ascontiguousarray(a).tobytes()).hexdigest()


def digest(config, k, buckets, max_reads, threads):
    tmp = tempfile.mkdtemp(prefix="sb200_full_", dir=os.environ.get("SB200_TMP", "/tmp"))
    try:
        rp = os.path.join(tmp, "reads.txt")
        n_reads = 0
        with open(rp, "w") as f:
            for codes in synth.workload_chunks(config):
                if max_reads is not None and n_reads + len(codes) > max_reads:
                    codes = codes[:max_reads - n_reads]
                f.write("\n".join(synth.codes_to_strings(codes)) + "\n")
                n_reads += len(codes)
                if max_reads is not None and n_reads >= max_reads:
                    break
        out = os.path.join(tmp, "out")
        t0 = time.time()
        subprocess.check_call([DRIVER, "--mode", "gbuilder", "--reads", rp, "--out", out, "-k", str(k), "-t", str(threads), "--buckets", str(buckets),
                               "--coverage", "--quiet"], stdout=subprocess.DEVNULL)
        secs = time.time() - t0
        d = {"config": config, "k": k, "buckets": buckets, "n_reads": n_reads, "reference_seconds_here": round(secs, 1), "threads": threads}
        d["kpomers_md5"] = md5_files([os.path.join(out, "kpomers.%d" % b) for b in range(buckets)])
        d["kp_bucket_sizes_md5"] = md5_files([os.path.join(out, "kpomer_bucket_sizes.u64")])
        d["coverage_md5"] = md5_files([os.path.join(out, "coverage.u32")])
        d["kmers_md5"] = md5_files([os.path.join(out, "final_kmers")])
        d["masks_idx_md5"] = md5_files([os.path.join(out, "masks_idx.u8")])
        d["index_bin_md5"] = md5_files([os.path.join(out, "index.bin")])
        d["n_kpomers"] = os.path.getsize(os.path.join(out, "coverage.u32")) // 4
        d["n_kmers"] = os.path.getsize(os.path.join(out, "masks_idx.u8"))
        with open(os.path.join(out, "unitigs.txt"), "rb") as f:
            words, word_off, lens = synth.pack_text_sequences(f.read())
        d["n_unitigs"] = int(len(lens))
        d["unitig_bases"] = int(lens.astype(np.int64).sum())
        d["unitig_words_md5"] = md5_arr(words)
        d["unitig_len_md5"] = md5_arr(lens)
        return d
    finally:
        subprocess.call(["rm", "-rf", tmp])


def main():
    alld = json.load(open(OUT)) if os.path.exists(OUT) else {}
    for spec in sys.argv[1:]:
        config, _, mr = spec.partition(":")
        config = int(config)
        max_reads = int(mr) if mr else None
        w = synth.WORKLOADS[config]
        for k in w["ks"]:
            key = "config%d_k%d%s" % (config, k, "_first%d" % max_reads if max_reads else "")
            alld[key] = digest(config, k, 80, max_reads, os.cpu_count() or 1)
            print(key, json.dumps(alld[key]))
            json.dump(alld, open(OUT, "w"), indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
