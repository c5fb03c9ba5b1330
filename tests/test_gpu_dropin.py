"""The drop-in at the reference's OWN types (include/sb200_spades.hpp), checked by oracle/_ref/ref_dropin: a binary built in the build
container against the unmodified SPAdes headers + libspades_b200.so (oracle/Makefile `make dropin`).  It runs the reference's CPU
path and the GPU path side by side and lets the REFERENCE's own code consume the GPU's results: KMerCounter::Count bucket files,
ConstructKWH(kmer).idx() on the GPU-built KMerIndex, UnbranchingPathExtractor and FastGraphFromSequencesConstructor + GFAWriter on
the GPU-filled DeBruijnExtensionIndex, and the edge index's KMerDiskCounter over DeBruijnGraphKMerSplitter against the GPU counter on the
same edges (see the header of oracle/ref_dropin.cpp for checks A-E)."""
import os
import subprocess

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "oracle", "_ref", "ref_dropin")


@pytest.mark.parametrize("name,threads", [("ecoli1k_k21", 2), ("ecoli1k_k55", 1), ("multiword_k127", 1), ("loops_k21", 2), ("ctest_SplitThread2", 1)])
def test_reference_types_filled_from_the_gpu(tmp_path, name, threads):
    if not os.path.exists(BIN):
        pytest.skip("oracle/_ref/ref_dropin is not built (no /root/reference at build time)")
    from conftest import load_golden
    g = load_golden(name)
    rp = tmp_path / "reads.txt"
    rp.write_text("\n".join(g["reads"]) + "\n")
    r = subprocess.run([BIN, "--reads", str(rp), "--out", str(tmp_path / "out"), "-k", str(g["k"]), "-t", str(threads), "--buckets", str(g["buckets"])],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
    assert "DROPIN OK" in r.stdout
    for check in ("A_counter_bucket_files", "A_counter_final_kmers_after_merge", "B_kpomer_storage", "B_reference_lookup_on_gpu_index", "B_masks",
                  "B_kmer_index_serialize", "C_reference_extractor_on_gpu_index", "C_gpu_unitigs", "D_gfa_from_gpu_index_and_unitigs", "E_edge_index_counter_all_kmers",
                  "E_edge_index_counter_minimal_kmers"):
        assert check + " OK" in r.stdout, r.stdout
