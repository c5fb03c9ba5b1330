#!/bin/bash
# One GPU-box pass: parity tests, the N=1 bench line, ncu launch list and full captures of the top kernels.
# Usage (from the repo root, via gpurun): bash tools/gpu_round.sh <tag> [skip-tests|tests-only] [pytest args]
tag=${1:-r1}
mode=$2
mkdir -p gpurun_out
if [ "$mode" != "skip-tests" ]; then
  timeout 1200 python -m pytest tests -m gpu -x -q ${3} > gpurun_out/pytest_$tag.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_$tag.log
  tail -3 gpurun_out/pytest_$tag.log
fi
[ "$mode" == "tests-only" ] && exit 0
timeout 900 python bench.py --steps 5 --warmup 3 --breakdown > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err; echo "bench rc=$?"
tail -c 600 gpurun_out/bench_$tag.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/launches_$tag.csv \
  python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/ncu_launch_$tag.log 2>&1; echo "ncu launches rc=$?"
# full capture: the first launches of the heavy kernels of one step (report kept under 45 MB: gpurun_out/ is capped at 64 MiB)
REP=/tmp/full_$tag
timeout 900 ncu --set full --clock-control none \
  -k regex:"${NCU_KERNELS:-group_hash_kernel|group_chunk_kernel|rs_scatter_kernel|links_kernel|walk_measure_links_kernel|walk_emit_links_kernel|extract_reads_kernel|index_of_kmers_kernel|index_from_place_kernel|derive_kernel|mphf_level0_kernel|rs_hist_kernel}" \
  -c ${NCU_COUNT:-20} -o $REP -f python bench.py --steps 1 --warmup 0 --no-e2e --no-cpu-baseline > gpurun_out/ncu_full_$tag.log 2>&1; echo "ncu full rc=$?"
ncu -i $REP.ncu-rep --page raw --csv > gpurun_out/full_${tag}_raw.csv 2>/dev/null
ncu -i $REP.ncu-rep --page details --csv > gpurun_out/full_${tag}_details.csv 2>/dev/null
sz=$(stat -c %s $REP.ncu-rep 2>/dev/null || echo 0)
if [ "$sz" -gt 0 ] && [ "$sz" -lt 45000000 ]; then cp $REP.ncu-rep gpurun_out/; else echo "ncu-rep of $sz bytes left on the box"; fi
du -sh gpurun_out; ls -la gpurun_out | tail -12
