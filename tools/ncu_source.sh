#!/bin/bash
# Source-level ncu capture of ONE kernel: bash tools/ncu_source.sh <tag> <kernel regex> [launch count]
tag=$1; kern=$2; cnt=${3:-1}
REP=/tmp/src_$tag
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"$kern" -c $cnt -o $REP -f \
  python bench.py --steps 1 --warmup 0 --no-e2e --no-cpu-baseline > gpurun_out/ncu_src_$tag.log 2>&1; echo "ncu rc=$?"
ncu -i $REP.ncu-rep --page source --csv --print-source cuda,sass > gpurun_out/src_${tag}.csv 2>/dev/null || ncu -i $REP.ncu-rep --page source --csv > gpurun_out/src_${tag}.csv
ncu -i $REP.ncu-rep --page raw --csv > gpurun_out/src_${tag}_raw.csv
ls -la gpurun_out/src_${tag}*
