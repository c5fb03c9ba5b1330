#!/bin/bash
# bench each prebuilt library variant lib_<tag>.so (kept at the repo root): bash tools/try_libs.sh a b c
cp spades_for_blackbird_b200/libspades_b200.so /tmp/lib_orig.so
for tag in "$@"; do
  cp lib_$tag.so spades_for_blackbird_b200/libspades_b200.so
  timeout 600 python bench.py --steps 4 --warmup 2 --breakdown --no-cpu-baseline --no-e2e > gpurun_out/try_$tag.json 2> gpurun_out/try_$tag.err
  echo "== $tag rc=$? $(python -c "import json;d=json.loads(open('gpurun_out/try_$tag.json').read().strip().splitlines()[-1]);print('ms_per_step',d['ms_per_step'], d['stage_ms'])")"
  head -${LINES_SHOWN:-8} gpurun_out/try_$tag.err
done
cp /tmp/lib_orig.so spades_for_blackbird_b200/libspades_b200.so
