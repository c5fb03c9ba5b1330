#!/bin/bash
# bench under different environment settings: bash tools/try_env.sh "A=1" "B=2 C=3" ...   ("-" = no extra variables)
i=0
for e in "$@"; do
  i=$((i+1))
  [ "$e" == "-" ] && e=""
  env $e timeout 600 python bench.py --steps 4 --warmup 2 --breakdown --no-cpu-baseline --no-e2e > gpurun_out/env_$i.json 2> gpurun_out/env_$i.err
  echo "== [$e] rc=$? $(python -c "import json;d=json.loads(open('gpurun_out/env_$i.json').read().strip().splitlines()[-1]);print('ms_per_step',d['ms_per_step'], d['stage_ms'])")"
  head -${LINES_SHOWN:-12} gpurun_out/env_$i.err
done
