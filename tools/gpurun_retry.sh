#!/bin/bash
# gpurun with retries while the pod answers "busy" (exit code 3 / status=transient): bash tools/gpurun_retry.sh <timeout> '<command>' [gpus]
t=$1; cmd=$2; gpus=${3:-1}
for i in $(seq 1 20); do
  if [ "$gpus" -gt 1 ]; then out=$(/usr/local/graft/bin/gpurun --gpus $gpus --timeout $t -- "$cmd" 2>&1); else out=$(/usr/local/graft/bin/gpurun --timeout $t -- "$cmd" 2>&1); fi
  if echo "$out" | grep -q "status=transient"; then sleep 90; continue; fi
  echo "$out"; exit 0
done
echo "$out"; exit 3
