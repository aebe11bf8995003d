#!/bin/bash
# usage: tools/gpurun_retry.sh <timeout-seconds> '<command>' [gpus]
# Retries a gpurun call while the pod answers busy (exit 3 / status=transient).
t=$1; cmd=$2; gpus=${3:-1}
for i in $(seq 1 40); do
  if [ "$gpus" = "1" ]; then
    out=$(/usr/local/graft/bin/gpurun --timeout "$t" -- "$cmd" 2>&1); rc=$?
  else
    out=$(/usr/local/graft/bin/gpurun --gpus "$gpus" --timeout "$t" -- "$cmd" 2>&1); rc=$?
  fi
  if echo "$out" | grep -q "status=transient\|nothing was charged"; then
    echo "[retry $i] busy, sleeping 90 s"; sleep 90; continue
  fi
  echo "$out"; exit $rc
done
echo "gave up"; exit 3
