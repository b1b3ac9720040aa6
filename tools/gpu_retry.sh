#!/usr/bin/env bash
# Retries a gpurun call while the pod answers "transient" (no box/slot free; nothing charged).
# usage: tools/gpu_retry.sh <timeout_s> '<command>' [max_tries]
T="$1"; CMD="$2"; MAX="${3:-15}"
for i in $(seq 1 "$MAX"); do
  OUT="$(/usr/local/graft/bin/gpurun --timeout "$T" -- "$CMD" 2>&1)"
  if echo "$OUT" | grep -q "status=transient"; then
    echo "[gpu_retry] try $i: transient, sleeping 150 s"; sleep 150; continue
  fi
  echo "$OUT" | tail -120
  exit 0
done
echo "[gpu_retry] gave up after $MAX tries"; exit 3
