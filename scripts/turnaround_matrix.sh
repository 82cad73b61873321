#!/bin/bash
# A/B of the step-to-step turnaround switches: graphs x sync-spin x copies-outside-graph
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
B=${1:-32}
for cfg in "1 1 1" "1 0 1" "1 1 0" "1 0 0" "0 1 0" "0 0 0"; do
  set -- $cfg
  out=gpurun_out/turn_b${B}_g$1_s$2_o$3
  TGIS_CUDA_GRAPHS=$1 TGIS_SYNC_SPIN=$2 TGIS_GRAPH_COPY_OUTSIDE=$3 timeout 300 python bench.py --no-cpu-baseline --steps 2 --warmup 2 --batch $B > $out.log 2> $out.err
  python - <<PY
import json
try:
    d = json.loads(open("$out.log").read().strip().split("\n")[-1])
    print("B=$B graphs=$1 spin=$2 copies_outside=$3: %.0f tok/s, %.3f ms/step, e2e %.0f, ttft %.0f ms" % (d["value"], d["roofline"]["decode_step_ms"], d["e2e"]["value"], d["ttft_p50_ms"]))
except Exception as e:
    print("B=$B $cfg failed", e)
PY
done
