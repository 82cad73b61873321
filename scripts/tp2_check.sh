#!/bin/bash
# 2-GPU validation of the tensor-parallel path: parity (exchange modes, graphs, prompt logprobs), bench with in-run
# parity, and exchange timing at a 70B-shaped large-T step.  Every stage has its own short timeout.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1"
i=0
for spec in "auto tiny" "auto tp8" "twoshot tiny" "twoshot tp8" "nccl tiny"; do
  set -- $spec
  i=$((i+1))
  TGIS_TP_EXCHANGE=$1 timeout 100 $TR --master-port $((29600+i)) scripts/tp_check.py $2 > gpurun_out/tp2_$1_$2.log 2>&1
  echo "tp_check $1 $2 rc=$? : $(grep -h 'tp=2\|TP_CHECK' gpurun_out/tp2_$1_$2.log | tr '\n' ' ')"
done
timeout 300 $TR --master-port 29660 bench.py --gpus 2 --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/bench_tp2.log 2> gpurun_out/bench_tp2.err; echo "bench tp2 rc=$?"
for mode in twoshot oneshot nccl; do
  i=$((i+1))
  TGIS_TP_EXCHANGE=$mode timeout 240 $TR --master-port $((29670+i)) bench.py --gpus 2 --steps 2 --warmup 1 --no-cpu-baseline --parallel tp --named-configs 0 --model llama3-70b --layers 4 --batch 256 > gpurun_out/bench_tp2_70b4l_$mode.log 2> gpurun_out/bench_tp2_70b4l_$mode.err; echo "70b-4l b256 $mode rc=$?"
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/bench_tp2*.log")):
    try:
        d = json.loads(open(f).read().strip().split("\n")[-1])
        r = d["roofline"]
        print(f.split("/")[-1], "%.0f tok/s  step %.3f ms  exch %.3f ms/step (%s/step) gemm %.3f ms/step  parity %s  dp %s" % (
            d["value"], r["decode_step_ms"], r.get("exchange_ms_per_step", 0), r.get("exchanges_per_step"), r.get("gemm_ms_per_step", 0),
            (d.get("tp_parity") or {}).get("matching_prefix_tokens"), (d.get("dp") or {}).get("value")))
    except Exception as e:
        print(f, "unparsed", e)
PY
tail -4 gpurun_out/bench_tp2.err
