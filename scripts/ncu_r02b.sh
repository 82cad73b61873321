#!/bin/bash
# ncu --set full captures of the kernels added late in round 2 (one GPU): lora.cu shrink / expand at a mixed-adapter decode
# step and a prefill chunk, and the masked instantiation of the sampling kernel.  Same conventions as scripts/ncu_r02.sh;
# python scripts/summarize_ncu_csv.py r02 turns gpurun_out/ncu/*.csv into profiles/r02_ncu_<name>.csv.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out/ncu
NCU="ncu --set full --clock-control none --import-source on -f"
run() {  # name, kernel regex, skip, count, command...
  name=$1; k=$2; s=$3; c=$4; shift 4
  timeout 400 $NCU -k "regex:$k" -s $s -c $c -o gpurun_out/ncu/$name "$@" > gpurun_out/ncu/$name.log 2>&1
  echo "ncu $name rc=$?"
  ncu -i gpurun_out/ncu/$name.ncu-rep --page raw --csv > gpurun_out/ncu/$name.csv 2>/dev/null
  rm -f gpurun_out/ncu/$name.ncu-rep
}
LORA_BENCH_ONLY=1 run lora "lora_shrink|lora_expand" 0 24 python scripts/lora_guided_bench.py
LORA_BENCH_ONLY=1 run sampler_masked "tgis_sampler" 0 6 python scripts/lora_guided_bench.py
ls -la gpurun_out/ncu | head
