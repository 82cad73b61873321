#!/bin/bash
# ncu --set full captures of the OPT kernels (csrc/opt.cu) at the shapes of tests/test_zz_opt_gpu.py's kernel tests (one GPU):
# LayerNorm rows of 256...8192 (T = 2048 x 768 = an opt-125m prefill chunk), bias + ReLU over [2048, 3072] (fc1 of that chunk),
# the embedding add.  Same conventions as scripts/ncu_r02.sh; python scripts/summarize_ncu_csv.py r02 turns
# gpurun_out/ncu/*.csv into profiles/r02_ncu_<name>.csv.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out/ncu
NCU="ncu --set full --clock-control none --import-source on -f"
timeout 300 $NCU -k "regex:opt_" -c 40 -o gpurun_out/ncu/opt python -m pytest tests/test_zz_opt_gpu.py -q -k "kernel" > gpurun_out/ncu/opt.log 2>&1
echo "ncu opt rc=$?"
ncu -i gpurun_out/ncu/opt.ncu-rep --page raw --csv > gpurun_out/ncu/opt.csv 2>/dev/null
rm -f gpurun_out/ncu/opt.ncu-rep
ls -la gpurun_out/ncu | head
