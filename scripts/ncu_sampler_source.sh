#!/bin/bash
# Source-level ncu capture of ONE cfg3 (typical-p + penalties, sampling) launch of the sampling kernel: which lines the
# warps are sampled on.  Output: gpurun_out/ncu/sampler_cfg3_source.csv (ncu --page source --csv).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out/ncu
SAMPLER_BENCH_ONLY="64:cfg3" timeout 400 ncu --set full --section SourceCounters --clock-control none --import-source on -f \
  -k "regex:tgis_sampler" -s 1 -c 1 -o gpurun_out/ncu/sampler_cfg3_src python scripts/sampler_bench.py > gpurun_out/ncu/sampler_cfg3_src.log 2>&1
echo "ncu rc=$?"
ncu -i gpurun_out/ncu/sampler_cfg3_src.ncu-rep --page source --csv > gpurun_out/ncu/sampler_cfg3_source.csv 2>gpurun_out/ncu/sampler_cfg3_source.err
ncu -i gpurun_out/ncu/sampler_cfg3_src.ncu-rep --page raw --csv > gpurun_out/ncu/sampler_cfg3_raw.csv 2>/dev/null
# the .ncu-rep (a few MB) travels back too: ncu -i ... --page source reads it in the build container
ls -la gpurun_out/ncu | tail -5
head -c 600 gpurun_out/ncu/sampler_cfg3_source.err
