"""Micro-benchmark of the fused sampling kernel at V = 128256 on bf16 logits (the product path's dtype): device time per
launch (CUDA events around `iters` launches, C-ABI hook tgis_k_sampler_ex) for the BASELINE row mixes, against the HBM
floor of ONE read of the rows' logits.  Writes gpurun_out/sampler_bench.json."""
import json
import os
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent / "tests"))
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import tgis_gpu_utils as g  # noqa: E402
from oracle.sampler_oracle import len_penalty_factor_m1  # noqa: E402

V = 128256
HBM = 6566.7e9
peaks = Path(__file__).resolve().parent.parent / "MEASURED_PEAKS.json"
if peaks.exists():
    HBM = json.loads(peaks.read_text())["hbm_gbs"] * 1e9


def rows_for(n, kind):
    r = np.zeros(n, dtype=g.SAMPLE_ROW_DTYPE)
    r["temperature"], r["top_p"], r["rep_penalty"] = 1.0, 1.0, 1.0
    r["eos_id"], r["seq_slot"] = 2, np.arange(n)
    r["logits_row"] = np.arange(n)
    if kind == "greedy":
        r["flags"] = g.SAMPLE_GREEDY
    elif kind == "greedy_lp1":
        r["flags"] = g.SAMPLE_GREEDY | g.SAMPLE_LOGPROBS
        r["n_topn"] = 1
    elif kind == "greedy_lp11":
        r["flags"] = g.SAMPLE_GREEDY | g.SAMPLE_LOGPROBS
        r["n_topn"] = 11
    elif kind == "cfg3":   # BASELINE configs[2]: typical-p 0.9 + repetition penalty 1.2 + length penalty (64, 1.05), sampling
        r["flags"] = g.SAMPLE_TYPICAL | g.SAMPLE_LENPEN
        r["typical_p"], r["rep_penalty"] = 0.9, 1.2
        r["n_out"], r["min_tokens"], r["step"] = 100, 128, 100
        r["len_decay_factor"] = len_penalty_factor_m1(100, 64, 1.05)
        r["seed_lo"] = 1234
    elif kind == "topk_topp":
        r["temperature"], r["top_k"], r["top_p"] = 0.8, 40, 0.9
        r["seed_lo"] = 99
    return r


res = []
torch.manual_seed(0)
ONLY = os.environ.get("SAMPLER_BENCH_ONLY")       # e.g. "64:cfg3,32:greedy" (ncu captures): just these cases, 2 launches each
ITERS = 2 if ONLY else 50
CASES = [(32, "greedy"), (64, "greedy"), (128, "greedy"), (256, "greedy"), (32, "greedy_lp1"), (32, "greedy_lp11"),
                (64, "cfg3"), (32, "cfg3"), (64, "topk_topp")]
if ONLY:
    CASES = [(int(c.split(":")[0]), c.split(":")[1]) for c in ONLY.split(",")]
for n, kind in CASES:
    logits = (torch.randn(n, V, device="cuda") * 1.3).to(torch.bfloat16)
    words = (V + 31) // 32
    bitmap = torch.zeros(n, words, dtype=torch.int32, device="cuda")
    bitmap[:, :20] = 0x55555555
    rows = rows_for(n, kind)
    variants = [None] + ([] if ONLY else [1, 2, 4, 8] if kind == "greedy" else [2, 4] if kind == "cfg3" else [])
    for ncl in variants:
        if ncl is None:
            os.environ.pop("TGIS_SAMPLER_CLUSTER", None)
        else:
            os.environ["TGIS_SAMPLER_CLUSTER"] = str(ncl)
        out, us = g.run_sampler(logits, rows, bitmap, iters=ITERS, return_us=True)
        floor_us = n * V * 2 / HBM * 1e6
        res.append({"rows": n, "kind": kind, "cluster": ncl or "auto", "us": us, "hbm_floor_us": floor_us,
                    "logits_GBps": n * V * 2 / (us * 1e-6) / 1e9})
        print(json.dumps(res[-1]), flush=True)
os.environ.pop("TGIS_SAMPLER_CLUSTER", None)
Path("gpurun_out").mkdir(exist_ok=True)
Path("gpurun_out/sampler_bench.json").write_text(json.dumps(res, indent=1))
