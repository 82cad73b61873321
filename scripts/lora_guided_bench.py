"""Micro-benchmark of the round-2 additions at Llama-3-8B shapes (device time per launch, CUDA events around `iters`
launches through the C-ABI test hooks):
  * lora.cu shrink + expand for every adapted projection (rank capacity 16; gate/up as one module of capacity 32) on a
    decode-shaped step (T = 32, four adapters mixed) and a prefill chunk (T = 2048, one adapter), against the bytes the
    kernels must move (A and B of the adapters present once, x and y once);
  * the fused sampling kernel with and without a guided-decoding bitmask (V = 128256, bf16 logits, greedy rows).
Writes gpurun_out/lora_guided_bench.json.  LORA_BENCH_ONLY=1: two launches per case (ncu captures)."""
import ctypes as C
import json
import os
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent / "tests"))
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import tgis_gpu_utils as g  # noqa: E402

HBM = 6566.7e9
peaks = Path(__file__).resolve().parent.parent / "MEASURED_PEAKS.json"
if peaks.exists():
    HBM = json.loads(peaks.read_text())["hbm_gbs"] * 1e9
ITERS = 2 if os.environ.get("LORA_BENCH_ONLY") else 50
H, QD, KVD, F, R = 4096, 4096, 1024, 14336, 16
MODULES = [("q_proj", H, QD, R), ("k_proj", H, KVD, R), ("o_proj", QD, H, R), ("gate_up (one module, 2R)", H, 2 * F, 2 * R),
           ("down_proj", F, H, R)]
res = {"hbm_peak_gbs": HBM / 1e9, "iters": ITERS, "lora": [], "sampler": []}
torch.manual_seed(0)
CASES = ((32, "4 adapters mixed"), (32, "1 adapter"), (256, "1 adapter"), (2048, "1 adapter"))
if os.environ.get("LORA_BENCH_ONLY"):   # ncu captures: two step shapes, the two largest modules
    CASES = ((32, "4 adapters mixed"), (2048, "1 adapter"))
    MODULES = MODULES[3:]
for T, mix in CASES:
    slots = 4
    tok = (torch.randint(1, slots + 1, (T,), dtype=torch.int32) if "mixed" in mix else torch.ones(T, dtype=torch.int32)).cuda()
    n_present = len(set(tok.tolist()))
    for name, K, N, Rm in MODULES:
        x = torch.randn(T, K, device="cuda").to(torch.bfloat16)
        y = torch.randn(T, N, device="cuda").to(torch.bfloat16)
        a = (torch.randn(slots, Rm, K, device="cuda") * 0.05).to(torch.bfloat16)
        b = (torch.randn(slots, N, Rm, device="cuda") * 0.05).to(torch.bfloat16)
        us_s, us_e = C.c_float(0), C.c_float(0)
        rc = g.lib().tgis_k_lora_bench(g.ptr(x), K, g.ptr(tok), g.ptr(a), g.ptr(b), K, N, Rm, g.ptr(y), N, T, ITERS,
                                       C.byref(us_s), C.byref(us_e))
        assert rc == 0, g.kerr()
        shrink_bytes = n_present * Rm * K * 2 + T * K * 2 + T * Rm * 4
        expand_bytes = n_present * N * Rm * 2 + 2 * T * N * 2 + T * Rm * 4
        res["lora"].append({"T": T, "mix": mix, "module": name, "K": K, "N": N, "rank_capacity": Rm,
                            "shrink_us": round(us_s.value, 2), "expand_us": round(us_e.value, 2),
                            "shrink_algorithmic_MB": round(shrink_bytes / 1e6, 3),
                            "expand_algorithmic_MB": round(expand_bytes / 1e6, 3),
                            "shrink_frac_of_hbm": round(shrink_bytes / (us_s.value * 1e-6) / HBM, 3),
                            "expand_frac_of_hbm": round(expand_bytes / (us_e.value * 1e-6) / HBM, 3)})
        print(res["lora"][-1], flush=True)

V = 128256
words = (V + 31) // 32
for n in ((32,) if os.environ.get("LORA_BENCH_ONLY") else (32, 64, 256)):
    logits = (torch.randn(n, V, device="cuda") * 1.3).to(torch.bfloat16)
    rows = np.zeros(n, dtype=g.SAMPLE_ROW_DTYPE)
    rows["temperature"], rows["top_p"], rows["rep_penalty"] = 1.0, 1.0, 1.0
    rows["eos_id"], rows["seq_slot"], rows["logits_row"] = 2, np.arange(n), np.arange(n)
    rows["flags"] = g.SAMPLE_GREEDY
    allow = torch.from_numpy(np.random.RandomState(1).randint(-2**31, 2**31 - 1, size=(n, words), dtype=np.int64)
                             .astype(np.int32)).cuda()
    _, us_plain = g.run_sampler(logits, rows, iters=ITERS, return_us=True)
    rows_m = rows.copy()
    rows_m["flags"] |= g.SAMPLE_MASKED
    _, us_mask = g.run_sampler(logits, rows_m, iters=ITERS, return_us=True, allow=allow)
    floor_us = n * V * 2 / HBM * 1e6
    res["sampler"].append({"rows": n, "greedy_us": round(us_plain, 2), "greedy_masked_us": round(us_mask, 2),
                           "hbm_floor_us_one_read_of_the_logits": round(floor_us, 2),
                           "mask_bytes_per_row": words * 4})
    print(res["sampler"][-1], flush=True)
os.makedirs("gpurun_out", exist_ok=True)
Path("gpurun_out/lora_guided_bench.json").write_text(json.dumps(res, indent=1))
