"""How much HBM bandwidth can k CTAs (one per SM) pull?  Sweeps TGIS_GEMM_MAX_CTAS on an lm_head-sized stream."""
import ctypes as C
import json
import os
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
sys.path.insert(0, str(Path(__file__).resolve().parent.parent / "tests"))
import tgis_gpu_utils as g  # noqa: E402

N, K, T = 128256, 4096, 32
w = (torch.randn(N, K, device="cuda") * 0.02).bfloat16()
x = (torch.randn(256, K, device="cuda") * 0.5).bfloat16()
y = torch.empty(T, N, dtype=torch.bfloat16, device="cuda")
res = []
for ctas in (16, 32, 48, 64, 74, 96, 111, 128, 148):
    os.environ["TGIS_GEMM_MAX_CTAS"] = str(ctas)
    ms = C.c_float(0)
    for _ in range(2):
        assert g.lib().tgis_k_gemm(g.ptr(x), g.ptr(w), g.ptr(y), T, N, K, 256, 0, 1, C.byref(ms), 0) == 0, g.kerr()
    by = N * K * 2
    r = {"ctas": ctas, "us": ms.value * 1e3, "GBps": by / ms.value / 1e6, "GBps_per_cta": by / ms.value / 1e6 / ctas}
    res.append(r)
    print(json.dumps(r), flush=True)
Path("gpurun_out/gemm_cta_sweep.json").write_text(json.dumps(res, indent=1))
