"""One GEMM launch through the C ABI (for ncu).  Usage: python scripts/gemm_one.py T N K [iters]"""
import ctypes as C
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
sys.path.insert(0, str(Path(__file__).resolve().parent.parent / "tests"))
import tgis_gpu_utils as g  # noqa: E402

T, N, K = (int(v) for v in sys.argv[1:4])
iters = int(sys.argv[4]) if len(sys.argv) > 4 else 3
w = (torch.randn(N, K, device="cuda") * 0.02).bfloat16()
rows = max(T, 256)
x = (torch.randn(rows, K, device="cuda") * 0.5).bfloat16()
y = torch.empty(T, N, dtype=torch.bfloat16, device="cuda")
ms = C.c_float(0)
assert g.lib().tgis_k_gemm(g.ptr(x), g.ptr(w), g.ptr(y), T, N, K, rows, 0, iters, C.byref(ms), 0) == 0, g.kerr()
print(f"T={T} N={N} K={K}: {ms.value * 1e3:.1f} us per launch, {2.0 * T * N * K / ms.value / 1e9:.1f} TFLOP/s")
