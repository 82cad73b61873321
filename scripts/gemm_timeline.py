"""Where does a decode-shaped GEMM launch spend its time?  Needs a -DTGIS_GEMM_TIMELINE build."""
import ctypes as C
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
sys.path.insert(0, str(Path(__file__).resolve().parent.parent / "tests"))
import tgis_gpu_utils as g  # noqa: E402

NAMES = ["entry", "setup", "w_prefetch", "dep_wait", "last_tma", "acc_ready", "drained", "atomic", "reduced", "exit"]
for name, N, K in [("qkv", 6144, 4096), ("o", 4096, 4096), ("gate_up", 28672, 4096), ("down", 4096, 14336)]:
    T = 32
    ws = [(torch.randn(N, K, device="cuda") * 0.02).bfloat16() for _ in range(3)]
    x = (torch.randn(256, K, device="cuda") * 0.5).bfloat16()
    y = torch.empty(T, N, dtype=torch.bfloat16, device="cuda")
    ms = C.c_float(0)
    for w in ws:
        assert g.lib().tgis_k_gemm(g.ptr(x), g.ptr(w), g.ptr(y), T, N, K, 256, 0, 1, C.byref(ms), 0) == 0, g.kerr()
    out = np.zeros((4, 16), dtype=np.uint64)
    rc = g.lib().tgis_k_gemm_timeline(out.ctypes.data_as(C.POINTER(C.c_uint64)))
    print(f"{name}: event time {ms.value * 1e3:.1f} us, timeline rc={rc}")
    for cta in range(2):
        t0 = int(out[cta][0])
        print("   cta", "0" if cta == 0 else "mid", " ".join(f"{n}={(int(v) - t0) / 1e3:.2f}" for n, v in zip(NAMES, out[cta]) if v))
