"""Micro-benchmark of the paged decode-attention kernel through the C ABI on the Llama-3-8B head layout (32 q / 8 kv
heads of 128).  The cache is replicated over enough "layers" that every launch streams its KV from HBM, not L2."""
import ctypes as C
import json
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
sys.path.insert(0, str(Path(__file__).resolve().parent.parent / "tests"))
import tgis_gpu_utils as g  # noqa: E402

PEAK = 6566.7
NQ, NKV, HD, BLK = 32, 8, 128, 32
res = []
cases = [(32, 576), (64, 576), (128, 576), (256, 576), (32, 2048), (8, 8192), (32, 100), (256, 130)]
if len(sys.argv) > 1:
    cases = [tuple(int(v) for v in a.split("x")) for a in sys.argv[1:]]
for B, kv in cases:
    nblk_seq = (kv + BLK - 1) // BLK
    n_blocks = B * nblk_seq
    layer_elems = n_blocks * NKV * BLK * HD
    n_layers = max(2, int(400e6 // (layer_elems * 2 * 2)) + 1)
    kc = (torch.randn(n_layers * layer_elems, device="cuda") * 0.5).bfloat16()
    vc = (torch.randn(n_layers * layer_elems, device="cuda") * 0.5).bfloat16()
    qkv = (torch.randn(B, (NQ + 2 * NKV) * HD, device="cuda") * 0.5).bfloat16()
    out = torch.empty(B, NQ * HD, dtype=torch.bfloat16, device="cuda")
    perm = np.random.default_rng(0).permutation(n_blocks).astype(np.int32)
    bt = perm.reshape(B, nblk_seq)
    seqs = np.array([[i, 1, kv, i] for i in range(B)], dtype=np.int32)
    us = C.c_float(0)
    rc = g.lib().tgis_k_attention_bench(g.ptr(qkv), g.ptr(kc), g.ptr(vc), g.i32p(seqs.reshape(-1)), B,
                                        g.i32p(bt.reshape(-1)), B, nblk_seq, g.ptr(out), NQ, NKV,
                                        C.c_float(HD ** -0.5), n_layers, C.c_int64(layer_elems * 2), 40, C.byref(us))
    assert rc == 0, g.kerr()
    by = B * kv * NKV * HD * 2 * 2
    r = {"batch": B, "kv_len": kv, "us": us.value, "GBps": by / us.value / 1e3, "hbm_frac": by / us.value / 1e3 / PEAK,
         "roofline_us": by / PEAK / 1e3}
    res.append(r)
    print(json.dumps(r), flush=True)
    del kc, vc
Path("gpurun_out").mkdir(exist_ok=True)
Path("gpurun_out/attn_bench.json").write_text(json.dumps(res, indent=1))
