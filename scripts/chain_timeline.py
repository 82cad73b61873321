"""Timeline of one per-layer chain launch (o-proj -> norm -> gate_up -> down -> norm -> qkv) inside a real decode step.
Needs a -DTGIS_GEMM_TIMELINE build (TGIS_ENGINE_LIB).  Usage: python scripts/chain_timeline.py [batch] [graphs]"""
import ctypes as C
import dataclasses
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from vllm_tgis_adapter_b200.engine import _lib  # noqa: E402
from vllm_tgis_adapter_b200.engine.core import PRESETS, NativeEngine, make_sampling_params  # noqa: E402
from vllm_tgis_adapter_b200.engine.loader import load_synthetic_weights, rope_cos_sin  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
graphs = bool(int(sys.argv[2])) if len(sys.argv) > 2 else True
mc = dataclasses.replace(PRESETS["llama3-8b"], n_layers=6, max_model_len=1024)
eng = NativeEngine(mc, max_num_seqs=B, max_batched_tokens=max(8192, B * 512), kv_cache_bytes=2 << 30, use_cuda_graphs=graphs)
load_synthetic_weights(eng, mc, 0, 0)
eng.load_weight("tgis.rope_cos_sin", rope_cos_sin(mc))
rs = np.random.RandomState(0)
prompts = [rs.randint(1000, mc.vocab - 1000, size=512).tolist() for _ in range(B)]
eng.generate_sync(prompts, make_sampling_params(greedy=True, max_tokens=12, min_tokens=12))
st = eng.status()
print(f"decode step {st.gpu_decode_ms / max(1, st.decode_steps) * 1e3:.1f} us for 6 layers, batch {B}")
out = np.zeros((2, 64), dtype=np.uint64)
rc = _lib.load_library().tgis_k_chain_timeline(out.ctypes.data_as(C.POINTER(C.c_uint64)))
print("timeline rc", rc)
STEPS = ["o", "norm2", "gate_up", "down", "norm1", "qkv"]
EV = ["w_first", "x_first(gate)", "x_last", "acc_last", "done_last", "row_begin", "row_end", "w_last"]
t0 = int(min(v for v in out.reshape(-1) if v))
for cta in range(2):
    print("CTA", "0" if cta == 0 else "mid")
    for s, name in enumerate(STEPS):
        evs = [(EV[e], (int(out[cta][s * 8 + e]) - t0) / 1e3) for e in range(8) if out[cta][s * 8 + e]]
        print(f"  {name:8s}", "  ".join(f"{n}={t:.2f}" for n, t in sorted(evs, key=lambda kv: kv[1])))
eng.close()
