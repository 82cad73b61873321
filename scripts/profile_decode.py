"""Short engine run for ncu: Llama-3-8B shaped layers (default 4 of them), B requests with P-token prompts, a few
decode steps.  Usage: python scripts/profile_decode.py [layers] [batch] [prompt] [new_tokens] [graphs]"""
import dataclasses
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from vllm_tgis_adapter_b200.engine.core import PRESETS, NativeEngine, make_sampling_params  # noqa: E402
from vllm_tgis_adapter_b200.engine.loader import load_synthetic_weights, rope_cos_sin  # noqa: E402

L = int(sys.argv[1]) if len(sys.argv) > 1 else 4
B = int(sys.argv[2]) if len(sys.argv) > 2 else 32
P = int(sys.argv[3]) if len(sys.argv) > 3 else 512
G = int(sys.argv[4]) if len(sys.argv) > 4 else 6
graphs = bool(int(sys.argv[5])) if len(sys.argv) > 5 else False
mc = dataclasses.replace(PRESETS["llama3-8b"], n_layers=L, max_model_len=1024)
import os
MBT = int(os.environ.get("MBT", max(8192, B * P)))
eng = NativeEngine(mc, max_num_seqs=B, max_batched_tokens=MBT, kv_cache_bytes=int(float(os.environ.get('KVGB', '2')) * (1 << 30)), use_cuda_graphs=graphs)
load_synthetic_weights(eng, mc, 0, 0)
eng.load_weight("tgis.rope_cos_sin", rope_cos_sin(mc))
rs = np.random.RandomState(0)
prompts = [rs.randint(1000, mc.vocab - 1000, size=P).tolist() for _ in range(B)]
sp = make_sampling_params(greedy=True, max_tokens=G, min_tokens=G)
outs = eng.generate_sync(prompts, sp)
st = eng.status()
print(f"steps={st.steps} decode_steps={st.decode_steps} decode_ms={st.gpu_decode_ms:.3f} mixed_ms={st.gpu_mixed_ms:.3f} "
      f"launches={st.kernel_launches} graph_launches={st.graph_launches}")
if st.decode_steps:
    print(f"ms per decode step = {st.gpu_decode_ms / st.decode_steps:.4f} ({L} layers)")
eng.close()
