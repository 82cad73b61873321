"""Golden vectors for the LoRA restatement (oracle/llama_oracle.py::lora_add), produced by vLLM 0.22's own reference
implementation of the punica ops (vllm/lora/ops/torch_ops/lora_ops.py bgmv_shrink + bgmv_expand, imported unmodified)
in the configuration punica_gpu.add_lora_linear uses: fp32 shrink buffer, scale 1.0, expand with add_inputs=True into the
bf16 layer output.  Run in the build container (vLLM is importable there): python oracle/gen_lora_golden.py"""
import json
from pathlib import Path

import torch
from vllm.lora.ops.torch_ops.lora_ops import bgmv_expand, bgmv_shrink


def main() -> None:
    torch.manual_seed(5)
    T, K, N, R, n_adapters = 7, 96, 80, 8, 3
    x = torch.randn(T, K).to(torch.bfloat16)
    y = torch.randn(T, N).to(torch.bfloat16)
    a = (torch.randn(n_adapters, R, K) * 0.3).to(torch.bfloat16)
    b = (torch.randn(n_adapters, N, R) * 0.3).to(torch.bfloat16)
    idx = torch.tensor([0, 2, 1, 1, 0, 2, 2])
    buf = torch.zeros(T, R, dtype=torch.float32)
    bgmv_shrink(x, a, buf, idx, 1.0)
    out = y.clone()
    bgmv_expand(buf, b, out, idx, add_inputs=True)
    fix = {"x": x.float().tolist(), "y": y.float().tolist(), "a": a.float().tolist(), "b": b.float().tolist(),
           "idx": idx.tolist(), "buffer": buf.tolist(), "out": out.float().tolist(),
           "source": "vllm 0.22.0 vllm/lora/ops/torch_ops/lora_ops.py bgmv_shrink(scale=1.0) + bgmv_expand(add_inputs=True)"}
    path = Path(__file__).resolve().parent.parent / "tests" / "golden" / "lora_torch_ops.json"
    path.write_text(json.dumps(fix))
    print(path, path.stat().st_size)


if __name__ == "__main__":
    main()
