"""LoRA adapters, host side: PEFT checkpoint -> engine adapter slot.

Reference seam: /root/reference/src/vllm_tgis_adapter/grpc/adapters.py:63-163 resolves `adapter_id` to a directory under
the adapter cache, reads `adapter_config.json`, and hands a `LoRARequest` to vLLM, which loads the weights
(vllm/lora/lora_model.py `from_local_checkpoint` / `from_lora_tensors`, peft_helper.py) and applies them per request.
What is restated here, with the vLLM lines it follows:
  * accepted configuration: peft_type LORA, bias "none", no DoRA, no modules_to_save (peft_helper.py `_validate_features`),
    r <= max_lora_rank (`validate_legal`);
  * scaling = lora_alpha / r, or lora_alpha / sqrt(r) with use_rslora (peft_helper.py `__post_init__`);
  * tensors: `...layers.{i}.{self_attn|mlp}.{module}.lora_A.weight` [r, in], `lora_B.weight` [out, r], cast to the model
    dtype, then the scaling is folded into lora_B IN THE MODEL DTYPE (lora_weights.py `optimize`: lora_b *= scaling) --
    one more bf16 rounding of B, reproduced here so that engine, oracle and vLLM see the same B.
PyTorch is used for the checkpoint tensors only ("PyTorch tensors for weights only")."""
from __future__ import annotations

import json
import math
import re
import threading
from pathlib import Path

from .types import LoRARequest

MODULES = ("q_proj", "k_proj", "v_proj", "o_proj", "gate_proj", "up_proj", "down_proj")
_KEY = re.compile(r"(?:^|\.)layers\.(\d+)\.(?:self_attn|mlp)\.(\w+)\.lora_([AB])\.weight$")


class LoRAConfigError(ValueError):
    pass


def read_adapter(path: str, *, n_layers: int, max_rank: int):
    """-> (rank, {(layer, module): (A [r, in] bf16, B_scaled [out, r] bf16)}) from a PEFT LoRA directory."""
    import torch

    p = Path(path)
    cfg = json.loads((p / "adapter_config.json").read_text())
    if cfg.get("peft_type") != "LORA":
        raise LoRAConfigError(f"adapter type {cfg.get('peft_type')} is not a LoRA adapter")
    r = int(cfg["r"])
    alpha = float(cfg.get("lora_alpha", r))
    problems = []
    if cfg.get("use_dora"):
        problems.append("vLLM does not yet support DoRA.")
    if cfg.get("modules_to_save"):
        problems.append("vLLM only supports modules_to_save being None.")
    if cfg.get("bias", "none") != "none":
        problems.append("Adapter bias is not supported.")
    if r > max_rank:
        problems.append(f"LoRA rank {r} is greater than max_lora_rank {max_rank}.")
    if problems:
        raise LoRAConfigError(" ".join(problems))
    scaling = alpha / math.sqrt(r) if cfg.get("use_rslora") else alpha / r
    st, binf = p / "adapter_model.safetensors", p / "adapter_model.bin"
    if st.exists():
        from safetensors.torch import load_file

        tensors = load_file(str(st))
    elif binf.exists():
        tensors = torch.load(str(binf), map_location="cpu", weights_only=True)
    else:
        raise LoRAConfigError(f"{path} has neither adapter_model.safetensors nor adapter_model.bin")
    halves: dict[tuple[int, str], dict[str, torch.Tensor]] = {}
    for name, t in tensors.items():
        m = _KEY.search(name)
        if m is None:
            raise LoRAConfigError(f"{name} is unsupported LoRA weight")   # lora_model.py rejects unexpected modules
        layer, module, which = int(m.group(1)), m.group(2), m.group(3)
        if module not in MODULES:
            raise LoRAConfigError(f"{name} is unsupported LoRA weight")
        if layer >= n_layers:
            raise LoRAConfigError(f"{name}: the model has {n_layers} layers")
        halves.setdefault((layer, module), {})[which] = t.to(torch.bfloat16)
    out = {}
    for key, ab in halves.items():
        if "A" not in ab or "B" not in ab:
            raise LoRAConfigError(f"layer {key[0]} {key[1]}: lora_A / lora_B pair incomplete")
        a, b = ab["A"], ab["B"]
        if a.shape[0] != r or b.shape[1] != r:
            raise LoRAConfigError(f"layer {key[0]} {key[1]}: rank {a.shape[0]}/{b.shape[1]} != r = {r}")
        out[key] = (a.contiguous(), (b * scaling).contiguous())   # bf16 multiply: lora_weights.py optimize()
    return r, out


class LoRAManager:
    """adapter name -> engine slot (1..max_loras).  A slot is pinned while a request uses it; idle slots are recycled in
    least-recently-used order (vLLM: LRUCacheWorkerLoRAManager)."""

    def __init__(self, engine, *, n_layers: int, max_loras: int, max_rank: int):
        self.engine = engine
        self.n_layers, self.max_loras, self.max_rank = n_layers, max_loras, max_rank
        self._lock = threading.Lock()
        self._slot_of: dict[str, int] = {}
        self._users: dict[int, int] = {}
        self._clock = 0
        self._last_use: dict[int, int] = {}
        self.lora_requests: dict[str, LoRARequest] = {}    # what adapters.py looks up (OpenAIServingModels.lora_requests)
        self.loads = 0

    def register(self, req: LoRARequest) -> None:
        """Validate the checkpoint now (a bad adapter must fail the request that names it, adapters.py:146-155)."""
        read_adapter(req.lora_path, n_layers=self.n_layers, max_rank=self.max_rank)
        self.lora_requests[req.lora_name] = req

    def acquire(self, req: LoRARequest) -> int:
        with self._lock:
            self._clock += 1
            slot = self._slot_of.get(req.lora_name)
            if slot is None:
                used = set(self._slot_of.values())
                free = [s for s in range(1, self.max_loras + 1) if s not in used]
                if free:
                    slot = free[0]
                else:
                    idle = [s for s in used if self._users.get(s, 0) == 0]
                    if not idle:
                        raise RuntimeError(f"all {self.max_loras} LoRA slots are in use by running requests")
                    slot = min(idle, key=lambda s: self._last_use.get(s, 0))
                    for name in [n for n, s in self._slot_of.items() if s == slot]:
                        del self._slot_of[name]
                _, weights = read_adapter(req.lora_path, n_layers=self.n_layers, max_rank=self.max_rank)
                self.engine.load_adapter(slot, weights)
                self.loads += 1
                self._slot_of[req.lora_name] = slot
            self._users[slot] = self._users.get(slot, 0) + 1
            self._last_use[slot] = self._clock
            return slot

    def release(self, slot: int) -> None:
        with self._lock:
            self._users[slot] = max(0, self._users.get(slot, 0) - 1)
