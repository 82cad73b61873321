"""Thin object wrapper over the C ABI (include/tgis_engine.h).  No arithmetic happens in Python."""
from __future__ import annotations

import ctypes as C
import dataclasses
from typing import Iterable

from . import _lib
from ._lib import TgisConfig, TgisSamplingParams, TgisStatus, TgisStepOutput


class EngineError(RuntimeError):
    pass


@dataclasses.dataclass
class ModelConfig:
    """Decoder dims (HF config.json names in comments).  arch "llama": RMSNorm / RoPE / GQA / SwiGLU; arch "opt": learned
    positions / LayerNorm / biased projections / ReLU (n_kv_heads == n_q_heads, head_dim 64 or 128, ffn = ffn_dim,
    rms_eps carries the LayerNorm epsilon, rope_theta unused) -- include/tgis_engine.h TGIS_ARCH_*."""

    n_layers: int        # num_hidden_layers
    hidden: int          # hidden_size
    n_q_heads: int       # num_attention_heads
    n_kv_heads: int      # num_key_value_heads
    ffn: int             # intermediate_size
    vocab: int           # vocab_size
    head_dim: int = 128
    rope_theta: float = 500000.0
    rms_eps: float = 1e-5
    max_model_len: int = 2048
    arch: str = "llama"


# Named architectures (real dims; this offline build runs them on synthetic weights — SURVEY.md §0)
PRESETS: dict[str, ModelConfig] = {
    "tiny": ModelConfig(n_layers=2, hidden=256, n_q_heads=4, n_kv_heads=2, ffn=512, vocab=1024, max_model_len=512),
    "small": ModelConfig(n_layers=4, hidden=512, n_q_heads=4, n_kv_heads=1, ffn=1536, vocab=4096, max_model_len=1024),
    "125m": ModelConfig(n_layers=12, hidden=768, n_q_heads=6, n_kv_heads=2, ffn=3072, vocab=50272, max_model_len=2048),
    # facebook/opt-125m (the reference's own test model; BASELINE configs[0]) and a CPU-oracle-sized sibling
    "opt-125m": ModelConfig(n_layers=12, hidden=768, n_q_heads=12, n_kv_heads=12, ffn=3072, vocab=50272, head_dim=64,
                            max_model_len=2048, arch="opt"),
    "opt-tiny": ModelConfig(n_layers=2, hidden=256, n_q_heads=4, n_kv_heads=4, ffn=1024, vocab=1024, head_dim=64,
                            max_model_len=512, arch="opt"),
    "llama3-8b": ModelConfig(n_layers=32, hidden=4096, n_q_heads=32, n_kv_heads=8, ffn=14336, vocab=128256,
                             max_model_len=8192),
    "llama3-70b": ModelConfig(n_layers=80, hidden=8192, n_q_heads=64, n_kv_heads=8, ffn=28672, vocab=128256,
                              max_model_len=8192),
}


@dataclasses.dataclass
class StepOutput:
    request_id: str
    new_token: int | None
    logprob: float
    rank: int
    topn: list[tuple[int, float]]
    finish_reason: int
    stop_token_id: int
    n_prompt_tokens: int
    n_output_tokens: int
    ts_arrival: float
    ts_first_scheduled: float
    ts_first_token: float
    ts_last_token: float
    prompt_pos: int = -1   # >= 1: prompt-logprob record for that prompt position
    token_id: int = -1     # the token the record describes (generated token, or prompt token for prompt records)


def make_sampling_params(*, greedy: bool = True, temperature: float = 1.0, top_k: int = 0, top_p: float = 1.0,
                         typical_p: float = 0.0, repetition_penalty: float = 1.0,
                         length_penalty: tuple[int, float] | None = None, eos_token_id: int = 2, min_tokens: int = 0,
                         max_tokens: int = 16, num_logprobs: int = 0, prompt_logprobs: int = 0, seed: int | None = None,
                         stop_token_ids: Iterable[int] = (), guided: bool = False,
                         lora_slot: int = 0) -> TgisSamplingParams:
    sp = TgisSamplingParams()
    sp.greedy = 1 if greedy else 0
    sp.temperature = float(temperature)
    sp.top_k = int(top_k)
    sp.top_p = float(top_p)
    sp.typical_p = float(typical_p)
    sp.repetition_penalty = float(repetition_penalty)
    sp.has_length_penalty = 1 if length_penalty is not None else 0
    if length_penalty is not None:
        sp.lp_start_index = int(length_penalty[0])
        sp.lp_decay_factor = float(length_penalty[1])
    sp.eos_token_id = int(eos_token_id)
    sp.min_tokens = int(min_tokens)
    sp.max_tokens = int(max_tokens)
    sp.num_logprobs = int(num_logprobs)
    sp.prompt_logprobs = int(prompt_logprobs)
    sp.has_seed = 1 if seed is not None else 0
    sp.seed = int(seed or 0)
    ids = list(stop_token_ids)
    sp.n_stop_token_ids = len(ids)
    for i, t in enumerate(ids):
        sp.stop_token_ids[i] = int(t)
    sp.guided = 1 if guided else 0
    sp.lora_slot = int(lora_slot)
    return sp


class NativeEngine:
    """Owns one `tgis_engine*`."""

    def __init__(self, model: ModelConfig, *, max_num_seqs: int = 64, max_batched_tokens: int = 2048,
                 kv_cache_bytes: int = 0, gpu_mem_fraction: float = 0.85, device: int = 0, seed: int = 0,
                 debug_gemm_ref: bool = False, use_cuda_graphs: bool | None = None, tp_size: int = 1, tp_rank: int = 0,
                 nccl_id: bytes | None = None, shm_name: str = "", max_loras: int = 0, max_lora_rank: int = 16):
        self.lib = _lib.load_library()
        self.model = model
        if use_cuda_graphs is None:   # default on; TGIS_CUDA_GRAPHS=0 turns decode-step graph replay off
            import os

            use_cuda_graphs = os.environ.get("TGIS_CUDA_GRAPHS", "1") != "0"
        cfg = TgisConfig()
        cfg.abi_version = _lib.ABI_VERSION
        cfg.n_layers, cfg.hidden, cfg.n_q_heads, cfg.n_kv_heads = model.n_layers, model.hidden, model.n_q_heads, model.n_kv_heads
        cfg.head_dim, cfg.ffn, cfg.vocab = model.head_dim, model.ffn, model.vocab
        cfg.rope_theta, cfg.rms_eps, cfg.max_model_len = model.rope_theta, model.rms_eps, model.max_model_len
        cfg.max_num_seqs, cfg.max_batched_tokens = max_num_seqs, max_batched_tokens
        cfg.kv_cache_bytes, cfg.gpu_mem_fraction = kv_cache_bytes, gpu_mem_fraction
        cfg.device, cfg.tp_size, cfg.tp_rank = device, tp_size, tp_rank
        if tp_size > 1:
            import torch  # noqa: F401  (load torch's NCCL before the engine dlopens libnccl.so.2)

            if nccl_id is None or len(nccl_id) != 128:
                raise EngineError("tensor parallelism needs the 128-byte NCCL unique id of rank 0")
            C.memmove(cfg.nccl_id, nccl_id, 128)
            cfg.shm_name = shm_name.encode()
        cfg.use_cuda_graphs, cfg.debug_gemm_ref, cfg.seed = int(use_cuda_graphs), 1 if debug_gemm_ref else 0, seed
        cfg.max_loras, cfg.max_lora_rank = int(max_loras), int(max_lora_rank) if max_loras else 0
        if model.arch not in ("llama", "opt"):
            raise EngineError(f"unknown architecture {model.arch!r}")
        cfg.arch = _lib.ARCH_OPT if model.arch == "opt" else _lib.ARCH_LLAMA
        self.max_loras, self.max_lora_rank = int(max_loras), int(max_lora_rank)
        self._h = C.c_void_p()
        if self.lib.tgis_engine_create(C.byref(cfg), C.byref(self._h)) != 0:
            raise EngineError(f"tgis_engine_create failed: {_lib.last_error(self.lib)}")
        self._poll_buf = (TgisStepOutput * 512)()

    # -- weights ------------------------------------------------------------------------------------------------
    def load_weight(self, name: str, tensor) -> None:
        """tensor: contiguous torch bf16 tensor (host or device) — "PyTorch tensors for weights only"."""
        import torch

        t = tensor.detach()
        if t.dtype != torch.bfloat16:
            t = t.to(torch.bfloat16)
        t = t.contiguous()
        rows = t.shape[0]
        cols = t.numel() // rows if rows else 0
        if self.lib.tgis_engine_load_weight(self._h, name.encode(), C.c_void_p(t.data_ptr()), rows, cols, 0) != 0:
            raise EngineError(f"load_weight({name}) failed: {_lib.last_error(self.lib)}")

    def load_weights(self, weights: dict) -> None:
        for k, v in weights.items():
            self.load_weight(k, v)

    # -- control ------------------------------------------------------------------------------------------------
    def start(self) -> None:
        if self.lib.tgis_engine_start(self._h) != 0:
            raise EngineError(_lib.last_error(self.lib))

    def add_request(self, request_id: str, prompt_ids: list[int], params: TgisSamplingParams) -> None:
        arr = (C.c_int32 * len(prompt_ids))(*prompt_ids)
        if self.lib.tgis_engine_add_request(self._h, request_id.encode(), arr, len(prompt_ids), C.byref(params)) != 0:
            raise EngineError(f"add_request failed: {_lib.last_error(self.lib)}")

    def abort(self, request_id: str) -> None:
        self.lib.tgis_engine_abort(self._h, request_id.encode())

    def load_adapter(self, slot: int, weights: dict) -> None:
        """weights: {(layer, module): (A [r, in] bf16, B_scaled [out, r] bf16)} as engine/lora.py `read_adapter` returns
        them.  The slot is zeroed first (modules the adapter does not target contribute nothing)."""
        import torch

        if self.lib.tgis_engine_clear_adapter(self._h, slot) != 0:
            raise EngineError(f"clear_adapter({slot}) failed: {_lib.last_error(self.lib)}")
        for (layer, module), (a, b) in weights.items():
            for which, t in (("lora_A", a), ("lora_B", b)):
                t = t.detach().to(torch.bfloat16).contiguous()
                name = f"layers.{layer}.{module}.{which}".encode()
                if self.lib.tgis_engine_load_adapter_weight(self._h, slot, name, C.c_void_p(t.data_ptr()), t.shape[0],
                                                            t.shape[1]) != 0:
                    raise EngineError(f"load_adapter_weight({name.decode()}) failed: {_lib.last_error(self.lib)}")

    def set_mask_provider(self, callback) -> None:
        """callback: a `guided.MASK_FN` C function pointer (the caller keeps it alive), or None to remove it."""
        self._mask_cb = callback   # the engine stores the raw pointer: keep the ctypes object alive with the engine
        fn = C.cast(callback, C.c_void_p) if callback is not None else None
        if self.lib.tgis_engine_set_mask_provider(self._h, fn, None) != 0:
            raise EngineError(_lib.last_error(self.lib))

    def poll(self, timeout_ms: int = 0) -> list[StepOutput]:
        n = self.lib.tgis_engine_poll(self._h, self._poll_buf, len(self._poll_buf), timeout_ms)
        if n < 0:
            raise EngineError(_lib.last_error(self.lib))
        outs = []
        for i in range(n):
            o = self._poll_buf[i]
            outs.append(StepOutput(
                request_id=o.request_id.decode(), new_token=o.token_id if o.n_new_tokens else None,
                logprob=o.logprob, rank=o.rank,
                topn=[(o.topn_ids[j], o.topn_logprobs[j]) for j in range(o.n_topn)],
                finish_reason=o.finish_reason, stop_token_id=o.stop_token_id, n_prompt_tokens=o.n_prompt_tokens,
                n_output_tokens=o.n_output_tokens, ts_arrival=o.ts_arrival, ts_first_scheduled=o.ts_first_scheduled,
                ts_first_token=o.ts_first_token, ts_last_token=o.ts_last_token, prompt_pos=o.prompt_pos, token_id=o.token_id))
        return outs

    def run_until_idle(self) -> int:
        n = self.lib.tgis_engine_run_until_idle(self._h)
        if n < 0:
            raise EngineError(f"run_until_idle failed: {_lib.last_error(self.lib)}")
        return n

    def status(self) -> TgisStatus:
        st = TgisStatus()
        self.lib.tgis_engine_status(self._h, C.byref(st))
        return st

    @staticmethod
    def nccl_unique_id() -> bytes:
        import torch  # noqa: F401  (torch's bundled libnccl.so.2 must be the copy in the process: see csrc/engine.cu)

        buf = (C.c_uint8 * 128)()
        lib = _lib.load_library()
        if lib.tgis_nccl_unique_id(C.byref(buf)) != 0:
            raise EngineError(_lib.last_error(lib))
        return bytes(buf)

    def worker_run(self) -> None:
        """Tensor-parallel ranks > 0: execute rank 0's step plans until it shuts down (blocking)."""
        if self.lib.tgis_engine_worker_run(self._h) != 0:
            raise EngineError(f"worker_run failed: {_lib.last_error(self.lib)}")

    def set_profiling(self, on: bool | int) -> None:
        """True/1: time every GEMM launch; 2: only those of pure-decode steps."""
        self.lib.tgis_engine_set_profiling(self._h, int(on))

    @property
    def max_model_len(self) -> int:
        return self.lib.tgis_engine_max_model_len(self._h)

    def close(self) -> None:
        if self._h:
            self.lib.tgis_engine_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):  # noqa: D105
        try:
            self.close()
        except Exception:  # noqa: BLE001
            pass

    # -- convenience for tests / bench --------------------------------------------------------------------------
    def generate_sync(self, prompts: list[list[int]], params: list[TgisSamplingParams] | TgisSamplingParams
                      ) -> list[list[StepOutput]]:
        """Submit all prompts, drive the step loop on this thread, return per-request output records."""
        if not isinstance(params, list):
            params = [params] * len(prompts)
        for i, (p, sp) in enumerate(zip(prompts, params)):
            self.add_request(f"r{i}", p, sp)
        self.run_until_idle()
        res: list[list[StepOutput]] = [[] for _ in prompts]
        while True:
            outs = self.poll(0)
            if not outs:
                break
            for o in outs:
                res[int(o.request_id[1:])].append(o)
        return res
