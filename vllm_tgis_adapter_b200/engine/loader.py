"""Build the engine the way `build_async_engine_client(args)` does for the reference
(/root/reference/src/vllm_tgis_adapter/__main__.py:48): read the HF model directory (config.json + *.safetensors),
stream each tensor into libtgis_engine.so ("PyTorch tensors for weights only"), attach a tokenizer.

Offline builds have no checkpoints: `--model <preset> --synthetic-weights` creates seeded N(0, 0.02)
weights of the named architecture on the device instead."""
from __future__ import annotations

import dataclasses
import json
import logging
from pathlib import Path

from .async_engine import AsyncTGISEngine
from .core import PRESETS, ModelConfig, NativeEngine
from .tokenizer import load_tokenizer

logger = logging.getLogger("vllm_tgis_adapter.engine")


def model_config_from_hf(path: Path, max_model_len: int | None) -> ModelConfig:
    cfg = json.loads((path / "config.json").read_text())
    archs = cfg.get("architectures") or []
    if cfg.get("model_type") == "opt" or any("OPT" in a for a in archs):
        return _opt_model_config(cfg, max_model_len)
    if cfg.get("model_type") != "llama" and not any("Llama" in a for a in archs):
        raise ValueError(f"unsupported model architecture {archs or cfg.get('model_type')}: only Llama-family "
                         "decoders (RMSNorm, RoPE, GQA, SwiGLU) and OPT are implemented")
    hidden, heads = cfg["hidden_size"], cfg["num_attention_heads"]
    head_dim = cfg.get("head_dim") or hidden // heads
    if head_dim != 128:
        raise ValueError(f"head_dim {head_dim} unsupported (kernels are specialised for 128)")
    if cfg.get("rope_scaling"):
        raise ValueError("rope_scaling is not supported yet")
    derived = cfg.get("max_position_embeddings", 8192)
    if max_model_len is not None and max_model_len > derived:
        raise ValueError(f"max_model_len {max_model_len} exceeds the model's max_position_embeddings {derived}")
    return ModelConfig(n_layers=cfg["num_hidden_layers"], hidden=hidden, n_q_heads=heads,
                       n_kv_heads=cfg.get("num_key_value_heads", heads), ffn=cfg["intermediate_size"],
                       vocab=cfg["vocab_size"], head_dim=head_dim, rope_theta=float(cfg.get("rope_theta", 10000.0)),
                       rms_eps=float(cfg.get("rms_norm_eps", 1e-5)), max_model_len=max_model_len or min(derived, 8192))


def _opt_model_config(cfg: dict, max_model_len: int | None) -> ModelConfig:
    """facebook/opt-* config.json -> ModelConfig(arch="opt").  The variants vLLM's OPT model handles with extra modules
    (opt-350m: post-LayerNorm + project_in/out) are refused rather than silently mis-served."""
    hidden, heads = cfg["hidden_size"], cfg["num_attention_heads"]
    head_dim = hidden // heads
    problems = []
    if head_dim not in (64, 128):
        problems.append(f"head_dim {head_dim} (64 or 128)")
    if not cfg.get("do_layer_norm_before", True):
        problems.append("do_layer_norm_before=false")
    if cfg.get("word_embed_proj_dim", hidden) != hidden:
        problems.append("word_embed_proj_dim != hidden_size")
    if cfg.get("activation_function", "relu") != "relu":
        problems.append(f"activation_function {cfg.get('activation_function')}")
    if not cfg.get("enable_bias", True) or not cfg.get("layer_norm_elementwise_affine", True):
        problems.append("enable_bias / layer_norm_elementwise_affine = false")
    if cfg.get("_remove_final_layer_norm", False):
        problems.append("_remove_final_layer_norm")
    if problems:
        raise ValueError("unsupported OPT variant: " + ", ".join(problems))
    derived = cfg.get("max_position_embeddings", 2048)
    if max_model_len is not None and max_model_len > derived:
        raise ValueError(f"max_model_len {max_model_len} exceeds the model's max_position_embeddings {derived}")
    return ModelConfig(n_layers=cfg["num_hidden_layers"], hidden=hidden, n_q_heads=heads, n_kv_heads=heads,
                       ffn=cfg["ffn_dim"], vocab=cfg["vocab_size"], head_dim=head_dim, rms_eps=1e-5,  # nn.LayerNorm default
                       max_model_len=max_model_len or derived, arch="opt")


def rope_cos_sin(mc: ModelConfig):
    """HF LlamaRotaryEmbedding / vllm rotary_embedding base.py: fp32 table -> bf16."""
    import torch

    d = mc.head_dim
    inv_freq = 1.0 / (mc.rope_theta ** (torch.arange(0, d, 2, dtype=torch.float32) / d))
    freqs = torch.outer(torch.arange(mc.max_model_len, dtype=torch.float32), inv_freq)
    return torch.cat([freqs.cos(), freqs.sin()], dim=-1).to(torch.bfloat16)


def load_synthetic_weights(eng: NativeEngine, mc: ModelConfig, seed: int, device: int) -> None:
    import torch

    gen = torch.Generator(device=f"cuda:{device}").manual_seed(seed)

    def rnd(r, c):
        return (torch.randn(r, c, generator=gen, device=f"cuda:{device}", dtype=torch.float32) * 0.02).to(torch.bfloat16)

    if mc.arch == "opt":
        # HF OPT parameter names; LayerNorm weights around 1, small random biases (HF's init has zeros: random ones exercise
        # the bias paths).  lm_head is tied to embed_tokens (not loaded).
        H = mc.hidden

        def vec(n, s=0.02, base=0.0):
            return (base + torch.randn(n, generator=gen, device=f"cuda:{device}", dtype=torch.float32) * s).to(torch.bfloat16)

        eng.load_weight("model.decoder.embed_tokens.weight", rnd(mc.vocab, H))
        eng.load_weight("model.decoder.embed_positions.weight", rnd(mc.max_model_len + 2, H))
        eng.load_weight("model.decoder.final_layer_norm.weight", vec(H, 0.05, 1.0))
        eng.load_weight("model.decoder.final_layer_norm.bias", vec(H))
        for i in range(mc.n_layers):
            p = f"model.decoder.layers.{i}."
            for m in ("q_proj", "k_proj", "v_proj", "out_proj"):
                eng.load_weight(p + f"self_attn.{m}.weight", rnd(H, H))
                eng.load_weight(p + f"self_attn.{m}.bias", vec(H))
            eng.load_weight(p + "fc1.weight", rnd(mc.ffn, H))
            eng.load_weight(p + "fc1.bias", vec(mc.ffn))
            eng.load_weight(p + "fc2.weight", rnd(H, mc.ffn))
            eng.load_weight(p + "fc2.bias", vec(H))
            for ln in ("self_attn_layer_norm", "final_layer_norm"):
                eng.load_weight(p + ln + ".weight", vec(H, 0.05, 1.0))
                eng.load_weight(p + ln + ".bias", vec(H))
        return

    q_dim, kv_dim = mc.n_q_heads * mc.head_dim, mc.n_kv_heads * mc.head_dim
    ones = torch.ones(mc.hidden, dtype=torch.bfloat16, device=f"cuda:{device}")
    eng.load_weight("model.embed_tokens.weight", rnd(mc.vocab, mc.hidden))
    eng.load_weight("lm_head.weight", rnd(mc.vocab, mc.hidden))
    eng.load_weight("model.norm.weight", ones)
    for i in range(mc.n_layers):
        p = f"model.layers.{i}."
        eng.load_weight(p + "self_attn.q_proj.weight", rnd(q_dim, mc.hidden))
        eng.load_weight(p + "self_attn.k_proj.weight", rnd(kv_dim, mc.hidden))
        eng.load_weight(p + "self_attn.v_proj.weight", rnd(kv_dim, mc.hidden))
        eng.load_weight(p + "self_attn.o_proj.weight", rnd(mc.hidden, q_dim))
        eng.load_weight(p + "mlp.gate_proj.weight", rnd(mc.ffn, mc.hidden))
        eng.load_weight(p + "mlp.up_proj.weight", rnd(mc.ffn, mc.hidden))
        eng.load_weight(p + "mlp.down_proj.weight", rnd(mc.hidden, mc.ffn))
        eng.load_weight(p + "input_layernorm.weight", ones)
        eng.load_weight(p + "post_attention_layernorm.weight", ones)


def load_safetensors_dir(eng: NativeEngine, path: Path) -> None:
    from safetensors import safe_open

    files = sorted(path.glob("*.safetensors"))
    if not files:
        bins = sorted(path.glob("pytorch_model*.bin"))   # facebook/opt-125m ships pytorch_model.bin
        if not bins:
            raise ValueError(f"no *.safetensors or pytorch_model*.bin files under {path}")
        import torch

        for f in bins:
            for name, t in torch.load(str(f), map_location="cpu", weights_only=True).items():
                eng.load_weight(name, t)
        return
    for f in files:
        with safe_open(str(f), framework="pt", device="cpu") as sf:
            for name in sf.keys():  # noqa: SIM118
                if name.endswith("rotary_emb.inv_freq"):
                    continue
                eng.load_weight(name, sf.get_tensor(name))


def _resolve_model(args):
    """(ModelConfig, checkpoint dir or None) from --model: a HF-format directory or a preset name + --synthetic-weights."""
    if not args.model:
        raise ValueError("--model / --model-name is required")
    path = Path(args.model)
    if path.is_dir():
        return model_config_from_hf(path, args.max_model_len), path
    if args.model in PRESETS:
        if not args.synthetic_weights:
            raise ValueError(f"--model {args.model} is a preset name: pass --synthetic-weights or a model directory")
        mc = dataclasses.replace(PRESETS[args.model])
        if args.max_model_len:
            mc.max_model_len = args.max_model_len
        return mc, None
    raise ValueError(f"model path {args.model} does not exist (no network access: hub ids cannot be resolved)")


def _make_native_engine(args, mc: ModelConfig, path, device: int, **tp_kw) -> NativeEngine:
    eng = NativeEngine(mc, max_num_seqs=args.max_num_seqs, max_batched_tokens=args.max_num_batched_tokens,
                       gpu_mem_fraction=args.gpu_memory_utilization, device=device, seed=args.seed,
                       max_loras=args.max_loras if getattr(args, "enable_lora", False) else 0,
                       max_lora_rank=getattr(args, "max_lora_rank", 16), **tp_kw)
    if path is not None:
        load_safetensors_dir(eng, path)      # full tensors: a tensor-parallel engine keeps its rank's shard
    else:
        load_synthetic_weights(eng, mc, args.seed, device)
    if mc.arch != "opt":   # OPT has learned positions (part of the checkpoint), no rotary table
        eng.load_weight("tgis.rope_cos_sin", rope_cos_sin(mc))
    return eng


def _tp_worker_main(args, rank: int, tp: int, nccl_id: bytes, shm_name: str) -> None:
    """Ranks 1..tp-1 of a tensor-parallel server (one spawned process per GPU): build the rank's shard of the same
    model and follow rank 0's step plans (shared-memory control block, csrc/engine.cu worker_loop) until it shuts down."""
    import torch  # noqa: F401  (first: resolves libnccl for the engine library)

    mc, path = _resolve_model(args)
    eng = _make_native_engine(args, mc, path, args.device + rank, tp_size=tp, tp_rank=rank, nccl_id=nccl_id,
                              shm_name=shm_name)
    try:
        eng.worker_run()
    finally:
        eng.close()


def build_engine(args) -> AsyncTGISEngine:
    """`--tensor-parallel-size N` (reference: tgis_utils/args.py:139-148 maps --num-gpus / --num-shard onto it): this
    process is rank 0 (scheduler, sampler, gRPC); ranks 1..N-1 are spawned here, one process per GPU."""
    tp = args.tensor_parallel_size or 1
    mc, path = _resolve_model(args)
    tp_kw = {}
    if tp > 1:
        import multiprocessing as mp
        import os

        import torch  # noqa: F401

        if mc.n_kv_heads % tp or mc.ffn % tp or mc.vocab % tp:
            raise ValueError(f"tensor_parallel_size {tp} does not divide kv heads / ffn / vocab of {mc}")
        nccl_id = NativeEngine.nccl_unique_id()
        shm_name = f"/tgis_tp_{os.getpid()}"
        ctx = mp.get_context("spawn")
        workers = [ctx.Process(target=_tp_worker_main, args=(args, r, tp, nccl_id, shm_name), daemon=True,
                               name=f"tgis-tp-rank{r}") for r in range(1, tp)]
        for w in workers:
            w.start()
        tp_kw = dict(tp_size=tp, tp_rank=0, nccl_id=nccl_id, shm_name=shm_name)
    eng = _make_native_engine(args, mc, path, args.device, **tp_kw)
    # the synthetic vocabulary only for synthetic models (preset + --synthetic-weights); a checkpoint directory without
    # tokenizer files is a start-up error
    tokenizer = load_tokenizer(args.tokenizer or (str(path) if path is not None else None), mc.vocab,
                               allow_synthetic=path is None)
    logger.info("engine ready: %s (tensor_parallel_size=%d)", mc, tp)
    return AsyncTGISEngine(eng, tokenizer, mc)
