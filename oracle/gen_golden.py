"""Generate tests/golden/*.json by running the REFERENCE's own code (and the third-party code it delegates to) on
seeded inputs, in the build container.  Run:  python oracle/gen_golden.py

Sources of truth exercised here:
  * /root/reference/src/vllm_tgis_adapter/tgis_utils/logits_processors.py  (imported unmodified: ExpDecayLengthPenaltyWarper,
    TypicalLogitsWarperWrapper -> transformers TypicalLogitsWarper)
  * vllm 0.22.0 CPU-runnable pieces: _custom_ops.apply_repetition_penalties_torch, v1.sample.ops.topk_topp_sampler.
    apply_top_k_top_p, v1.sample.sampler.Sampler.gather_logprobs / compute_logprobs, v1.core.sched.utils.check_stop
    semantics
  * transformers 5.5.0 LlamaForCausalLM (fp32, eager attention) for the model half
The fixtures travel to the GPU box (which has no /root/reference); tests/test_oracle_cpu.py pins oracle/ against them.
"""
from __future__ import annotations

import json
import os

os.environ.setdefault("TORCHDYNAMO_DISABLE", "1")  # vLLM wraps its rank count in torch.compile: run it eagerly
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
REF_SRC = "/root/reference/src"
OUT = ROOT / "tests" / "golden"


def f32list(t: torch.Tensor) -> list[float]:
    return [float(x) for x in t.flatten().tolist()]


def sampler_fixtures() -> dict:
    sys.path.insert(0, REF_SRC)
    # the package __init__ imports nothing heavy; the module itself needs only torch + transformers
    from vllm_tgis_adapter.tgis_utils.logits_processors import ExpDecayLengthPenaltyWarper, TypicalLogitsWarperWrapper

    V = 256
    g = torch.Generator().manual_seed(20260921)
    fx: dict = {"vocab": V, "exp_decay": [], "typical": [], "rep_penalty": [], "topk_topp": [], "logprobs": []}

    # --- ExpDecayLengthPenaltyWarper (reference-owned)
    for (start, decay, n_out, eos_val) in [(2, 1.5, 1, 1.25), (2, 1.5, 2, 1.25), (2, 1.5, 5, 1.25), (2, 1.5, 5, -3.5),
                                           (64, 1.05, 100, 0.37), (0, 10.0, 3, -0.001), (10, 1.0, 50, 2.0)]:
        logits = torch.randn(V, generator=g)
        eos = 7
        logits[eos] = eos_val
        inp = logits.clone()
        out = ExpDecayLengthPenaltyWarper((start, decay), eos)(list(range(n_out)), logits)
        fx["exp_decay"].append({"start": start, "decay": decay, "n_out": n_out, "eos": eos, "logits": f32list(inp),
                                "eos_out": float(out[eos])})

    # --- TypicalLogitsWarperWrapper (reference-owned wrapper around transformers)
    for mass, scale in [(0.9, 1.0), (0.5, 2.0), (0.2, 3.0), (0.95, 0.3), (0.99, 1.0)]:
        logits = torch.randn(V, generator=g) * scale
        out = TypicalLogitsWarperWrapper(mass=mass)([], logits.clone())
        fx["typical"].append({"mass": mass, "logits": f32list(logits),
                              "removed": torch.nonzero(torch.isinf(out)).flatten().tolist()})

    # --- vLLM repetition penalty (torch reference implementation of the op)
    from vllm._custom_ops import apply_repetition_penalties_torch

    for pen in (1.2, 1.7, 0.8):
        logits = torch.randn(1, V, generator=g) * 2
        seen = torch.zeros(1, V, dtype=torch.bool)
        seen[0, torch.randint(0, V, (40,), generator=g)] = True
        out = logits.clone()
        apply_repetition_penalties_torch(out, seen, torch.zeros_like(seen), torch.tensor([pen]))
        fx["rep_penalty"].append({"penalty": pen, "logits": f32list(logits), "seen": torch.nonzero(seen[0]).flatten().tolist(),
                                  "out": f32list(out)})

    # --- vLLM top-k / top-p
    from vllm.v1.sample.ops.topk_topp_sampler import apply_top_k_top_p

    for k, p in [(10, None), (None, 0.8), (50, 0.5), (1, None), (None, 0.05), (200, 0.95)]:
        logits = torch.randn(1, V, generator=g) * 2.5
        out = apply_top_k_top_p(logits.clone(), None if k is None else torch.tensor([k]),
                                None if p is None else torch.tensor([p]))
        fx["topk_topp"].append({"k": k, "p": p, "logits": f32list(logits),
                                "kept": torch.nonzero(torch.isfinite(out[0])).flatten().tolist()})

    # --- vLLM logprobs / rank / top-n (raw_logprobs mode)
    from vllm.v1.sample.sampler import Sampler

    s = Sampler()
    for n in (1, 3, 11):
        logits = torch.randn(2, V, generator=g) * 2
        lp = s.compute_logprobs(logits)
        tok = torch.argmax(logits, dim=-1)
        tok[1] = 123
        res = s.gather_logprobs(lp, n, tok.to(torch.int64))
        fx["logprobs"].append({"n": n, "logits": f32list(logits), "tokens": tok.tolist(),
                               "ids": res.logprob_token_ids.tolist(), "lps": [f32list(r) for r in res.logprobs],
                               "ranks": res.selected_token_ranks.tolist()})
    return fx


def llama_fixture() -> dict:
    from transformers import LlamaConfig as HFConfig
    from transformers import LlamaForCausalLM

    from oracle.llama_oracle import CONFIGS, synthetic_weights

    cfg = CONFIGS["tiny"]
    hf_cfg = HFConfig(vocab_size=cfg.vocab, hidden_size=cfg.hidden, intermediate_size=cfg.ffn,
                      num_hidden_layers=cfg.n_layers, num_attention_heads=cfg.n_q_heads,
                      num_key_value_heads=cfg.n_kv_heads, head_dim=cfg.head_dim, rms_norm_eps=cfg.rms_eps,
                      rope_theta=cfg.rope_theta, max_position_embeddings=cfg.max_model_len, tie_word_embeddings=False,
                      attn_implementation="eager")
    try:
        hf_cfg.rope_parameters = {"rope_type": "default", "rope_theta": cfg.rope_theta}
    except Exception:  # noqa: BLE001
        pass
    model = LlamaForCausalLM(hf_cfg).to(torch.float32).eval()
    w = synthetic_weights(cfg, seed=11, dtype=torch.float32)
    missing, unexpected = model.load_state_dict(w, strict=False)
    assert not unexpected and all("rotary" in m or "inv_freq" in m for m in missing), (missing, unexpected)
    prompt = [5, 17, 1000, 3, 250, 77, 9, 640, 31, 2, 999, 512, 13]
    with torch.no_grad():
        logits = model(torch.tensor([prompt])).logits[0]   # [T, V]
    return {"config": "tiny", "weights_seed": 11, "prompt": prompt,
            "last_logits_head": f32list(logits[-1, :32]), "argmax_per_pos": logits.argmax(-1).tolist(),
            "logsumexp_per_pos": f32list(torch.logsumexp(logits, -1)),
            "rope_theta_used": float(getattr(model.config, "rope_theta", cfg.rope_theta) or cfg.rope_theta)}


def llama_bf16_fixture() -> dict:
    """HF LlamaForCausalLM run in bf16 on CPU (sdpa attention: fp32 softmax inside the fused kernel, like the flash
    kernels vLLM uses): pins WHERE the oracle rounds to the model dtype (after every linear, after the norm product,
    after RoPE products, after attention, logits) -- the fp32 fixture above cannot see those points."""
    from transformers import LlamaConfig as HFConfig
    from transformers import LlamaForCausalLM

    from oracle.llama_oracle import CONFIGS, synthetic_weights

    out = {"cases": []}
    for name, seed, prompt_len in (("tiny", 21, 48), ("small", 22, 40)):
        cfg = CONFIGS[name]
        hf_cfg = HFConfig(vocab_size=cfg.vocab, hidden_size=cfg.hidden, intermediate_size=cfg.ffn,
                          num_hidden_layers=cfg.n_layers, num_attention_heads=cfg.n_q_heads,
                          num_key_value_heads=cfg.n_kv_heads, head_dim=cfg.head_dim, rms_norm_eps=cfg.rms_eps,
                          rope_theta=cfg.rope_theta, max_position_embeddings=cfg.max_model_len,
                          tie_word_embeddings=False, attn_implementation="sdpa")
        try:
            hf_cfg.rope_parameters = {"rope_type": "default", "rope_theta": cfg.rope_theta}
        except Exception:  # noqa: BLE001
            pass
        model = LlamaForCausalLM(hf_cfg).to(torch.bfloat16).eval()
        w = synthetic_weights(cfg, seed=seed)   # bf16
        missing, unexpected = model.load_state_dict(w, strict=False)
        assert not unexpected and all("rotary" in m or "inv_freq" in m for m in missing), (missing, unexpected)
        g = torch.Generator().manual_seed(seed)
        prompt = torch.randint(3, cfg.vocab, (prompt_len,), generator=g).tolist()
        with torch.no_grad():
            logits = model(torch.tensor([prompt])).logits[0]   # [T, V] bf16
        assert logits.dtype == torch.bfloat16
        lf = logits.float()
        top2 = torch.topk(lf, 2, dim=-1).values
        out["cases"].append({
            "config": name, "weights_seed": seed, "prompt": prompt,
            "argmax_per_pos": lf.argmax(-1).tolist(), "top2_margin_per_pos": f32list(top2[:, 0] - top2[:, 1]),
            "logsumexp_per_pos": f32list(torch.logsumexp(lf, -1)),
            "last_logits": f32list(lf[-1]),          # full last row (bf16 values, exactly representable)
            "mid_logits_head": f32list(lf[prompt_len // 2, :256])})
    return out


def main() -> None:
    OUT.mkdir(parents=True, exist_ok=True)
    import transformers
    import vllm

    meta = {"generated_by": "oracle/gen_golden.py", "torch": torch.__version__, "transformers": transformers.__version__,
            "vllm": vllm.__version__, "reference": "opendatahub-io/vllm-tgis-adapter @ df3596ae (/root/reference)"}
    s = sampler_fixtures()
    s["meta"] = meta
    (OUT / "sampler_reference.json").write_text(json.dumps(s))
    m = llama_fixture()
    m["meta"] = meta
    (OUT / "llama_tiny_hf_fp32.json").write_text(json.dumps(m))
    b = llama_bf16_fixture()
    b["meta"] = meta
    (OUT / "llama_hf_bf16.json").write_text(json.dumps(b))
    print("wrote", [p.name for p in OUT.iterdir()])


if __name__ == "__main__":
    main()
