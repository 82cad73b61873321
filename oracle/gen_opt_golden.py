"""Golden vectors for the OPT oracle (oracle/opt_oracle.py) from transformers' OPTForCausalLM — run in THIS container
(CPU), outputs committed as tests/golden/opt_hf_{fp32,bf16}.json.  TEST INFRASTRUCTURE ONLY.

    python -m oracle.gen_opt_golden

The reference's OPT fixture is the hub checkpoint facebook/opt-125m (/root/reference/tests/conftest.py:79-91 (no --model: vLLM's default, facebook/opt-125m) and tests/test_hub.py:17), which is
not obtainable here (no network); the fixtures use seeded random weights of the same architecture
(oracle/opt_oracle.py::synthetic_opt_weights) loaded into an HF OPTForCausalLM built from the same config keys
facebook/opt-125m's config.json carries (do_layer_norm_before, enable_bias, layer_norm_elementwise_affine, relu, tied
embeddings, word_embed_proj_dim == hidden)."""
from __future__ import annotations

import json
from pathlib import Path

import torch

OUT = Path(__file__).resolve().parent.parent / "tests" / "golden"


def f32list(t: torch.Tensor) -> list[float]:
    return [float(x) for x in t.detach().float().flatten().tolist()]


def hf_model(cfg, weights: dict[str, torch.Tensor], dtype: torch.dtype, attn: str):
    from transformers import OPTConfig as HFConfig
    from transformers import OPTForCausalLM

    hf_cfg = HFConfig(vocab_size=cfg.vocab, hidden_size=cfg.hidden, num_hidden_layers=cfg.n_layers, ffn_dim=cfg.ffn,
                      num_attention_heads=cfg.n_heads, max_position_embeddings=cfg.max_positions,
                      word_embed_proj_dim=cfg.hidden, do_layer_norm_before=True, enable_bias=True,
                      layer_norm_elementwise_affine=True, activation_function="relu", dropout=0.0,
                      attention_dropout=0.0, layerdrop=0.0, tie_word_embeddings=True, attn_implementation=attn)
    model = OPTForCausalLM(hf_cfg).to(dtype).eval()
    sd = {k: v.to(dtype) for k, v in weights.items()}
    missing, unexpected = model.load_state_dict(sd, strict=False)
    assert not unexpected, unexpected
    assert all(m == "lm_head.weight" for m in missing), missing
    model.tie_weights()
    assert model.lm_head.weight.data_ptr() == model.model.decoder.embed_tokens.weight.data_ptr()
    return model


def fixture(dtype: torch.dtype, attn: str) -> dict:
    from oracle.opt_oracle import OPT_CONFIGS, synthetic_opt_weights

    out = {"cases": []}
    for name, seed, prompt_len, n_dec in (("opt-tiny", 31, 37, 6), ("opt-125m-2l", 32, 24, 3)):
        if name == "opt-125m-2l":  # facebook/opt-125m dims, 2 of its 12 layers (keeps the fixture + CPU test small)
            import dataclasses

            cfg = dataclasses.replace(OPT_CONFIGS["opt-125m"], n_layers=2)
        else:
            cfg = OPT_CONFIGS[name]
        w = synthetic_opt_weights(cfg, seed=seed, dtype=torch.bfloat16)  # bf16-representable in both runs
        model = hf_model(cfg, w, dtype, attn)
        g = torch.Generator().manual_seed(seed)
        prompt = torch.randint(3, cfg.vocab, (prompt_len,), generator=g).tolist()
        with torch.no_grad():
            res = model(torch.tensor([prompt]), use_cache=True)
            logits = res.logits[0].float()
            # greedy continuation through HF's KV cache (the decode path: one new position per call)
            past, toks, dec_rows = res.past_key_values, [], []
            nxt = int(logits[-1].argmax())
            for _ in range(n_dec):
                toks.append(nxt)
                r = model(torch.tensor([[nxt]]), past_key_values=past, use_cache=True)
                past = r.past_key_values
                row = r.logits[0, -1].float()
                dec_rows.append(row)
                nxt = int(row.argmax())
        top2 = torch.topk(logits, 2, dim=-1).values
        out["cases"].append({
            "config": name, "weights_seed": seed, "prompt": prompt,
            "argmax_per_pos": logits.argmax(-1).tolist(), "top2_margin_per_pos": f32list(top2[:, 0] - top2[:, 1]),
            "logsumexp_per_pos": f32list(torch.logsumexp(logits, -1)),
            "last_logits_head": f32list(logits[-1, :512]),
            "mid_logits_head": f32list(logits[prompt_len // 2, :256]),
            "decode_tokens": toks,
            "decode_logsumexp": f32list(torch.stack([torch.logsumexp(r, -1) for r in dec_rows])),
            "decode_logits_head": [f32list(r[:128]) for r in dec_rows],
            "decode_top2_margin": f32list(torch.stack([torch.topk(r, 2).values[0] - torch.topk(r, 2).values[1]
                                                       for r in dec_rows]))})
    return out


def main() -> None:
    import transformers

    OUT.mkdir(parents=True, exist_ok=True)
    meta = {"generated_by": "oracle/gen_opt_golden.py", "torch": torch.__version__,
            "transformers": transformers.__version__,
            "reference": "opendatahub-io/vllm-tgis-adapter (/root/reference) tests/conftest.py:79-91 (no --model: vLLM's default, facebook/opt-125m) and tests/test_hub.py:17 (facebook/opt-125m)"}
    f = fixture(torch.float32, "eager")
    f["meta"] = meta
    (OUT / "opt_hf_fp32.json").write_text(json.dumps(f))
    b = fixture(torch.bfloat16, "sdpa")
    b["meta"] = meta
    (OUT / "opt_hf_bf16.json").write_text(json.dumps(b))
    print("wrote opt_hf_fp32.json, opt_hf_bf16.json")


if __name__ == "__main__":
    main()
