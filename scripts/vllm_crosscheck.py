#!/usr/bin/env python
"""Cross-check of this engine and the CPU oracle against the reference's REAL engine path, and the GPU-vLLM baseline.

The reference adapter's hot path is one call into vLLM (`/root/reference/src/vllm_tgis_adapter/grpc/grpc_server.py:205-225`
`engine.generate(prompt=TokensPrompt(prompt_token_ids=...), sampling_params=..., request_id=...)`; parameters built at
`:560-622`).  vLLM 0.22.0 is installed on the GPU box (not importable with CUDA here), so this script runs THERE:

  python scripts/vllm_crosscheck.py check  [--configs tiny small 8b2l]   # parity: vLLM vs engine vs oracle
  python scripts/vllm_crosscheck.py bench  [--batches 32 64]             # GPU-vLLM tokens/s on the BASELINE configs

`check` builds seeded synthetic Llama checkpoints (HF format: config.json + model.safetensors + the synthetic WordLevel
tokenizer) in a scratch directory, runs the same request sets through
  (1) vLLM `LLM(model=<dir>, dtype=bfloat16, enforce_eager=True)` -- in a subprocess, twice: all requests in one batch,
      and one request at a time (the difference between the two is vLLM's OWN batch-composition noise),
  (2) this repo's engine (C ABI) loading the same safetensors,
  (3) the CPU oracle,
and writes  tests/golden/vllm_<cfg>.json  (vLLM's outputs: the fixture the -m gpu and CPU tests compare against) and
profiles/r02_vllm_crosscheck.json (the |dlogprob| / flip / rank tables of DESIGN.md section 5).
The reference's ExpDecayLengthPenaltyWarper rides along as a vLLM V1 `AdapterLogitsProcessor`: the
per-request callable is the oracle's restatement, which tests/test_oracle_cpu.py pins bit-exactly to the reference's own
class (the reference source tree does not exist on the GPU box).

Request sets (seeded; the tests rebuild them from the same seeds):
  greedy   6 prompts (5..250 tokens), 24 new tokens, logprobs=3        -> token ids, logprob, rank, top-3
  plp      4 prompts of 96 tokens, prompt_logprobs=2, 1 new token      -> teacher-forced logprob / rank per position
  lenpen   3 prompts, greedy + ExpDecay length penalty (start 6, 1.35) -> EOS position and token ids
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

SCRATCH = Path(os.environ.get("TGIS_XCHECK_DIR", "/tmp/tgis_xcheck"))
EOS = 2
WEIGHT_SEED = 1


def configs():
    import dataclasses

    from oracle.llama_oracle import CONFIGS

    c = dict(CONFIGS)
    # OPT family (the reference's own test model is facebook/opt-125m): its dims in full depth, and the CPU-suite sibling
    from oracle.opt_oracle import OPT_CONFIGS

    c["opt-tiny"] = OPT_CONFIGS["opt-tiny"]
    c["opt-125m"] = dataclasses.replace(OPT_CONFIGS["opt-125m"], max_model_len=1024)
    # Llama-3-8B dims (hidden 4096, 32/8 heads, ffn 14336, V 128256), 2 layers: BASELINE shapes at a size the CPU
    # oracle and a fixture can carry
    c["8b2l"] = dataclasses.replace(CONFIGS["llama3-8b"], n_layers=2, max_model_len=1024)
    return c


def is_opt(cfg_name: str) -> bool:
    return cfg_name.startswith("opt-")


def make_engine(cfg_name: str):
    """This repo's engine over the checkpoint directory vLLM loads.  OPT goes through the product loader end to end
    (config.json -> ModelConfig(arch="opt"), HF parameter names -> the engine's head-padded layout)."""
    from vllm_tgis_adapter_b200.engine.core import ModelConfig, NativeEngine
    from vllm_tgis_adapter_b200.engine.loader import load_safetensors_dir, model_config_from_hf, rope_cos_sin

    cfg = configs()[cfg_name]
    d = make_model_dir(cfg_name)
    if is_opt(cfg_name):
        mc = model_config_from_hf(d, cfg.max_model_len)
        eng = NativeEngine(mc, max_num_seqs=16, max_batched_tokens=2048, kv_cache_bytes=1 << 30)
        load_safetensors_dir(eng, d)
        return eng
    mc = ModelConfig(n_layers=cfg.n_layers, hidden=cfg.hidden, n_q_heads=cfg.n_q_heads, n_kv_heads=cfg.n_kv_heads,
                     ffn=cfg.ffn, vocab=cfg.vocab, rope_theta=cfg.rope_theta, rms_eps=cfg.rms_eps,
                     max_model_len=cfg.max_model_len)
    eng = NativeEngine(mc, max_num_seqs=16, max_batched_tokens=2048, kv_cache_bytes=256 << 20)
    load_safetensors_dir(eng, d)
    eng.load_weight("tgis.rope_cos_sin", rope_cos_sin(mc))
    return eng


def request_sets(cfg_name: str, vocab: int) -> dict:
    import numpy as np

    rng = np.random.RandomState(0)
    greedy = [rng.randint(3, vocab, size=n).tolist() for n in (5, 33, 64, 100, 17, 250)]
    rng = np.random.RandomState(7)
    plp = [rng.randint(3, vocab, size=96).tolist() for _ in range(4)]
    rng = np.random.RandomState(11)
    lenpen = [rng.randint(3, vocab, size=n).tolist() for n in (12, 40, 77)]
    return {"greedy": {"prompts": greedy, "n_new": 24, "logprobs": 3},
            "plp": {"prompts": plp, "prompt_logprobs": 2},
            "lenpen": {"prompts": lenpen, "max_tokens": 48, "start": 6, "decay": 1.35, "logprobs": 1}}


def make_model_dir(cfg_name: str) -> Path:
    """HF-format checkpoint of seeded synthetic weights + the synthetic tokenizer (SURVEY.md section 8c fixtures)."""
    import torch
    from safetensors.torch import save_file

    from oracle.llama_oracle import synthetic_weights
    from vllm_tgis_adapter_b200.engine.tokenizer import build_synthetic_tokenizer

    cfg = configs()[cfg_name]
    d = SCRATCH / cfg_name
    if (d / "model.safetensors").exists() and (d / "config.json").exists():
        return d
    d.mkdir(parents=True, exist_ok=True)
    if is_opt(cfg_name):
        from oracle.opt_oracle import synthetic_opt_weights

        hf = {   # the keys of facebook/opt-125m's config.json
            "architectures": ["OPTForCausalLM"], "model_type": "opt", "torch_dtype": "bfloat16", "dtype": "bfloat16",
            "vocab_size": cfg.vocab, "hidden_size": cfg.hidden, "ffn_dim": cfg.ffn, "num_hidden_layers": cfg.n_layers,
            "num_attention_heads": cfg.n_heads, "max_position_embeddings": cfg.max_positions,
            "word_embed_proj_dim": cfg.hidden, "do_layer_norm_before": True, "activation_function": "relu",
            "enable_bias": True, "layer_norm_elementwise_affine": True, "_remove_final_layer_norm": False,
            "dropout": 0.0, "attention_dropout": 0.0, "layerdrop": 0.0, "init_std": 0.02, "tie_word_embeddings": True,
            "bos_token_id": 1, "eos_token_id": EOS, "pad_token_id": 0, "prefix": "</s>", "use_cache": True,
        }
        (d / "config.json").write_text(json.dumps(hf, indent=1))
        (d / "generation_config.json").write_text(json.dumps({"bos_token_id": 1, "eos_token_id": EOS, "do_sample": False}))
        w = synthetic_opt_weights(cfg, seed=WEIGHT_SEED)
        save_file({k: v.contiguous() for k, v in w.items()}, str(d / "model.safetensors"), metadata={"format": "pt"})
        build_synthetic_tokenizer(cfg.vocab).save_pretrained(str(d))
        return d
    hf = {
        "architectures": ["LlamaForCausalLM"], "model_type": "llama", "torch_dtype": "bfloat16", "dtype": "bfloat16",
        "vocab_size": cfg.vocab, "hidden_size": cfg.hidden, "intermediate_size": cfg.ffn,
        "num_hidden_layers": cfg.n_layers, "num_attention_heads": cfg.n_q_heads,
        "num_key_value_heads": cfg.n_kv_heads, "head_dim": cfg.head_dim, "hidden_act": "silu",
        "rms_norm_eps": cfg.rms_eps, "rope_theta": cfg.rope_theta,
        "rope_parameters": {"rope_type": "default", "rope_theta": cfg.rope_theta}, "rope_scaling": None,
        "max_position_embeddings": cfg.max_model_len, "tie_word_embeddings": False, "attention_bias": False,
        "mlp_bias": False, "bos_token_id": 1, "eos_token_id": EOS, "pad_token_id": 0, "use_cache": True,
        "initializer_range": 0.02, "attention_dropout": 0.0, "pretraining_tp": 1,
    }
    (d / "config.json").write_text(json.dumps(hf, indent=1))
    (d / "generation_config.json").write_text(json.dumps({"bos_token_id": 1, "eos_token_id": EOS, "do_sample": False}))
    w = synthetic_weights(cfg, seed=WEIGHT_SEED)
    save_file({k: v.contiguous() for k, v in w.items()}, str(d / "model.safetensors"), metadata={"format": "pt"})
    build_synthetic_tokenizer(cfg.vocab).save_pretrained(str(d))
    return d


# ======================================================================================================== vLLM side
def _lenpen_processor_cls():
    """The reference's ExpDecayLengthPenaltyWarper as a vLLM V1 per-request processor."""
    import torch  # noqa: F401
    from vllm.v1.sample.logits_processor import AdapterLogitsProcessor

    from oracle.sampler_oracle import exp_decay_length_penalty

    class ExpDecayAdapter(AdapterLogitsProcessor):
        def is_argmax_invariant(self) -> bool:
            return False

        def new_req_logits_processor(self, params):
            lp = (params.extra_args or {}).get("length_penalty")
            if not lp:
                return None
            start, decay = int(lp[0]), float(lp[1])

            def call(output_ids, logits):
                # reference: tgis_utils/logits_processors.py:33-47 (oracle restatement, pinned bit-exactly by the golden)
                exp_decay_length_penalty(logits, len(output_ids), start, decay, EOS)
                return logits

            return call

    return ExpDecayAdapter


def _lp_entry(d: dict, token: int) -> dict:
    lp = d[token]
    items = sorted(((int(t), float(v.logprob), int(v.rank) if v.rank is not None else -1) for t, v in d.items()),
                   key=lambda x: (x[2] if x[2] > 0 else 1 << 30, x[0]))
    return {"logprob": float(lp.logprob), "rank": int(lp.rank), "top": [[t, l, r] for t, l, r in items]}


def vllm_run(cfg_name: str, out_path: str) -> None:
    os.environ.setdefault("VLLM_ENABLE_V1_MULTIPROCESSING", "0")
    os.environ.setdefault("HF_HUB_OFFLINE", "1")
    os.environ.setdefault("TRANSFORMERS_OFFLINE", "1")
    os.environ.setdefault("VLLM_NO_USAGE_STATS", "1")
    os.environ.setdefault("VLLM_LOGGING_LEVEL", "WARNING")
    import torch
    import vllm
    from vllm import LLM, SamplingParams

    cfg = configs()[cfg_name]
    d = make_model_dir(cfg_name)
    rs = request_sets(cfg_name, cfg.vocab)
    t0 = time.time()
    llm = LLM(model=str(d), dtype="bfloat16", enforce_eager=True, max_model_len=cfg.max_model_len,
              gpu_memory_utilization=0.30, seed=0, enable_prefix_caching=False, max_num_seqs=16, max_logprobs=20,
              logits_processors=[_lenpen_processor_cls()])
    t_load = time.time() - t0

    def gen(prompts, sps):
        reqs = [{"prompt_token_ids": p} for p in prompts]
        return llm.generate(reqs, sps, use_tqdm=False)

    def greedy_set(batched: bool):
        g = rs["greedy"]
        sp = SamplingParams(temperature=0.0, max_tokens=g["n_new"], min_tokens=g["n_new"], logprobs=g["logprobs"],
                            detokenize=False)
        outs = gen(g["prompts"], sp) if batched else [gen([p], sp)[0] for p in g["prompts"]]
        res = []
        for o in outs:
            c = o.outputs[0]
            toks = list(c.token_ids)
            res.append({"tokens": toks, "steps": [_lp_entry(c.logprobs[i], toks[i]) for i in range(len(toks))],
                        "finish_reason": c.finish_reason})
        return res

    def plp_set(batched: bool):
        g = rs["plp"]
        sp = SamplingParams(temperature=0.0, max_tokens=1, prompt_logprobs=g["prompt_logprobs"], detokenize=False)
        outs = gen(g["prompts"], sp) if batched else [gen([p], sp)[0] for p in g["prompts"]]
        res = []
        for p, o in zip(g["prompts"], outs):
            pl = o.prompt_logprobs
            assert pl[0] is None and len(pl) == len(p)
            res.append({"positions": [_lp_entry(pl[i], p[i]) for i in range(1, len(p))]})
        return res

    def lenpen_set():
        g = rs["lenpen"]
        sp = SamplingParams(temperature=0.0, max_tokens=g["max_tokens"], logprobs=g["logprobs"], detokenize=False,
                            extra_args={"length_penalty": [g["start"], g["decay"]]})
        outs = gen(g["prompts"], sp)
        res = []
        for o in outs:
            c = o.outputs[0]
            toks = list(c.token_ids)
            res.append({"tokens": toks, "finish_reason": c.finish_reason, "stop_reason": c.stop_reason,
                        "steps": [_lp_entry(c.logprobs[i], toks[i]) for i in range(len(toks))]})
        return res

    result = {"meta": {"vllm": vllm.__version__, "torch": torch.__version__, "gpu": torch.cuda.get_device_name(0),
                       "config": cfg_name, "weights_seed": WEIGHT_SEED, "dtype": "bfloat16", "enforce_eager": True,
                       "load_s": t_load, "generated_by": "scripts/vllm_crosscheck.py (on the GPU box)"},
              "greedy": greedy_set(True), "greedy_single": greedy_set(False),
              "plp": plp_set(True), "plp_single": plp_set(False)}
    try:
        result["lenpen"] = lenpen_set()
    except Exception as e:  # noqa: BLE001  (custom processors are the least stable part of the vLLM API)
        result["lenpen_error"] = repr(e)
    try:
        result["meta"]["attention_backend"] = str(
            llm.llm_engine.engine_core.engine_core.model_executor.driver_worker.worker.model_runner.attn_groups[0][0]
            .backend.get_name())
    except Exception:  # noqa: BLE001
        result["meta"]["attention_backend"] = "unknown"
    Path(out_path).write_text(json.dumps(result))
    print(f"[vllm-run] {cfg_name}: wrote {out_path} (load {t_load:.1f}s)", flush=True)


# ======================================================================================================== our side
def engine_run(cfg_name: str) -> dict:
    from vllm_tgis_adapter_b200.engine.core import make_sampling_params

    cfg = configs()[cfg_name]
    rs = request_sets(cfg_name, cfg.vocab)
    eng = make_engine(cfg_name)

    def entry(r):
        return {"logprob": r.logprob, "rank": r.rank, "top": [[t, l, i + 1] for i, (t, l) in enumerate(r.topn)]}

    g = rs["greedy"]
    sp = make_sampling_params(greedy=True, max_tokens=g["n_new"], min_tokens=g["n_new"], num_logprobs=g["logprobs"],
                              eos_token_id=EOS)
    outs = eng.generate_sync(g["prompts"], sp)
    greedy = [{"tokens": [r.new_token for r in recs if r.new_token is not None],
               "steps": [entry(r) for r in recs if r.new_token is not None]} for recs in outs]
    # teacher-forced replay of vLLM's tokens is done by the caller through `plp`-style prompts (see compare)
    p = rs["plp"]
    sp = make_sampling_params(greedy=True, max_tokens=1, num_logprobs=p["prompt_logprobs"],
                              prompt_logprobs=p["prompt_logprobs"], eos_token_id=EOS)
    outs = eng.generate_sync(p["prompts"], sp)
    plp = []
    for pr, recs in zip(p["prompts"], outs):
        pos = {r.prompt_pos: entry(r) for r in recs if r.prompt_pos >= 1}
        plp.append({"positions": [pos[i] for i in range(1, len(pr))]})
    lpn = rs["lenpen"]
    sp = make_sampling_params(greedy=True, max_tokens=lpn["max_tokens"], num_logprobs=lpn["logprobs"], eos_token_id=EOS,
                              length_penalty=(lpn["start"], lpn["decay"]))
    outs = eng.generate_sync(lpn["prompts"], sp)
    lenpen = [{"tokens": [r.new_token for r in recs if r.new_token is not None],
               "finish_reason": recs[-1].finish_reason,
               "steps": [entry(r) for r in recs if r.new_token is not None]} for recs in outs]
    eng.close()
    return {"greedy": greedy, "plp": plp, "lenpen": lenpen}


def engine_teacher_forced(cfg_name: str, seqs: list[list[int]], n_prompt: list[int]) -> list[list[dict]]:
    """logprob / rank of tokens seqs[i][n_prompt[i]:] given their prefix, through the engine's prompt-logprob pass."""
    from vllm_tgis_adapter_b200.engine.core import make_sampling_params

    eng = make_engine(cfg_name)
    sp = make_sampling_params(greedy=True, max_tokens=1, num_logprobs=1, prompt_logprobs=1, eos_token_id=EOS)
    outs = eng.generate_sync(seqs, sp)
    eng.close()
    res = []
    for s, n, recs in zip(seqs, n_prompt, outs):
        pos = {r.prompt_pos: {"logprob": r.logprob, "rank": r.rank} for r in recs if r.prompt_pos >= 1}
        res.append([pos[i] for i in range(n, len(s))])
    return res


def oracle_run(cfg_name: str, vllm_res: dict) -> dict:
    """Oracle: greedy continuation teacher-forced on vLLM's tokens (per-step oracle argmax, margin, logprob/rank of the
    forced token) and the prompt-logprob set."""
    import torch

    from oracle.llama_oracle import LlamaOracle, synthetic_weights

    cfg = configs()[cfg_name]
    if is_opt(cfg_name):
        from oracle.opt_oracle import OPTOracle, synthetic_opt_weights

        ora = OPTOracle(cfg, synthetic_opt_weights(cfg, seed=WEIGHT_SEED))
    else:
        w = synthetic_weights(cfg, seed=WEIGHT_SEED)
        ora = LlamaOracle(cfg, w)
    rs = request_sets(cfg_name, cfg.vocab)
    torch.set_num_threads(min(32, os.cpu_count() or 8))
    greedy = []
    for p, vr in zip(rs["greedy"]["prompts"], vllm_res["greedy"]):
        st = ora.new_seq()
        logits = ora.step([(st, p)])[0]
        steps = []
        for tok in vr["tokens"]:
            lp = torch.log_softmax(logits, -1)
            top2 = torch.topk(logits, 2).values
            steps.append({"argmax": int(torch.argmax(logits)), "margin": float(top2[0] - top2[1]),
                          "logprob": float(lp[tok]), "rank": int((lp >= lp[tok]).sum())})
            logits = ora.step([(st, [tok])])[0]
        greedy.append({"steps": steps})
    plp = []
    for p in rs["plp"]["prompts"]:
        logits = ora.step([(ora.new_seq(), p)], want_all_logits=True)
        lp = torch.log_softmax(logits, -1)
        pos = []
        for i in range(1, len(p)):
            row = lp[i - 1]
            pos.append({"logprob": float(row[p[i]]), "rank": int((row >= row[p[i]]).sum()),
                        "argmax": int(torch.argmax(row))})
        plp.append({"positions": pos})
    return {"greedy": greedy, "plp": plp}


# ======================================================================================================== compare
def _stats(diffs: list[float]) -> dict:
    import numpy as np

    a = np.asarray(diffs, dtype=np.float64)
    if a.size == 0:
        return {"n": 0}
    return {"n": int(a.size), "mean": float(a.mean()), "p50": float(np.percentile(a, 50)),
            "p95": float(np.percentile(a, 95)), "max": float(a.max()), "frac_below_1e-3": float((a < 1e-3).mean()),
            "frac_exact": float((a == 0).mean())}


def compare(cfg_name: str, v: dict, e: dict, o: dict, e_tf: list[list[dict]]) -> dict:
    """Tables for DESIGN.md section 5.  Teacher-forced logprobs: every stack scores vLLM's own (batched-run) tokens."""
    out: dict = {"config": cfg_name}
    # ---- greedy generation: free-running token agreement
    for name, other in (("engine", e["greedy"]), ("vllm_single", v["greedy_single"])):
        n_same_prefix, n_total, first_div = 0, 0, []
        for a, b in zip(v["greedy"], other):
            ta, tb = a["tokens"], b["tokens"]
            k = 0
            while k < min(len(ta), len(tb)) and ta[k] == tb[k]:
                k += 1
            n_same_prefix += k
            n_total += len(ta)
            first_div.append(k if k < len(ta) else None)
        out[f"greedy_tokens_vllm_vs_{name}"] = {"matching_prefix_tokens": n_same_prefix, "total": n_total,
                                               "first_divergence_step_per_request": first_div}
    # decode path, free-running: the engine's own sampled-token logprobs/ranks vs vLLM's while the prefixes are identical
    d_free, rank_free = [], 0
    for a, b in zip(v["greedy"], e["greedy"]):
        for k, (x, y) in enumerate(zip(a["tokens"], b["tokens"])):
            if x != y:
                break
            d_free.append(abs(a["steps"][k]["logprob"] - b["steps"][k]["logprob"]))
            rank_free += int(a["steps"][k]["rank"] != b["steps"][k]["rank"])
    out["decode_free_running_engine_vs_vllm"] = {"logprob_absdiff": _stats(d_free), "rank_mismatch": rank_free}
    # margins at the engine's first divergences (from the oracle's teacher-forced run on vLLM's tokens)
    div_margins = []
    for a, b, os_ in zip(v["greedy"], e["greedy"], o["greedy"]):
        for k, (x, y) in enumerate(zip(a["tokens"], b["tokens"])):
            if x != y:
                div_margins.append(os_["steps"][k]["margin"])
                break
    out["engine_first_divergence_oracle_margins"] = div_margins
    # ---- teacher-forced logprobs on vLLM's tokens
    d_e, d_o, d_vv, rank_e, rank_o, rank_vv, n = [], [], [], 0, 0, 0, 0
    for a, s_, tf, os_ in zip(v["greedy"], v["greedy_single"], e_tf, o["greedy"]):
        for k, st in enumerate(a["steps"]):
            d_e.append(abs(st["logprob"] - tf[k]["logprob"]))
            d_o.append(abs(st["logprob"] - os_["steps"][k]["logprob"]))
            rank_e += int(st["rank"] != tf[k]["rank"])
            rank_o += int(st["rank"] != os_["steps"][k]["rank"])
            n += 1
            if k < len(s_["tokens"]) and s_["tokens"][:k + 1] == a["tokens"][:k + 1]:
                d_vv.append(abs(st["logprob"] - s_["steps"][k]["logprob"]))
                rank_vv += int(st["rank"] != s_["steps"][k]["rank"])
    out["decode_logprob_absdiff"] = {"engine_vs_vllm": _stats(d_e), "oracle_vs_vllm": _stats(d_o),
                                     "vllm_batch_vs_vllm_single": _stats(d_vv)}
    out["decode_rank_mismatch"] = {"engine_vs_vllm": rank_e, "oracle_vs_vllm": rank_o,
                                   "vllm_batch_vs_vllm_single": rank_vv, "steps": n}
    # ---- prompt logprobs (prefill path, teacher-forced by construction)
    d_e, d_o, d_vv, rank_e, rank_o, rank_vv, n = [], [], [], 0, 0, 0, 0
    for a, s_, b, c in zip(v["plp"], v["plp_single"], e["plp"], o["plp"]):
        for pa, ps, pb, pc in zip(a["positions"], s_["positions"], b["positions"], c["positions"]):
            d_e.append(abs(pa["logprob"] - pb["logprob"]))
            d_o.append(abs(pa["logprob"] - pc["logprob"]))
            d_vv.append(abs(pa["logprob"] - ps["logprob"]))
            rank_e += int(pa["rank"] != pb["rank"])
            rank_o += int(pa["rank"] != pc["rank"])
            rank_vv += int(pa["rank"] != ps["rank"])
            n += 1
    out["prompt_logprob_absdiff"] = {"engine_vs_vllm": _stats(d_e), "oracle_vs_vllm": _stats(d_o),
                                     "vllm_batch_vs_vllm_single": _stats(d_vv)}
    out["prompt_rank_mismatch"] = {"engine_vs_vllm": rank_e, "oracle_vs_vllm": rank_o,
                                   "vllm_batch_vs_vllm_single": rank_vv, "positions": n}
    # ---- ExpDecay length penalty through vLLM's V1 processor hook
    if "lenpen" in v:
        rows = []
        for a, b in zip(v["lenpen"], e["lenpen"]):
            rows.append({"vllm_n_tokens": len(a["tokens"]), "engine_n_tokens": len(b["tokens"]),
                         "vllm_finish": a["finish_reason"], "engine_finish": b["finish_reason"],
                         "tokens_equal": a["tokens"] == b["tokens"],
                         "vllm_last_is_eos": bool(a["tokens"] and a["tokens"][-1] == EOS),
                         "engine_last_is_eos": bool(b["tokens"] and b["tokens"][-1] == EOS)})
        out["lenpen"] = rows
    else:
        out["lenpen_error"] = v.get("lenpen_error")
    # ---- are vLLM's logits bf16-rounded?  Ties: rank > 1 for the greedy (argmax) token can only come from exact ties
    tie_steps = sum(1 for a in v["greedy"] for st in a["steps"] if st["rank"] > 1)
    out["vllm_greedy_steps_with_rank_gt_1"] = tie_steps
    return out


def cmd_check(args) -> None:
    SCRATCH.mkdir(parents=True, exist_ok=True)
    gold = ROOT / "tests" / "golden"
    outdir = ROOT / "gpurun_out"
    outdir.mkdir(exist_ok=True)
    summary = {}
    for name in args.configs:
        t0 = time.time()
        make_model_dir(name)
        print(f"[check] {name}: model dir ready ({time.time() - t0:.1f}s)", flush=True)
        vpath = outdir / f"vllm_{name}.json"
        r = subprocess.run([sys.executable, __file__, "vllm-run", name, str(vpath)], capture_output=True, text=True,
                           timeout=args.vllm_timeout)
        (outdir / f"vllm_run_{name}.log").write_text(r.stdout[-20000:] + "\n---- stderr ----\n" + r.stderr[-40000:])
        if r.returncode != 0 or not vpath.exists():
            print(f"[check] {name}: vLLM run FAILED rc={r.returncode}; see gpurun_out/vllm_run_{name}.log", flush=True)
            summary[name] = {"error": f"vllm run failed rc={r.returncode}", "stderr_tail": r.stderr[-2000:]}
            continue
        v = json.loads(vpath.read_text())
        print(f"[check] {name}: vLLM done ({time.time() - t0:.1f}s)", flush=True)
        e = engine_run(name)
        seqs = [p + a["tokens"] for p, a in zip(request_sets(name, configs()[name].vocab)["greedy"]["prompts"], v["greedy"])]
        n_prompt = [len(p) for p in request_sets(name, configs()[name].vocab)["greedy"]["prompts"]]
        e_tf = engine_teacher_forced(name, seqs, n_prompt)
        print(f"[check] {name}: engine done ({time.time() - t0:.1f}s)", flush=True)
        o = oracle_run(name, v)
        print(f"[check] {name}: oracle done ({time.time() - t0:.1f}s)", flush=True)
        summary[name] = compare(name, v, e, o, e_tf)
        summary[name]["vllm_meta"] = v["meta"]
        # fixture: vLLM's batched-run outputs (what tests compare against); the single-request run only feeds the
        # noise table
        fx = {"meta": v["meta"], "greedy": v["greedy"], "plp": v["plp"]}
        if "lenpen" in v:
            fx["lenpen"] = v["lenpen"]
        # logprob floats are kept to 6 decimals: 1e-6 is far below every tolerance in play
        (gold / f"vllm_{name}.json").write_text(json.dumps(_round(fx)))
        # only gpurun_out/ travels back from the GPU box: the fixture is copied from there into tests/golden/ and committed
        (outdir / "golden").mkdir(exist_ok=True)
        (outdir / "golden" / f"vllm_{name}.json").write_text(json.dumps(_round(fx)))
        tag = "_opt" if all(is_opt(n) for n in args.configs) else ""
        (outdir / f"vllm_crosscheck{tag}.json").write_text(json.dumps(summary, indent=1))
    print(json.dumps(summary, indent=1))


def _round(o):
    if isinstance(o, float):
        return round(o, 6)
    if isinstance(o, list):
        return [_round(x) for x in o]
    if isinstance(o, dict):
        return {k: _round(x) for k, x in o.items()}
    return o


# ======================================================================================================== vLLM bench
def cmd_bench_inner(args) -> None:
    """GPU vLLM 0.22.0 on BASELINE.json configs[1]/[2]-shaped jobs: llama3-8b dims, dummy (random) weights, default
    engine settings (torch.compile + CUDA graphs, its own attention backend, cuBLAS GEMMs) -- 'the Blackwell kernels to
    beat' (BASELINE.md section 4).  decode tokens/s = B * (G - 1) / (t[G tokens] - t[1 token])."""
    os.environ.setdefault("VLLM_ENABLE_V1_MULTIPROCESSING", "0")
    os.environ.setdefault("HF_HUB_OFFLINE", "1")
    os.environ.setdefault("TRANSFORMERS_OFFLINE", "1")
    os.environ.setdefault("VLLM_NO_USAGE_STATS", "1")
    import numpy as np
    import torch
    import vllm
    from vllm import LLM, SamplingParams

    from vllm_tgis_adapter_b200.engine.tokenizer import build_synthetic_tokenizer

    from oracle.llama_oracle import CONFIGS

    cfg = CONFIGS[args.model]
    d = SCRATCH / f"{args.model}_dummy"
    d.mkdir(parents=True, exist_ok=True)
    hf = {"architectures": ["LlamaForCausalLM"], "model_type": "llama", "torch_dtype": "bfloat16", "dtype": "bfloat16",
          "vocab_size": cfg.vocab, "hidden_size": cfg.hidden, "intermediate_size": cfg.ffn,
          "num_hidden_layers": cfg.n_layers, "num_attention_heads": cfg.n_q_heads,
          "num_key_value_heads": cfg.n_kv_heads, "head_dim": cfg.head_dim, "hidden_act": "silu",
          "rms_norm_eps": cfg.rms_eps, "rope_theta": cfg.rope_theta,
          "rope_parameters": {"rope_type": "default", "rope_theta": cfg.rope_theta}, "rope_scaling": None,
          "max_position_embeddings": 8192, "tie_word_embeddings": False, "attention_bias": False, "mlp_bias": False,
          "bos_token_id": 1, "eos_token_id": EOS, "pad_token_id": 0}
    (d / "config.json").write_text(json.dumps(hf))
    (d / "generation_config.json").write_text(json.dumps({"bos_token_id": 1, "eos_token_id": EOS}))
    build_synthetic_tokenizer(cfg.vocab).save_pretrained(str(d))
    P, G = args.prompt_len, args.gen_len
    t0 = time.time()
    llm = LLM(model=str(d), load_format="dummy", dtype="bfloat16", max_model_len=1024, seed=0,
              gpu_memory_utilization=0.6, enable_prefix_caching=False, max_num_seqs=max(args.batches),
              max_num_batched_tokens=args.max_batched_tokens, enforce_eager=bool(args.eager))
    t_load = time.time() - t0
    res = {"meta": {"vllm": vllm.__version__, "torch": torch.__version__, "gpu": torch.cuda.get_device_name(0),
                    "model": args.model, "weights": "load_format=dummy", "load_s": t_load, "enforce_eager": bool(args.eager),
                    "max_num_batched_tokens": args.max_batched_tokens,
                    "method": "llm.generate wall clock; decode tok/s = B*(G-1)/(t_G - t_1), median of 3"},
           "runs": []}
    rs = np.random.RandomState(1234)
    for B in args.batches:
        prompts = [{"prompt_token_ids": rs.randint(1000, cfg.vocab - 1000, size=P).tolist()} for _ in range(B)]

        def timed(n_tok):
            sp = SamplingParams(temperature=0.0, max_tokens=n_tok, min_tokens=n_tok, ignore_eos=True, detokenize=False)
            torch.cuda.synchronize()
            t = time.perf_counter()
            outs = llm.generate(prompts, sp, use_tqdm=False)
            torch.cuda.synchronize()
            dt = time.perf_counter() - t
            assert all(len(o.outputs[0].token_ids) == n_tok for o in outs)
            return dt

        timed(4)
        timed(G)   # warm-up (graph capture sizes, autotune)
        t1 = sorted(timed(1) for _ in range(3))[1]
        tg = sorted(timed(G) for _ in range(3))[1]
        run = {"batch": B, "prompt_len": P, "gen_len": G, "t_prefill_burst_s": t1, "t_job_s": tg,
               "decode_tokens_per_s": B * (G - 1) / (tg - t1), "job_output_tokens_per_s": B * G / tg,
               "decode_ms_per_step": 1e3 * (tg - t1) / (G - 1)}
        print("[vllm-bench]", json.dumps(run), flush=True)
        res["runs"].append(run)
    Path(args.out).write_text(json.dumps(res, indent=1))


def cmd_bench(args) -> None:
    outdir = ROOT / "gpurun_out"
    outdir.mkdir(exist_ok=True)
    out = outdir / "vllm_baseline.json"
    cmd = [sys.executable, __file__, "bench-inner", "--model", args.model, "--out", str(out), "--prompt-len",
           str(args.prompt_len), "--gen-len", str(args.gen_len), "--max-batched-tokens", str(args.max_batched_tokens),
           "--eager", str(args.eager), "--batches", *map(str, args.batches)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=args.vllm_timeout)
    (outdir / "vllm_bench.log").write_text(r.stdout[-20000:] + "\n---- stderr ----\n" + r.stderr[-40000:])
    print(r.stdout[-3000:])
    if r.returncode != 0:
        print("[vllm-bench] FAILED rc=", r.returncode, r.stderr[-3000:])


def main() -> None:
    ap = argparse.ArgumentParser()
    sub = ap.add_subparsers(dest="cmd", required=True)
    c = sub.add_parser("check")
    c.add_argument("--configs", nargs="+", default=["tiny", "small", "8b2l"])
    c.add_argument("--vllm-timeout", type=int, default=900)
    v = sub.add_parser("vllm-run")
    v.add_argument("config")
    v.add_argument("out")
    for name in ("bench", "bench-inner"):
        b = sub.add_parser(name)
        b.add_argument("--model", default="llama3-8b")
        b.add_argument("--batches", nargs="+", type=int, default=[32, 64])
        b.add_argument("--prompt-len", type=int, default=512)
        b.add_argument("--gen-len", type=int, default=128)
        b.add_argument("--max-batched-tokens", type=int, default=2048)
        b.add_argument("--eager", type=int, default=0)
        b.add_argument("--vllm-timeout", type=int, default=1200)
        b.add_argument("--out", default=str(ROOT / "gpurun_out" / "vllm_baseline.json"))
    args = ap.parse_args()
    if args.cmd == "check":
        cmd_check(args)
    elif args.cmd == "vllm-run":
        vllm_run(args.config, args.out)
    elif args.cmd == "bench":
        cmd_bench(args)
    else:
        cmd_bench_inner(args)


if __name__ == "__main__":
    main()
