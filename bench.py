#!/usr/bin/env python
"""bench.py — decode tokens/sec (+ p50 TTFT) of the TGIS hot path on B200, per the driver contract.

A "step" = one pass of the hot path over one batch of synthetic input = one whole batched generation job:
B concurrent requests x (512-token prompt -> 128 generated tokens), greedy, on synthetic seeded weights of the named
architecture (BASELINE.json configs[1]: Llama-3-8B bf16, 1xB200, batch 32, 512-in/128-out).

  value  : decode tokens/s with inputs resident in HBM — tokens sampled in pure-decode engine steps divided by the
           CUDA-event time of those steps (events on the engine's launching stream), max over ranks.
  e2e    : the same metric through the public C-ABI call with HOST buffers (tgis_engine_add_request /
           run_until_idle / poll): wall clock of the decode phase, i.e. including the per-step H2D metadata copy, the
           D2H result copy and the host scheduler.  Also reports whole-job output tokens/s and p50 TTFT.
  roofline: dominant kernel = the tcgen05 GEMM; achieved = algorithmic bytes / CUDA-event time per launch, measured in
           a separate short profiled pass (events around every GEMM launch), against MEASURED_PEAKS.json hbm_gbs.
  cpu_baseline / --impl reference: the CPU oracle (oracle/llama_oracle.py — the port of the reference's vLLM-CPU path,
           which cannot be built offline) timed on the host cores on a bounded sample.

N > 1 (torchrun): data-parallel replicas, one engine per GPU, no data-path collective (requests are independent:
SURVEY.md §8e (1)); weak scaling.  (The engine also has tensor parallelism — DESIGN.md §6 — which the 8B config
does not need: it fits one GPU, so replicas are the faster way to use N GPUs for it.)
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--model", default="llama3-8b")
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--prompt-len", type=int, default=512)
    ap.add_argument("--gen-len", type=int, default=128)
    ap.add_argument("--parallel", default="dp", choices=["dp", "tp"],
                    help="N>1: dp = independent replicas (default, weak scaling); tp = ONE tensor-parallel engine over N GPUs")
    ap.add_argument("--max-batched-tokens", type=int, default=2048,
                    help="scheduler token budget per step (= chunked-prefill size); 2048 = this engine's and vLLM's "
                         "online-serving default")
    ap.add_argument("--sampling", default="greedy", choices=["greedy", "cfg3"],
                    help="cfg3 = repetition penalty 1.2 + length penalty (64, 1.05) + typical_p 0.9 sampling")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-decode-steps", type=int, default=3)
    ap.add_argument("--cpu-layers", type=int, default=2, help="layers timed by the CPU baseline (extrapolated)")
    ap.add_argument("--cpu-budget-s", type=float, default=45.0, help="wall-clock bound of the CPU baseline sample")
    return ap.parse_args()


_T0 = time.time()


def log(msg: str) -> None:
    print(f"[bench +{time.time() - _T0:7.1f}s] {msg}", file=sys.stderr, flush=True)


def peaks() -> tuple[float, str]:
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        return float(json.loads(p.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index: int):
        self.index = index
        self.samples: list[tuple[float, float, str]] = []
        self._stop = threading.Event()
        self._t: threading.Thread | None = None

    def _run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i",
                                      str(self.index)], capture_output=True, text=True, timeout=5).stdout.strip()
                f = [x.strip() for x in out.split(",")]
                reasons = [n for n, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown",
                                              "sw_power_cap"), f[2:6]) if v.lower().startswith("active")]
                self.samples.append((float(f[0]), float(f[1]), ",".join(reasons)))
            except Exception:  # noqa: BLE001
                pass
            self._stop.wait(0.2)

    def __enter__(self):
        self._t = threading.Thread(target=self._run, daemon=True)
        self._t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        if self._t:
            self._t.join(timeout=6)

    def summary(self) -> dict:
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        reasons = sorted({r for s in self.samples for r in s[2].split(",") if r})
        return {"sm_mhz": statistics.median(s[0] for s in self.samples), "sm_max_mhz": self.samples[0][1],
                "reasons": reasons, "samples": len(self.samples)}


# ----------------------------------------------------------------------------------------------------- CPU reference
def cpu_reference(args, steps: int, warmup: int) -> dict:
    """The oracle port of the reference's CPU path on the host cores, bounded sample: B sequences with a
    prompt_len-token (synthetic) KV history; `cpu_layers` of the model's layers + final norm + lm_head are timed over
    batched decode steps and the per-layer time is extrapolated linearly to the full depth (stated in `sample`)."""
    import dataclasses

    import torch

    from oracle.llama_oracle import CONFIGS, LlamaOracle, SeqState

    cfg_full = CONFIGS[args.model]
    L_s = min(cfg_full.n_layers, args.cpu_layers)
    cfg = dataclasses.replace(cfg_full, n_layers=L_s)
    cores = os.cpu_count() or 1
    g = torch.Generator().manual_seed(0)
    # vLLM's CPU backend runs bf16 where the host has AMX / avx512_bf16 and fp32 otherwise: pick by a micro-benchmark
    xa, wa = torch.randn(32, 4096), torch.randn(4096, 4096)

    def _t(dt):
        x_, w_ = xa.to(dt), wa.to(dt)
        torch.nn.functional.linear(x_, w_)
        t0 = time.perf_counter()
        for _ in range(3):
            torch.nn.functional.linear(x_, w_)
        return time.perf_counter() - t0

    torch.set_num_threads(min(cores, 32))
    t16, t32 = _t(torch.bfloat16), _t(torch.float32)
    dtype = torch.bfloat16 if t16 <= 1.5 * t32 else torch.float32
    log(f"cpu reference: linear micro-bench bf16 {1e3 * t16:.1f} ms vs fp32 {1e3 * t32:.1f} ms -> {dtype}")
    tile = (torch.randn(1024, 1024, generator=g) * 0.02).to(dtype)

    def fake(rows: int, cols: int) -> torch.Tensor:  # values irrelevant for timing; avoids minutes of randn
        r = (rows + 1023) // 1024
        c = (cols + 1023) // 1024
        return tile.repeat(r, c)[:rows, :cols].contiguous()

    w = {"model.embed_tokens.weight": fake(cfg.vocab, cfg.hidden), "lm_head.weight": fake(cfg.vocab, cfg.hidden),
         "model.norm.weight": torch.ones(cfg.hidden, dtype=dtype)}
    for i in range(cfg.n_layers):
        p = f"model.layers.{i}."
        w[p + "self_attn.q_proj.weight"] = fake(cfg.q_dim, cfg.hidden)
        w[p + "self_attn.k_proj.weight"] = fake(cfg.kv_dim, cfg.hidden)
        w[p + "self_attn.v_proj.weight"] = fake(cfg.kv_dim, cfg.hidden)
        w[p + "self_attn.o_proj.weight"] = fake(cfg.hidden, cfg.q_dim)
        w[p + "mlp.gate_proj.weight"] = fake(cfg.ffn, cfg.hidden)
        w[p + "mlp.up_proj.weight"] = fake(cfg.ffn, cfg.hidden)
        w[p + "mlp.down_proj.weight"] = fake(cfg.hidden, cfg.ffn)
        w[p + "input_layernorm.weight"] = torch.ones(cfg.hidden, dtype=dtype)
        w[p + "post_attention_layernorm.weight"] = torch.ones(cfg.hidden, dtype=dtype)
    ora = LlamaOracle(cfg, w, dtype=dtype)
    del w
    B, ctx = args.batch, args.prompt_len

    def fresh_states():
        sts = []
        for _ in range(B):
            st = SeqState(cfg, dtype)
            st.k = [fake(ctx, cfg.kv_dim).view(ctx, cfg.n_kv_heads, cfg.head_dim) for _ in range(cfg.n_layers)]
            st.v = [fake(ctx, cfg.kv_dim).view(ctx, cfg.n_kv_heads, cfg.head_dim) for _ in range(cfg.n_layers)]
            st.n = ctx
            sts.append(st)
        return sts

    def decode_steps(n: int) -> float:
        sts = fresh_states()
        toks = [5] * B
        t0 = time.perf_counter()
        for _ in range(n):
            logits = ora.step([(st, [t]) for st, t in zip(sts, toks)])
            toks = torch.argmax(logits, dim=-1).tolist()
        return (time.perf_counter() - t0) / n

    # thread count: all host threads unless fewer are faster (OpenMP fork/join cost on very wide hosts)
    best_thr, best_t = cores, None
    for thr in sorted({min(cores, 16), min(cores, 32), min(cores, 64), cores}):
        torch.set_num_threads(thr)
        t = decode_steps(1)
        log(f"cpu reference: {thr} threads -> {t:.3f} s per {L_s}-layer decode step (probe)")
        if best_t is None or t < best_t:
            best_thr, best_t = thr, t
        elif t > 1.5 * best_t:
            break   # wider is getting slower (fork/join cost): do not pay for the even wider probes
    torch.set_num_threads(best_thr)
    nd = args.cpu_decode_steps
    times = []
    t_budget = time.perf_counter() + args.cpu_budget_s
    for it in range(warmup + steps):
        t = decode_steps(nd)
        log(f"cpu reference: iteration {it}: {t:.3f} s per {L_s}-layer decode step")
        if it >= warmup or time.perf_counter() > t_budget:
            times.append(t)
        if time.perf_counter() > t_budget:
            break
    t_sample = sum(times) / len(times)
    # head (final norm + lm_head + argmax) timed alone so the layer part can be scaled to the full depth
    xh = torch.randn(B, cfg.hidden).to(dtype)
    ora.head(xh)
    t0 = time.perf_counter()
    for _ in range(3):
        ora.head(xh)
    t_head = (time.perf_counter() - t0) / 3
    t_layer = max(t_sample - t_head, 0.0) / L_s
    t_full = t_head + t_layer * cfg_full.n_layers
    val = B / t_full
    return {"value": val, "unit": "tokens/s", "cores": best_thr, "kind": "port",
            "sample": f"{args.model} B={B} ctx={ctx} (synthetic KV history): {L_s} of {cfg_full.n_layers} layers + lm_head "
                      f"timed over {nd * len(times)} batched decode steps ({t_sample * 1e3:.0f} ms/step, head "
                      f"{t_head * 1e3:.0f} ms), per-layer time extrapolated linearly to {cfg_full.n_layers} layers; torch "
                      f"CPU {str(dtype).split('.')[-1]} oracle (oracle/llama_oracle.py), {best_thr} of {cores} threads",
            "ms_per_step": 1e3 * t_full}


def ncu_traffic_per_launch(n_layers: int):
    """DRAM bytes (read + write) per launch of the dominant kernel from the committed `ncu --set full` capture
    (profiles/, produced by scripts/gpu_check.sh prof_gemm: the qkv, o, gate_up, down and lm_head launches of one decode
    step at batch 32), weighted like the timed mix: n_layers x the four layer GEMMs + one lm_head."""
    import csv
    import glob
    files = sorted(glob.glob(str(Path(__file__).resolve().parent / "profiles" / "r*_ncu_full_gemm_tcgen05.csv")))
    if not files:
        return None, None
    try:
        with open(files[-1]) as f:
            rows = list(csv.DictReader(f))
        def col(row, name):
            for k, v in row.items():
                if k.startswith(name):
                    unit = k[k.index("[") + 1:k.index("]")] if "[" in k else "byte"
                    mult = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}[unit]
                    return float(v) * mult
            raise KeyError(name)
        per = [col(r, "dram__bytes_read.sum") + col(r, "dram__bytes_write.sum") for r in rows]
        if len(per) != 5:
            return None, None
        total = n_layers * sum(per[:4]) + per[4]
        return total / (4 * n_layers + 1), f"{Path(files[-1]).name} (ncu --set full, cold-cache replay)"
    except Exception:  # noqa: BLE001
        return None, None


def reduce_over_ranks(max_vals: list[float], sum_vals: list[float], device: str) -> tuple[list[float], list[float]]:
    """Whole-job aggregation for the N>1 (data-parallel replicas) path: times -> MAX over ranks, counts -> SUM.
    Works on any initialised torch.distributed backend (NCCL on the GPU box, gloo in tests/test_multirank_cpu.py)."""
    import torch
    import torch.distributed as dist

    vals = torch.tensor(max_vals, dtype=torch.float64, device=device)
    sums = torch.tensor(sum_vals, dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(vals, op=dist.ReduceOp.MAX)
        dist.all_reduce(sums, op=dist.ReduceOp.SUM)
    return vals.tolist(), sums.tolist()


# ----------------------------------------------------------------------------------------------------------- ours
def run_ours(args) -> dict | None:
    import torch

    from vllm_tgis_adapter_b200.engine.core import PRESETS, ModelConfig, NativeEngine, make_sampling_params

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        import torch.distributed as dist

        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    import dataclasses

    B, P, G = args.batch, args.prompt_len, args.gen_len
    cfg = mc = dataclasses.replace(PRESETS[args.model], max_model_len=max(1024, P + G + 64))
    cfg.q_dim, cfg.kv_dim = cfg.n_q_heads * 128, cfg.n_kv_heads * 128
    blocks = B * ((P + G + 31) // 32 + 2)
    kv_bytes = int(blocks * 2 * cfg.n_layers * cfg.n_kv_heads * 32 * 128 * 2 * 1.1)
    torch.cuda.set_device(local)
    tp = world if (args.parallel == "tp" and world > 1) else 1
    tp_kw = {}
    if tp > 1:
        ids = [NativeEngine.nccl_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(ids, src=0, device=torch.device("cuda", local))
        tp_kw = dict(tp_size=tp, tp_rank=rank, nccl_id=ids[0], shm_name=f"/tgis_bench_{os.environ.get('MASTER_PORT', '0')}")
        kv_bytes //= tp
    eng = NativeEngine(mc, max_num_seqs=B, max_batched_tokens=args.max_batched_tokens, kv_cache_bytes=kv_bytes, device=local, seed=1234,
                       **tp_kw)
    # synthetic N(0, 0.02) weights generated on the device, one tensor at a time ("PyTorch tensors for weights only");
    # under tp every rank draws the SAME full tensor (same seed) and the engine keeps its shard
    gen = torch.Generator(device="cuda").manual_seed(1234 + (rank if tp == 1 else 0))

    def rnd(r, c):
        return (torch.randn(r, c, generator=gen, device="cuda", dtype=torch.float32) * 0.02).to(torch.bfloat16)

    ones = torch.ones(cfg.hidden, dtype=torch.bfloat16, device="cuda")
    eng.load_weight("model.embed_tokens.weight", rnd(cfg.vocab, cfg.hidden))
    eng.load_weight("lm_head.weight", rnd(cfg.vocab, cfg.hidden))
    eng.load_weight("model.norm.weight", ones)
    for i in range(cfg.n_layers):
        p = f"model.layers.{i}."
        eng.load_weight(p + "self_attn.q_proj.weight", rnd(cfg.q_dim, cfg.hidden))
        eng.load_weight(p + "self_attn.k_proj.weight", rnd(cfg.kv_dim, cfg.hidden))
        eng.load_weight(p + "self_attn.v_proj.weight", rnd(cfg.kv_dim, cfg.hidden))
        eng.load_weight(p + "self_attn.o_proj.weight", rnd(cfg.hidden, cfg.q_dim))
        eng.load_weight(p + "mlp.gate_proj.weight", rnd(cfg.ffn, cfg.hidden))
        eng.load_weight(p + "mlp.up_proj.weight", rnd(cfg.ffn, cfg.hidden))
        eng.load_weight(p + "mlp.down_proj.weight", rnd(cfg.hidden, cfg.ffn))
        eng.load_weight(p + "input_layernorm.weight", ones)
        eng.load_weight(p + "post_attention_layernorm.weight", ones)
    torch.cuda.synchronize()
    torch.cuda.empty_cache()
    log(f"engine built and {cfg.n_layers}-layer synthetic weights loaded")
    if tp > 1 and rank != 0:   # tensor-parallel worker: follow rank 0's step plans until it closes its engine
        eng.worker_run()
        eng.close()
        dist.barrier()
        dist.destroy_process_group()
        return None

    import numpy as np

    rs = np.random.RandomState(1234 + rank)
    prompts = [rs.randint(1000, cfg.vocab - 1000, size=P).tolist() for _ in range(B)]
    if args.sampling == "cfg3":   # BASELINE.json configs[2]: repetition penalty + ExpDecayLengthPenalty + typical-p sampling
        sp = make_sampling_params(greedy=False, temperature=1.0, typical_p=0.9, repetition_penalty=1.2,
                                  length_penalty=(64, 1.05), seed=1234, max_tokens=G, min_tokens=G, eos_token_id=2)
    else:
        sp = make_sampling_params(greedy=True, max_tokens=G, min_tokens=G, eos_token_id=2)

    def job():
        """One bench step through the C ABI with host buffers.  Returns (decode_wall_s, ttfts, n_tokens,
        n_tokens produced inside the decode wall): the decode phase starts when the LAST request has its first token;
        tokens that early requests produced before that (chunked prefill interleaves them) are not counted in it."""
        for i, pr in enumerate(prompts):
            eng.add_request(f"r{i}", pr, sp)
        eng.run_until_idle()
        t_end = time.monotonic()
        first, n_tok, stamps = {}, 0, []
        while True:
            outs = eng.poll(0)
            if not outs:
                break
            for o in outs:
                if o.new_token is not None:
                    n_tok += 1
                    stamps.append(o.ts_last_token)
                first[o.request_id] = (o.ts_first_token, o.ts_arrival)
        t_all_first = max(v[0] for v in first.values())
        ttfts = [v[0] - v[1] for v in first.values()]
        return t_end - t_all_first, ttfts, n_tok, sum(1 for t in stamps if t > t_all_first)

    def barrier():
        torch.cuda.synchronize()
        if world > 1 and tp == 1:
            dist.barrier()

    for i in range(args.warmup):
        dw, tt, nt, _ = job()
        log(f"warmup job {i}: {nt} tokens, decode wall {dw:.3f}s, ttft p50 {1e3 * statistics.median(tt):.1f} ms")
    barrier()
    st0 = eng.status()
    decode_wall, ttfts, n_tok, n_tok_dec = 0.0, [], 0, 0
    with ClockSampler(local) as clocks:
        t0 = time.perf_counter()
        for _ in range(args.steps):
            dw, tt, nt, nd = job()
            decode_wall += dw
            ttfts += tt
            n_tok += nt
            n_tok_dec += nd
        barrier()
        wall = time.perf_counter() - t0
    st1 = eng.status()
    log(f"timed region done: {args.steps} jobs in {wall:.3f}s")
    dec_ms = st1.gpu_decode_ms - st0.gpu_decode_ms
    dec_tok = st1.decode_tokens - st0.decode_tokens
    dec_steps = st1.decode_steps - st0.decode_steps
    launches = st1.kernel_launches - st0.kernel_launches
    h2d = (st1.h2d_bytes - st0.h2d_bytes) / args.steps
    d2h = (st1.d2h_bytes - st0.d2h_bytes) / args.steps

    # ---- roofline leg: profiled short pass (events around every GEMM launch)
    eng.set_profiling(2)   # decode steps only: the HBM-bound regime the roofline is quoted for
    sp_short = make_sampling_params(greedy=True, max_tokens=17, min_tokens=17, eos_token_id=2)
    g0 = eng.status()
    for i, pr in enumerate(prompts):
        eng.add_request(f"p{i}", pr, sp_short)
    eng.run_until_idle()
    while eng.poll(0):
        pass
    g1 = eng.status()
    eng.set_profiling(False)
    gemm_ms = g1.gemm_ms - g0.gemm_ms
    gemm_bytes = g1.gemm_bytes - g0.gemm_bytes
    gemm_calls = g1.gemm_calls - g0.gemm_calls
    log(f"profiled pass done: {gemm_calls} GEMM launches, {gemm_ms:.2f} ms")

    # ---- reduce over ranks (max time, sum tokens)
    if tp > 1:   # one engine: rank 0 holds the whole-job numbers
        (dec_ms_m, decode_wall_m, wall_m), (dec_tok_s, n_tok_s, launches_s, n_tok_dec_s) = (
            [dec_ms, decode_wall, wall], [dec_tok, n_tok, launches, n_tok_dec])
        eng.close()
        dist.barrier()
        dist.destroy_process_group()
    else:
        (dec_ms_m, decode_wall_m, wall_m), (dec_tok_s, n_tok_s, launches_s, n_tok_dec_s) = reduce_over_ranks(
            [dec_ms, decode_wall, wall], [dec_tok, n_tok, launches, n_tok_dec], "cuda" if world > 1 else "cpu")
        eng.close()
        if world > 1:
            dist.destroy_process_group()
    if rank != 0:
        return None
    n_rep = 1 if tp > 1 else world   # independent replicas
    peak, peak_src = peaks()
    n_params = cfg.n_layers * (cfg.hidden * (cfg.q_dim + 2 * cfg.kv_dim) + cfg.hidden * cfg.q_dim
                               + 3 * cfg.hidden * cfg.ffn) + cfg.vocab * cfg.hidden
    kv_tok = 2 * cfg.n_layers * cfg.n_kv_heads * 128 * 2
    # algorithmic bytes of one decode step PER GPU: weights/tp + KV/tp + one fp32 logits scan
    bytes_step = (n_params * 2 + B * (P + G / 2) * kv_tok) / tp + B * cfg.vocab * 4
    step_ms = dec_ms / max(dec_steps, 1)
    achieved = gemm_bytes / (gemm_ms * 1e-3) / 1e9 if gemm_ms > 0 else None
    traffic, traffic_src = ncu_traffic_per_launch(cfg.n_layers) if args.model == "llama3-8b" and tp == 1 else (None, None)
    out = {
        "metric": "decode tokens/sec + p50 TTFT, 512-in/128-out batch, 1/2/4/8xB200 vs CPU ref",
        "value": dec_tok_s / (dec_ms_m * 1e-3), "unit": "tokens/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * wall_m / args.steps, "higher_is_better": True,
        "scaling": "strong" if tp > 1 else "weak",
        "vs_baseline": None, "dtype": "bf16", "data": "synthetic (seeded N(0,0.02) weights, uniform random prompts)",
        "config": {"workload": f"{args.model} bf16, {B} concurrent requests/GPU, {P}-in/{G}-out, {args.sampling} "
                               f"(BASELINE.json configs[1])", "batch_per_gpu": B, "prompt_len": P, "gen_len": G,
                   "parallelism": (f"tp{tp} (one engine, NCCL all-reduce after o/down proj, all-gather of logits)"
                                   if tp > 1 else f"dp{world} (independent replicas, no collective)"),
                   "scheduler": f"continuous batching, chunked prefill, {args.max_batched_tokens} tokens per step",
                   "l2": "inputs larger than L2 (15 GB of weights streamed per decode step)",
                   "timing": "value: CUDA events on the engine stream over pure-decode steps; e2e: wall clock"},
        "e2e": {"value": n_tok_dec_s / decode_wall_m if decode_wall_m > 0 else None,
                "unit": "tokens/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "job_output_tokens_per_s": n_tok_s / wall_m, "ttft_p50_ms": 1e3 * statistics.median(ttfts),
                "ttft_max_ms": 1e3 * max(ttfts),
                "note": "through the C ABI with host buffers: add_request/run_until_idle/poll; tokens stamped after the last "
                        "request's first token / wall clock from that moment to the end of the job"},
        "gpu_launches": int(launches_s),
        "clocks": clocks.summary(),
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": (achieved / peak) if achieved else None, "traffic": traffic, "traffic_source": traffic_src,
                     "peak_source": peak_src,
                     "kernel": "gemm_bf16_tcgen05_kernel", "launches_timed": int(gemm_calls),
                     "algorithmic_bytes_per_launch": gemm_bytes / max(gemm_calls, 1),
                     "avg_launch_us": 1e3 * gemm_ms / max(gemm_calls, 1),
                     "decode_step_ms": step_ms, "decode_step_algorithmic_bytes": bytes_step,
                     "decode_step_frac_of_hbm_roofline": bytes_step / (step_ms * 1e-3) / 1e9 / peak if step_ms else None},
    }
    return out


def main():
    args = parse_args()
    rank = int(os.environ.get("RANK", "0"))
    if args.impl == "reference":
        if rank != 0:
            return
        ref = cpu_reference(args, args.steps, args.warmup)
        world = int(os.environ.get("WORLD_SIZE", "1"))
        print(json.dumps({
            "impl": "reference", "metric": "decode tokens/sec + p50 TTFT, 512-in/128-out batch, 1/2/4/8xB200 vs CPU ref",
            "value": ref["value"], "unit": "tokens/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ref["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16", "data": "synthetic",
            "config": {"workload": f"{args.model} bf16, {args.batch} concurrent requests, {args.prompt_len}-in, "
                                   f"decode steps on host cores (bounded sample)"},
            "cpu_baseline": {k: ref[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": ref["value"], "unit": "tokens/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}), flush=True)
        return
    out = run_ours(args)
    if out is None:
        return
    try:  # keep the GPU numbers even if the CPU baseline leg is cut short by a timeout
        (ROOT / "gpurun_out").mkdir(exist_ok=True)
        (ROOT / "gpurun_out" / "bench_ours_partial.json").write_text(json.dumps(out))
    except OSError:
        pass
    if not args.no_cpu_baseline and int(os.environ.get("WORLD_SIZE", "1")) == 1:
        ref = cpu_reference(args, steps=2, warmup=1)
        out["cpu_baseline"] = {k: ref[k] for k in ("value", "unit", "cores", "kind", "sample")}
    else:
        out["cpu_baseline"] = None
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
