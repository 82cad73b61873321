"""-m gpu: LoRA adapters on the device.  (1) lora.cu's shrink + expand against the oracle's restatement of vLLM's punica
arithmetic (oracle/llama_oracle.py::lora_add, pinned to vLLM's reference ops by tests/test_lora_cpu.py) on uniform and
mixed-adapter token tiles; (2) the engine through the C ABI with adapters loaded from PEFT directories: adapter requests
follow the oracle's LoRA forward, base-model requests that share their steps are bit-identical to a run without any
adapter, a zero adapter changes nothing; (3) the gRPC path with an adapter cache directory."""
import ctypes as C

import numpy as np
import pytest
import torch

from test_lora_cpu import make_peft_dir

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def g():
    import tgis_gpu_utils as g

    return g


@pytest.mark.parametrize("T,K,N,Rm,col0", [(37, 256, 512, 8, 0), (64, 4096, 1024, 16, 128), (9, 14336, 4096, 64, 0),
                                           (130, 512, 2048 + 64, 128, 64)])
def test_lora_kernels_match_oracle(g, T, K, N, Rm, col0):
    from oracle.llama_oracle import lora_add

    torch.manual_seed(T + K)
    slots = 3
    x = torch.randn(T, K).to(torch.bfloat16)
    ldy = col0 + N + 8
    y = torch.randn(T, ldy).to(torch.bfloat16)
    a = (torch.randn(slots, Rm, K) * 0.05).to(torch.bfloat16)
    b = (torch.randn(slots, N, Rm) * 0.2).to(torch.bfloat16)
    tok = torch.randint(0, slots + 1, (T,), dtype=torch.int32)
    tok[:16] = 2                      # two uniform 8-token tiles
    tok[16:24] = 0                    # a tile without any adapter
    xd, yd, ad, bd, td = x.cuda(), y.clone().cuda(), a.cuda(), b.cuda(), tok.cuda()
    rc = g.lib().tgis_k_lora(g.ptr(xd), K, g.ptr(td), g.ptr(ad), g.ptr(bd), K, N, Rm, col0, g.ptr(yd), ldy, T)
    assert rc == 0, g.kerr()
    got = yd.cpu()
    want = y.clone()
    for t in range(T):
        s = int(tok[t])
        if s > 0:
            want[t:t + 1, col0:col0 + N] = lora_add(y[t:t + 1, col0:col0 + N], x[t:t + 1], a[s - 1], b[s - 1])
    # untouched: tokens without an adapter, and every column outside the module's range
    assert torch.equal(got[tok == 0], y[tok == 0])
    assert torch.equal(got[:, :col0], y[:, :col0]) and torch.equal(got[:, col0 + N:], y[:, col0 + N:])
    d = g.bf16_ulp_diff(got[:, col0:col0 + N], want[:, col0:col0 + N])
    # two fp32 sums in a different order than torch's: a result may land one bf16 ulp away (of y, or of the delta)
    assert int(d.max()) <= 2 and float((d == 0).float().mean()) > 0.97, (int(d.max()), float((d == 0).float().mean()))


def test_silu_mul_interleaved_matches_split_layout(g):
    torch.manual_seed(3)
    T, F = 19, 1536
    gate, up = torch.randn(T, F).to(torch.bfloat16), torch.randn(T, F).to(torch.bfloat16)
    inter = torch.stack([gate, up], dim=-1).reshape(T, 2 * F).contiguous().cuda()
    act = torch.empty(T, F, dtype=torch.bfloat16, device="cuda")
    assert g.lib().tgis_k_silu_mul_interleaved(g.ptr(inter), g.ptr(act), T, F) == 0, g.kerr()
    want = torch.nn.functional.silu(gate) * up     # bf16 ops: silu rounded, product rounded
    assert torch.equal(act.cpu(), want)


# ---------------------------------------------------------------------------------------------------- engine level
def _engine(cfg_name, **kw):
    from oracle.llama_oracle import CONFIGS, rope_table, synthetic_weights
    from vllm_tgis_adapter_b200.engine.core import ModelConfig, NativeEngine

    cfg = CONFIGS[cfg_name]
    weights = synthetic_weights(cfg, seed=1)
    mc = ModelConfig(n_layers=cfg.n_layers, hidden=cfg.hidden, n_q_heads=cfg.n_q_heads, n_kv_heads=cfg.n_kv_heads,
                     ffn=cfg.ffn, vocab=cfg.vocab, rope_theta=cfg.rope_theta, rms_eps=cfg.rms_eps,
                     max_model_len=cfg.max_model_len)
    eng = NativeEngine(mc, **kw)
    eng.load_weights(weights)
    eng.load_weight("tgis.rope_cos_sin", rope_table(cfg))
    return cfg, weights, eng


def _peft_for(cfg, root, name, **kw):
    return make_peft_dir(root, name, n_layers=cfg.n_layers, hidden=cfg.hidden, q_dim=cfg.q_dim, kv_dim=cfg.kv_dim,
                         ffn=cfg.ffn, **kw)


def _run(eng, prompts, slots, n_new):
    from vllm_tgis_adapter_b200.engine.core import make_sampling_params

    sps = [make_sampling_params(greedy=True, max_tokens=n_new, min_tokens=n_new, num_logprobs=1, eos_token_id=2,
                                lora_slot=s) for s in slots]
    outs = eng.generate_sync(prompts, sps)
    return [[(r.new_token, r.logprob) for r in recs if r.new_token is not None] for recs in outs]


@pytest.mark.parametrize("cfg_name,chunk", [("tiny", 2048), ("small", 64)])
def test_engine_lora_matches_oracle_and_leaves_base_requests_untouched(tmp_path, cfg_name, chunk):
    from oracle.llama_oracle import CONFIGS, LlamaOracle
    from vllm_tgis_adapter_b200.engine.lora import read_adapter

    cfg = CONFIGS[cfg_name]
    da = _peft_for(cfg, tmp_path, "a", r=8, alpha=16, seed=1)
    db = _peft_for(cfg, tmp_path, "b", r=16, alpha=16, seed=2, modules=("q_proj", "v_proj", "down_proj"))
    dz = _peft_for(cfg, tmp_path, "z", r=8, alpha=16, seed=3, std=0.0)       # all-zero adapter
    ads = {n: read_adapter(str(d), n_layers=cfg.n_layers, max_rank=16)[1] for n, d in (("a", da), ("b", db), ("z", dz))}
    rng = np.random.RandomState(7)
    lens = [5, 33, 64, 100, 17, 70, 9]
    prompts = [rng.randint(3, cfg.vocab, size=n).tolist() for n in lens]
    n_new = 16
    kw = dict(max_num_seqs=8, max_batched_tokens=chunk, kv_cache_bytes=64 << 20)
    # run 1: no adapter anywhere (the fused production path)
    _, weights, eng = _engine(cfg_name, **kw)
    base = _run(eng, prompts, [0] * len(prompts), n_new)
    eng.close()
    # run 2: same prompts; requests 0,1 -> adapter a (slot 1), 2 -> b (slot 2), 3 -> zero adapter (slot 3), 4.. base
    _, _, eng = _engine(cfg_name, max_loras=3, max_lora_rank=16, **kw)
    eng.load_adapter(1, ads["a"])
    eng.load_adapter(2, ads["b"])
    eng.load_adapter(3, ads["z"])
    slots = [1, 1, 2, 3, 0, 0, 0]
    mixed = _run(eng, prompts, slots, n_new)
    st = eng.status()
    # slot reuse: overwrite slot 1 with adapter b and run request 2 there -- same result as from slot 2
    eng.load_adapter(1, ads["b"])
    again = _run(eng, [prompts[2]], [1], n_new)
    eng.close()
    assert st.errored == 0
    for i, s in enumerate(slots):
        if s in (0, 3):   # base-model and zero-adapter requests: bit-identical to the adapter-free run
            assert mixed[i] == base[i], i
    assert mixed[0] != base[0] and mixed[2] != base[2]        # the adapters do change the outputs
    # (a batch of one partitions the GEMMs differently from the batch of seven: same tokens up to bf16 near-ties, logprobs
    # within a bf16 ulp of the logits -- not bit-identical)
    assert again[0][0][0] == mixed[2][0][0] and abs(again[0][0][1] - mixed[2][0][1]) < 2e-2
    assert sum(int(x[0] == y[0]) for x, y in zip(again[0], mixed[2])) >= n_new // 2
    # adapter requests vs the oracle's LoRA forward, teacher-forced on the engine's tokens
    ora = LlamaOracle(cfg, weights)
    flips = total = 0
    diffs = []
    for i, name in ((0, "a"), (1, "a"), (2, "b")):
        st_ = ora.new_seq()
        logits = ora.step([(st_, prompts[i])], lora=[ads[name]])[0]
        for tok, lp in mixed[i]:
            olp = torch.log_softmax(logits, -1)
            top2 = torch.topk(logits, 2).values
            total += 1
            diffs.append(abs(lp - float(olp[tok])))
            if int(torch.argmax(logits)) != tok:
                assert float(top2[0] - top2[1]) < 0.03, (i, float(top2[0] - top2[1]))
                flips += 1
            logits = ora.step([(st_, [tok])], lora=[ads[name]])[0]
    diffs = np.array(diffs)
    assert flips <= max(1, total // 12), (flips, total)
    assert float(diffs.mean()) < 6e-3 and float(diffs.max()) < 4e-2, (float(diffs.mean()), float(diffs.max()))
    # and without the adapter the oracle does NOT explain those tokens (the test would otherwise be vacuous)
    st_ = ora.new_seq()
    logits = ora.step([(st_, prompts[0])])[0]
    off = 0
    for tok, lp in mixed[0]:
        off += int(int(torch.argmax(logits)) != tok)
        logits = ora.step([(st_, [tok])])[0]
    assert off >= 3, off


def test_engine_rejects_bad_lora_use(tmp_path):
    from vllm_tgis_adapter_b200.engine.core import EngineError, make_sampling_params
    from oracle.llama_oracle import CONFIGS
    from vllm_tgis_adapter_b200.engine.lora import read_adapter

    cfg = CONFIGS["tiny"]
    _, _, eng = _engine("tiny", max_num_seqs=4, max_batched_tokens=64, kv_cache_bytes=32 << 20)
    with pytest.raises(EngineError, match="lora_slot out of range"):
        eng.add_request("x", [5, 6], make_sampling_params(max_tokens=2, lora_slot=1))
    with pytest.raises(EngineError, match="without LoRA slots"):
        eng.load_adapter(1, {})
    eng.close()
    _, _, eng = _engine("tiny", max_num_seqs=4, max_batched_tokens=64, kv_cache_bytes=32 << 20, max_loras=1, max_lora_rank=8)
    big = read_adapter(str(_peft_for(cfg, tmp_path, "big", r=16)), n_layers=cfg.n_layers, max_rank=16)[1]
    with pytest.raises(EngineError, match="exceeds max_lora_rank"):
        eng.load_adapter(1, big)
    with pytest.raises(EngineError, match="slot out of range"):
        eng.load_adapter(2, {})
    eng.close()


def test_lora_generate_over_grpc_with_adapter_cache(tmp_path, monkeypatch):
    """adapter_id -> adapters.py -> LoRAManager slot -> lora.cu, through a real channel; unknown ids and prompt-tuning
    adapters come back as INVALID_ARGUMENT with the reference's strings."""
    import json

    import grpc
    from oracle.llama_oracle import CONFIGS, LlamaOracle
    from test_server_gpu import LiveServer
    from vllm_tgis_adapter_b200.engine.lora import read_adapter
    from vllm_tgis_adapter_b200.grpc.pb import generation_pb2 as pb

    cfg = CONFIGS["tiny"]
    cache = tmp_path / "cache"
    cache.mkdir()
    da = _peft_for(cfg, cache, "my-lora", r=8, alpha=16, seed=11)
    (cache / "pt").mkdir()
    (cache / "pt" / "adapter_config.json").write_text(json.dumps({"peft_type": "PROMPT_TUNING"}))
    monkeypatch.chdir(tmp_path)
    live = LiveServer(adapter_cache=str(cache), max_loras=2, max_lora_rank=8)
    try:
        call = live.channel.unary_unary("/fmaas.GenerationService/Generate",
                                        request_serializer=pb.BatchedGenerationRequest.SerializeToString,
                                        response_deserializer=pb.BatchedGenerationResponse.FromString)
        p = pb.Parameters()
        p.stopping.max_new_tokens = 12
        p.stopping.min_new_tokens = 12
        p.response.generated_tokens = True
        text = "t5 t6 t7 t100 t200 t300"
        ids = [5, 6, 7, 100, 200, 300]
        with_ad = call(pb.BatchedGenerationRequest(model_id="m", adapter_id="my-lora",
                                                   requests=[pb.GenerationRequest(text=text)], params=p), timeout=120)
        without = call(pb.BatchedGenerationRequest(model_id="m", requests=[pb.GenerationRequest(text=text)], params=p),
                       timeout=120)
        toks_ad = [int(t.text[1:]) for t in with_ad.responses[0].tokens]
        toks_base = [int(t.text[1:]) for t in without.responses[0].tokens]
        assert len(toks_ad) == 12 and toks_ad != toks_base
        ad = read_adapter(str(da), n_layers=cfg.n_layers, max_rank=8)[1]
        ora = LlamaOracle(live.cfg, live.weights)
        st = ora.new_seq()
        logits = ora.step([(st, ids)], lora=[ad])[0]
        for t in toks_ad:
            if int(torch.argmax(logits)) != t:
                top2 = torch.topk(logits, 2).values
                assert float(top2[0] - top2[1]) < 0.03
                break
            logits = ora.step([(st, [t])], lora=[ad])[0]
        assert live.engine._lora.loads == 1
        for bad, msg in (("nope", "can't retrieve adapter with id 'nope': directory does not exist"),
                         ("pt", "adapter type PROMPT_TUNING is not currently supported")):
            with pytest.raises(grpc.RpcError) as ei:
                call(pb.BatchedGenerationRequest(model_id="m", adapter_id=bad,
                                                 requests=[pb.GenerationRequest(text=text)], params=p), timeout=60)
            assert ei.value.code() == grpc.StatusCode.INVALID_ARGUMENT and ei.value.details() == msg
    finally:
        live.close()
