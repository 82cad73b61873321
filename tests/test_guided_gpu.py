"""-m gpu: guided decoding on the device.  (1) the fused sampling kernel under an allowed-token bitmask vs the oracle
(`oracle/sampler_oracle.py::apply_token_bitmask`, pinned to xgrammar's own kernel in tests/test_guided_cpu.py) on every
path of the kernel -- greedy / logprobs / top-n, every sampling stage, both logits dtypes, staged and unstaged slices;
(2) the engine through the C ABI with a real xgrammar mask provider: token ids equal to the oracle's constrained greedy
decode, the output is a member of the language, preemption does not double-feed the matcher; (3) the gRPC path."""
import re

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def g():
    import tgis_gpu_utils as g

    return g


def _rows(g, n):
    r = np.zeros(n, dtype=g.SAMPLE_ROW_DTYPE)
    r["temperature"], r["top_p"], r["rep_penalty"] = 1.0, 1.0, 1.0
    r["eos_id"] = 2
    r["seq_slot"] = np.arange(n)
    r["logits_row"] = np.arange(n)
    return r


def _random_allow(rng, n, V, densities):
    allow = np.zeros((n, V), dtype=bool)
    for i in range(n):
        d = densities[i % len(densities)]
        if d >= 1.0:
            allow[i] = True
        else:
            k = max(1, int(round(d * V)))
            allow[i, rng.choice(V, size=k, replace=False)] = True
    words = (V + 31) // 32
    packed = np.zeros((n, words * 32), dtype=bool)
    packed[:, :V] = allow
    bits = np.packbits(packed.reshape(n, words, 32), axis=-1, bitorder="little").view(np.uint32).reshape(n, words)
    return allow, torch.from_numpy(bits.view(np.int32).copy())


@pytest.mark.parametrize("V,n,dtype", [(4096, 8, torch.float32), (128256, 8, torch.bfloat16), (128256, 8, torch.float32),
                                       (128256, 160, torch.bfloat16)])
def test_sampler_greedy_under_mask_matches_oracle(g, V, n, dtype):
    """n = 160 rows at V = 128256: one CTA per row and a slice too large to stage -- the passes after the first re-read
    global memory and must re-apply the mask there."""
    from oracle.sampler_oracle import SamplingCase, sample_row

    rng = np.random.RandomState(V + n)
    torch.manual_seed(V + n)
    logits = (torch.randn(n, V) * 2).to(dtype)
    allow, bits = _random_allow(rng, n, V, [1e-4, 0.01, 0.5, 1.0, 3.0 / V])
    rows = _rows(g, n)
    rows["flags"] = g.SAMPLE_GREEDY | g.SAMPLE_LOGPROBS | g.SAMPLE_MASKED
    rows["n_topn"] = [1 + (i % 5) for i in range(n)]
    unmasked = [i for i in range(n) if i % 7 == 3]       # rows without the flag ignore their (restrictive) bitmap row
    for i in unmasked:
        rows["flags"][i] &= ~g.SAMPLE_MASKED
    out = g.run_sampler(logits.cuda(), rows, allow=bits.cuda())
    for i in range(n):
        a = None if i in unmasked else torch.from_numpy(allow[i])
        k = int(rows["n_topn"][i])
        o = sample_row(logits[i].float(), SamplingCase(greedy=True, num_logprobs=k), allow=a)
        assert out["token"][i] == o["token"], i
        if a is not None:
            assert allow[i, out["token"][i]]
        assert abs(out["logprob"][i] - o["logprob"]) < 1e-3
        assert out["rank"][i] == o["rank"]
        for j in range(k):
            want = o["topn_logprobs"][j]
            got = out["topn_lps"][i][j]
            assert (np.isinf(want) and np.isinf(got)) or abs(got - want) < 1e-3, (i, j, got, want)
            if np.isfinite(want) and out["topn_ids"][i][j] != o["topn_ids"][j]:
                assert abs(got - want) < 1e-5   # tie between equal bf16 logits (fp32 noise of the two log-sum-exps)


@pytest.mark.parametrize("V,dtype", [(2048, torch.float32), (128256, torch.bfloat16)])
def test_sampler_sampling_paths_under_mask_match_oracle(g, V, dtype):
    from oracle.sampler_oracle import SamplingCase, len_penalty_factor_m1, sample_row

    torch.manual_seed(V + 5)
    rng = np.random.RandomState(V + 5)
    cases = [
        SamplingCase(greedy=False, temperature=0.7, seed=1),
        SamplingCase(greedy=False, temperature=1.0, top_k=50, seed=2),
        SamplingCase(greedy=False, temperature=1.3, top_p=0.8, seed=3),
        SamplingCase(greedy=False, temperature=0.9, top_k=200, top_p=0.5, seed=4),
        SamplingCase(greedy=False, temperature=1.0, typical_p=0.9, seed=5),
        SamplingCase(greedy=False, temperature=1.0, typical_p=0.2, top_k=40, seed=6),
        SamplingCase(greedy=False, temperature=1.0, typical_p=0.9, repetition_penalty=1.2, length_penalty=(64, 1.05),
                     n_out=100, min_tokens=128, seed=1234 << 20),
        SamplingCase(greedy=True, typical_p=0.5),   # method SAMPLE at temperature 0: typical-p then argmax
    ]
    n = len(cases)
    logits = (torch.randn(n, V) * 3).to(dtype)
    allow, bits = _random_allow(rng, n, V, [0.02, 0.3, 0.001, 0.5])
    rows = _rows(g, n)
    words = (V + 31) // 32
    bitmap = torch.zeros(n, words, dtype=torch.int32)
    seen = torch.zeros(n, V, dtype=torch.bool)
    seen[:, 10:20] = True
    bitmap[:, 0] = sum(1 << b for b in range(10, 20))
    for i, c in enumerate(cases):
        c.num_logprobs = 2
        rows["flags"][i] = (g.SAMPLE_LOGPROBS | g.SAMPLE_MASKED | (g.SAMPLE_TYPICAL if 0 < c.typical_p < 1 else 0)
                            | (g.SAMPLE_GREEDY if c.greedy else 0))
        rows["n_topn"][i] = 2
        rows["temperature"][i], rows["top_k"][i], rows["top_p"][i] = c.temperature, c.top_k, c.top_p
        rows["typical_p"][i], rows["rep_penalty"][i] = c.typical_p, c.repetition_penalty
        rows["n_out"][i], rows["min_tokens"][i], rows["step"][i] = c.n_out, c.min_tokens, c.n_out
        rows["seed_lo"][i], rows["seed_hi"][i] = c.seed & 0xFFFFFFFF, c.seed >> 32
        if c.length_penalty:
            f = len_penalty_factor_m1(c.n_out, *c.length_penalty)
            if f != 0.0:
                rows["flags"][i] |= g.SAMPLE_LENPEN
                rows["len_decay_factor"][i] = f
    out = g.run_sampler(logits.cuda(), rows, bitmap.cuda(), allow=bits.cuda())
    same = 0
    for i, c in enumerate(cases):
        a = torch.from_numpy(allow[i])
        o = sample_row(logits[i].float(), c, seen[i], allow=a)
        tok = int(out["token"][i])
        assert allow[i, tok], (i, tok)
        if o["allowed"] is not None:
            assert bool(o["allowed"][tok]), (i, tok)
        same += int(tok == o["token"])
        lp = torch.log_softmax(logits[i].float().masked_fill(~a, float("-inf")), -1)
        assert abs(out["logprob"][i] - float(lp[tok])) < 1e-3
        assert out["rank"][i] == int((lp >= lp[tok]).sum())
    assert same >= n - 1, same


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


def _oracle_guided_greedy(cfg, weights, prompt, matcher, n_max, eos=2):
    """Constrained greedy decode with the CPU oracle: the same xgrammar matcher class fills the bitmask, the oracle
    masks the logits (apply_token_bitmask) and takes the argmax.  Returns [(token, logprob, margin between the two best
    ALLOWED logits)]."""
    import xgrammar as xgr
    from oracle.llama_oracle import LlamaOracle
    from oracle.sampler_oracle import apply_token_bitmask, unpack_token_bitmask

    ora = LlamaOracle(cfg, weights)
    st = ora.new_seq()
    logits = ora.step([(st, prompt)])[0]
    bm = xgr.allocate_token_bitmask(1, cfg.vocab)
    out = []
    for _ in range(n_max):
        matcher.fill_next_token_bitmask(bm, 0)
        masked = apply_token_bitmask(logits.float(), unpack_token_bitmask(bm.numpy()[0], cfg.vocab))
        lp = torch.log_softmax(masked, -1)
        top2 = torch.topk(masked, 2).values
        t = int(torch.argmax(masked))
        out.append((t, float(lp[t]), float(top2[0] - top2[1])))
        if t == eos:
            break
        assert matcher.accept_token(t)
        logits = ora.step([(st, [t])])[0]
    return out


def test_engine_guided_greedy_matches_oracle_and_language():
    from vllm_tgis_adapter_b200.engine.core import make_sampling_params
    from vllm_tgis_adapter_b200.engine.guided import GrammarCompiler, MaskProvider
    from vllm_tgis_adapter_b200.engine.tokenizer import build_synthetic_tokenizer
    from vllm_tgis_adapter_b200.engine.types import StructuredOutputsParams

    cfg, weights, eng = _engine("tiny", max_num_seqs=8, max_batched_tokens=64, kv_cache_bytes=64 << 20)
    tok = build_synthetic_tokenizer(cfg.vocab)
    compiler = GrammarCompiler(tok, cfg.vocab)
    prov = MaskProvider(compiler)
    eng.set_mask_provider(prov.callback)
    rng = np.random.RandomState(4)
    specs = [
        StructuredOutputsParams(regex=r"(t1\d){6}"),
        StructuredOutputsParams(choice=["t10t11t12", "t13", "t14t15"]),
        StructuredOutputsParams(grammar='root ::= "t10" ("t11" | "t12")+ "t13"'),
        StructuredOutputsParams(regex=r"(t1\d){3}(t2\d)*t13"),
        None,                                              # an unguided request rides in the same batch
        StructuredOutputsParams(regex=r"(t[12]\d)+"),       # 20 allowed tokens, + EOS after the first
    ]
    # (every pattern is dead-end free on the synthetic vocabulary: no allowed token is a proper prefix of a longer
    # alternative that only a non-existent bare-digit token could complete -- "t1" / "t2" are not tokens, "t10" is)
    prompts = [rng.randint(3, cfg.vocab, size=n).tolist() for n in (5, 40, 17, 70, 9, 33)]
    n_new = 12
    for i, (p, spec) in enumerate(zip(prompts, specs)):
        if spec is not None:
            prov.register(f"r{i}", spec)
        eng.add_request(f"r{i}", p, make_sampling_params(greedy=True, max_tokens=n_new, num_logprobs=1, eos_token_id=2,
                                                         guided=spec is not None))
    eng.run_until_idle()
    res = [[] for _ in prompts]
    while True:
        outs = eng.poll(0)
        if not outs:
            break
        for o in outs:
            res[int(o.request_id[1:])].append(o)
    st = eng.status()
    eng.close()
    assert st.errored == 0
    assert prov.calls >= sum(1 for i, r in enumerate(res) if specs[i] is not None for o in r if o.new_token is not None)
    lang = [r"(t1\d){6}", r"t10t11t12|t13|t14t15", r"t10(t11|t12)+t13", r"(t1\d){3}(t2\d)*t13", None, r"(t[12]\d)+"]
    checked = 0
    for i, (p, spec, recs) in enumerate(zip(prompts, specs, res)):
        toks = [r.new_token for r in recs if r.new_token is not None]
        assert toks and recs[-1].finish_reason != 0
        if spec is None:
            assert len(toks) == n_new or toks[-1] == 2
            continue
        assert prov.error_of(f"r{i}") is None
        body = toks[:-1] if toks[-1] == 2 else toks
        text = "".join(f"t{t}" for t in body)
        if toks[-1] == 2:      # finished by the grammar: a complete member of the language
            assert recs[-1].finish_reason == 2 and re.fullmatch(lang[i], text), (i, text)
        else:                  # cut by max_tokens: a prefix of a member (every token was allowed when it was sampled)
            assert recs[-1].finish_reason == 1 and len(toks) == n_new
        ora = _oracle_guided_greedy(cfg, weights, p, compiler.compile(spec), n_new)
        for (ot, olp, margin), r in zip(ora, recs):
            if r.new_token != ot:
                assert margin < 0.02, (i, margin, r.new_token, ot)   # a near-tie between two allowed tokens may flip
                break
            assert abs(r.logprob - olp) < 2e-2, (i, r.logprob, olp)
            checked += 1
    assert checked >= 20, checked


def test_engine_guided_survives_preemption_and_provider_failure():
    """A KV cache too small for the batch forces preemption + recomputation: the matcher must see every token exactly
    once (a double feed would be rejected and abort the request).  A request whose provider entry is missing ends with
    TGIS_FINISH_ABORT instead of hanging the engine."""
    from vllm_tgis_adapter_b200.engine.core import make_sampling_params
    from vllm_tgis_adapter_b200.engine.guided import GrammarCompiler, MaskProvider
    from vllm_tgis_adapter_b200.engine.tokenizer import build_synthetic_tokenizer
    from vllm_tgis_adapter_b200.engine.types import StructuredOutputsParams

    # 6 sequences of 120 + 40 tokens need 30 blocks of 32 tokens; the cache holds 22, so the scheduler has to evict
    from oracle.llama_oracle import CONFIGS

    c = CONFIGS["tiny"]
    bytes_per_block = 2 * c.n_layers * c.n_kv_heads * 128 * 2 * 32   # K+V, bf16, KV_BLOCK = 32 tokens
    cfg, weights, eng = _engine("tiny", max_num_seqs=8, max_batched_tokens=256, kv_cache_bytes=22 * bytes_per_block + 4096)
    tok = build_synthetic_tokenizer(cfg.vocab)
    prov = MaskProvider(GrammarCompiler(tok, cfg.vocab))
    eng.set_mask_provider(prov.callback)
    rng = np.random.RandomState(9)
    n_new = 40
    for i in range(6):
        prov.register(f"r{i}", StructuredOutputsParams(regex=r"(t[12]\d)+"))
        eng.add_request(f"r{i}", rng.randint(3, cfg.vocab, size=120).tolist(),
                        make_sampling_params(greedy=True, max_tokens=n_new, min_tokens=n_new, eos_token_id=2, guided=True))
    eng.add_request("r6", [5, 6, 7], make_sampling_params(greedy=True, max_tokens=4, eos_token_id=2, guided=True))
    eng.run_until_idle()
    res = {}
    while True:
        outs = eng.poll(0)
        if not outs:
            break
        for o in outs:
            res.setdefault(o.request_id, []).append(o)
    st = eng.status()
    eng.close()
    assert st.errored == 0 and st.preemptions > 0, st.preemptions
    for i in range(6):
        toks = [r.new_token for r in res[f"r{i}"] if r.new_token is not None]
        assert prov.error_of(f"r{i}") is None
        assert len(toks) == n_new and all(10 <= t <= 29 for t in toks), toks
        assert res[f"r{i}"][-1].finish_reason == 1
    assert res["r6"][-1].finish_reason == 4   # ABORT: nobody registered a grammar for it


def test_guided_generate_over_grpc_on_the_real_engine():
    from test_server_gpu import LiveServer
    from vllm_tgis_adapter_b200.grpc.pb import generation_pb2 as pb

    live = LiveServer()
    try:
        call = live.channel.unary_unary("/fmaas.GenerationService/Generate",
                                        request_serializer=pb.BatchedGenerationRequest.SerializeToString,
                                        response_deserializer=pb.BatchedGenerationResponse.FromString)
        p = pb.Parameters()
        p.stopping.max_new_tokens = 16
        p.decoding.choice.choices.extend(["t10t11", "t12", "t13t14t15"])
        reqs = [pb.GenerationRequest(text=t) for t in ("t5 t6 t7", "t100 t200", "t9")]
        resp = call(pb.BatchedGenerationRequest(model_id="m", requests=reqs, params=p), timeout=120)
        for r in resp.responses:
            assert r.text.replace(" ", "") in ("t10t11", "t12", "t13t14t15"), r.text
            assert r.stop_reason == pb.StopReason.EOS_TOKEN
        p = pb.Parameters()
        p.method = pb.DecodingMethod.SAMPLE
        p.sampling.temperature = 1.0
        p.sampling.seed = 7
        p.stopping.max_new_tokens = 8
        p.decoding.regex = r"(t1\d){1,4}"
        resp = call(pb.BatchedGenerationRequest(model_id="m", requests=reqs[:1], params=p), timeout=120)
        assert re.fullmatch(r"(t1\d){1,4}", resp.responses[0].text.replace(" ", "")), resp.responses[0].text
    finally:
        live.close()
