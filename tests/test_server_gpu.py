"""-m gpu: the full drop-in path — grpc client -> TextGenerationService -> AsyncTGISEngine -> libtgis_engine.so (real
kernels) — against the CPU oracle on the same BatchedGenerationRequest (BASELINE.json configs[0] shape: one greedy
request through `Generate`, plus the streaming and sampling-parameter cases)."""
import argparse
import asyncio
import threading

import grpc
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

CFG = "tiny"


class LiveServer:
    def __init__(self, cfg_name=CFG, seed=21, max_new_tokens=64, adapter_cache=None, max_loras=0, max_lora_rank=16):
        from oracle.llama_oracle import CONFIGS, rope_table, synthetic_weights
        from vllm_tgis_adapter_b200.engine.async_engine import AsyncTGISEngine
        from vllm_tgis_adapter_b200.engine.core import ModelConfig, NativeEngine
        from vllm_tgis_adapter_b200.engine.tokenizer import build_synthetic_tokenizer
        from vllm_tgis_adapter_b200.grpc import grpc_server

        if cfg_name.startswith("opt-"):   # OPT family (tests/test_zz_opt_gpu.py): HF OPT names, no rotary table
            from oracle.opt_oracle import OPT_CONFIGS, synthetic_opt_weights

            self.cfg = c = OPT_CONFIGS[cfg_name]
            self.weights = synthetic_opt_weights(c, seed=seed)
            mc = ModelConfig(n_layers=c.n_layers, hidden=c.hidden, n_q_heads=c.n_heads, n_kv_heads=c.n_heads, ffn=c.ffn,
                             vocab=c.vocab, head_dim=c.head_dim, rms_eps=c.ln_eps, max_model_len=c.max_model_len,
                             arch="opt")
            # 12 kv heads in 128-dim slots: one 2048-token sequence alone is 151 MB of cache at opt-125m's depth
            native = NativeEngine(mc, max_num_seqs=16, max_batched_tokens=256, kv_cache_bytes=512 << 20)
            native.load_weights(self.weights)
        else:
            self.cfg = CONFIGS[cfg_name]
            self.weights = synthetic_weights(self.cfg, seed=seed)
            c = self.cfg
            mc = ModelConfig(n_layers=c.n_layers, hidden=c.hidden, n_q_heads=c.n_q_heads, n_kv_heads=c.n_kv_heads,
                             ffn=c.ffn, vocab=c.vocab, rope_theta=c.rope_theta, rms_eps=c.rms_eps,
                             max_model_len=c.max_model_len)
            native = NativeEngine(mc, max_num_seqs=16, max_batched_tokens=256, kv_cache_bytes=64 << 20,
                                  max_loras=max_loras, max_lora_rank=max_lora_rank)
            native.load_weights(self.weights)
            native.load_weight("tgis.rope_cos_sin", rope_table(c))
        self.tok = build_synthetic_tokenizer(c.vocab)
        self.args = argparse.Namespace(max_new_tokens=max_new_tokens, output_special_tokens=False, default_include_stop_seqs=True,
                                       disable_prompt_logprobs=False, adapter_cache=adapter_cache, prefix_store_path=None,
                                       host="127.0.0.1", grpc_port=0, ssl_keyfile=None, ssl_certfile=None,
                                       ssl_ca_certs=None)
        self.loop = asyncio.new_event_loop()
        self.ready = threading.Event()

        def run():
            asyncio.set_event_loop(self.loop)

            async def main():
                self.engine = AsyncTGISEngine(native, self.tok, mc)
                self.engine.start(self.loop)
                self.stop_event = asyncio.Event()
                self.server = await grpc_server.start_grpc_server(self.args, self.engine, self.stop_event)
                self.port = self.server.bound_port
                self.ready.set()
                await self.stop_event.wait()
                await self.server.stop(0)
                self.engine.shutdown()

            self.loop.run_until_complete(main())

        self.thread = threading.Thread(target=run, daemon=True)
        self.thread.start()
        assert self.ready.wait(60)
        self.channel = grpc.insecure_channel(f"127.0.0.1:{self.port}")

    def close(self):
        self.channel.close()
        self.loop.call_soon_threadsafe(self.stop_event.set)
        self.thread.join(10)


@pytest.fixture(scope="module")
def live():
    s = LiveServer()
    yield s
    s.close()


def _oracle_greedy(live, prompt, n):
    from oracle.llama_oracle import LlamaOracle

    ora = LlamaOracle(live.cfg, live.weights)
    st = ora.new_seq()
    logits = ora.step([(st, prompt)])[0]
    out = []
    for _ in range(n):
        lp = torch.log_softmax(logits, -1)
        top2 = torch.topk(logits, 2).values
        t = int(torch.argmax(logits))
        out.append((t, float(lp[t]), float(top2[0] - top2[1])))
        logits = ora.step([(st, [t])])[0]
    return out


def test_generate_batch_matches_oracle(live):
    from vllm_tgis_adapter_b200.engine.tokenizer import synthetic_prompt
    from vllm_tgis_adapter_b200.grpc.pb import generation_pb2 as pb

    rs = np.random.RandomState(5)
    prompts = [rs.randint(3, live.cfg.vocab, size=n).tolist() for n in (12, 40, 7)]
    params = pb.Parameters()
    params.stopping.max_new_tokens = 12
    params.stopping.min_new_tokens = 12
    params.response.generated_tokens = True
    params.response.token_logprobs = True
    params.response.token_ranks = True
    call = live.channel.unary_unary("/fmaas.GenerationService/Generate",
                                    request_serializer=pb.BatchedGenerationRequest.SerializeToString,
                                    response_deserializer=pb.BatchedGenerationResponse.FromString)
    resp = call(pb.BatchedGenerationRequest(model_id="m", requests=[pb.GenerationRequest(text=synthetic_prompt(p))
                                                                    for p in prompts], params=params), timeout=60)
    assert len(resp.responses) == 3
    for p, r in zip(prompts, resp.responses):
        assert r.input_token_count == len(p) and r.generated_token_count == 12
        assert r.stop_reason == pb.StopReason.MAX_TOKENS
        ora = _oracle_greedy(live, p, 12)
        got = [int(t.text[1:]) for t in r.tokens]
        for i, ((otok, olp, margin), tok, ti) in enumerate(zip(ora, got, r.tokens)):
            if tok != otok:
                assert margin < 0.02, (i, margin)
                break  # continuation after a legitimate near-tie flip is a different sequence
            assert abs(ti.logprob - olp) < 1e-2
            if margin > 0.02:       # bf16 logits tie exactly now and then: rank 1 only off ties
                assert ti.rank == 1
        assert r.text == " " + " ".join(t.text for t in r.tokens)


def test_input_tokens_with_logprobs(live):
    """ResponseOptions.input_tokens + token_logprobs + token_ranks -> prompt logprobs (first prompt token has none)."""
    from vllm_tgis_adapter_b200.engine.tokenizer import synthetic_prompt
    from vllm_tgis_adapter_b200.grpc.pb import generation_pb2 as pb

    prompt = list(range(40, 70))
    params = pb.Parameters()
    params.stopping.max_new_tokens = 2
    params.response.input_tokens = True
    params.response.token_logprobs = True
    params.response.token_ranks = True
    params.response.top_n_tokens = 2
    call = live.channel.unary_unary("/fmaas.GenerationService/Generate",
                                    request_serializer=pb.BatchedGenerationRequest.SerializeToString,
                                    response_deserializer=pb.BatchedGenerationResponse.FromString)
    r = call(pb.BatchedGenerationRequest(model_id="m", requests=[pb.GenerationRequest(text=synthetic_prompt(prompt))],
                                         params=params), timeout=60).responses[0]
    assert len(r.input_tokens) == len(prompt)
    assert [t.text for t in r.input_tokens] == [f"t{i}" for i in prompt]
    assert r.input_tokens[0].logprob == 0.0 and r.input_tokens[0].rank == 0 and not r.input_tokens[0].top_tokens
    for t in r.input_tokens[1:]:
        assert t.logprob < 0.0 and t.rank >= 1 and len(t.top_tokens) == 2
        assert t.top_tokens[0].logprob >= t.top_tokens[1].logprob >= t.logprob - 1e-6 or t.rank <= 2


def test_generate_stream_protocol_shape_and_sampling_params(live):
    from vllm_tgis_adapter_b200.engine.tokenizer import synthetic_prompt
    from vllm_tgis_adapter_b200.grpc.pb import generation_pb2 as pb

    params = pb.Parameters()
    params.method = pb.DecodingMethod.SAMPLE
    params.sampling.temperature = 0.8
    params.sampling.top_k = 40
    params.sampling.top_p = 0.9
    params.sampling.typical_p = 0.9
    params.sampling.seed = 1234
    params.decoding.repetition_penalty = 1.2
    params.decoding.length_penalty.start_index = 4
    params.decoding.length_penalty.decay_factor = 1.05
    params.stopping.max_new_tokens = 10
    params.stopping.min_new_tokens = 10
    params.response.generated_tokens = True
    params.response.token_logprobs = True
    params.response.top_n_tokens = 2
    call = live.channel.unary_stream("/fmaas.GenerationService/GenerateStream",
                                     request_serializer=pb.SingleGenerationRequest.SerializeToString,
                                     response_deserializer=pb.GenerationResponse.FromString)
    req = pb.SingleGenerationRequest(model_id="m", request=pb.GenerationRequest(text=synthetic_prompt(range(10, 30))),
                                     params=params)
    chunks = list(call(req, timeout=60))
    assert len(chunks) == 11 and chunks[0].input_token_count == 20 and chunks[0].seed == 1234
    assert chunks[-1].generated_token_count == 10 and chunks[-1].stop_reason == pb.StopReason.MAX_TOKENS
    toks1 = [c.tokens[0].text for c in chunks[1:]]
    assert all(len(c.tokens[0].top_tokens) == 2 for c in chunks[1:])
    # seeded sampling is reproducible within the engine (own counter-based RNG)
    toks2 = [c.tokens[0].text for c in list(call(req, timeout=60))[1:]]
    assert toks1 == toks2
    assert "</s>" not in toks1   # min_new_tokens masks EOS


def test_generate_stream_greedy_tokens_match_oracle(live):
    """R2 on the oracle: the DELTA stream (1 input-details message + 1 message per token) carries the oracle's greedy
    token ids, logprobs and ranks -- not just a self-consistent sequence."""
    from vllm_tgis_adapter_b200.engine.tokenizer import synthetic_prompt
    from vllm_tgis_adapter_b200.grpc.pb import generation_pb2 as pb

    rs = np.random.RandomState(9)
    prompt = rs.randint(3, live.cfg.vocab, size=57).tolist()
    params = pb.Parameters()
    params.stopping.max_new_tokens = 14
    params.stopping.min_new_tokens = 14
    params.response.generated_tokens = True
    params.response.token_logprobs = True
    params.response.token_ranks = True
    call = live.channel.unary_stream("/fmaas.GenerationService/GenerateStream",
                                     request_serializer=pb.SingleGenerationRequest.SerializeToString,
                                     response_deserializer=pb.GenerationResponse.FromString)
    chunks = list(call(pb.SingleGenerationRequest(model_id="m", request=pb.GenerationRequest(text=synthetic_prompt(prompt)),
                                                  params=params), timeout=60))
    assert len(chunks) == 15 and chunks[0].input_token_count == 57 and not chunks[0].tokens
    ora = _oracle_greedy(live, prompt, 14)
    text = ""
    for i, (c, (otok, olp, margin)) in enumerate(zip(chunks[1:], ora)):
        assert len(c.tokens) == 1
        tok = int(c.tokens[0].text[1:])
        if tok != otok:
            assert margin < 0.02, (i, margin)
            break
        assert abs(c.tokens[0].logprob - olp) < 1e-2
        if margin > 0.02:
            assert c.tokens[0].rank == 1
        text += c.text
    assert chunks[-1].stop_reason == pb.StopReason.MAX_TOKENS and chunks[-1].generated_token_count == 14
    assert text.startswith(" t")


def test_cfg0_125m_single_greedy_generate_matches_oracle():
    """BASELINE.json configs[0]: the 125m-class model (12 layers, hidden 768; the Llama-architecture stand-in for
    facebook/opt-125m, whose weights/architecture are not available offline), ONE greedy request through `Generate` --
    the reference's own fixture path (/root/reference/tests/test_grpc_server.py:42-49: text, token count, stop reason) --
    checked token by token against the CPU oracle run in full."""
    from vllm_tgis_adapter_b200.engine.tokenizer import synthetic_prompt
    from vllm_tgis_adapter_b200.grpc.pb import generation_pb2 as pb

    srv = LiveServer("125m", seed=31)
    try:
        rs = np.random.RandomState(31)
        prompt = rs.randint(3, srv.cfg.vocab, size=64).tolist()
        params = pb.Parameters()
        params.stopping.max_new_tokens = 20
        params.stopping.min_new_tokens = 20
        params.response.generated_tokens = True
        params.response.token_logprobs = True
        call = srv.channel.unary_unary("/fmaas.GenerationService/Generate",
                                       request_serializer=pb.BatchedGenerationRequest.SerializeToString,
                                       response_deserializer=pb.BatchedGenerationResponse.FromString)
        resp = call(pb.BatchedGenerationRequest(model_id="m", requests=[pb.GenerationRequest(text=synthetic_prompt(prompt))],
                                                params=params), timeout=120)
        assert len(resp.responses) == 1
        r = resp.responses[0]
        assert r.text and r.generated_token_count == 20 and r.input_token_count == 64
        assert r.stop_reason == pb.StopReason.MAX_TOKENS
        ora = _oracle_greedy(srv, prompt, 20)
        n_same = 0
        for i, ((otok, olp, margin), ti) in enumerate(zip(ora, r.tokens)):
            if int(ti.text[1:]) != otok:
                assert margin < 0.04, (i, margin)
                break
            assert abs(ti.logprob - olp) < 2e-2
            n_same += 1
        assert n_same >= 4      # where the sequences part ways is asserted above: only at a near-tie of the oracle
    finally:
        srv.close()
