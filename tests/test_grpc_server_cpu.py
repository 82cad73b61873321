"""Host-logic tests (no GPU): the real gRPC servicer + async engine wrapper + detokenizer on a deterministic fake
native engine, driven over a real grpc channel.  Mirrors what the reference's tests pin at this boundary
(/root/reference/tests/test_grpc_server.py:42-131: non-empty text, token counts, 11 stream chunks for 10 tokens,
batch of 2 -> 2 responses, Tokenize count, correlation id) and adds the numeric/semantic cases the reference never
tests (stop sequences, stop reasons, logprobs/ranks/top-n, validation errors)."""
import argparse
import asyncio
import threading
import time

import grpc
import pytest

from fakes import FakeNativeEngine, fake_next_token
from vllm_tgis_adapter_b200.engine.async_engine import AsyncTGISEngine
from vllm_tgis_adapter_b200.engine.core import ModelConfig
from vllm_tgis_adapter_b200.engine.tokenizer import build_synthetic_tokenizer, synthetic_prompt
from vllm_tgis_adapter_b200.grpc import grpc_server
from vllm_tgis_adapter_b200.grpc.health import HealthCheckRequest, HealthCheckResponse
from vllm_tgis_adapter_b200.grpc.pb import generation_pb2 as pb

VOCAB = 1024


class Server:
    def __init__(self, script=None, max_model_len=128, **argkw):
        self.mc = ModelConfig(n_layers=1, hidden=128, n_q_heads=1, n_kv_heads=1, ffn=128, vocab=VOCAB,
                              max_model_len=max_model_len)
        self.fake = FakeNativeEngine(self.mc, script=script)
        self.fake.max_loras = argkw.pop("max_loras", 0)
        self.tok = build_synthetic_tokenizer(VOCAB)
        self.args = argparse.Namespace(max_new_tokens=64, output_special_tokens=False, default_include_stop_seqs=True,
                                       disable_prompt_logprobs=False, adapter_cache=None, prefix_store_path=None,
                                       host="127.0.0.1", grpc_port=0, ssl_keyfile=None, ssl_certfile=None,
                                       ssl_ca_certs=None)
        for k, v in argkw.items():
            setattr(self.args, k, v)
        self.loop = asyncio.new_event_loop()
        self.ready = threading.Event()
        self.thread = threading.Thread(target=self._run, daemon=True)
        self.thread.start()
        assert self.ready.wait(20)
        self.channel = grpc.insecure_channel(f"127.0.0.1:{self.port}")

    def _run(self):
        asyncio.set_event_loop(self.loop)

        async def main():
            self.engine = AsyncTGISEngine(self.fake, self.tok, self.mc)
            self.engine.start(self.loop)
            self.stop_event = asyncio.Event()
            self.server = await grpc_server.start_grpc_server(self.args, self.engine, self.stop_event)
            self.port = self.server.bound_port
            self.ready.set()
            await self.stop_event.wait()
            await self.server.stop(0)

        self.loop.run_until_complete(main())

    def close(self):
        self.channel.close()
        self.loop.call_soon_threadsafe(self.stop_event.set)
        self.thread.join(5)
        self.fake.close()

    # client helpers (what generated stubs do)
    def generate(self, texts, params=None, metadata=None):
        call = self.channel.unary_unary("/fmaas.GenerationService/Generate",
                                        request_serializer=pb.BatchedGenerationRequest.SerializeToString,
                                        response_deserializer=pb.BatchedGenerationResponse.FromString)
        req = pb.BatchedGenerationRequest(model_id="m", requests=[pb.GenerationRequest(text=t) for t in texts],
                                          params=params)
        return call(req, metadata=metadata, timeout=30)

    def stream(self, text, params=None):
        call = self.channel.unary_stream("/fmaas.GenerationService/GenerateStream",
                                         request_serializer=pb.SingleGenerationRequest.SerializeToString,
                                         response_deserializer=pb.GenerationResponse.FromString)
        return list(call(pb.SingleGenerationRequest(model_id="m", request=pb.GenerationRequest(text=text),
                                                    params=params), timeout=30))


@pytest.fixture()
def srv():
    s = Server()
    yield s
    s.close()


def _params(**kw):
    p = pb.Parameters()
    st = kw.pop("stopping", {})
    for k, v in st.items():
        if k == "stop_sequences":
            p.stopping.stop_sequences.extend(v)
        else:
            setattr(p.stopping, k, v)
    for k, v in kw.pop("response", {}).items():
        setattr(p.response, k, v)
    for k, v in kw.pop("sampling", {}).items():
        setattr(p.sampling, k, v)
    for k, v in kw.items():
        setattr(p, k, v)
    return p


def expected_tokens(prompt_ids, n):
    return [fake_next_token(prompt_ids, i, VOCAB) for i in range(n)]


def test_generate_batch_and_counts(srv):
    prompts = [[10, 11, 12], [500, 600]]
    resp = srv.generate([synthetic_prompt(p) for p in prompts], _params(stopping={"max_new_tokens": 10}))
    assert len(resp.responses) == 2
    for p, r in zip(prompts, resp.responses):
        assert r.generated_token_count == 10 and r.input_token_count == len(p)
        assert r.stop_reason == pb.StopReason.MAX_TOKENS
        toks = expected_tokens(p, 10)
        assert r.text == " " + " ".join(f"t{t}" for t in toks)


def test_generate_stream_chunk_count_and_delta_text(srv):
    p = [7, 8, 9, 10]
    chunks = srv.stream(synthetic_prompt(p), _params(stopping={"max_new_tokens": 10},
                                                     response={"generated_tokens": True, "token_logprobs": True,
                                                               "token_ranks": True}))
    assert len(chunks) == 11                          # reference tests/test_grpc_server.py:60-69
    assert chunks[0].input_token_count == 4 and chunks[0].generated_token_count == 0
    text = "".join(c.text for c in chunks[1:])
    toks = expected_tokens(p, 10)
    assert text == " " + " ".join(f"t{t}" for t in toks)
    assert [c.generated_token_count for c in chunks[1:]] == list(range(1, 11))
    assert chunks[-1].stop_reason == pb.StopReason.MAX_TOKENS
    assert all(c.stop_reason == pb.StopReason.NOT_FINISHED for c in chunks[1:-1])
    assert [c.tokens[0].text for c in chunks[1:]] == [f"t{t}" for t in toks]
    assert all(c.tokens[0].logprob < 0 and c.tokens[0].rank >= 1 for c in chunks[1:])


def test_stream_coalescing_merges_deltas_only_when_the_consumer_is_behind(monkeypatch):
    """Stream coalescing (vLLM's RequestOutputCollector behaviour: a DELTA output that has not been taken yet absorbs the next
    one; on under load by default, TGIS_STREAM_COALESCE=1 / 0 forces it): the stream carries the same tokens, text, logprobs
    and final counts in FEWER messages when the engine runs ahead of the consumer, and exactly one message per token when
    it does not."""
    import time as _time

    monkeypatch.setenv("TGIS_STREAM_COALESCE", "1")
    p = [7, 8, 9, 10]
    toks = expected_tokens(p, 40)
    # (1) the records of all 40 steps reach the request's queue together (one poll): ONE merged DELTA output
    import types

    from vllm_tgis_adapter_b200.engine import _lib
    from vllm_tgis_adapter_b200.engine.core import StepOutput
    from vllm_tgis_adapter_b200.engine.types import RequestOutputKind, SamplingParams

    class BurstEngine:
        max_loras = 0

        def __init__(self):
            self.lib = types.SimpleNamespace(tgis_last_error=lambda: b"")
            self.batch = None

        def start(self):
            pass

        def add_request(self, rid, prompt_ids, sp):
            self.batch = [StepOutput(request_id=rid, new_token=t, logprob=-0.5, rank=1, topn=[(t, -0.5)],
                                     finish_reason=_lib.FINISH_LENGTH if k == 39 else _lib.FINISH_NONE, stop_token_id=-1,
                                     n_prompt_tokens=len(prompt_ids), n_output_tokens=k + 1, ts_arrival=1.0,
                                     ts_first_scheduled=1.0, ts_first_token=1.0, ts_last_token=1.0, token_id=t)
                          for k, t in enumerate(toks)]

        def poll(self, timeout_ms=0):
            if self.batch is None:
                _time.sleep(0.002)
                return []
            b, self.batch = self.batch, None
            return b

        def abort(self, rid):
            pass

        def status(self):
            return types.SimpleNamespace(errored=0, is_running=1)

        def close(self):
            pass

    mc = ModelConfig(n_layers=1, hidden=128, n_q_heads=1, n_kv_heads=1, ffn=128, vocab=VOCAB, max_model_len=128)

    async def run(coalesce):
        eng = AsyncTGISEngine(BurstEngine(), build_synthetic_tokenizer(VOCAB), mc, coalesce_streams=coalesce)
        eng.start(asyncio.get_running_loop())
        sp = SamplingParams(max_tokens=40, min_tokens=40, logprobs=1, output_kind=RequestOutputKind.DELTA)
        outs = [o async for o in eng.generate({"prompt_token_ids": p}, sp, request_id="r1")]
        eng.shutdown()
        return outs

    merged = asyncio.run(run(True))
    assert len(merged) == 1 and merged[0].finished and merged[0].outputs[0].finish_reason == "length"
    assert list(merged[0].outputs[0].token_ids) == toks and len(merged[0].outputs[0].logprobs) == 40
    assert merged[0].outputs[0].text == " " + " ".join(f"t{t}" for t in toks)
    plain = asyncio.run(run(False))
    assert len(plain) == 40 and [t for o in plain for t in o.outputs[0].token_ids] == toks
    assert "".join(o.outputs[0].text for o in plain) == merged[0].outputs[0].text
    # the default ("auto"): by load -- merging starts once TGIS_STREAM_COALESCE_MIN_STREAMS requests are in flight
    monkeypatch.setenv("TGIS_STREAM_COALESCE", "auto")
    assert len(asyncio.run(run(None))) == 40                    # one request in flight, threshold 64: one message per token
    monkeypatch.setenv("TGIS_STREAM_COALESCE_MIN_STREAMS", "1")
    auto = asyncio.run(run(None))
    assert len(auto) == 1 and list(auto[0].outputs[0].token_ids) == toks
    monkeypatch.setenv("TGIS_STREAM_COALESCE", "1")
    # (2) an engine slower than the consumer: nothing to merge, one message per token as without the switch
    s = Server()
    s.fake.step_delay = 0.01
    try:
        chunks = s.stream(synthetic_prompt(p), _params(stopping={"max_new_tokens": 10}, response={"generated_tokens": True}))
        assert len(chunks) == 11 and [c.generated_token_count for c in chunks[1:]] == list(range(1, 11))
    finally:
        s.close()


def test_stop_reasons_eos_token_limit_and_stop_sequence():
    p = [20, 21]
    s = Server(script={tuple(p): [40, 41, 2, 50]})
    try:
        r = s.generate([synthetic_prompt(p)], _params(stopping={"max_new_tokens": 10})).responses[0]
        assert r.stop_reason == pb.StopReason.EOS_TOKEN and r.stop_sequence == "</s>"
        assert r.generated_token_count == 3 and r.text == " t40 t41"    # EOS text skipped (skip_special_tokens)
        # min_new_tokens keeps going past the EOS the engine would otherwise stop on
        r = s.generate([synthetic_prompt(p)], _params(stopping={"max_new_tokens": 4, "min_new_tokens": 4})).responses[0]
        assert r.generated_token_count == 4 and r.stop_reason == pb.StopReason.MAX_TOKENS
    finally:
        s.close()
    p2 = [30, 31, 32]
    toks = expected_tokens(p2, 8)
    stop = f"t{toks[3]} t{toks[4]}"
    s = Server()
    try:
        r = s.generate([synthetic_prompt(p2)], _params(stopping={"max_new_tokens": 8, "stop_sequences": [stop]})).responses[0]
        assert r.stop_reason == pb.StopReason.STOP_SEQUENCE and r.stop_sequence == stop
        assert r.text.endswith(stop) and r.generated_token_count == 5          # default include_stop_sequence = True
        r = s.generate([synthetic_prompt(p2)], _params(stopping={"max_new_tokens": 8, "stop_sequences": [stop],
                                                                 "include_stop_sequence": False})).responses[0]
        assert r.stop_reason == pb.StopReason.STOP_SEQUENCE and stop not in r.text
        assert r.text == " " + " ".join(f"t{t}" for t in toks[:3]) + " "
        # unset max_new_tokens near the context limit -> TOKEN_LIMIT (grpc_server.py:787-798)
        long_prompt = list(range(100, 220))
        r = s.generate([synthetic_prompt(long_prompt)], _params()).responses[0]
        assert r.stop_reason == pb.StopReason.TOKEN_LIMIT and r.generated_token_count == 128 - 120
    finally:
        s.close()


def test_token_details_topn_and_input_text(srv):
    p = [60, 61]
    r = srv.generate([synthetic_prompt(p)], _params(stopping={"max_new_tokens": 3},
                                                    response={"generated_tokens": True, "token_logprobs": True,
                                                              "top_n_tokens": 2, "input_text": True})).responses[0]
    assert r.text.startswith(synthetic_prompt(p))
    assert len(r.tokens) == 3
    for t in r.tokens:
        assert len(t.top_tokens) == 2 and t.top_tokens[0].logprob >= t.top_tokens[1].logprob


def test_validation_errors_match_tgis_strings(srv):
    cases = [
        (_params(stopping={"max_new_tokens": 1000}), "max_new_tokens must be <= 64"),
        (_params(stopping={"max_new_tokens": 4, "min_new_tokens": 5}), "min_new_tokens must be <= max_new_tokens"),
        (_params(response={"token_logprobs": True}), "must request input and/or generated tokens to request extra token detail"),
        (_params(response={"generated_tokens": True, "top_n_tokens": 11}), "top_n_tokens (11) must be <= 10"),
        (_params(sampling={"top_p": 1.5}), "top_p must be > 0.0 and <= 1.0"),
        (_params(stopping={"stop_sequences": ["a"] * 7}), "can specify at most 6 non-empty stop sequences, each not more than 240 UTF8 bytes"),
    ]
    for params, msg in cases:
        with pytest.raises(grpc.RpcError) as ei:
            srv.generate(["t5 t6"], params)
        assert ei.value.code() == grpc.StatusCode.INVALID_ARGUMENT and ei.value.details() == msg
    with pytest.raises(grpc.RpcError) as ei:
        srv.generate([synthetic_prompt(range(3, 3 + 130))], _params(stopping={"max_new_tokens": 2}))
    assert ei.value.details() == "input tokens (130) plus prefix length (0) must be < 128"
    d = pb.Parameters()
    d.decoding.length_penalty.start_index = 1
    d.decoding.length_penalty.decay_factor = 11.0
    with pytest.raises(grpc.RpcError) as ei:
        srv.generate(["t5"], d)
    assert ei.value.details() == "length_penalty.decay_factor must be >= 1.0 and <= 10.0"
    req = pb.BatchedGenerationRequest(model_id="m", adapter_id="x", requests=[pb.GenerationRequest(text="t5")])
    call = srv.channel.unary_unary("/fmaas.GenerationService/Generate",
                                   request_serializer=pb.BatchedGenerationRequest.SerializeToString,
                                   response_deserializer=pb.BatchedGenerationResponse.FromString)
    with pytest.raises(grpc.RpcError) as ei:
        call(req, timeout=10)
    assert ei.value.details() == "adapter_id supplied but no adapter store was configured"


def test_tokenize_model_info_health(srv):
    call = srv.channel.unary_unary("/fmaas.GenerationService/Tokenize",
                                   request_serializer=pb.BatchedTokenizeRequest.SerializeToString,
                                   response_deserializer=pb.BatchedTokenizeResponse.FromString)
    text = "t5 t6 t7 t8"
    r = call(pb.BatchedTokenizeRequest(model_id="m", requests=[pb.TokenizeRequest(text=text)], return_tokens=True,
                                       return_offsets=True, truncate_input_tokens=3), timeout=10).responses[0]
    assert r.token_count == 3 and list(r.tokens) == ["t6", "t7", "t8"]         # left truncation (:872-874)
    assert [(o.start, o.end) for o in r.offsets] == [(3, 5), (6, 8), (9, 11)]
    info = srv.channel.unary_unary("/fmaas.GenerationService/ModelInfo",
                                   request_serializer=pb.ModelInfoRequest.SerializeToString,
                                   response_deserializer=pb.ModelInfoResponse.FromString)(pb.ModelInfoRequest(model_id="m"), timeout=10)
    assert info.max_sequence_length == 128 and info.max_new_tokens == 64
    assert info.model_kind == pb.ModelInfoResponse.ModelKind.DECODER_ONLY
    health = srv.channel.unary_unary("/grpc.health.v1.Health/Check",
                                     request_serializer=HealthCheckRequest.SerializeToString,
                                     response_deserializer=HealthCheckResponse.FromString)
    assert health(HealthCheckRequest(service="fmaas.GenerationService"), timeout=10).status == 1


def test_healthcheck_cli_against_a_live_server(srv, capsys):
    """The k8s probe CLI (reference healthcheck.py:1-96, console script grpc_healthcheck): exit status and the two output
    shapes against a live server, an unknown service and a dead port; same flags and defaults as the reference's."""
    from vllm_tgis_adapter.healthcheck import cli, health_check, parse_args

    d = parse_args([])
    assert (d.server_url, d.timeout, d.service_name, d.insecure) == ("localhost:8033", 1, "fmaas.GenerationService", True)
    assert parse_args(["--secure"]).insecure is False
    with pytest.raises(SystemExit):      # --insecure and --secure exclude each other
        parse_args(["--secure", "--insecure"])
    url = f"127.0.0.1:{srv.port}"
    assert health_check(server_url=url, service="fmaas.GenerationService", timeout=10) is True
    assert capsys.readouterr().out == "health check...status: SERVING\n"
    assert health_check(server_url=url, service=None, timeout=10) is True          # the server as a whole
    cli(["--server-url", url, "--timeout", "10"])                                  # exit status 0: returns
    capsys.readouterr()
    assert health_check(server_url=url, service="no.such.Service", timeout=10) is False
    out = capsys.readouterr().out
    assert out.startswith("health check...Health.Check failed: code=StatusCode.NOT_FOUND")
    with pytest.raises(SystemExit) as ei:
        cli(["--server-url", "127.0.0.1:1", "--timeout", "0.5"])
    assert ei.value.code == 1
    assert "Health.Check failed: code=StatusCode." in capsys.readouterr().out


def test_time_limit_stream_aborts_engine_request():
    s = Server()
    s.fake.step_delay = 0.05
    try:
        chunks = s.stream(synthetic_prompt([70, 71]), _params(stopping={"max_new_tokens": 60, "time_limit_millis": 200}))
        assert chunks[-1].stop_reason == pb.StopReason.TIME_LIMIT
        assert 1 <= chunks[-1].generated_token_count < 60
        import time as _t
        _t.sleep(0.2)
        assert s.fake.aborted, "engine request must be aborted after the deadline (grpc_server.py:387-389)"
    finally:
        s.close()


def test_correlation_id_is_request_id_and_logged(srv, caplog):
    import logging

    with caplog.at_level(logging.INFO, logger="vllm_tgis_adapter.tgis_utils.logs"):
        srv.generate(["t5 t6"], _params(stopping={"max_new_tokens": 2}), metadata=[("x-correlation-id", "corr-42")])
    msgs = [r.getMessage() for r in caplog.records]
    assert any("request_id=corr-42-0" in m and "correlation_id=corr-42" in m for m in msgs), msgs
    assert any(m.startswith("Finished processing request") and "Generated 2 tokens" in m for m in msgs)


@pytest.mark.parametrize("version", ["v1alpha", "v1"])
def test_server_reflection_lists_services_and_serves_descriptors(srv, version):
    """Server reflection (reference grpc_server.py:919-926: health + fmaas.GenerationService + reflection): what
    `grpcurl -plaintext host:port list / describe fmaas.GenerationService` does -- list_services, then
    file_containing_symbol, and the returned FileDescriptorProto must describe the real wire surface."""
    from google.protobuf import descriptor_pb2

    from vllm_tgis_adapter_b200.grpc import reflection

    pkg = reflection.V1ALPHA if version == "v1alpha" else reflection.V1
    call = srv.channel.stream_stream(f"/{pkg.service_name}/ServerReflectionInfo",
                                     request_serializer=pkg.Request.SerializeToString,
                                     response_deserializer=pkg.Response.FromString)
    reqs = [pkg.Request(list_services=""), pkg.Request(file_containing_symbol="fmaas.GenerationService"),
            pkg.Request(file_containing_symbol="fmaas.GenerationService.GenerateStream"),
            pkg.Request(file_containing_symbol="grpc.health.v1.Health"),
            pkg.Request(file_by_filename="no/such.proto"),
            pkg.Request(file_containing_symbol="fmaas.Parameters")]
    resps = list(call(iter(reqs), timeout=10))
    assert len(resps) == len(reqs)
    names = {s.name for s in resps[0].list_services_response.service}
    assert {"fmaas.GenerationService", "grpc.health.v1.Health", pkg.service_name} <= names
    fd = descriptor_pb2.FileDescriptorProto.FromString(resps[1].file_descriptor_response.file_descriptor_proto[0])
    assert fd.package == "fmaas"
    methods = {m.name: (m.client_streaming, m.server_streaming) for m in fd.service[0].method}
    assert methods == {"Generate": (False, False), "GenerateStream": (False, True), "Tokenize": (False, False),
                       "ModelInfo": (False, False)}
    assert resps[2].file_descriptor_response.file_descriptor_proto[0] == resps[1].file_descriptor_response.file_descriptor_proto[0]
    hfd = descriptor_pb2.FileDescriptorProto.FromString(resps[3].file_descriptor_response.file_descriptor_proto[0])
    assert hfd.package == "grpc.health.v1" and {m.name for m in hfd.service[0].method} == {"Check", "Watch"}
    assert resps[4].WhichOneof("message_response") == "error_response" and resps[4].error_response.error_code == 5
    assert resps[5].original_request.file_containing_symbol == "fmaas.Parameters"
    assert resps[5].file_descriptor_response.file_descriptor_proto


def test_http_sidecar_health_and_vllm_named_metrics(srv):
    """/health and /metrics of the HTTP side-car (reference http.py:41-99, tests/test_http_server.py:4-34): after a few
    requests through gRPC the Prometheus text must carry vLLM's metric names with the request counts and histograms."""
    import urllib.request

    from vllm_tgis_adapter_b200 import http as sidecar

    srv.generate(["t1 t2 t3", "t4 t5"], _params(stopping={"max_new_tokens": 5, "min_new_tokens": 5}))
    srv.stream("t9 t8 t7", _params(stopping={"max_new_tokens": 3, "min_new_tokens": 3}))
    args = argparse.Namespace(host="127.0.0.1", port=0)
    fut = asyncio.run_coroutine_threadsafe(sidecar.run_http_server(args, srv.engine), srv.loop)
    for _ in range(200):
        if getattr(sidecar.run_http_server, "bound_port", 0):
            break
        time.sleep(0.01)
    port = sidecar.run_http_server.bound_port
    try:
        with urllib.request.urlopen(f"http://127.0.0.1:{port}/health", timeout=5) as r:
            assert r.status == 200
        with urllib.request.urlopen(f"http://127.0.0.1:{port}/metrics", timeout=5) as r:
            text = r.read().decode()
    finally:
        fut.cancel()
    def value(name):
        return next(float(line.rsplit(" ", 1)[1]) for line in text.splitlines() if line.startswith(name))
    import re
    m = re.search(r'vllm:request_success_total\{[^}]*finished_reason="length"[^}]*\} (\S+)', text)
    assert m and float(m.group(1)) == 3.0
    assert value("vllm:generation_tokens_total{") == 13.0
    assert value("vllm:prompt_tokens_total{") == 8.0
    assert value("vllm:time_to_first_token_seconds_count{") == 3.0
    assert value("vllm:e2e_request_latency_seconds_count{") == 3.0
    assert value("vllm:request_generation_tokens_sum{") == 13.0
    assert "vllm:num_requests_running{" in text and "vllm:kv_cache_usage_perc{" in text


# ---------------------------------------------------------------------------------------------- wire schema golden
def test_proto_descriptor_matches_reference_schema():
    """The hand-built run-time descriptor (grpc/pb/generation_pb2.py; no protoc in this image) against the golden that
    oracle/gen_proto_golden.py extracted from the reference's generation.proto: every message, field name, NUMBER, type,
    label, proto3-optional presence, oneof membership, enum value and RPC signature."""
    import json
    from pathlib import Path

    from google.protobuf import descriptor as D

    gold = json.loads((Path(__file__).parent / "golden" / "generation_proto_schema.json").read_text())
    fd = pb.DESCRIPTOR
    assert fd.package == gold["package"] == "fmaas"
    scalar = {D.FieldDescriptor.TYPE_STRING: "string", D.FieldDescriptor.TYPE_UINT32: "uint32",
              D.FieldDescriptor.TYPE_UINT64: "uint64", D.FieldDescriptor.TYPE_FLOAT: "float",
              D.FieldDescriptor.TYPE_BOOL: "bool", D.FieldDescriptor.TYPE_INT32: "int32",
              D.FieldDescriptor.TYPE_INT64: "int64", D.FieldDescriptor.TYPE_DOUBLE: "double",
              D.FieldDescriptor.TYPE_BYTES: "bytes"}

    def walk(msgs, prefix=""):
        for m in msgs:
            yield prefix + m.name, m
            yield from walk(m.nested_types, prefix + m.name + ".")

    ours = dict(walk(fd.message_types_by_name.values()))
    assert sorted(ours) == sorted(gold["messages"])
    n_checked = 0
    for name, m in ours.items():
        gf = gold["messages"][name]
        assert sorted(f.name for f in m.fields) == sorted(gf), name
        for f in m.fields:
            g = gf[f.name]
            assert f.number == g["number"], (name, f.name)
            if f.type in (D.FieldDescriptor.TYPE_MESSAGE, D.FieldDescriptor.TYPE_ENUM):
                tname = (f.message_type or f.enum_type).name
                assert tname == g["type"].split(".")[-1], (name, f.name, tname, g["type"])
            else:
                assert scalar[f.type] == g["type"], (name, f.name)
            rep = f.is_repeated if hasattr(f, "is_repeated") else f.label == D.FieldDescriptor.LABEL_REPEATED
            assert bool(rep() if callable(rep) else rep) == (g["label"] == "repeated"), (name, f.name)
            assert bool(f.has_presence and f.containing_oneof is not None and f.containing_oneof.name.startswith("_")) \
                == (g["label"] == "optional"), (name, f.name)
            real_oneof = f.containing_oneof.name if f.containing_oneof is not None and \
                not f.containing_oneof.name.startswith("_") else None
            assert real_oneof == g["oneof"], (name, f.name)
            n_checked += 1
    assert n_checked == gold["meta"]["n_fields"]

    def walk_enums():
        for e in fd.enum_types_by_name.values():
            yield e.name, e
        for name, m in ours.items():
            for e in m.enum_types:
                yield name + "." + e.name, e

    enums = dict(walk_enums())
    assert sorted(enums) == sorted(gold["enums"])
    for name, e in enums.items():
        assert {v.name: v.number for v in e.values} == gold["enums"][name], name
    svc = fd.services_by_name["GenerationService"]
    assert sorted(m.name for m in svc.methods) == sorted(gold["services"]["GenerationService"])
    for m in svc.methods:
        g = gold["services"]["GenerationService"][m.name]
        assert (m.input_type.name, m.output_type.name) == (g["input"], g["output"])
        assert bool(m.server_streaming) == g["server_streaming"] and bool(m.client_streaming) == g["client_streaming"]


# ---------------------------------------------------------------------------------------------- TLS / mTLS
def _make_certs(tmp_path):
    """A throw-away CA, a server certificate for 127.0.0.1/localhost and a client certificate (what the reference's
    tests/utils.py:226-292 builds with openssl)."""
    import datetime
    import ipaddress

    from cryptography import x509
    from cryptography.hazmat.primitives import hashes, serialization
    from cryptography.hazmat.primitives.asymmetric import rsa
    from cryptography.x509.oid import NameOID

    def key():
        return rsa.generate_private_key(public_exponent=65537, key_size=2048)

    def name(cn):
        return x509.Name([x509.NameAttribute(NameOID.COMMON_NAME, cn)])

    now = datetime.datetime.now(datetime.timezone.utc)
    ca_key = key()
    ca = (x509.CertificateBuilder().subject_name(name("tgis-test-ca")).issuer_name(name("tgis-test-ca"))
          .public_key(ca_key.public_key()).serial_number(x509.random_serial_number())
          .not_valid_before(now - datetime.timedelta(days=1)).not_valid_after(now + datetime.timedelta(days=2))
          .add_extension(x509.BasicConstraints(ca=True, path_length=None), critical=True)
          .sign(ca_key, hashes.SHA256()))

    def leaf(cn, server):
        k = key()
        b = (x509.CertificateBuilder().subject_name(name(cn)).issuer_name(ca.subject).public_key(k.public_key())
             .serial_number(x509.random_serial_number()).not_valid_before(now - datetime.timedelta(days=1))
             .not_valid_after(now + datetime.timedelta(days=2)))
        if server:
            b = b.add_extension(x509.SubjectAlternativeName([x509.DNSName("localhost"),
                                                             x509.IPAddress(ipaddress.ip_address("127.0.0.1"))]),
                                critical=False)
        return k, b.sign(ca_key, hashes.SHA256())

    def pem_key(k):
        return k.private_bytes(serialization.Encoding.PEM, serialization.PrivateFormat.TraditionalOpenSSL,
                               serialization.NoEncryption())

    def pem_cert(c):
        return c.public_bytes(serialization.Encoding.PEM)

    sk, sc = leaf("localhost", True)
    ck, cc = leaf("tgis-test-client", False)
    paths = {}
    for fname, data in (("ca.pem", pem_cert(ca)), ("server.key", pem_key(sk)), ("server.pem", pem_cert(sc)),
                        ("client.key", pem_key(ck)), ("client.pem", pem_cert(cc))):
        p = tmp_path / fname
        p.write_bytes(data)
        paths[fname] = p
    return paths


@pytest.mark.parametrize("mtls", [False, True])
def test_tls_and_mtls_server(tmp_path, mtls):
    """--ssl-keyfile/--ssl-certfile (TLS) and --ssl-ca-certs (mTLS: client certificate required), reference
    grpc_server.py:934-962; a Generate call goes through the encrypted channel."""
    c = _make_certs(tmp_path)
    srv = Server(ssl_keyfile=str(c["server.key"]), ssl_certfile=str(c["server.pem"]),
                 ssl_ca_certs=str(c["ca.pem"]) if mtls else None)
    try:
        ca = c["ca.pem"].read_bytes()
        if mtls:
            creds = grpc.ssl_channel_credentials(ca, c["client.key"].read_bytes(), c["client.pem"].read_bytes())
        else:
            creds = grpc.ssl_channel_credentials(ca)
        opts = (("grpc.ssl_target_name_override", "localhost"),)
        with grpc.secure_channel(f"127.0.0.1:{srv.port}", creds, options=opts) as ch:
            call = ch.unary_unary("/fmaas.GenerationService/Generate",
                                  request_serializer=pb.BatchedGenerationRequest.SerializeToString,
                                  response_deserializer=pb.BatchedGenerationResponse.FromString)
            resp = call(pb.BatchedGenerationRequest(model_id="m", requests=[pb.GenerationRequest(text="t5 t6 t7")],
                                                    params=pb.Parameters(stopping=pb.StoppingCriteria(max_new_tokens=4))),
                        timeout=20)
            assert len(resp.responses) == 1 and resp.responses[0].generated_token_count == 4
        # a plaintext client cannot talk to the TLS port
        with grpc.insecure_channel(f"127.0.0.1:{srv.port}") as ch:
            call = ch.unary_unary("/fmaas.GenerationService/ModelInfo",
                                  request_serializer=pb.ModelInfoRequest.SerializeToString,
                                  response_deserializer=pb.ModelInfoResponse.FromString)
            with pytest.raises(grpc.RpcError):
                call(pb.ModelInfoRequest(model_id="m"), timeout=3)
        if mtls:   # TLS without a client certificate is rejected when the server demands one
            with grpc.secure_channel(f"127.0.0.1:{srv.port}", grpc.ssl_channel_credentials(ca), options=opts) as ch:
                call = ch.unary_unary("/fmaas.GenerationService/ModelInfo",
                                      request_serializer=pb.ModelInfoRequest.SerializeToString,
                                      response_deserializer=pb.ModelInfoResponse.FromString)
                with pytest.raises(grpc.RpcError):
                    call(pb.ModelInfoRequest(model_id="m"), timeout=3)
    finally:
        srv.close()


def test_tls_flag_with_unreadable_file_fails_like_reference(tmp_path):
    """grpc_server.py:940-950: `Error reading `ssl_keyfile` file: ...` as a ValueError at start-up."""
    args = argparse.Namespace(max_new_tokens=64, output_special_tokens=False, default_include_stop_seqs=True,
                              disable_prompt_logprobs=False, adapter_cache=None, prefix_store_path=None, host="127.0.0.1",
                              grpc_port=0, ssl_keyfile=str(tmp_path / "missing.key"), ssl_certfile=str(tmp_path / "m.pem"),
                              ssl_ca_certs=None)
    mc = ModelConfig(n_layers=1, hidden=128, n_q_heads=1, n_kv_heads=1, ffn=128, vocab=VOCAB, max_model_len=128)
    fake = FakeNativeEngine(mc)

    async def go():
        eng = AsyncTGISEngine(fake, build_synthetic_tokenizer(VOCAB), mc)
        eng.start(asyncio.get_running_loop())
        with pytest.raises(ValueError, match="Error reading `ssl_keyfile` file"):
            await grpc_server.start_grpc_server(args, eng, asyncio.Event())

    asyncio.run(go())
    fake.close()
