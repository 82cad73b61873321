"""Tensor-parallel SERVER smoke: the real entrypoint path (`parse_args` -> `build_engine` with
--tensor-parallel-size N: rank 0 in this process, ranks 1..N-1 spawned, one per GPU) behind the in-process gRPC
server; a handful of unary Generate calls and one GenerateStream, compared with a single-GPU engine built the same way.
Usage: python scripts/tp_server_smoke.py [tp] [model]"""
import asyncio
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))


async def run(tp: int, model: str) -> list[list[int]]:
    import grpc

    from vllm_tgis_adapter_b200.engine.loader import build_engine
    from vllm_tgis_adapter_b200.grpc import grpc_server
    from vllm_tgis_adapter_b200.grpc.pb import generation_pb2 as pb
    from vllm_tgis_adapter_b200.tgis_utils.args import parse_args

    argv = ["--model", model, "--synthetic-weights", "--max-model-len", "512", "--grpc-port", "0", "--seed", "7"]
    if tp > 1:
        argv += ["--tensor-parallel-size", str(tp)]
    args = parse_args(argv)
    engine = build_engine(args)
    loop = asyncio.get_running_loop()
    engine.start(loop)
    stop = asyncio.Event()
    server = await grpc_server.start_grpc_server(args, engine, stop)
    out: list[list[int]] = []
    try:
        async with grpc.aio.insecure_channel(f"127.0.0.1:{server.bound_port}") as ch:
            gen = ch.unary_unary("/fmaas.GenerationService/Generate", request_serializer=pb.BatchedGenerationRequest.SerializeToString,
                                 response_deserializer=pb.BatchedGenerationResponse.FromString)
            stream = ch.unary_stream("/fmaas.GenerationService/GenerateStream",
                                     request_serializer=pb.SingleGenerationRequest.SerializeToString,
                                     response_deserializer=pb.GenerationResponse.FromString)
            params = pb.Parameters(stopping=pb.StoppingCriteria(max_new_tokens=12, min_new_tokens=12),
                                   response=pb.ResponseOptions(generated_tokens=True, token_logprobs=True))
            texts = ["alpha beta gamma", "one two three four five six seven", "x"]
            resp = await gen(pb.BatchedGenerationRequest(model_id="m", requests=[pb.GenerationRequest(text=t) for t in texts],
                                                         params=params))
            for r in resp.responses:
                assert r.generated_token_count == 12 and len(r.tokens) == 12, r
                out.append([t.text for t in r.tokens])
            n_msgs = 0
            async for _ in stream(pb.SingleGenerationRequest(model_id="m", request=pb.GenerationRequest(text=texts[1]),
                                                             params=params)):
                n_msgs += 1
            assert n_msgs == 13, n_msgs   # N tokens -> N + 1 messages (reference tests/test_grpc_server.py:60-69)
    finally:
        stop.set()
        await server.stop(0)
        engine.shutdown()
    return out


if __name__ == "__main__":
    tp = int(sys.argv[1]) if len(sys.argv) > 1 else 2
    model = sys.argv[2] if len(sys.argv) > 2 else "tiny"
    got = asyncio.run(run(tp, model))
    ref = asyncio.run(run(1, model))
    same = sum(a == b for a, b in zip(got, ref))
    print(f"tp={tp} server: {len(got)} responses, {same} identical to the single-GPU server")
    # sum order differs (fp32 rank-order all-reduce vs one GEMM): near-ties may flip late tokens; the first must agree
    assert all(a[0] == b[0] for a, b in zip(got, ref)), (got, ref)
    print("TP_SERVER_SMOKE_PASS")
