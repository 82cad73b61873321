"""configs[1] end to end through the REAL wire surface: N concurrent `GenerateStream` RPCs (512-in / 128-out, greedy)
against the in-process fmaas.GenerationService on top of the native engine.  Client-side numbers: p50/max TTFT (arrival
of the second stream message), decode tokens/s, whole-job tokens/s.
Usage: python scripts/grpc_bench.py [model] [n_streams] [prompt_len] [gen_len] [rounds]"""
import argparse
import asyncio
import dataclasses
import json
import statistics
import sys
import threading
import time
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from vllm_tgis_adapter_b200.engine.async_engine import AsyncTGISEngine  # noqa: E402
from vllm_tgis_adapter_b200.engine.core import PRESETS, NativeEngine  # noqa: E402
from vllm_tgis_adapter_b200.engine.loader import load_synthetic_weights, rope_cos_sin  # noqa: E402
from vllm_tgis_adapter_b200.engine.tokenizer import build_synthetic_tokenizer, synthetic_prompt  # noqa: E402
from vllm_tgis_adapter_b200.grpc import grpc_server  # noqa: E402
from vllm_tgis_adapter_b200.grpc.pb import generation_pb2 as pb  # noqa: E402

model = sys.argv[1] if len(sys.argv) > 1 else "llama3-8b"
N = int(sys.argv[2]) if len(sys.argv) > 2 else 32
P = int(sys.argv[3]) if len(sys.argv) > 3 else 512
G = int(sys.argv[4]) if len(sys.argv) > 4 else 128
ROUNDS = int(sys.argv[5]) if len(sys.argv) > 5 else 3

mc = dataclasses.replace(PRESETS[model], max_model_len=max(1024, P + G + 64))
native = NativeEngine(mc, max_num_seqs=N, max_batched_tokens=8192,
                      kv_cache_bytes=int(N * ((P + G) // 32 + 3) * 2 * mc.n_layers * mc.n_kv_heads * 32 * 128 * 2 * 1.1))
load_synthetic_weights(native, mc, 1234, 0)
native.load_weight("tgis.rope_cos_sin", rope_cos_sin(mc))
tok = build_synthetic_tokenizer(mc.vocab)
args = argparse.Namespace(max_new_tokens=1024, output_special_tokens=False, default_include_stop_seqs=True,
                          disable_prompt_logprobs=False, adapter_cache=None, prefix_store_path=None, host="127.0.0.1",
                          grpc_port=0, ssl_keyfile=None, ssl_certfile=None, ssl_ca_certs=None)
srv_loop = asyncio.new_event_loop()
ready = threading.Event()
state = {}


def serve():
    asyncio.set_event_loop(srv_loop)

    async def main():
        eng = AsyncTGISEngine(native, tok, mc)
        eng.start(srv_loop)
        state["stop"] = asyncio.Event()
        server = await grpc_server.start_grpc_server(args, eng, state["stop"])
        state["port"] = server.bound_port
        ready.set()
        await state["stop"].wait()
        await server.stop(0)
        eng.shutdown()

    srv_loop.run_until_complete(main())


threading.Thread(target=serve, daemon=True).start()
assert ready.wait(120)

rs = np.random.RandomState(1234)
texts = [synthetic_prompt(rs.randint(1000, mc.vocab - 1000, size=P)) for _ in range(N)]
params = pb.Parameters()
params.stopping.max_new_tokens = G
params.stopping.min_new_tokens = G


async def client_round():
    import grpc

    async with grpc.aio.insecure_channel(f"127.0.0.1:{state['port']}") as ch:
        call = ch.unary_stream("/fmaas.GenerationService/GenerateStream",
                               request_serializer=pb.SingleGenerationRequest.SerializeToString,
                               response_deserializer=pb.GenerationResponse.FromString)

        async def one(text):
            t0 = time.perf_counter()
            ttft, n_msg, n_tok = None, 0, 0
            async for msg in call(pb.SingleGenerationRequest(model_id="m", request=pb.GenerationRequest(text=text),
                                                             params=params)):
                n_msg += 1
                if n_msg == 2:
                    ttft = time.perf_counter() - t0
                n_tok = msg.generated_token_count
            return ttft, n_tok, time.perf_counter() - t0

        t0 = time.perf_counter()
        res = await asyncio.gather(*[one(t) for t in texts])
        wall = time.perf_counter() - t0
    ttfts = [r[0] for r in res]
    toks = sum(r[1] for r in res)
    t_first_all = max(ttfts)
    return {"wall_s": wall, "tokens": toks, "ttft_p50_ms": 1e3 * statistics.median(ttfts), "ttft_max_ms": 1e3 * max(ttfts),
            "job_tokens_per_s": toks / wall, "decode_tokens_per_s": (toks - N) / (wall - t_first_all)}


rounds = [asyncio.run(client_round()) for _ in range(ROUNDS + 1)][1:]   # first round = warm-up
best = max(rounds, key=lambda r: r["decode_tokens_per_s"])
out = {"surface": "fmaas.GenerationService/GenerateStream over grpc.aio (in-process server, loopback)", "model": model,
       "streams": N, "prompt_len": P, "gen_len": G, "rounds": rounds,
       "decode_tokens_per_s_median": statistics.median(r["decode_tokens_per_s"] for r in rounds),
       "ttft_p50_ms_median": statistics.median(r["ttft_p50_ms"] for r in rounds), "best": best}
print(json.dumps(out), flush=True)
Path("gpurun_out").mkdir(exist_ok=True)
Path("gpurun_out/grpc_bench.json").write_text(json.dumps(out, indent=1))
srv_loop.call_soon_threadsafe(state["stop"].set)
time.sleep(1.0)
