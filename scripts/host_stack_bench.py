"""Ceiling of the HOST stack alone (no GPU): how many streamed tokens per second the Python side -- poller thread ->
AsyncTGISEngine dispatch -> incremental detokenizer -> TextGenerationService.GenerateStream -> grpc.aio -- can carry when the
engine behind it answers instantly.  The native engine is replaced by a replay double that emits one pre-built record per
active request per "step" with no arithmetic; N concurrent GenerateStream RPCs come from a SEPARATE client process (so the
client's own Python cost does not share the server's GIL).

    python scripts/host_stack_bench.py [n_streams] [gen_len] [rounds] [step_interval_ms]  ->  one JSON line

step_interval_ms = 0: the engine answers instantly (the ceiling); > 0: it paces its steps like a GPU engine of that step
time would (e.g. 256 streams at 12.9 ms = the 19 800 tokens/s of batch 256), and the question becomes whether the host
keeps up.

Read it against the engine's decode rate at the same concurrency (8 000 tokens/s at 32 streams,
19 800 at 256): the host stack must stay above it, since it runs concurrently with the GPU step, not inside it."""
import argparse
import asyncio
import json
import multiprocessing as mp
import os
import queue
import sys
import threading
import time
import types
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from vllm_tgis_adapter_b200.engine import _lib  # noqa: E402
from vllm_tgis_adapter_b200.engine.core import ModelConfig, StepOutput  # noqa: E402

VOCAB = 32000
ROUND_SLOT_S = 8.0   # wall-clock slot per round (longer than any round)


class ReplayEngine:
    """engine.core.NativeEngine's surface.  Every active request advances by one token per step; all records of a request
    are built when it is added, and there is NO engine thread: `poll` (called by AsyncTGISEngine's poller thread, like
    tgis_engine_poll) hands out the records of every step whose wall-clock deadline has passed -- a C++ engine that is paced
    by the GPU and never waits for the GIL, which is what the product's engine thread is."""

    def __init__(self, model, step_interval_s=0.0):
        self.model = model
        self.lib = types.SimpleNamespace(tgis_last_error=lambda: b"")
        self._new = queue.Queue()
        self._active = {}
        self.step_interval_s = step_interval_s
        self._deadline = 0.0
        self.max_loras, self.max_lora_rank = 0, 16
        self.steps = 0

    def start(self):
        pass

    def add_request(self, request_id, prompt_ids, params):
        n_prompt, ts = len(prompt_ids), time.monotonic()
        recs = []
        for k in range(1, params.max_tokens + 1):
            fin = _lib.FINISH_LENGTH if k >= params.max_tokens else _lib.FINISH_NONE
            tok = 3 + (k * 7919 + n_prompt) % (VOCAB - 3)
            recs.append(StepOutput(request_id=request_id, new_token=tok, logprob=-0.5, rank=1, topn=[], finish_reason=fin,
                                   stop_token_id=-1, n_prompt_tokens=n_prompt, n_output_tokens=k, ts_arrival=ts,
                                   ts_first_scheduled=ts, ts_first_token=ts, ts_last_token=ts, token_id=tok))
        self._new.put((request_id, iter(recs)))

    def abort(self, request_id):
        pass

    def set_mask_provider(self, cb):
        pass

    def _step(self, outs):
        try:
            while True:
                rid, it = self._new.get_nowait()
                self._active[rid] = it
        except queue.Empty:
            pass
        if not self._active:
            return False
        self.steps += 1
        done = []
        for rid, it in self._active.items():
            r = next(it)
            outs.append(r)
            if r.finish_reason:
                done.append(rid)
        for rid in done:
            del self._active[rid]
        return True

    def poll(self, timeout_ms=0):
        outs = []
        t_end = time.monotonic() + timeout_ms * 1e-3
        while True:
            now = time.monotonic()
            if not self.step_interval_s:
                if self._step(outs) or now >= t_end:
                    return outs
                time.sleep(0.0005)
                continue
            if self._deadline < now - 20 * self.step_interval_s:   # idle period: restart the clock
                self._deadline = now
            while self._deadline <= now:                            # every step that is due by now
                if not self._step(outs):
                    self._deadline = now + self.step_interval_s
                    break
                self._deadline += self.step_interval_s
            if outs or now >= t_end:
                return outs
            time.sleep(max(0.0, min(self._deadline, t_end) - now))

    def status(self):
        return types.SimpleNamespace(errored=0, is_running=1, n_running=len(self._active), n_waiting=0, free_blocks=10,
                                     total_blocks=10, tokens_generated=0, steps=self.steps, kernel_launches=0, gpu_busy_ms=0.0)

    def close(self):
        pass


def client_main(port, n_streams, gen_len, prompt_len, rounds, q, start_at):
    import grpc

    from vllm_tgis_adapter_b200.engine.tokenizer import synthetic_prompt
    from vllm_tgis_adapter_b200.grpc.pb import generation_pb2 as pb

    async def one(stub, text):
        params = pb.Parameters()
        params.stopping.max_new_tokens = gen_len
        params.stopping.min_new_tokens = gen_len
        n_tok = n_msg = 0
        async for m in stub(pb.SingleGenerationRequest(model_id="m", request=pb.GenerationRequest(text=text), params=params)):
            n_tok = max(n_tok, m.generated_token_count)   # cumulative per stream
            n_msg += 1
        return n_tok, n_msg

    async def run():
        res = []
        async with grpc.aio.insecure_channel(f"127.0.0.1:{port}") as ch:
            stub = ch.unary_stream("/fmaas.GenerationService/GenerateStream",
                                   request_serializer=pb.SingleGenerationRequest.SerializeToString,
                                   response_deserializer=pb.GenerationResponse.FromString)
            text = synthetic_prompt(range(10, 10 + prompt_len))
            for r in range(rounds):
                # rounds start on a shared wall-clock grid so that the clients' rounds overlap
                await asyncio.sleep(max(0.0, start_at + r * ROUND_SLOT_S - time.time()))
                t0 = time.perf_counter()
                ns = await asyncio.gather(*[one(stub, text) for _ in range(n_streams)])
                res.append((sum(a for a, _ in ns), time.perf_counter() - t0, sum(b for _, b in ns)))
        return res

    q.put(asyncio.run(run()))


def main():
    n_streams = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    gen_len = int(sys.argv[2]) if len(sys.argv) > 2 else 128
    rounds = int(sys.argv[3]) if len(sys.argv) > 3 else 3
    step_ms = float(sys.argv[4]) if len(sys.argv) > 4 else 0.0
    prompt_len = 64
    from vllm_tgis_adapter_b200.engine.async_engine import AsyncTGISEngine
    from vllm_tgis_adapter_b200.engine.tokenizer import build_synthetic_tokenizer
    from vllm_tgis_adapter_b200.grpc import grpc_server

    mc = ModelConfig(n_layers=1, hidden=128, n_q_heads=1, n_kv_heads=1, ffn=128, vocab=VOCAB, max_model_len=2048)
    eng = ReplayEngine(mc, step_interval_s=step_ms * 1e-3)
    tok = build_synthetic_tokenizer(VOCAB)
    args = argparse.Namespace(max_new_tokens=1024, output_special_tokens=False, default_include_stop_seqs=True,
                              disable_prompt_logprobs=False, adapter_cache=None, prefix_store_path=None, host="127.0.0.1",
                              grpc_port=0, ssl_keyfile=None, ssl_certfile=None, ssl_ca_certs=None)
    use_uvloop = os.environ.get("TGIS_HOST_BENCH_UVLOOP", "1") != "0"   # the entrypoint installs uvloop (__main__.py, as the
    if use_uvloop:                                                       # reference does: __main__.py:128)
        import uvloop

        loop = uvloop.new_event_loop()
    else:
        loop = asyncio.new_event_loop()
    ready = threading.Event()
    box = {}

    def run():
        asyncio.set_event_loop(loop)

        async def amain():
            engine = AsyncTGISEngine(eng, tok, mc)   # TGIS_STREAM_COALESCE=auto (default: by load) | 1 | 0
            engine.start(loop)
            box["stop"] = asyncio.Event()
            server = await grpc_server.start_grpc_server(args, engine, box["stop"])
            box["port"] = server.bound_port
            ready.set()
            await box["stop"].wait()
            await server.stop(0)

        if os.environ.get("TGIS_HOST_BENCH_PROFILE"):   # where the event-loop thread spends its time
            import cProfile
            import pstats

            pr = cProfile.Profile()
            pr.enable()
            loop.run_until_complete(amain())
            pr.disable()
            pstats.Stats(pr, stream=sys.stderr).sort_stats("tottime").print_stats(28)
        else:
            loop.run_until_complete(amain())

    th = threading.Thread(target=run, daemon=True)
    th.start()
    assert ready.wait(30)
    # the clients are Python too: N_CLIENTS processes share the streams so that the number measured is the server's
    n_clients = int(os.environ.get("TGIS_HOST_BENCH_CLIENTS", "4"))
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    share = [n_streams // n_clients + (1 if i < n_streams % n_clients else 0) for i in range(n_clients)]
    start_at = time.time() + 8.0   # all client processes begin their first round together (imports take seconds)
    procs = [ctx.Process(target=client_main, args=(box["port"], k, gen_len, prompt_len, rounds, q, start_at))
             for k in share if k > 0]
    for p in procs:
        p.start()
    per_client = [q.get(timeout=600) for _ in procs]
    for p in procs:
        p.join(10)
    # round r: tokens of all clients / the longest client time of that round (rounds start together: barrier by clock)
    res = [(sum(c[r][0] for c in per_client), max(c[r][1] for c in per_client), sum(c[r][2] for c in per_client))
           for r in range(rounds)]
    loop.call_soon_threadsafe(box["stop"].set)
    th.join(5)
    eng.close()
    best = max(n / t for n, t, _ in res)
    print(json.dumps({"host_stack_tokens_per_s": best, "n_streams": n_streams, "gen_len": gen_len,
                      "stream_coalesce": os.environ.get("TGIS_STREAM_COALESCE", "auto"), "client_processes": n_clients, "event_loop": "uvloop" if use_uvloop else "asyncio", "engine_step_interval_ms": step_ms,
                      "engine_paced_tokens_per_s": (n_streams / (step_ms * 1e-3)) if step_ms else None,
                      "rounds": [{"tokens": n, "messages": m, "s": round(t, 3)} for n, t, m in res],
                      "note": "replay engine (no GPU): streamed tokens delivered per second by the Python host stack"}))


if __name__ == "__main__":
    main()
