"""`EngineClient`-shaped façade over the native engine: the object the TGIS gRPC servicer holds where the reference
holds vLLM's AsyncLLM (/root/reference/src/vllm_tgis_adapter/grpc/grpc_server.py:166-176).

Members implemented = exactly the ones the reference touches: generate / abort / get_tokenizer /
is_tracing_enabled / errored / is_running / vllm_config.model_config.max_model_len / get_model_config.

Threading: the C++ engine runs its own step-loop thread; one Python poller thread blocks in `tgis_engine_poll`
(GIL released inside ctypes) and hands each batch of step records to the asyncio loop with ONE
`call_soon_threadsafe` per engine step (not per request), so 256 concurrent streams cost one wake-up per step."""
from __future__ import annotations

import asyncio
import collections
import itertools
import threading
import types as _types
from collections.abc import AsyncGenerator

from . import _lib
from .core import ModelConfig, NativeEngine, StepOutput, make_sampling_params
from .detokenizer import IncrementalDetokenizer
from ..metrics import EngineMetrics
from .types import (CompletionOutput, Logprob, RequestMetrics, RequestOutput, RequestOutputKind, SamplingParams,
                    TokensPrompt)

_FINISH = {_lib.FINISH_LENGTH: "length", _lib.FINISH_STOP_EOS: "stop", _lib.FINISH_STOP_TOKEN: "stop",
           _lib.FINISH_ABORT: "abort", _lib.FINISH_ERROR: "abort"}


class EngineDeadError(RuntimeError):
    pass


class _ReqState:
    def __init__(self, request_id: str, prompt_ids: list[int], sp: SamplingParams, detok: IncrementalDetokenizer,
                 queue: asyncio.Queue):
        self.request_id = request_id
        self.prompt_ids = prompt_ids
        self.sp = sp
        self.detok = detok
        self.queue = queue
        self.token_ids: list[int] = []
        self.logprobs: list[dict[int, Logprob]] = []
        # vLLM layout: one entry per prompt token, entry 0 is None (grpc_server.py:438-449,721-724)
        self.prompt_logprobs: list[dict[int, Logprob] | None] = [None] * len(prompt_ids)
        self.done = False


class AsyncTGISEngine:
    def __init__(self, engine: NativeEngine, tokenizer, model_config: ModelConfig, coalesce_streams: bool | None = None):
        self.engine = engine
        # DELTA (GenerateStream) outputs of several engine steps merge into one RequestOutput when the consumer is behind,
        # as vLLM's RequestOutputCollector does (vllm v1/engine/output_processor.py `put`: aggregate when the previous output
        # has not been taken).  The per-message cost of grpc.aio (~60 us) caps the host stack near 14 k streamed messages/s
        # (scripts/host_stack_bench.py); merging lifts the cap exactly when it binds.
        #   coalesce_streams / TGIS_STREAM_COALESCE = None / "auto" (default): merge only while at least
        #       TGIS_STREAM_COALESCE_MIN_STREAMS (64) requests are in flight -- below that the engine's decode rate is under
        #       the host's ceiling, and a lightly loaded server keeps sending exactly one message per token, which the
        #       reference's tests count (tests/test_grpc_server.py:60-69);
        #   True / "1": always;  False / "0": never.
        import os

        if coalesce_streams is None:
            env = os.environ.get("TGIS_STREAM_COALESCE", "auto").strip().lower()
            coalesce_streams = None if env in ("", "auto") else env not in ("0", "false", "off")
        self._coalesce_streams = coalesce_streams          # None: by load
        self._coalesce_min_streams = int(os.environ.get("TGIS_STREAM_COALESCE_MIN_STREAMS", "64"))
        self.tokenizer = tokenizer
        self._model_config = model_config
        mc = _types.SimpleNamespace(max_model_len=model_config.max_model_len)
        self.vllm_config = _types.SimpleNamespace(model_config=mc)   # grpc_server.py:196-199
        self._states: dict[str, _ReqState] = {}
        self._ids = itertools.count()
        self._loop: asyncio.AbstractEventLoop | None = None
        self._poller: threading.Thread | None = None
        self._stopping = False
        self._dead_error: str | None = None
        self._pending: collections.deque = collections.deque()   # poller thread -> event loop (see _poll_loop)
        self._drain_scheduled = False
        self.metrics = EngineMetrics()
        self._mask_provider = None   # guided decoding: created with the first guided request (imports xgrammar)
        # LoRA: adapter name -> engine slot; `lora_requests` is what grpc/adapters.py consults (the reference asks
        # vLLM's OpenAIServingModels.lora_requests, adapters.py:157-172)
        self._lora = None
        max_loras = getattr(engine, "max_loras", 0)
        if max_loras > 0:
            from .lora import LoRAManager

            self._lora = LoRAManager(engine, n_layers=model_config.n_layers, max_loras=max_loras,
                                     max_rank=getattr(engine, "max_lora_rank", 16))

    # -- lifecycle ------------------------------------------------------------------------------------------------
    def start(self, loop: asyncio.AbstractEventLoop | None = None) -> None:
        self._loop = loop or asyncio.get_event_loop()
        self.engine.start()
        self._poller = threading.Thread(target=self._poll_loop, name="tgis-poller", daemon=True)
        self._poller.start()

    def shutdown(self) -> None:
        self._stopping = True
        if self._poller is not None:
            self._poller.join(timeout=5)
        self.engine.close()

    def _poll_loop(self) -> None:
        while not self._stopping:
            try:
                outs = self.engine.poll(timeout_ms=100)
            except Exception as e:  # noqa: BLE001
                self._dead_error = str(e)
                outs = []
            if self._dead_error is None and not outs:
                st = self.engine.status()
                if st.errored:
                    self._dead_error = _lib.last_error(self.engine.lib) or "engine errored"
            if (outs or self._dead_error) and self._loop is not None:
                # hand-off to the event loop: records wait in `_pending`, and at most ONE drain callback is in the loop's
                # ready queue.  When the loop is behind, the records of several engine steps therefore reach the per-request
                # queues together (which is what lets stream coalescing see a backlog) instead of queueing one callback
                # per step behind the coroutines of the previous one.
                self._pending.append((outs, self._dead_error))
                if not self._drain_scheduled:
                    self._drain_scheduled = True
                    try:
                        self._loop.call_soon_threadsafe(self._drain)
                    except RuntimeError:
                        return  # loop closed
            if self._dead_error:
                return

    def _drain(self) -> None:
        # cleared FIRST: a record appended after this line schedules the next drain; one appended before it is popped below
        self._drain_scheduled = False
        while self._pending:
            outs, dead = self._pending.popleft()
            self._dispatch(outs, dead)

    def _dispatch(self, outs: list[StepOutput], dead: str | None) -> None:
        for o in outs:
            st = self._states.get(o.request_id)
            if st is not None:
                st.queue.put_nowait(o)
        if dead:
            for st in self._states.values():
                st.queue.put_nowait(EngineDeadError(dead))

    # -- EngineClient subset ----------------------------------------------------------------------------------------
    @property
    def errored(self) -> bool:
        return self._dead_error is not None or bool(self.engine.status().errored)

    @property
    def is_running(self) -> bool:
        return self._dead_error is None and bool(self.engine.status().is_running)

    @property
    def dead_error(self) -> BaseException:
        return EngineDeadError(self._dead_error or "engine dead")

    @property
    def supports_guided_decoding(self) -> bool:
        return True

    @property
    def lora_requests(self) -> dict:
        return self._lora.lora_requests if self._lora is not None else {}

    async def load_lora_adapter(self, *, lora_name: str, lora_path: str, lora_int_id: int) -> None:
        """vLLM `OpenAIServingModels.load_lora_adapter` as adapters.py:139-155 uses it: validate the checkpoint and
        remember the LoRARequest; a bad adapter is a ValueError.  The weights move into an engine slot when the first
        request that names the adapter is scheduled (LoRAManager.acquire)."""
        from .types import LoRARequest

        if self._lora is None:
            raise ValueError("LoRA is not enabled: start the server with --enable-lora")
        req = LoRARequest(lora_name=lora_name, lora_int_id=lora_int_id, lora_path=lora_path)
        await asyncio.get_running_loop().run_in_executor(None, self._lora.register, req)

    def _guided(self):
        if self._mask_provider is None:
            from .guided import GrammarCompiler, MaskProvider

            self._mask_provider = MaskProvider(GrammarCompiler(self.tokenizer, self._model_config.vocab))
            self.engine.set_mask_provider(self._mask_provider.callback)
        return self._mask_provider

    async def get_model_config(self):
        return self.vllm_config.model_config

    async def get_tokenizer(self, lora_request=None):  # noqa: ARG002
        return self.tokenizer

    async def is_tracing_enabled(self) -> bool:
        return False

    async def abort(self, request_id: str) -> None:
        st = self._states.get(request_id)
        if st is not None and not st.done:
            self.engine.abort(st.request_id)   # st.request_id is the engine-side id

    async def generate(self, prompt: TokensPrompt | dict, sampling_params: SamplingParams, request_id: str,
                       lora_request=None, trace_headers=None, **_: object  # noqa: ARG002
                       ) -> AsyncGenerator[RequestOutput, None]:
        if self._dead_error is not None:
            raise EngineDeadError(self._dead_error)
        lora_slot = 0
        if lora_request is not None and self._lora is None:
            raise ValueError("LoRA is not enabled: start the server with --enable-lora")
        prompt_ids = list(prompt["prompt_token_ids"] if isinstance(prompt, dict) else prompt.prompt_token_ids)
        sp = sampling_params
        eos = sp.eos_token_id if sp.eos_token_id is not None else getattr(self.tokenizer, "eos_token_id", None)
        max_tokens = sp.max_tokens if sp.max_tokens is not None else self._model_config.max_model_len - len(prompt_ids)
        native = make_sampling_params(
            greedy=sp.greedy, temperature=sp.temperature if not sp.greedy else 1.0,
            top_k=sp.top_k if sp.top_k and sp.top_k > 0 else 0, top_p=sp.top_p, typical_p=sp.typical_p,
            repetition_penalty=sp.repetition_penalty, length_penalty=sp.length_penalty,
            eos_token_id=eos if eos is not None else -1, min_tokens=sp.min_tokens, max_tokens=max_tokens,
            num_logprobs=sp.logprobs or 0, prompt_logprobs=sp.prompt_logprobs or 0,
            seed=sp.seed if not sp.greedy else None,
            stop_token_ids=sp.stop_token_ids, guided=sp.structured_outputs is not None)
        nid = f"q{next(self._ids)}"
        queue: asyncio.Queue = asyncio.Queue()
        detok = IncrementalDetokenizer(self.tokenizer, prompt_ids, stop=sp.stop, min_tokens=sp.min_tokens,
                                       include_stop_str_in_output=sp.include_stop_str_in_output,
                                       skip_special_tokens=sp.skip_special_tokens)
        st = _ReqState(nid, prompt_ids, sp, detok, queue)
        if request_id in self._states:   # vLLM rejects duplicate in-flight request ids as well
            raise ValueError(f"request id {request_id!r} is already in flight")
        self._states[nid] = st
        self._states[request_id] = st
        delta = sp.output_kind == RequestOutputKind.DELTA
        final_only = sp.output_kind == RequestOutputKind.FINAL_ONLY
        try:
            if sp.structured_outputs is not None:   # compile before the request exists: a bad spec is a ValueError
                self._guided().register(nid, sp.structured_outputs)
            if lora_request is not None:   # pins the adapter's slot (loading the weights if it has none) until `finally`
                lora_slot = await asyncio.get_running_loop().run_in_executor(None, self._lora.acquire, lora_request)
                native.lora_slot = lora_slot
            self.engine.add_request(nid, prompt_ids, native)
            sent_tokens = 0
            while True:
                item = await queue.get()
                if isinstance(item, BaseException):
                    raise item
                batch = [item]
                # DELTA streams emit one RequestOutput per engine step (the reference's tests pin "N tokens -> N+1
                # messages", tests/test_grpc_server.py:60-69) unless stream coalescing is on; FINAL_ONLY may swallow
                # whatever has already arrived.
                while not queue.empty() and (not delta or self._coalesce_streams or (
                        self._coalesce_streams is None and len(self._states) >= 2 * self._coalesce_min_streams)):
                    nxt = queue.get_nowait()
                    if isinstance(nxt, BaseException):
                        raise nxt
                    batch.append(nxt)
                finish_reason: str | None = None
                stop_reason: int | str | None = None
                last = batch[-1]
                for o in batch:   # one engine step at a time: stop strings are evaluated per step (S10)
                    if o.prompt_pos >= 1:   # prompt-logprob record (arrives before the first generated token)
                        plp: dict[int, Logprob] = {o.token_id: Logprob(o.logprob, o.rank)}
                        for r, (tid, tlp) in enumerate(o.topn, start=1):
                            plp.setdefault(tid, Logprob(tlp, r))
                        st.prompt_logprobs[o.prompt_pos] = plp
                        continue
                    step_finish: str | None = None
                    if o.finish_reason != _lib.FINISH_NONE:
                        if (o.finish_reason == _lib.FINISH_ABORT and self._mask_provider is not None
                                and (gerr := self._mask_provider.error_of(nid))):
                            raise RuntimeError(f"guided decoding failed: {gerr}")
                        if o.finish_reason == _lib.FINISH_ERROR:
                            # the message lives in the engine object (set on the engine thread); status() copies it
                            # into THIS thread's last-error slot
                            self.engine.status()
                            raise EngineDeadError(_lib.last_error(self.engine.lib) or "engine error")
                        step_finish = _FINISH[o.finish_reason]
                    if o.new_token is not None:
                        st.token_ids.append(o.new_token)
                        if sp.logprobs:
                            lp: dict[int, Logprob] = {o.new_token: Logprob(o.logprob, o.rank)}
                            for r, (tid, tlp) in enumerate(o.topn, start=1):
                                lp.setdefault(tid, Logprob(tlp, r))    # vllm logprobs.py:175-206 (dict merge dedups)
                            st.logprobs.append(lp)
                        stop_str = detok.update([o.new_token], stop_terminated=step_finish == "stop")
                        if stop_str is not None:
                            # stop string hit: later tokens (if any were already produced) are discarded and the
                            # engine-side sequence is released
                            finish_reason, stop_reason = "stop", stop_str
                            if step_finish is None:
                                self.engine.abort(nid)
                            last = o
                            break
                    if step_finish is not None:
                        finish_reason = step_finish
                        if o.finish_reason == _lib.FINISH_STOP_TOKEN:
                            stop_reason = o.stop_token_id
                        last = o
                        break
                if all(o.prompt_pos >= 1 for o in batch):
                    continue
                finished = finish_reason is not None
                if final_only and not finished:
                    continue
                text = detok.next_text(finished, delta)
                ids = st.token_ids[sent_tokens:] if delta else list(st.token_ids)
                lps = (st.logprobs[sent_tokens:] if delta else list(st.logprobs)) if sp.logprobs else None
                sent_tokens = len(st.token_ids)
                metrics = RequestMetrics(arrival_time=last.ts_arrival, first_scheduled_time=last.ts_first_scheduled,
                                         first_token_time=last.ts_first_token, last_token_time=last.ts_last_token,
                                         time_in_queue=(last.ts_first_scheduled - last.ts_arrival)
                                         if last.ts_first_scheduled else None,
                                         finished_time=last.ts_last_token if finished else None)
                yield RequestOutput(
                    request_id=request_id, prompt=None, prompt_token_ids=prompt_ids,
                    prompt_logprobs=st.prompt_logprobs if sp.prompt_logprobs else None,
                    outputs=[CompletionOutput(index=0, text=text, token_ids=ids, logprobs=lps,
                                              finish_reason=finish_reason, stop_reason=stop_reason)],
                    finished=finished, metrics=metrics)
                if finished:
                    st.done = True
                    self.metrics.observe_finished(
                        n_prompt=len(prompt_ids), n_generated=len(st.token_ids), finish_reason=finish_reason,
                        arrival=last.ts_arrival, first_scheduled=last.ts_first_scheduled,
                        first_token=last.ts_first_token, last_token=last.ts_last_token)
                    return
        finally:
            if not st.done:   # client went away / generator closed early: free the sequence in the engine
                self.engine.abort(nid)
                st.done = True
            if self._mask_provider is not None:
                self._mask_provider.unregister(nid)
            if lora_slot:
                self._lora.release(lora_slot)
            self._states.pop(nid, None)
            if self._states.get(request_id) is st:
                del self._states[request_id]
