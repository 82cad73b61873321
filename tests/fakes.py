"""CPU test double for the native engine (tests only): same surface as engine.core.NativeEngine, deterministic token
stream, so the whole host stack (async engine wrapper, detokenizer, gRPC servicer) runs on a GPU-less box."""
from __future__ import annotations

import queue
import threading
import time
import types

from vllm_tgis_adapter_b200.engine import _lib
from vllm_tgis_adapter_b200.engine.core import ModelConfig, StepOutput


def fake_next_token(prompt: list[int], n_out: int, vocab: int) -> int:
    return 3 + (sum(prompt) * 31 + n_out * 7919) % (vocab - 3)


class FakeNativeEngine:
    def __init__(self, model: ModelConfig, step_delay: float = 0.0, script: dict | None = None):
        self.model = model
        self.lib = types.SimpleNamespace(tgis_last_error=lambda: b"")
        self._in: queue.Queue = queue.Queue()
        self._out: queue.Queue = queue.Queue()
        self._active: dict[str, dict] = {}
        self._stop = False
        self._thread: threading.Thread | None = None
        self.step_delay = step_delay
        self.script = script or {}          # prompt tuple -> explicit token list
        self.errored = False
        self.aborted: list[str] = []
        self.steps = 0
        self._mask_cb = None
        self.max_loras, self.max_lora_rank = 0, 16   # tests that exercise adapters raise max_loras
        self.adapter_loads: list[tuple[int, int]] = []
        self.lora_slots_seen: list[int] = []

    def load_adapter(self, slot, weights) -> None:    # engine.core.NativeEngine.load_adapter
        self.adapter_loads.append((slot, len(weights)))

    def set_mask_provider(self, callback) -> None:   # engine.core.NativeEngine.set_mask_provider
        self._mask_cb = callback

    def _guided_token(self, rid, st, tok):
        """What the native engine does for a guided row (csrc/engine.cu run_batch): report the tokens generated since
        the last call, take the bitmask, and sample under it -- here: the first allowed id at or after `tok`, cyclically.
        Returns None when the provider failed (the request is aborted)."""
        import ctypes as C

        words = (self.model.vocab + 31) // 32
        bits = (C.c_uint32 * words)()
        fed = st.setdefault("fed", 0)
        new = st["out"][fed:]
        arr = (C.c_int32 * max(1, len(new)))(*new)
        rc = self._mask_cb(None, rid.encode(), arr, len(new), bits, words)
        st["fed"] = len(st["out"])
        if rc < 0:
            return None
        if rc == 1:
            return tok
        for d in range(self.model.vocab):
            i = (tok + d) % self.model.vocab
            if (bits[i >> 5] >> (i & 31)) & 1:
                return i
        return None

    def start(self) -> None:
        self._thread = threading.Thread(target=self._loop, daemon=True)
        self._thread.start()

    def add_request(self, request_id, prompt_ids, params) -> None:
        self._in.put(("add", request_id, list(prompt_ids), params, time.monotonic()))

    def abort(self, request_id) -> None:
        self._in.put(("abort", request_id))

    def _loop(self) -> None:
        while not self._stop:
            try:
                while True:
                    msg = self._in.get_nowait()
                    if msg[0] == "add":
                        _, rid, prompt, sp, ts = msg
                        self.lora_slots_seen.append(getattr(sp, "lora_slot", 0))
                        self._active[rid] = {"prompt": prompt, "sp": sp, "out": [], "ts": ts, "first": 0.0}
                    else:
                        if msg[1] in self._active:
                            st = self._active.pop(msg[1])
                            self.aborted.append(msg[1])
                            self._emit(msg[1], st, None, _lib.FINISH_ABORT, -1)
            except queue.Empty:
                pass
            if not self._active:
                time.sleep(0.002)
                continue
            self.steps += 1
            if self.step_delay:
                time.sleep(self.step_delay)
            for rid in list(self._active):
                st = self._active[rid]
                sp = st["sp"]
                scripted = self.script.get(tuple(st["prompt"]))
                if scripted is not None and len(st["out"]) < len(scripted):
                    tok = scripted[len(st["out"])]
                else:
                    tok = fake_next_token(st["prompt"], len(st["out"]), self.model.vocab)
                if len(st["out"]) < sp.min_tokens and tok == sp.eos_token_id:
                    tok = 3
                if getattr(sp, "guided", 0) and self._mask_cb is not None:
                    tok = self._guided_token(rid, st, tok)
                    if tok is None:
                        self._active.pop(rid)
                        self.aborted.append(rid)
                        self._emit(rid, st, None, _lib.FINISH_ABORT, -1)
                        continue
                st["out"].append(tok)
                now = time.monotonic()
                st["first"] = st["first"] or now
                n_out = len(st["out"])
                finish, stop_tok = _lib.FINISH_NONE, -1
                if n_out >= sp.min_tokens:
                    if tok == sp.eos_token_id:
                        finish = _lib.FINISH_STOP_EOS
                    elif tok in list(sp.stop_token_ids)[: sp.n_stop_token_ids]:
                        finish, stop_tok = _lib.FINISH_STOP_TOKEN, tok
                    elif len(st["prompt"]) + n_out >= self.model.max_model_len or n_out >= sp.max_tokens:
                        finish = _lib.FINISH_LENGTH
                self._emit(rid, st, tok, finish, stop_tok)
                if finish != _lib.FINISH_NONE:
                    self._active.pop(rid)

    def _emit(self, rid, st, tok, finish, stop_tok) -> None:
        n = st["sp"].num_logprobs
        now = time.monotonic()
        self._out.put(StepOutput(
            request_id=rid, new_token=tok, logprob=-0.5 - 0.01 * len(st["out"]), rank=1 + len(st["out"]) % 3,
            topn=[((tok or 0) + j, -0.5 - j) for j in range(n)] if tok is not None else [],
            finish_reason=finish, stop_token_id=stop_tok, n_prompt_tokens=len(st["prompt"]),
            n_output_tokens=len(st["out"]), ts_arrival=st["ts"], ts_first_scheduled=st["ts"] + 1e-4,
            ts_first_token=st["first"] or now, ts_last_token=now, token_id=tok if tok is not None else -1))

    def poll(self, timeout_ms: int = 0):
        outs = []
        try:
            outs.append(self._out.get(timeout=timeout_ms / 1e3 if timeout_ms else 0.0001))
            while True:
                outs.append(self._out.get_nowait())
        except queue.Empty:
            pass
        return outs

    def status(self):
        return types.SimpleNamespace(errored=int(self.errored), is_running=int(not self._stop), n_running=len(self._active),
                                     n_waiting=0, free_blocks=10, total_blocks=10, tokens_generated=0, steps=self.steps,
                                     kernel_launches=0, gpu_busy_ms=0.0)

    def close(self) -> None:
        self._stop = True
