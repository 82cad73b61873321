"""Guided decoding, host side: grammar matchers that feed the engine's mask provider hook.

Reference seam: /root/reference/src/vllm_tgis_adapter/tgis_utils/structured_outputs.py:14-38 turns the request's
`DecodingParameters.guided` oneof into vLLM `StructuredOutputsParams`; vLLM then (v1/structured_output/backend_xgrammar.py)
compiles the spec with the xgrammar library, keeps one `GrammarMatcher` per request, fills a token bitmask per step
(v1/structured_output/__init__.py:204-300) and masks the logits before its sampler.  Here the split is the same: xgrammar
(third-party library, exactly as in the reference stack) owns the grammar automaton; the engine calls
`MaskProvider._on_step` through the C ABI (`tgis_engine_set_mask_provider`, include/tgis_engine.h) right before a step
that samples for a guided request, and the fused sampling kernel applies the bits (csrc/sampler.cu, SAMPLE_MASKED rows).

Spec -> grammar rules restate vLLM 0.22 (`backend_xgrammar.py:compile_grammar`, `utils.py:choice_as_grammar`):
  json         compile_json_schema(schema, any_whitespace=True)
  json_object  compile_json_schema('{"type": "object"}', any_whitespace=True)
  regex        compile_regex(pattern)
  choice       EBNF  root ::= "a" | "b" | ...   (quotes and backslashes escaped)
  grammar      compile_grammar(ebnf)            (lark-style grammars are converted by vLLM's `convert_lark_to_ebnf`;
                                                 here only EBNF/GBNF text is accepted, which is what the reference's own
                                                 test passes: tests/test_grpc_server.py:15-29)
"""
from __future__ import annotations

import ctypes as C
import json
import re
import threading

import numpy as np

from .types import StructuredOutputsParams

MASK_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_char_p, C.POINTER(C.c_int32), C.c_int32, C.POINTER(C.c_uint32),
                      C.c_int32)


def choice_as_grammar(choices: list[str]) -> str:
    """vllm v1/structured_output/utils.py:451-459."""
    quoted = ['"' + re.sub(r'(["\\])', r"\\\1", c) + '"' for c in choices]
    return "root ::= " + " | ".join(quoted)


class GrammarCompiler:
    """One xgrammar compiler per tokenizer (compiled grammars are cached inside xgrammar)."""

    def __init__(self, tokenizer, vocab_size: int):
        import xgrammar as xgr

        self._xgr = xgr
        self.vocab_size = vocab_size
        info = xgr.TokenizerInfo.from_huggingface(tokenizer, vocab_size=vocab_size)
        self._compiler = xgr.GrammarCompiler(info, max_threads=8, cache_enabled=True)

    def compile(self, p: StructuredOutputsParams):
        c = self._compiler
        if p.json is not None:
            spec = p.json if isinstance(p.json, str) else json.dumps(p.json)
            ctx = c.compile_json_schema(spec, any_whitespace=True)
        elif p.json_object:
            ctx = c.compile_json_schema('{"type": "object"}', any_whitespace=True)
        elif p.regex is not None:
            ctx = c.compile_regex(p.regex)
        elif p.choice is not None:
            ctx = c.compile_grammar(choice_as_grammar(p.choice))
        elif p.grammar is not None:
            ctx = c.compile_grammar(p.grammar)
        else:
            raise ValueError("empty structured output specification")
        return self._xgr.GrammarMatcher(ctx)


class _Guide:
    __slots__ = ("matcher", "terminated", "n_tokens", "error")

    def __init__(self, matcher):
        self.matcher = matcher
        self.terminated = False
        self.n_tokens = 0
        self.error: str | None = None


class MaskProvider:
    """request id -> matcher; `callback` is the C function pointer the engine calls on its step thread."""

    def __init__(self, compiler: GrammarCompiler):
        self.compiler = compiler
        self._guides: dict[str, _Guide] = {}
        self._lock = threading.Lock()
        self.callback = MASK_FN(self._on_step)   # keep a reference: the engine holds the raw pointer
        self.calls = 0

    def register(self, request_id: str, params: StructuredOutputsParams) -> None:
        """Compile (a ValueError / RuntimeError from xgrammar surfaces to the caller as an invalid request)."""
        try:
            matcher = self.compiler.compile(params)
        except Exception as e:  # noqa: BLE001  xgrammar raises RuntimeError for malformed specs
            raise ValueError(f"invalid structured output specification: {e}") from e
        with self._lock:
            self._guides[request_id] = _Guide(matcher)

    def unregister(self, request_id: str) -> None:
        with self._lock:
            self._guides.pop(request_id, None)

    def error_of(self, request_id: str) -> str | None:
        with self._lock:
            g = self._guides.get(request_id)
        return g.error if g is not None else None

    def _on_step(self, _user, request_id, new_tokens, n_new, allow_bits, n_words) -> int:
        """tgis_mask_fn (include/tgis_engine.h): advance over the tokens generated since the last call, then write the
        allowed-token bits of the next step.  0 = written, 1 = unconstrained, < 0 = failure (the engine aborts the request)."""
        try:
            import torch

            with self._lock:
                g = self._guides.get(request_id.decode())
            if g is None:
                return -1
            self.calls += 1
            # vllm backend_xgrammar.py:148-166 accept_tokens: a token the automaton rejects is an error, and nothing is
            # fed after termination
            for i in range(n_new):
                if g.terminated:
                    break
                if not g.matcher.accept_token(int(new_tokens[i])):
                    g.error = f"token {int(new_tokens[i])} rejected by the grammar after {g.n_tokens} tokens"
                    return -2
                g.n_tokens += 1
                g.terminated = g.matcher.is_terminated()
            if g.terminated:
                return 1   # the stop token was accepted: the engine's EOS check ends the request
            words = np.ctypeslib.as_array(allow_bits, shape=(1, n_words)).view(np.int32)
            g.matcher.fill_next_token_bitmask(torch.from_numpy(words), 0)
            return 0
        except Exception as e:  # noqa: BLE001  never unwind into the C++ engine thread
            try:
                with self._lock:
                    g = self._guides.get(request_id.decode())
                if g is not None:
                    g.error = f"{type(e).__name__}: {e}"
            except Exception:  # noqa: BLE001
                pass
            return -3
