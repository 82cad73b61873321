"""Incremental detokenisation + stop-string matching, host side (needs the HF tokenizer, so it stays in Python;
SURVEY.md §8b).  Semantics restated from vLLM 0.22 (what the reference gets behind `engine.generate`):

  update / get_next_output_text   vllm v1/engine/detokenizer.py:95-165
  check_stop_strings              vllm v1/engine/detokenizer.py:309-344
  incremental decode              vllm tokenizers/detokenizer_utils.py (prefix_offset / read_offset scheme, 5-token
                                  prompt suffix as initial context)
"""
from __future__ import annotations

INITIAL_INCREMENTAL_DETOKENIZATION_OFFSET = 5


def check_stop_strings(output_text: str, new_char_count: int, stop: list[str], include_in_output: bool
                       ) -> tuple[str, int] | None:
    """First stop string IN LIST ORDER that occurs in the not-yet-searched tail wins.  Returns (stop, truncate_to)
    with truncate_to == -1 meaning "keep everything"."""
    if not new_char_count or not stop:
        return None
    for s in stop:
        idx = output_text.find(s, 1 - new_char_count - len(s))
        if idx == -1:
            continue
        if include_in_output:
            idx += len(s)
            if idx >= len(output_text):
                return s, -1
        return s, idx
    return None


class IncrementalDetokenizer:
    def __init__(self, tokenizer, prompt_token_ids: list[int], *, stop: list[str] | None, min_tokens: int,
                 include_stop_str_in_output: bool, skip_special_tokens: bool):
        self.tokenizer = tokenizer
        self.stop = list(stop or [])
        self.min_tokens = min_tokens
        self.include_stop = include_stop_str_in_output
        self.skip_special = skip_special_tokens
        self.stop_buffer_length = (max(len(s) for s in self.stop) - 1) if (self.stop and not self.include_stop) else 0
        self._last_offset = 0
        self.output_text = ""
        self.n_prompt = len(prompt_token_ids)
        ctx = prompt_token_ids[-INITIAL_INCREMENTAL_DETOKENIZATION_OFFSET:]
        self.ids = list(ctx)                 # decoding window source: prompt suffix + all output ids
        self.prefix_offset = 0
        self.read_offset = len(ctx)
        self.n_out = 0
        # Two decodes per generated token sit on the serving loop's critical path.  For a fast (Rust) tokenizer call the
        # backend directly: PreTrainedTokenizerFast.decode(..., clean_up_tokenization_spaces=False) is exactly
        # backend.decode(ids, skip_special_tokens=...) plus ~15 us of Python wrapping per call (what vLLM's own fast
        # incremental detokenizer bypasses too, v1/engine/detokenizer.py FastIncrementalDetokenizer).
        backend = getattr(tokenizer, "backend_tokenizer", None)
        self._backend_decode = backend.decode if (backend is not None and getattr(tokenizer, "is_fast", False)) else None

    def _decode(self, ids: list[int]) -> str:
        if self._backend_decode is not None:
            return self._backend_decode(ids, skip_special_tokens=self.skip_special)
        return self.tokenizer.decode(ids, skip_special_tokens=self.skip_special,
                                     clean_up_tokenization_spaces=False)

    def _decode_next(self, token_id: int) -> str:
        self.ids.append(token_id)
        prefix_text = self._decode(self.ids[self.prefix_offset:self.read_offset])
        new_text = self._decode(self.ids[self.prefix_offset:])
        if len(new_text) > len(prefix_text) and not new_text.endswith("�"):
            self.prefix_offset = self.read_offset
            self.read_offset = len(self.ids)
            return new_text[len(prefix_text):]
        return ""

    def update(self, new_token_ids: list[int], stop_terminated: bool) -> str | None:
        """Returns the matched stop string, if any (and truncates output_text accordingly)."""
        if not new_token_ids:
            return None
        skipped = None
        if stop_terminated and not self.include_stop:
            skipped = new_token_ids[-1]
            new_token_ids = new_token_ids[:-1]
        stop_check_offset = len(self.output_text)
        for t in new_token_ids:
            self.n_out += 1
            self.output_text += self._decode_next(t)
            if self.min_tokens and self.n_out <= self.min_tokens:
                stop_check_offset = len(self.output_text)
        if skipped is not None:
            self.n_out += 1
        if self.stop and self.n_out > self.min_tokens:
            hit = check_stop_strings(self.output_text, len(self.output_text) - stop_check_offset, self.stop,
                                     self.include_stop)
            if hit is not None:
                s, cut = hit
                if cut != -1:
                    self.output_text = self.output_text[:cut]
                return s
        return None

    def next_text(self, finished: bool, delta: bool) -> str:
        buf = 0 if finished else self.stop_buffer_length
        if not delta:
            return self.output_text if not buf else self.output_text[:-buf]
        length = len(self.output_text) - buf
        if self._last_offset < length:
            out = self.output_text[self._last_offset:length]
            self._last_offset = length
            return out
        return ""
