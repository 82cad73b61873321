"""Tokenizers for the server: a HF tokenizer directory when one exists, otherwise the synthetic WordLevel tokenizer
that the offline build uses everywhere (no tokenizer files exist in this environment).

Synthetic vocabulary: id 0 `<unk>`, 1 `<s>`, 2 `</s>`, id n>=3 -> word `t{n}`; whitespace pre-tokenisation, so a prompt
of 512 whitespace-separated words is exactly 512 ids and `Tokenize` offsets are exact.  truncation_side is "left" like
vLLM's generate-runner tokenizers (vllm tokenizers/registry.py:143-147)."""
from __future__ import annotations

import functools
from pathlib import Path

UNK, BOS, EOS = 0, 1, 2


@functools.lru_cache(maxsize=4)
def build_synthetic_tokenizer(vocab_size: int):
    from tokenizers import Tokenizer, models, pre_tokenizers
    from transformers import PreTrainedTokenizerFast

    vocab = {"<unk>": UNK, "<s>": BOS, "</s>": EOS}
    for i in range(3, vocab_size):
        vocab[f"t{i}"] = i
    tok = Tokenizer(models.WordLevel(vocab=vocab, unk_token="<unk>"))
    tok.pre_tokenizer = pre_tokenizers.WhitespaceSplit()
    # no decoder: WordLevel's default decode joins tokens with a single space
    fast = PreTrainedTokenizerFast(tokenizer_object=tok, unk_token="<unk>", bos_token="<s>", eos_token="</s>",
                                   truncation_side="left", clean_up_tokenization_spaces=False)
    return fast


def load_tokenizer(path: str | None, vocab_size: int, *, allow_synthetic: bool = False):
    """HF tokenizer from `path`.  The synthetic tokenizer is only handed out when the caller says the model itself is
    synthetic (`--synthetic-weights` / a preset): a real checkpoint with a missing or mistyped tokenizer path must fail
    at start-up like the reference does (vLLM raises from `get_tokenizer`), not serve `t{n}` words."""
    if path and Path(path).is_dir() and any((Path(path) / f).exists()
                                            for f in ("tokenizer.json", "tokenizer.model", "tokenizer_config.json")):
        from transformers import AutoTokenizer

        tok = AutoTokenizer.from_pretrained(path, truncation_side="left")
        return tok
    if not allow_synthetic:
        raise ValueError(f"no tokenizer files (tokenizer.json / tokenizer.model / tokenizer_config.json) under "
                         f"{path!r}: pass --tokenizer <dir>, or --synthetic-weights for the synthetic vocabulary")
    return build_synthetic_tokenizer(vocab_size)


def synthetic_prompt(token_ids) -> str:
    """Text that the synthetic tokenizer maps back to exactly `token_ids` (ids must be >= 3)."""
    return " ".join(f"t{int(i)}" for i in token_ids)
