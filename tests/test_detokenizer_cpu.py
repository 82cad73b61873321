"""Stop-string semantics (S10) checked against vLLM's own function when importable, plus streaming hold-back."""
import pytest

from vllm_tgis_adapter_b200.engine.detokenizer import IncrementalDetokenizer, check_stop_strings
from vllm_tgis_adapter_b200.engine.tokenizer import build_synthetic_tokenizer

CASES = [
    ("hello world stop here", 5, ["stop"], True), ("hello world stop here", 21, ["stop"], False),
    ("abcabc", 3, ["bc", "abc"], True), ("abcabc", 6, ["zz", "ca"], False), ("x", 1, ["x"], True),
    ("the end.", 4, ["end", "."], False), ("aaa", 1, ["aa"], True), ("no match", 8, ["qq"], True),
]


def test_check_stop_strings_matches_vllm():
    try:
        from vllm.v1.engine.detokenizer import check_stop_strings as ref
    except Exception:  # noqa: BLE001
        pytest.skip("vllm detokenizer not importable here")
    for text, new_chars, stop, include in CASES:
        assert check_stop_strings(text, new_chars, stop, include) == ref(text, new_chars, stop, include)


def test_first_stop_in_list_order_wins_and_truncation():
    assert check_stop_strings("abcabc", 6, ["ca", "bc"], False) == ("ca", 2)
    assert check_stop_strings("abcabc", 6, ["bc", "ca"], True) == ("bc", 3)
    assert check_stop_strings("abc", 3, ["bc"], True) == ("bc", -1)


def test_streaming_holds_back_partial_stop_text():
    tok = build_synthetic_tokenizer(256)
    d = IncrementalDetokenizer(tok, [5, 6], stop=["t9 t10"], min_tokens=0, include_stop_str_in_output=False,
                               skip_special_tokens=True)
    emitted = ""
    for t in (7, 8, 9, 10):
        s = d.update([t], False)
        emitted += d.next_text(s is not None, True)
        if s:
            break
    assert s == "t9 t10" and emitted == " t7 t8 " and "t9" not in emitted


def test_min_tokens_suppresses_stop_match():
    tok = build_synthetic_tokenizer(256)
    d = IncrementalDetokenizer(tok, [5], stop=["t7"], min_tokens=3, include_stop_str_in_output=True,
                               skip_special_tokens=True)
    assert d.update([7], False) is None and d.update([8], False) is None and d.update([9], False) is None
    assert d.update([7], False) == "t7"


def test_backend_fast_path_is_text_identical_to_the_wrapper():
    """The incremental detokenizer calls the Rust tokenizer's decode directly (skipping transformers' Python wrapper);
    with clean_up_tokenization_spaces=False both must yield the same text, also with special tokens in the stream."""
    import random

    from vllm_tgis_adapter_b200.engine.detokenizer import IncrementalDetokenizer
    from vllm_tgis_adapter_b200.engine.tokenizer import build_synthetic_tokenizer

    tok = build_synthetic_tokenizer(512)
    rnd = random.Random(3)
    specials = [i for i in (tok.bos_token_id, tok.eos_token_id) if i is not None]
    for skip in (True, False):
        d = IncrementalDetokenizer(tok, [5, 6, 7], stop=None, min_tokens=0, include_stop_str_in_output=True,
                                   skip_special_tokens=skip)
        assert d._backend_decode is not None
        for _ in range(50):
            ids = [rnd.randrange(3, 512) for _ in range(rnd.randrange(1, 6))] + rnd.sample(specials, k=min(1, len(specials)))
            rnd.shuffle(ids)
            assert d._decode(ids) == tok.decode(ids, skip_special_tokens=skip, clean_up_tokenization_spaces=False)
