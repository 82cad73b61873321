"""Guided decoding, host side (no GPU).  What the reference pins at this seam
(/root/reference/tests/test_grpc_server.py:170-232 `test_guided_decoding_request`): each member of the `guided` oneof
reaches the engine as the matching StructuredOutputsParams field.  Added here: the mask-provider hook the native engine
calls (include/tgis_engine.h `tgis_mask_fn`), the oracle's restatement of the logits masking pinned to xgrammar's own CPU
kernel, and the whole host path over a real gRPC channel on the fake engine (the output obeys the constraint)."""
import ctypes as C
import re

import grpc
import numpy as np
import pytest
import torch

xgr = pytest.importorskip("xgrammar")

from test_grpc_server_cpu import Server, _params, VOCAB  # noqa: E402
from vllm_tgis_adapter_b200.engine.guided import GrammarCompiler, MaskProvider, choice_as_grammar  # noqa: E402
from vllm_tgis_adapter_b200.engine.tokenizer import build_synthetic_tokenizer  # noqa: E402
from vllm_tgis_adapter_b200.engine.types import SamplingParams, StructuredOutputsParams  # noqa: E402
from vllm_tgis_adapter_b200.grpc import grpc_server  # noqa: E402
from vllm_tgis_adapter_b200.grpc.pb import generation_pb2 as pb  # noqa: E402

SQL_GRAMMAR = """
    root ::= select_statement
    select_statement ::= "SELECT " column " from " table " where " condition
    column ::= "col_1 " | "col_2 "
    table ::= "table_1 " | "table_2 "
    condition ::= column "= " number
    number ::= "1 " | "2 "
"""


def _decoding(**kw):
    d = pb.DecodingParameters()
    for k, v in kw.items():
        if k == "choice":
            d.choice.choices.extend(v)
        else:
            setattr(d, k, v)
    return d


def test_guided_oneof_maps_to_structured_outputs_params():
    f = grpc_server._structured_output_params
    assert f(pb.DecodingParameters()) is None
    assert f(_decoding(repetition_penalty=1.2)) is None
    schema = '{"type": "object", "properties": {"name": {"type": "string"}, "age": {"type": "integer"}}}'
    assert f(_decoding(json_schema=schema)) == StructuredOutputsParams(json=schema)
    assert f(_decoding(regex=r"\d.\d+")) == StructuredOutputsParams(regex=r"\d.\d+")
    assert f(_decoding(choice=["1", "2", "3", "4"])) == StructuredOutputsParams(choice=["1", "2", "3", "4"])
    assert f(_decoding(grammar=SQL_GRAMMAR)) == StructuredOutputsParams(grammar=SQL_GRAMMAR)
    assert f(_decoding(format=pb.DecodingParameters.JSON)) == StructuredOutputsParams(json_object=True)
    with pytest.raises(ValueError, match="Must provide at least two choices"):
        f(_decoding(choice=["only"]))
    with pytest.raises(ValueError, match="format"):
        f(_decoding(format=pb.DecodingParameters.TEXT))
    with pytest.raises(ValueError, match="only use one kind"):
        StructuredOutputsParams(regex="a", choice=["a", "b"])
    with pytest.raises(ValueError, match="none are specified"):
        StructuredOutputsParams()


def test_choice_grammar_matches_vllm():
    vllm_utils = pytest.importorskip("vllm.v1.structured_output.utils")
    for choices in (["1", "2"], ['say "hi"', "back\\slash", "plain"], ["a b", "c"]):
        assert choice_as_grammar(choices) == vllm_utils.choice_as_grammar(choices)


def test_oracle_mask_restatement_matches_xgrammar_kernel():
    from oracle.sampler_oracle import apply_token_bitmask, unpack_token_bitmask

    rng = np.random.RandomState(3)
    for V in (1000, 1024, 4096 + 8):
        words = (V + 31) // 32
        bm = rng.randint(-2**31, 2**31 - 1, size=(3, words), dtype=np.int64).astype(np.int32)
        logits = torch.randn(3, V)
        ref = logits.clone()
        xgr.apply_token_bitmask_inplace(ref, torch.from_numpy(bm))
        for r in range(3):
            got = apply_token_bitmask(logits[r], unpack_token_bitmask(bm[r], V))
            assert torch.equal(got, ref[r])


def _call(provider, rid, new_tokens, words):
    bits = (C.c_uint32 * words)()
    arr = (C.c_int32 * max(1, len(new_tokens)))(*new_tokens)
    rc = provider.callback(None, rid.encode(), arr, len(new_tokens), bits, words)
    allowed = [i for i in range(words * 32) if (bits[i >> 5] >> (i & 31)) & 1]
    return rc, allowed


def test_mask_provider_hook_walks_the_grammar():
    tok = build_synthetic_tokenizer(VOCAB)
    prov = MaskProvider(GrammarCompiler(tok, VOCAB))
    words = (VOCAB + 31) // 32
    # the synthetic vocabulary spells token n as "t{n}": t1\dt2\d\d = one token of 10..19 then one of 200..299
    prov.register("a", StructuredOutputsParams(regex=r"t1\dt2\d\d"))
    rc, allowed = _call(prov, "a", [], words)
    assert rc == 0 and allowed == list(range(10, 20))
    rc, allowed = _call(prov, "a", [13], words)
    assert rc == 0 and allowed == [20, 21, 22, 23, 24, 25, 26, 27, 28, 29] + list(range(200, 300))
    rc, allowed = _call(prov, "a", [], words)      # a recomputed (preempted) sequence asks again: same answer
    assert rc == 0 and 250 in allowed and 20 in allowed
    rc, allowed = _call(prov, "a", [250], words)
    assert rc == 0 and allowed == [2]              # complete: only the stop token (</s> = 2) remains
    rc, _ = _call(prov, "a", [2], words)
    assert rc == 1                                 # terminated: nothing left to constrain
    # choice / grammar / json compile through the same door
    prov.register("b", StructuredOutputsParams(choice=["t5", "t77t78"]))
    rc, allowed = _call(prov, "b", [], words)
    assert rc == 0 and allowed == [5, 7, 77]       # "t7" is a proper prefix of "t77t78"
    rc, allowed = _call(prov, "b", [77], words)
    assert allowed == [7, 78]
    prov.register("c", StructuredOutputsParams(json_object=True))
    assert _call(prov, "c", [], words)[0] == 0
    # a token the automaton rejects is a failure the engine turns into an abort; the message is kept for the caller
    rc, _ = _call(prov, "b", [999], words)
    assert rc < 0 and "rejected" in prov.error_of("b")
    assert _call(prov, "nobody", [], words)[0] < 0
    with pytest.raises(ValueError, match="invalid structured output"):
        prov.register("d", StructuredOutputsParams(grammar="root ::= ("))
    prov.unregister("a")
    assert _call(prov, "a", [], words)[0] < 0


@pytest.fixture()
def srv():
    s = Server()
    yield s
    s.close()


def test_guided_requests_over_grpc_obey_their_constraint(srv):
    """Generate + GenerateStream with each guided kind on the fake engine (which samples under the provider's bitmask the
    way the native engine does): the decoded text -- synthetic tokens joined by spaces -- spells a member of the language."""
    def text_of(params):
        r = srv.generate(["t5 t6 t7"], params).responses[0]
        return r.text.replace(" ", ""), r

    p = _params(stopping={"max_new_tokens": 16})
    p.decoding.regex = r"t1\dt1\d"   # (no dead ends in the synthetic vocabulary: there is no bare-digit token)
    text, r = text_of(p)
    assert re.fullmatch(r"t1\dt1\d", text) and r.stop_reason == pb.StopReason.EOS_TOKEN and r.generated_token_count == 3
    p = _params(stopping={"max_new_tokens": 16})
    p.decoding.choice.choices.extend(["t10t11", "t12"])
    text, r = text_of(p)
    assert text in ("t10t11", "t12") and r.stop_reason == pb.StopReason.EOS_TOKEN
    p = _params(stopping={"max_new_tokens": 16})
    p.decoding.grammar = 'root ::= "t10" ("t11" | "t12") "t13"'
    text, r = text_of(p)
    assert text in ("t10t11t13", "t10t12t13")
    chunks = srv.stream("t5 t6 t7", p)
    assert "".join(c.text for c in chunks).replace(" ", "") in ("t10t11t13", "t10t12t13")
    # error cases travel as the reference's: fewer than two choices is a ValueError inside the handler (UNKNOWN)
    p = _params(stopping={"max_new_tokens": 4})
    p.decoding.choice.choices.extend(["one"])
    with pytest.raises(grpc.RpcError) as ei:
        srv.generate(["t5"], p)
    assert "Must provide at least two choices" in ei.value.details()
    p = _params(stopping={"max_new_tokens": 4})
    p.decoding.grammar = "root ::= ("
    with pytest.raises(grpc.RpcError) as ei:
        srv.generate(["t5"], p)
    assert "invalid structured output specification" in ei.value.details()
    # the provider forgets finished requests
    assert not srv.engine._mask_provider._guides


def test_sampling_params_carry_structured_outputs():
    sp = SamplingParams(temperature=0.0, structured_outputs=StructuredOutputsParams(regex="a+"))
    assert sp.structured_outputs.regex == "a+"
