"""LoRA adapters, host side (no GPU).  Mirrors /root/reference/tests/test_adapters.py (a LoRA adapter_id becomes a
`lora_request`, the load happens once, later requests reuse the engine's registry entry) and adds what the reference never
tests: the error strings of adapters.py, the PEFT checkpoint reader's vLLM rules, slot recycling, and the oracle's LoRA
arithmetic pinned to vLLM's own reference ops."""
import asyncio
import json
import types
from pathlib import Path

import pytest
import torch

from vllm_tgis_adapter_b200.engine.lora import LoRAConfigError, LoRAManager, read_adapter
from vllm_tgis_adapter_b200.engine.types import LoRARequest
from vllm_tgis_adapter_b200.grpc.adapters import AdapterStore, validate_adapters

GOLD = Path(__file__).parent / "golden"


def make_peft_dir(root: Path, name: str, *, n_layers=2, hidden=256, q_dim=512, kv_dim=256, ffn=512, r=8, alpha=16,
                  modules=("q_proj", "k_proj", "v_proj", "o_proj", "gate_proj", "up_proj", "down_proj"), seed=0, std=0.05,
                  **cfg_extra) -> Path:
    from safetensors.torch import save_file

    d = root / name
    d.mkdir(parents=True)
    cfg = {"peft_type": "LORA", "r": r, "lora_alpha": alpha, "target_modules": list(modules), "bias": "none",
           "use_dora": False, "use_rslora": False, "modules_to_save": None, "task_type": "CAUSAL_LM"}
    cfg.update(cfg_extra)
    (d / "adapter_config.json").write_text(json.dumps(cfg))
    shapes = {"q_proj": (hidden, q_dim), "k_proj": (hidden, kv_dim), "v_proj": (hidden, kv_dim), "o_proj": (q_dim, hidden),
              "gate_proj": (hidden, ffn), "up_proj": (hidden, ffn), "down_proj": (ffn, hidden)}
    g = torch.Generator().manual_seed(seed)
    tensors = {}
    for li in range(n_layers):
        for m in modules:
            fin, fout = shapes[m]
            blk = "self_attn" if m in ("q_proj", "k_proj", "v_proj", "o_proj") else "mlp"
            base = f"base_model.model.model.layers.{li}.{blk}.{m}"
            tensors[base + ".lora_A.weight"] = (torch.randn(r, fin, generator=g) * std).to(torch.bfloat16)
            tensors[base + ".lora_B.weight"] = (torch.randn(fout, r, generator=g) * std).to(torch.bfloat16)
    save_file(tensors, str(d / "adapter_model.safetensors"))
    return d


def test_oracle_lora_add_matches_vllm_reference_ops():
    from oracle.llama_oracle import lora_add

    fix = json.loads((GOLD / "lora_torch_ops.json").read_text())
    x = torch.tensor(fix["x"]).to(torch.bfloat16)
    y = torch.tensor(fix["y"]).to(torch.bfloat16)
    a = torch.tensor(fix["a"]).to(torch.bfloat16)
    b = torch.tensor(fix["b"]).to(torch.bfloat16)
    want = torch.tensor(fix["out"])
    buf = torch.tensor(fix["buffer"])
    exact = 0
    for t, s in enumerate(fix["idx"]):
        got = lora_add(y[t:t + 1], x[t:t + 1], a[s], b[s]).float()[0]
        # fp32 shrink sums agree to fp32 rounding; the bf16 results may differ by one ulp where a sum lands on a tie
        assert torch.allclose((x[t:t + 1].float() @ a[s].float().t())[0], buf[t], rtol=1e-5, atol=1e-5)
        ulp = torch.maximum(want[t].abs(), torch.tensor(1e-3)) * 2.0 ** -7
        assert bool(((got - want[t]).abs() <= ulp).all())
        exact += int((got == want[t]).sum())
    assert exact >= 0.97 * want.numel(), exact


def test_read_adapter_follows_vllm_rules(tmp_path):
    d = make_peft_dir(tmp_path, "ok", r=8, alpha=32)
    r, w = read_adapter(str(d), n_layers=2, max_rank=16)
    assert r == 8 and len(w) == 2 * 7
    from safetensors.torch import load_file

    raw = load_file(str(d / "adapter_model.safetensors"))
    a, b = w[(1, "down_proj")]
    assert torch.equal(a, raw["base_model.model.model.layers.1.mlp.down_proj.lora_A.weight"])
    # alpha / r = 4 folded into B with a bf16 multiply (lora_weights.py optimize)
    assert torch.equal(b, raw["base_model.model.model.layers.1.mlp.down_proj.lora_B.weight"] * 4.0)
    assert a.dtype == b.dtype == torch.bfloat16
    r2, w2 = read_adapter(str(make_peft_dir(tmp_path, "rs", r=4, alpha=8, use_rslora=True, modules=("q_proj",))),
                          n_layers=2, max_rank=16)
    raw2 = load_file(str(tmp_path / "rs" / "adapter_model.safetensors"))
    assert torch.equal(w2[(0, "q_proj")][1], raw2["base_model.model.model.layers.0.self_attn.q_proj.lora_B.weight"] * 4.0)
    with pytest.raises(LoRAConfigError, match="greater than max_lora_rank"):
        read_adapter(str(d), n_layers=2, max_rank=4)
    with pytest.raises(LoRAConfigError, match="DoRA"):
        read_adapter(str(make_peft_dir(tmp_path, "dora", use_dora=True)), n_layers=2, max_rank=16)
    with pytest.raises(LoRAConfigError, match="bias"):
        read_adapter(str(make_peft_dir(tmp_path, "bias", bias="all")), n_layers=2, max_rank=16)
    with pytest.raises(LoRAConfigError, match="has 1 layers"):
        read_adapter(str(d), n_layers=1, max_rank=16)
    bad = make_peft_dir(tmp_path, "badmod", modules=("q_proj",))
    from safetensors.torch import save_file

    t = load_file(str(bad / "adapter_model.safetensors"))
    t["base_model.model.model.layers.0.self_attn.rotary.lora_A.weight"] = torch.zeros(8, 8, dtype=torch.bfloat16)
    save_file(t, str(bad / "adapter_model.safetensors"))
    with pytest.raises(LoRAConfigError, match="unsupported LoRA weight"):
        read_adapter(str(bad), n_layers=2, max_rank=16)


class _Engine:
    """What adapters.py needs from the engine façade (the reference mocks OpenAIServingModels the same way)."""

    def __init__(self):
        self.lora_requests: dict[str, LoRARequest] = {}
        self.load_calls = []

    async def load_lora_adapter(self, *, lora_name, lora_path, lora_int_id):
        self.load_calls.append((lora_name, lora_path))
        self.lora_requests[lora_name] = LoRARequest(lora_name=lora_name, lora_int_id=lora_int_id, lora_path=lora_path)


def _req(adapter_id=None, prefix_id=None):
    return types.SimpleNamespace(adapter_id=adapter_id or "", prefix_id=prefix_id or "", model_id="m")


def test_validate_adapters_lora_is_loaded_once_and_cached_by_the_engine(tmp_path, monkeypatch):
    make_peft_dir(tmp_path, "my-lora")
    monkeypatch.chdir(tmp_path)
    eng, store = _Engine(), AdapterStore(cache_path=str(tmp_path), adapters={})

    async def go():
        a1 = await validate_adapters(_req("my-lora"), store, eng)
        a2 = await validate_adapters(_req(prefix_id="my-lora"), store, eng)   # deprecated alias
        return a1, a2

    a1, a2 = asyncio.run(go())
    assert a1["lora_request"].lora_name == "my-lora" and a1["lora_request"].lora_path == str(tmp_path / "my-lora")
    assert eng.load_calls == [("my-lora", str(tmp_path / "my-lora"))]            # one load (reference: awaited once)
    assert len(store.adapters) == 0                                               # metadata isn't cached locally
    assert a1["lora_request"].lora_int_id == a2["lora_request"].lora_int_id == 1000001
    assert asyncio.run(validate_adapters(_req(), store, eng)) == {}
    assert asyncio.run(validate_adapters(_req(), None, eng)) == {}


def test_validate_adapters_error_strings(tmp_path, monkeypatch):
    monkeypatch.chdir(tmp_path)
    eng, store = _Engine(), AdapterStore(cache_path=str(tmp_path), adapters={})
    with pytest.raises(ValueError, match="adapter_id supplied but no adapter store was configured"):
        asyncio.run(validate_adapters(_req("x"), None, eng))
    with pytest.raises(ValueError, match="can't retrieve adapter with id 'nope': directory does not exist"):
        asyncio.run(validate_adapters(_req("nope"), store, eng))
    (tmp_path / "empty").mkdir()
    with pytest.raises(ValueError, match="invalid adapter: no adapter_config.json found"):
        asyncio.run(validate_adapters(_req("empty"), store, eng))
    for bad in ("a b", "../up", "x;y"):
        with pytest.raises(ValueError, match="Invalid adapter id"):
            asyncio.run(validate_adapters(_req(bad), store, eng))
    (tmp_path / "pt").mkdir()
    (tmp_path / "pt" / "adapter_config.json").write_text(json.dumps({"peft_type": "PROMPT_TUNING"}))
    for _ in range(2):   # the second time the store's own cache answers
        with pytest.raises(ValueError, match="adapter type PROMPT_TUNING is not currently supported"):
            asyncio.run(validate_adapters(_req("pt"), store, eng))
    assert list(store.adapters) == ["pt"] and eng.load_calls == []


def test_lora_manager_slots_are_pinned_while_in_use_and_recycled_lru(tmp_path):
    class Native:
        def __init__(self):
            self.loaded = []

        def load_adapter(self, slot, weights):
            self.loaded.append((slot, len(weights)))

    native = Native()
    mgr = LoRAManager(native, n_layers=2, max_loras=2, max_rank=16)
    reqs = [LoRARequest(n, i, str(make_peft_dir(tmp_path, n, seed=i))) for i, n in enumerate("abc")]
    for r in reqs:
        mgr.register(r)
    sa = mgr.acquire(reqs[0])
    sb = mgr.acquire(reqs[1])
    assert {sa, sb} == {1, 2} and mgr.acquire(reqs[0]) == sa and mgr.loads == 2
    with pytest.raises(RuntimeError, match="in use"):
        mgr.acquire(reqs[2])
    mgr.release(sb)
    assert mgr.acquire(reqs[2]) == sb            # b's slot was idle: recycled
    mgr.release(sa); mgr.release(sa); mgr.release(sb)
    assert mgr.acquire(reqs[1]) == sa            # least recently used idle slot
    assert native.loaded == [(sa, 14), (sb, 14), (sb, 14), (sa, 14)]
    with pytest.raises(LoRAConfigError):
        mgr.register(LoRARequest("big", 9, str(make_peft_dir(tmp_path, "big", r=32))))


def test_adapter_id_over_grpc_reaches_the_engine_as_a_pinned_slot(tmp_path, monkeypatch):
    """Host path on the fake engine: adapter cache flag -> AdapterStore -> validate_adapters -> LoRAManager -> add_request
    with lora_slot; one weight load for two requests; the slot is released when the request ends."""
    import grpc
    from test_grpc_server_cpu import Server, _params
    from vllm_tgis_adapter_b200.grpc.pb import generation_pb2 as pb

    cache = tmp_path / "cache"
    cache.mkdir()
    make_peft_dir(cache, "my-lora", n_layers=1, hidden=128, q_dim=128, kv_dim=128, ffn=128)
    monkeypatch.chdir(tmp_path)
    srv = Server(adapter_cache=str(cache), max_loras=2)
    try:
        call = srv.channel.unary_unary("/fmaas.GenerationService/Generate",
                                       request_serializer=pb.BatchedGenerationRequest.SerializeToString,
                                       response_deserializer=pb.BatchedGenerationResponse.FromString)
        p = _params(stopping={"max_new_tokens": 3})
        for _ in range(2):
            r = call(pb.BatchedGenerationRequest(model_id="m", adapter_id="my-lora",
                                                 requests=[pb.GenerationRequest(text="t5 t6")], params=p), timeout=30)
            assert r.responses[0].generated_token_count == 3
        call(pb.BatchedGenerationRequest(model_id="m", requests=[pb.GenerationRequest(text="t5 t6")], params=p), timeout=30)
        assert srv.fake.lora_slots_seen == [1, 1, 0]
        assert srv.fake.adapter_loads == [(1, 7)]
        assert srv.engine._lora._users == {1: 0}
        with pytest.raises(grpc.RpcError) as ei:
            call(pb.BatchedGenerationRequest(model_id="m", adapter_id="nope",
                                             requests=[pb.GenerationRequest(text="t5")], params=p), timeout=30)
        assert ei.value.code() == grpc.StatusCode.INVALID_ARGUMENT
        assert ei.value.details() == "can't retrieve adapter with id 'nope': directory does not exist"
        # Tokenize validates the adapter id too (grpc_server.py:828)
        tok = srv.channel.unary_unary("/fmaas.GenerationService/Tokenize",
                                      request_serializer=pb.BatchedTokenizeRequest.SerializeToString,
                                      response_deserializer=pb.BatchedTokenizeResponse.FromString)
        assert tok(pb.BatchedTokenizeRequest(model_id="m", adapter_id="my-lora",
                                             requests=[pb.TokenizeRequest(text="t5 t6")]), timeout=30).responses[0].token_count == 2
    finally:
        srv.close()
