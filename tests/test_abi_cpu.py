"""The C-ABI library loads on a GPU-less box and exports every symbol include/*.h declare (no compute calls)."""
import ctypes as C
import re
from pathlib import Path

from vllm_tgis_adapter_b200.engine import _lib

ROOT = Path(__file__).resolve().parent.parent


def _declared(header: str) -> set[str]:
    text = (ROOT / "include" / header).read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return set(re.findall(r"\b(tgis_[a-z0-9_]+)\s*\(", text))


def test_library_exports_every_declared_symbol():
    lib = _lib.load_library()
    declared = _declared("tgis_engine.h") | _declared("tgis_kernels.h")
    assert declared == set(_lib.ENGINE_SYMBOLS) | set(_lib.KERNEL_SYMBOLS)
    for sym in declared:
        assert hasattr(lib, sym), sym
    assert lib.tgis_abi_version() == _lib.ABI_VERSION


def test_struct_layouts_match_the_header_sizes():
    lib = _lib.load_library()
    assert lib.tgis_k_sizeof_sample_row() == 64
    assert lib.tgis_k_sizeof_sample_out() == 112
    assert lib.tgis_k_kv_block() == 32
    assert C.sizeof(_lib.TgisSamplingParams) == 112
    # id, 5 ints, 2 arrays, 4 ints, pad, 4 doubles, prompt_pos + reserved
    assert C.sizeof(_lib.TgisStepOutput) == 96 + 4 * 5 + 48 + 48 + 4 * 4 + 4 + 32 + 8


def test_no_gpu_means_loud_failure_not_fallback():
    """Product path has no CPU fallback: creating an engine without a device must fail with a message."""
    import torch

    if torch.cuda.is_available():
        return
    lib = _lib.load_library()
    cfg = _lib.TgisConfig()
    cfg.abi_version = _lib.ABI_VERSION
    cfg.n_layers, cfg.hidden, cfg.n_q_heads, cfg.n_kv_heads, cfg.head_dim = 1, 128, 1, 1, 128
    cfg.ffn, cfg.vocab, cfg.max_model_len, cfg.max_num_seqs, cfg.max_batched_tokens = 128, 64, 64, 1, 64
    cfg.tp_size = 1
    h = C.c_void_p()
    assert lib.tgis_engine_create(C.byref(cfg), C.byref(h)) != 0
    assert b"no CUDA device" in lib.tgis_last_error() or b"CUDA" in lib.tgis_last_error()
