"""adapter_id -> `lora_request` kwarg of engine.generate.

Behavioural mirror of /root/reference/src/vllm_tgis_adapter/grpc/adapters.py:63-223 (same entry point name, same error
strings through TGISValidationError, same caching rule: LoRA adapters are remembered by the ENGINE's registry, other
adapter types by the store).  The reference asks vLLM's `OpenAIServingModels.load_lora_adapter`; here the engine façade
(`AsyncTGISEngine.load_lora_adapter` / `.lora_requests`) plays that part."""
from __future__ import annotations

import asyncio
import concurrent.futures
import dataclasses
import json
import re
from pathlib import Path

from ..engine.types import LoRARequest
from .validation import TGISValidationError

_ID_OK = re.compile(r"[/\w\-]+")
_pool: concurrent.futures.ThreadPoolExecutor | None = None


@dataclasses.dataclass
class AdapterMetadata:
    unique_id: int
    adapter_type: str | None
    full_path: str
    full_config: dict


@dataclasses.dataclass
class AdapterStore:
    cache_path: str
    adapters: dict[str, AdapterMetadata]
    next_unique_id: int = 1000001
    load_locks: dict[str, asyncio.Lock] = dataclasses.field(default_factory=dict)


def _check_adapter_id(adapter_id: str) -> None:
    """:210-223 -- characters, and no escape from the working directory."""
    if not _ID_OK.fullmatch(adapter_id) or not Path(adapter_id).resolve().is_relative_to(Path().cwd()):
        TGISValidationError.InvalidAdapterID.error(adapter_id)


def _read_metadata(adapter_id: str, adapter_path: str, unique_id: int) -> AdapterMetadata:
    """:175-207 (runs on a worker thread: file access must not block the event loop)."""
    root = Path(adapter_path)
    if not root.exists():
        TGISValidationError.AdapterNotFound.error(adapter_id, "directory does not exist")
    cfg_file = root / "adapter_config.json"
    if not cfg_file.exists():
        TGISValidationError.AdapterNotFound.error(adapter_id, "invalid adapter: no adapter_config.json found")
    config = json.loads(cfg_file.read_text())
    return AdapterMetadata(unique_id=unique_id, adapter_type=config.get("peft_type"), full_path=adapter_path,
                           full_config=config)


async def validate_adapters(request, adapter_store: AdapterStore | None, engine) -> dict[str, LoRARequest]:
    """:63-137.  Returns the kwargs to add to engine.generate()."""
    global _pool  # noqa: PLW0603
    adapter_id = request.adapter_id or request.prefix_id   # prefix_id: backwards compatibility
    if adapter_id and not adapter_store:
        TGISValidationError.AdaptersDisabled.error()
    if not adapter_id or not adapter_store:
        return {}
    async with adapter_store.load_locks.setdefault(adapter_id, asyncio.Lock()):
        known = engine.lora_requests.get(adapter_id)
        if known is not None:
            return {"lora_request": known}
        meta = adapter_store.adapters.get(adapter_id)
        if meta is None:
            _check_adapter_id(adapter_id)
            path = str(Path(adapter_store.cache_path) / adapter_id)
            if _pool is None:
                _pool = concurrent.futures.ThreadPoolExecutor(max_workers=2)
            unique_id = adapter_store.next_unique_id
            adapter_store.next_unique_id += 1
            meta = await asyncio.get_running_loop().run_in_executor(_pool, _read_metadata, adapter_id, path, unique_id)
            if meta.adapter_type == "LORA":
                # a load failure is a ValueError with the loader's message (:150-151)
                await engine.load_lora_adapter(lora_name=adapter_id, lora_path=meta.full_path, lora_int_id=unique_id)
                loaded = engine.lora_requests.get(adapter_id)
                if loaded is None:
                    raise RuntimeError("engine failed to load LoRA adapter")
                return {"lora_request": loaded}
            adapter_store.adapters[adapter_id] = meta
    TGISValidationError.AdapterUnsupported.error(meta.adapter_type)   # prompt tuning etc.
    return {}
