"""TGIS-style request / response log lines.

Same two lines, same field layout as /root/reference/src/vllm_tgis_adapter/tgis_utils/logs.py:150-226
("Processing request: {...}" / "Finished processing request: {...}. Timing info: {...}. Generated N tokens ..."),
same correlation-id blackboard (:29-45).  The reference monkey-patches engine.generate (:48-114); here the wrapper is an
explicit async generator the servicer calls, and the timing fields come from the engine's CLOCK_MONOTONIC stamps."""
from __future__ import annotations

import asyncio
import logging
import time
from collections import OrderedDict
from contextlib import suppress

logger = logging.getLogger("vllm_tgis_adapter.tgis_utils.logs")

_MAX, _TTL = 2048, 600.0
_REQUEST_ID_TO_CORRELATION_ID: OrderedDict[str, tuple[float, str]] = OrderedDict()   # tiny TTL cache (:29)


def set_correlation_id(request_id: str, correlation_id: str | None) -> None:
    if correlation_id is None:
        return
    now = time.monotonic()
    _REQUEST_ID_TO_CORRELATION_ID[request_id] = (now, correlation_id)
    while len(_REQUEST_ID_TO_CORRELATION_ID) > _MAX:
        _REQUEST_ID_TO_CORRELATION_ID.popitem(last=False)
    for k in [k for k, (t, _) in _REQUEST_ID_TO_CORRELATION_ID.items() if now - t > _TTL]:
        _REQUEST_ID_TO_CORRELATION_ID.pop(k, None)


def get_correlation_id(request_id: str) -> str | None:
    hit = _REQUEST_ID_TO_CORRELATION_ID.get(request_id)
    if not hit:
        hit = _REQUEST_ID_TO_CORRELATION_ID.get("-".join(request_id.split("-")[1:-1]))   # :38-45
    return hit[1] if hit else None


def _safe_div(a: float, b: float, default: float = 0.0) -> float:
    return a / b if b else default


def _log_request(request_id, params, adapter_id, correlation_id, n_input_tokens) -> None:
    logger.info("Processing request: {request_id=%s, correlation_id=%s, adapter_id=%s, input_tokens=%d, params=%s}",
                request_id, correlation_id, adapter_id, n_input_tokens, params)


def _log_response(request_id, correlation_id, response, start_mono: float) -> None:
    if not response.outputs:
        return
    out = response.outputs[0]
    generated = len(out.token_ids)
    m = response.metrics
    if m is None or m.first_scheduled_time is None:
        logger.warning("No engine metrics for request, cannot log timing info")
        inference = queue = per_token = total = 0.0
    else:
        inference = m.last_token_time - m.first_scheduled_time
        queue = m.time_in_queue or 0.0
        per_token = _safe_div(inference, generated)
        total = m.last_token_time - start_mono
    level = logging.WARNING if out.finish_reason == "abort" else logging.INFO
    logger.log(level,
               "Finished processing request: {request_id=%s, correlation_id=%s}. "
               "Timing info: {queue_time=%.2fms, inference_time=%.2fms, time_per_token=%.2fms, total_time=%.2fms}. "
               "Generated %d tokens before finish reason: %s, output %d chars",
               request_id, correlation_id, queue * 1e3, inference * 1e3, per_token * 1e3, total * 1e3, generated,
               out.finish_reason, len(out.text))


async def logged_generate(make_generator, *, prompt, prompt_token_ids, sampling_params, request_id, **kwargs):
    """engine.generate(...) with the request/response/error/cancel log lines around it (logs.py:55-114)."""
    start = time.monotonic()
    correlation_id = get_correlation_id(request_id)
    with suppress(BaseException):
        _log_request(request_id, sampling_params, None, correlation_id, len(prompt_token_ids))
    last = None
    total_tokens, text = [], ""
    try:
        async for response in make_generator(prompt=prompt, prompt_token_ids=prompt_token_ids,
                                             sampling_params=sampling_params, request_id=request_id, **kwargs):
            last = response
            yield response
    except asyncio.CancelledError:
        logger.info("Request cancelled: request_id=%s correlation_id=%s", request_id, correlation_id)
        raise
    except BaseException as e:
        logger.error("Request failed: request_id=%s correlation_id=%s error=%s", request_id, correlation_id, e)
        raise
    if last is not None:
        with suppress(BaseException):
            _log_response(request_id, correlation_id, last, start)
    del total_tokens, text
