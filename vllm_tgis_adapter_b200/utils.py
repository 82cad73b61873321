"""Termination log + task helpers (reference: /root/reference/src/vllm_tgis_adapter/utils.py:1-44)."""
from __future__ import annotations

import asyncio
import logging
import os
from collections.abc import Iterable

logger = logging.getLogger("vllm_tgis_adapter")


def check_for_failed_tasks(tasks: Iterable[asyncio.Task]) -> asyncio.Task | None:
    for task in tasks:
        try:
            if task.exception():
                return task
        except (asyncio.InvalidStateError, asyncio.CancelledError):
            pass
    return None


def write_termination_log(msg: str, file: str = "/dev/termination-log") -> None:
    """k8s termination message: only written when the file already exists (utils.py:20-40)."""
    if not os.path.exists(file):  # noqa: PTH110
        logger.debug("Not writing to termination log %s since it does not exist", file)
        return
    try:
        with open(file, "w") as f:
            f.write(f"{msg}\n")
    except Exception:  # noqa: BLE001
        logger.exception("Unable to write termination logs to %s", file)
