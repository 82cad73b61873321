"""`python -m vllm_tgis_adapter` entrypoint (shim in /root/repo/vllm_tgis_adapter/__main__.py).

Same supervision structure as /root/reference/src/vllm_tgis_adapter/__main__.py:38-131: build the engine, run the
HTTP and gRPC servers as tasks, exit as soon as either stops or the engine dies (:66-97), write the termination log
(:100-111)."""
from __future__ import annotations

import asyncio
import logging
import os
import traceback

from .engine.loader import build_engine
from .grpc.grpc_server import run_grpc_server
from .http import run_http_server
from .tgis_utils.args import parse_args
from .utils import check_for_failed_tasks, write_termination_log

logger = logging.getLogger("vllm_tgis_adapter")


async def start_servers(args) -> None:
    loop = asyncio.get_running_loop()
    engine = build_engine(args)
    engine.start(loop)
    tasks: list[asyncio.Task] = []
    try:
        tasks.append(loop.create_task(run_http_server(args, engine), name="http_server"))
        tasks.append(loop.create_task(run_grpc_server(args, engine), name="grpc_server"))
        runtime_error = None
        await asyncio.wait(tasks, return_when=asyncio.FIRST_COMPLETED)
        if engine.errored and not engine.is_running:                      # __main__.py:71-79
            runtime_error = RuntimeError("engine failed: " + str(engine.dead_error))
        failed_task = check_for_failed_tasks(tasks)
        for task in tasks:
            task.cancel()
        await asyncio.gather(*tasks, return_exceptions=True)
        if failed_task is not None:
            name, exc = failed_task.get_name(), failed_task.exception()
            raise RuntimeError(f"Failed task={name} ({exc})") from exc
        if runtime_error:
            raise runtime_error
    finally:
        engine.shutdown()


def run_and_catch_termination_cause(loop: asyncio.AbstractEventLoop, task: asyncio.Task) -> None:
    try:
        loop.run_until_complete(task)
    except Exception:
        write_termination_log(traceback.format_exc(), os.getenv("TERMINATION_LOG_DIR", "/dev/termination-log"))
        raise


def main(argv: list[str] | None = None) -> None:
    logging.basicConfig(level=os.getenv("VLLM_LOGGING_LEVEL", "INFO"),
                        format="%(levelname)s %(asctime)s %(name)s] %(message)s")
    args = parse_args(argv)
    try:
        import uvloop

        asyncio.set_event_loop_policy(uvloop.EventLoopPolicy())
    except ImportError:
        pass
    loop = asyncio.new_event_loop()
    asyncio.set_event_loop(loop)
    task = loop.create_task(start_servers(args))
    run_and_catch_termination_cause(loop, task)


if __name__ == "__main__":
    main()
