"""Flag system: env-var fallback matrix (model: /root/reference/tests/test_tgis_utils.py:13-168) + TGIS legacy flags."""
import argparse

import pytest

from vllm_tgis_adapter_b200.tgis_utils.args import (EnvVarArgumentParser, StoreBoolean, add_tgis_args,
                                                    make_engine_arg_parser, parse_args, postprocess_tgis_args)

TRUE = ["true", "True", "TRUE", "1"]
FALSE = ["false", "False", "FALSE", "0"]


def _parser():
    base = argparse.ArgumentParser(add_help=False)
    base.add_argument("--string-arg", type=str, default="x")
    base.add_argument("--int-arg", type=int, default=1)
    base.add_argument("--bool-arg", type=bool, default=False)
    base.add_argument("--flag-on", action="store_true")
    base.add_argument("--flag-off", action="store_false")
    base.add_argument("--store-boolean", action=StoreBoolean, default=False)
    return base


def test_str_and_int_env_fallback(monkeypatch):
    monkeypatch.setenv("STRING_ARG", "from-env")
    monkeypatch.setenv("INT_ARG", "42")
    args = EnvVarArgumentParser(parser=_parser()).parse_args([])
    assert args.string_arg == "from-env" and args.int_arg == 42
    args = EnvVarArgumentParser(parser=_parser()).parse_args(["--string-arg", "cli", "--int-arg", "7"])
    assert args.string_arg == "cli" and args.int_arg == 7


@pytest.mark.parametrize("val", TRUE + FALSE)
@pytest.mark.parametrize("name", ["BOOL_ARG", "FLAG_ON", "FLAG_OFF", "STORE_BOOLEAN"])
def test_bool_env_fallback(monkeypatch, name, val):
    monkeypatch.setenv(name, val)
    args = EnvVarArgumentParser(parser=_parser()).parse_args([])
    assert getattr(args, name.lower()) is (val in TRUE)


def test_args_added_after_construction_also_read_env(monkeypatch):
    monkeypatch.setenv("GRPC_PORT", "9999")
    monkeypatch.setenv("MAX_NEW_TOKENS", "77")
    parser = add_tgis_args(EnvVarArgumentParser(parser=make_engine_arg_parser()))
    args = parser.parse_args([])
    assert args.grpc_port == 9999 and args.max_new_tokens == 77


def test_tgis_legacy_translation_and_defaults():
    args = parse_args(["--model-name", "m", "--max-sequence-length", "512", "--num-gpus", "1", "--dtype-str", "bfloat16",
                       "--tls-cert-path", "c", "--tls-key-path", "k", "--tls-client-ca-cert-path", "ca"])
    assert args.model == "m" and args.max_model_len == 512 and args.tensor_parallel_size == 1
    assert args.dtype == "bfloat16" and (args.ssl_certfile, args.ssl_keyfile, args.ssl_ca_certs) == ("c", "k", "ca")
    assert args.max_logprobs >= 11            # reference args.py:214-216
    assert args.grpc_port == 8033 and args.max_new_tokens == 1024
    assert args.default_include_stop_seqs is True and args.output_special_tokens is False


def test_inconsistent_legacy_flags_raise():
    with pytest.raises(ValueError, match="Inconsistent max_model_len"):
        parse_args(["--max-sequence-length", "512", "--max-model-len", "256"])
    with pytest.raises(ValueError, match="Inconsistent num_gpus"):
        parse_args(["--num-gpus", "2", "--num-shard", "4"])


def test_tensor_parallel_flags_reach_the_engine_builder():
    """--num-gpus / --num-shard are the TGIS spellings of --tensor-parallel-size (reference tgis_utils/args.py:139-148);
    the engine builder spawns one worker per extra GPU and refuses, before touching a device, a size that does not
    divide the model's kv heads / ffn / vocab."""
    import pytest

    from vllm_tgis_adapter_b200.engine.loader import build_engine
    from vllm_tgis_adapter_b200.tgis_utils.args import parse_args

    args = parse_args(["--model", "tiny", "--synthetic-weights", "--num-gpus", "4"])
    assert args.tensor_parallel_size == 4
    with pytest.raises(ValueError, match="does not divide"):      # tiny has 2 kv heads
        build_engine(args)
    args = parse_args(["--model", "no/such/dir", "--tensor-parallel-size", "2"])
    with pytest.raises(ValueError, match="does not exist"):
        build_engine(args)


def test_startup_failure_is_written_to_the_termination_log(tmp_path, monkeypatch):
    """Reference tests/test_termination_log.py: a set-up error (here: a model path that does not exist) must crash the
    entrypoint AND leave the traceback in the k8s termination log, which is only written when the file already exists."""
    import pytest

    from vllm_tgis_adapter_b200.__main__ import main
    from vllm_tgis_adapter_b200.utils import write_termination_log

    log = tmp_path / "termination_log.txt"
    log.touch()
    monkeypatch.setenv("TERMINATION_LOG_DIR", str(log))
    with pytest.raises(BaseException):  # noqa: B017,PT011  (ValueError from build_engine, re-raised by the entrypoint)
        main(["--model", str(tmp_path / "no-such-model"), "--grpc-port", "0", "--port", "0"])
    text = log.read_text()
    assert "does not exist" in text and "Traceback" in text
    missing = tmp_path / "absent.txt"
    write_termination_log("x", str(missing))          # no file -> nothing created (utils.py:20-40)
    assert not missing.exists()


def test_check_for_failed_tasks_picks_the_task_that_raised():
    import asyncio

    from vllm_tgis_adapter_b200.utils import check_for_failed_tasks

    async def run():
        async def ok():
            await asyncio.sleep(0)

        async def bad():
            raise RuntimeError("boom")

        async def forever():
            await asyncio.sleep(10)

        tasks = [asyncio.create_task(ok(), name="ok"), asyncio.create_task(bad(), name="bad"),
                 asyncio.create_task(forever(), name="pending")]
        await asyncio.wait(tasks[:2])
        failed = check_for_failed_tasks(tasks)
        tasks[2].cancel()
        return failed.get_name()

    assert asyncio.run(run()) == "bad"
