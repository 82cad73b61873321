"""`python -m vllm_tgis_adapter.healthcheck` / `grpc_healthcheck` under the reference's package name."""
from vllm_tgis_adapter_b200.healthcheck import cli, health_check, parse_args  # noqa: F401

if __name__ == "__main__":
    cli()
