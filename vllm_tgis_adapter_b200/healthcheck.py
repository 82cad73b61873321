"""k8s probe CLI: one unary `grpc.health.v1.Health/Check` against a running server, exit status 0 = SERVING.

Mirrors the reference's probe (/root/reference/src/vllm_tgis_adapter/healthcheck.py:1-96, console script
`grpc_healthcheck`, pyproject.toml:46): same flags (`--insecure` | `--secure`, `--server-url`, `--timeout`,
`--service-name`), same defaults (localhost:8033, 1 s, `fmaas.GenerationService`), same two output shapes
(`health check...status: SERVING` / `health check...Health.Check failed: code=..., details=...`) and exit codes, so a
deployment's readiness / liveness probe command works unchanged.  The reference uses grpc_health's generated stubs
(`grpcio-health-checking`, absent in this image); here the two messages come from grpc/health.py's run-time descriptors
and the call is a plain channel method."""
from __future__ import annotations

import argparse
import sys

import grpc

from .grpc.health import SERVING, HealthCheckRequest, HealthCheckResponse

DEFAULT_SERVICE = "fmaas.GenerationService"   # grpc_server.TextGenerationService.SERVICE_NAME (not imported: start-up cost)


def health_check(*, server_url: str = "localhost:8033", service: str | None = None, insecure: bool = True,
                 timeout: float = 1) -> bool:
    print("health check...", end="")
    channel = (grpc.insecure_channel(server_url) if insecure
               else grpc.secure_channel(server_url, grpc.ssl_channel_credentials()))
    try:
        check = channel.unary_unary("/grpc.health.v1.Health/Check",
                                    request_serializer=HealthCheckRequest.SerializeToString,
                                    response_deserializer=HealthCheckResponse.FromString)
        response = check(HealthCheckRequest(service=service or ""), timeout=timeout)
    except grpc.RpcError as e:
        print(f"Health.Check failed: code={e.code()}, details={e.details()}")
        return False
    finally:
        channel.close()
    print(str(response).strip())
    return response.status == SERVING


def parse_args(argv: list[str] | None = None) -> argparse.Namespace:
    p = argparse.ArgumentParser(formatter_class=argparse.ArgumentDefaultsHelpFormatter)
    mode = p.add_mutually_exclusive_group(required=False)
    mode.add_argument("--insecure", dest="insecure", action="store_true", help="Use an insecure connection")
    mode.add_argument("--secure", dest="insecure", action="store_false", help="Use a secure connection")
    p.set_defaults(insecure=True)
    p.add_argument("--server-url", type=str, default="localhost:8033", help="grpc server url (`host:port`)")
    p.add_argument("--timeout", type=float, default=1, help="Timeout for healthcheck request")
    p.add_argument("--service-name", type=str, default=DEFAULT_SERVICE, help="Name of the service to check")
    return p.parse_args(argv)


def cli(argv: list[str] | None = None) -> None:
    args = parse_args(argv)
    if not health_check(server_url=args.server_url, service=args.service_name, insecure=args.insecure,
                        timeout=args.timeout):
        sys.exit(1)


if __name__ == "__main__":
    cli()
