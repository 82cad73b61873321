"""grpc.health.v1.Health (Check + Watch) without the grpcio-health-checking package (absent in this image).

The reference registers grpc_health's HealthServicer and flips `fmaas.GenerationService` to SERVING in post_init
(/root/reference/src/vllm_tgis_adapter/grpc/grpc_server.py:195-203,907-908); its tests and the k8s probe CLI
(healthcheck.py:1-96) poll `Check`.  The two messages are tiny, so they are declared at run time like generation_pb2."""
from __future__ import annotations

import asyncio

import grpc
from google.protobuf import descriptor_pb2, descriptor_pool, message_factory

SERVICE_NAME = "grpc.health.v1.Health"
UNKNOWN, SERVING, NOT_SERVING, SERVICE_UNKNOWN = 0, 1, 2, 3


def _build():
    fd = descriptor_pb2.FileDescriptorProto()
    fd.name = "tgis_b200/health.proto"
    fd.package = "grpc.health.v1"
    fd.syntax = "proto3"
    req = fd.message_type.add()
    req.name = "HealthCheckRequest"
    f = req.field.add()
    f.name, f.number, f.type, f.label = "service", 1, f.TYPE_STRING, f.LABEL_OPTIONAL
    resp = fd.message_type.add()
    resp.name = "HealthCheckResponse"
    e = resp.enum_type.add()
    e.name = "ServingStatus"
    for n, v in (("UNKNOWN", 0), ("SERVING", 1), ("NOT_SERVING", 2), ("SERVICE_UNKNOWN", 3)):
        ev = e.value.add()
        ev.name, ev.number = n, v
    f = resp.field.add()
    f.name, f.number, f.type, f.label = "status", 1, f.TYPE_ENUM, f.LABEL_OPTIONAL
    f.type_name = ".grpc.health.v1.HealthCheckResponse.ServingStatus"
    svc = fd.service.add()
    svc.name = "Health"
    for name, streaming in (("Check", False), ("Watch", True)):
        m = svc.method.add()
        m.name, m.server_streaming = name, streaming
        m.input_type, m.output_type = ".grpc.health.v1.HealthCheckRequest", ".grpc.health.v1.HealthCheckResponse"
    pool = descriptor_pool.DescriptorPool()
    pool.Add(fd)
    return (message_factory.GetMessageClass(pool.FindMessageTypeByName("grpc.health.v1.HealthCheckRequest")),
            message_factory.GetMessageClass(pool.FindMessageTypeByName("grpc.health.v1.HealthCheckResponse")),
            fd.SerializeToString())


HealthCheckRequest, HealthCheckResponse, FILE_DESCRIPTOR_SERIALIZED = _build()


class HealthServicer:
    def __init__(self) -> None:
        self._status: dict[str, int] = {"": SERVING}
        self._changed = asyncio.Event()

    def set(self, service: str, status: int) -> None:
        self._status[service] = status
        self._changed.set()

    async def Check(self, request, context):  # noqa: N802
        status = self._status.get(request.service)
        if status is None:
            await context.abort(grpc.StatusCode.NOT_FOUND, "unknown service")
        return HealthCheckResponse(status=status)

    async def Watch(self, request, context):  # noqa: N802, ARG002
        last = None
        while True:
            cur = self._status.get(request.service, SERVICE_UNKNOWN)
            if cur != last:
                last = cur
                yield HealthCheckResponse(status=cur)
            self._changed.clear()
            await self._changed.wait()


def add_health_servicer(servicer: HealthServicer, server) -> None:
    handlers = {
        "Check": grpc.unary_unary_rpc_method_handler(
            servicer.Check, request_deserializer=HealthCheckRequest.FromString,
            response_serializer=HealthCheckResponse.SerializeToString),
        "Watch": grpc.unary_stream_rpc_method_handler(
            servicer.Watch, request_deserializer=HealthCheckRequest.FromString,
            response_serializer=HealthCheckResponse.SerializeToString),
    }
    server.add_generic_rpc_handlers((grpc.method_handlers_generic_handler(SERVICE_NAME, handlers),))
