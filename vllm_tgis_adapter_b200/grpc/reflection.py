"""gRPC server reflection (grpc.reflection.v1alpha and .v1 ServerReflection) without the grpcio-reflection package
(absent in this image).

The reference enables it for `health`, `fmaas.GenerationService` and the reflection service itself
(/root/reference/src/vllm_tgis_adapter/grpc/grpc_server.py:919-926) so that `grpcurl` (examples/inference.sh) can
discover the API.  The protocol is one bidi stream of small oneof messages; both package versions share the wire
layout, so one run-time-built descriptor per package serves both.  What is answered: list_services,
file_containing_symbol, file_by_filename; extension queries return NOT_FOUND / empty (proto3 files, no extensions)."""
from __future__ import annotations

import grpc
from google.protobuf import descriptor_pb2, descriptor_pool, message_factory

_STRING, _INT32, _BYTES, _MESSAGE = (descriptor_pb2.FieldDescriptorProto.TYPE_STRING,
                                     descriptor_pb2.FieldDescriptorProto.TYPE_INT32,
                                     descriptor_pb2.FieldDescriptorProto.TYPE_BYTES,
                                     descriptor_pb2.FieldDescriptorProto.TYPE_MESSAGE)
_OPT, _REP = descriptor_pb2.FieldDescriptorProto.LABEL_OPTIONAL, descriptor_pb2.FieldDescriptorProto.LABEL_REPEATED


def _file(package: str) -> descriptor_pb2.FileDescriptorProto:
    fd = descriptor_pb2.FileDescriptorProto()
    fd.name = f"tgis_b200/{package.replace('.', '_')}.proto"
    fd.package = package
    fd.syntax = "proto3"

    def msg(name: str, fields, oneof: str | None = None, oneof_from: int = 0):
        m = fd.message_type.add()
        m.name = name
        if oneof:
            m.oneof_decl.add().name = oneof
        for i, (fname, number, ftype, label, tname) in enumerate(fields):
            f = m.field.add()
            f.name, f.number, f.type, f.label = fname, number, ftype, label
            if tname:
                f.type_name = f".{package}.{tname}"
            if oneof and i >= oneof_from:
                f.oneof_index = 0
        return m

    msg("ExtensionRequest", [("containing_type", 1, _STRING, _OPT, ""), ("extension_number", 2, _INT32, _OPT, "")])
    msg("ServerReflectionRequest",
        [("host", 1, _STRING, _OPT, ""), ("file_by_filename", 3, _STRING, _OPT, ""),
         ("file_containing_symbol", 4, _STRING, _OPT, ""),
         ("file_containing_extension", 5, _MESSAGE, _OPT, "ExtensionRequest"),
         ("all_extension_numbers_of_type", 6, _STRING, _OPT, ""), ("list_services", 7, _STRING, _OPT, "")],
        oneof="message_request", oneof_from=1)
    msg("FileDescriptorResponse", [("file_descriptor_proto", 1, _BYTES, _REP, "")])
    msg("ExtensionNumberResponse", [("base_type_name", 1, _STRING, _OPT, ""), ("extension_number", 2, _INT32, _REP, "")])
    msg("ServiceResponse", [("name", 1, _STRING, _OPT, "")])
    msg("ListServiceResponse", [("service", 1, _MESSAGE, _REP, "ServiceResponse")])
    msg("ErrorResponse", [("error_code", 1, _INT32, _OPT, ""), ("error_message", 2, _STRING, _OPT, "")])
    msg("ServerReflectionResponse",
        [("valid_host", 1, _STRING, _OPT, ""), ("original_request", 2, _MESSAGE, _OPT, "ServerReflectionRequest"),
         ("file_descriptor_response", 4, _MESSAGE, _OPT, "FileDescriptorResponse"),
         ("all_extension_numbers_response", 5, _MESSAGE, _OPT, "ExtensionNumberResponse"),
         ("list_services_response", 6, _MESSAGE, _OPT, "ListServiceResponse"),
         ("error_response", 7, _MESSAGE, _OPT, "ErrorResponse")],
        oneof="message_response", oneof_from=2)
    svc = fd.service.add()
    svc.name = "ServerReflection"
    m = svc.method.add()
    m.name = "ServerReflectionInfo"
    m.input_type, m.output_type = f".{package}.ServerReflectionRequest", f".{package}.ServerReflectionResponse"
    m.client_streaming = m.server_streaming = True
    return fd


class _Package:
    def __init__(self, package: str):
        self.package = package
        self.service_name = f"{package}.ServerReflection"
        self.file_proto = _file(package)
        pool = descriptor_pool.DescriptorPool()
        pool.Add(self.file_proto)
        self.Request = message_factory.GetMessageClass(pool.FindMessageTypeByName(f"{package}.ServerReflectionRequest"))
        self.Response = message_factory.GetMessageClass(pool.FindMessageTypeByName(f"{package}.ServerReflectionResponse"))


V1ALPHA = _Package("grpc.reflection.v1alpha")
V1 = _Package("grpc.reflection.v1")
SERVICE_NAME = V1ALPHA.service_name   # what the reference lists (grpc_reflection.v1alpha.reflection.SERVICE_NAME)


class ReflectionServicer:
    """files: serialized FileDescriptorProto blobs of everything this server exposes."""

    def __init__(self, service_names: tuple[str, ...], files: list[bytes]):
        self._services = tuple(service_names)
        self._by_name: dict[str, bytes] = {}
        self._by_symbol: dict[str, bytes] = {}
        for blob in files:
            fd = descriptor_pb2.FileDescriptorProto.FromString(blob)
            self._by_name[fd.name] = blob
            prefix = fd.package + "." if fd.package else ""

            def walk(msgs, scope: str, blob=blob) -> None:
                for m in msgs:
                    full = scope + m.name
                    self._by_symbol[full] = blob
                    for f in m.field:
                        self._by_symbol[f"{full}.{f.name}"] = blob
                    for e in m.enum_type:
                        self._by_symbol[f"{full}.{e.name}"] = blob
                    walk(m.nested_type, full + ".")

            walk(fd.message_type, prefix)
            for e in fd.enum_type:
                self._by_symbol[prefix + e.name] = blob
            for s in fd.service:
                self._by_symbol[prefix + s.name] = blob
                for meth in s.method:
                    self._by_symbol[f"{prefix}{s.name}.{meth.name}"] = blob

    def _answer(self, pkg: _Package, req):
        resp = pkg.Response(valid_host=req.host, original_request=req)
        kind = req.WhichOneof("message_request")
        if kind == "list_services":
            for name in self._services:
                resp.list_services_response.service.add().name = name
        elif kind in ("file_containing_symbol", "file_by_filename"):
            key = getattr(req, kind)
            blob = (self._by_symbol if kind == "file_containing_symbol" else self._by_name).get(key)
            if blob is None:
                resp.error_response.error_code = grpc.StatusCode.NOT_FOUND.value[0]
                resp.error_response.error_message = f"{key} not found"
            else:
                resp.file_descriptor_response.file_descriptor_proto.append(blob)
        elif kind == "all_extension_numbers_of_type":
            if req.all_extension_numbers_of_type in self._by_symbol:
                resp.all_extension_numbers_response.base_type_name = req.all_extension_numbers_of_type
            else:
                resp.error_response.error_code = grpc.StatusCode.NOT_FOUND.value[0]
                resp.error_response.error_message = "type not found"
        else:   # file_containing_extension / nothing set
            resp.error_response.error_code = grpc.StatusCode.NOT_FOUND.value[0]
            resp.error_response.error_message = "no extensions are registered"
        return resp

    def handler(self, pkg: _Package):
        async def ServerReflectionInfo(request_iterator, context):  # noqa: N802, ARG001
            async for req in request_iterator:
                yield self._answer(pkg, req)

        return grpc.stream_stream_rpc_method_handler(ServerReflectionInfo, request_deserializer=pkg.Request.FromString,
                                                     response_serializer=pkg.Response.SerializeToString)


def enable_server_reflection(service_names: tuple[str, ...], server, files: list[bytes]) -> ReflectionServicer:
    """Same call shape as grpc_reflection.v1alpha.reflection.enable_server_reflection, plus the descriptor blobs."""
    names = tuple(dict.fromkeys((*service_names, V1ALPHA.service_name, V1.service_name)))
    servicer = ReflectionServicer(names, [*files, V1ALPHA.file_proto.SerializeToString(), V1.file_proto.SerializeToString()])
    for pkg in (V1ALPHA, V1):
        server.add_generic_rpc_handlers((grpc.method_handlers_generic_handler(
            pkg.service_name, {"ServerReflectionInfo": servicer.handler(pkg)}),))
    return servicer
