"""Wire protocol of `fmaas.GenerationService`, built at import time without protoc.

`grpc_tools.protoc` (which the reference runs in setup.py:22-36 to generate generation_pb2*.py) is not available in
this image, so the FileDescriptorProto is written out by hand, field for field, from
/root/reference/src/vllm_tgis_adapter/grpc/pb/generation.proto:1-279 (package, message names, field NUMBERS, proto3
`optional` presence, the `guided` oneof, enum values).  The resulting classes are byte-compatible with the
reference's generated module and expose the same names (`BatchedGenerationRequest`, `StopReason`, ...,
`DESCRIPTOR.services_by_name["GenerationService"]`).
"""
from __future__ import annotations

from google.protobuf import descriptor_pb2, descriptor_pool, message_factory
from google.protobuf.internal import enum_type_wrapper

_F = descriptor_pb2.FieldDescriptorProto
_T = {
    "string": _F.TYPE_STRING, "uint32": _F.TYPE_UINT32, "uint64": _F.TYPE_UINT64, "float": _F.TYPE_FLOAT,
    "bool": _F.TYPE_BOOL, "enum": _F.TYPE_ENUM, "message": _F.TYPE_MESSAGE,
}


def _field(msg, name, number, ftype, *, type_name=None, repeated=False, optional=False, oneof_index=None):
    f = msg.field.add()
    f.name, f.number, f.type = name, number, _T[ftype]
    f.label = _F.LABEL_REPEATED if repeated else _F.LABEL_OPTIONAL
    if type_name:
        f.type_name = type_name
    if optional:  # proto3 `optional`: synthetic oneof + proto3_optional flag
        f.proto3_optional = True
        od = msg.oneof_decl.add()
        od.name = "_" + name
        f.oneof_index = len(msg.oneof_decl) - 1
    elif oneof_index is not None:
        f.oneof_index = oneof_index
    return f


def _enum(container, name, values):
    e = container.enum_type.add()
    e.name = name
    for n, v in values:
        ev = e.value.add()
        ev.name, ev.number = n, v
    return e


def _build_file() -> descriptor_pb2.FileDescriptorProto:
    fd = descriptor_pb2.FileDescriptorProto()
    fd.name = "generation.proto"
    fd.package = "fmaas"
    fd.syntax = "proto3"

    _enum(fd, "DecodingMethod", [("GREEDY", 0), ("SAMPLE", 1)])                       # generation.proto:23-26
    _enum(fd, "StopReason", [("NOT_FINISHED", 0), ("MAX_TOKENS", 1), ("EOS_TOKEN", 2), ("CANCELLED", 3),
                             ("TIME_LIMIT", 4), ("STOP_SEQUENCE", 5), ("TOKEN_LIMIT", 6), ("ERROR", 7)])  # :185-202

    m = fd.message_type.add(); m.name = "BatchedGenerationRequest"                   # :28-36
    _field(m, "model_id", 1, "string")
    _field(m, "prefix_id", 2, "string", optional=True)
    _field(m, "adapter_id", 4, "string", optional=True)
    _field(m, "requests", 3, "message", type_name=".fmaas.GenerationRequest", repeated=True)
    _field(m, "params", 10, "message", type_name=".fmaas.Parameters")

    m = fd.message_type.add(); m.name = "SingleGenerationRequest"                    # :38-46
    _field(m, "model_id", 1, "string")
    _field(m, "prefix_id", 2, "string", optional=True)
    _field(m, "adapter_id", 4, "string", optional=True)
    _field(m, "request", 3, "message", type_name=".fmaas.GenerationRequest")
    _field(m, "params", 10, "message", type_name=".fmaas.Parameters")

    m = fd.message_type.add(); m.name = "BatchedGenerationResponse"                  # :48-50
    _field(m, "responses", 1, "message", type_name=".fmaas.GenerationResponse", repeated=True)

    m = fd.message_type.add(); m.name = "GenerationRequest"                          # :52-54
    _field(m, "text", 2, "string")

    m = fd.message_type.add(); m.name = "GenerationResponse"                         # :56-71
    _field(m, "input_token_count", 6, "uint32")
    _field(m, "generated_token_count", 2, "uint32")
    _field(m, "text", 4, "string")
    _field(m, "stop_reason", 7, "enum", type_name=".fmaas.StopReason")
    _field(m, "stop_sequence", 11, "string")
    _field(m, "seed", 10, "uint64")
    _field(m, "tokens", 8, "message", type_name=".fmaas.TokenInfo", repeated=True)
    _field(m, "input_tokens", 9, "message", type_name=".fmaas.TokenInfo", repeated=True)

    m = fd.message_type.add(); m.name = "Parameters"                                 # :73-90
    _field(m, "method", 1, "enum", type_name=".fmaas.DecodingMethod")
    _field(m, "sampling", 2, "message", type_name=".fmaas.SamplingParameters")
    _field(m, "stopping", 3, "message", type_name=".fmaas.StoppingCriteria")
    _field(m, "response", 4, "message", type_name=".fmaas.ResponseOptions")
    _field(m, "decoding", 5, "message", type_name=".fmaas.DecodingParameters")
    _field(m, "truncate_input_tokens", 6, "uint32")

    m = fd.message_type.add(); m.name = "DecodingParameters"                         # :92-131
    lp = m.nested_type.add(); lp.name = "LengthPenalty"
    _field(lp, "start_index", 1, "uint32")
    _field(lp, "decay_factor", 2, "float")
    _enum(m, "ResponseFormat", [("TEXT", 0), ("JSON", 1)])
    sc = m.nested_type.add(); sc.name = "StringChoices"
    _field(sc, "choices", 1, "string", repeated=True)
    # declaration order in the .proto: real oneof `guided` is declared before the synthetic oneof of the proto3
    # optional field only matters for oneof indices; real oneofs must come first
    od = m.oneof_decl.add(); od.name = "guided"
    _field(m, "repetition_penalty", 1, "float")
    _field(m, "format", 3, "enum", type_name=".fmaas.DecodingParameters.ResponseFormat", oneof_index=0)
    _field(m, "json_schema", 4, "string", oneof_index=0)
    _field(m, "regex", 5, "string", oneof_index=0)
    _field(m, "choice", 6, "message", type_name=".fmaas.DecodingParameters.StringChoices", oneof_index=0)
    _field(m, "grammar", 7, "string", oneof_index=0)
    _field(m, "length_penalty", 2, "message", type_name=".fmaas.DecodingParameters.LengthPenalty", optional=True)

    m = fd.message_type.add(); m.name = "SamplingParameters"                         # :134-146
    _field(m, "temperature", 1, "float", optional=True)
    _field(m, "top_k", 2, "uint32")
    _field(m, "top_p", 3, "float")
    _field(m, "typical_p", 4, "float")
    _field(m, "seed", 5, "uint64", optional=True)

    m = fd.message_type.add(); m.name = "StoppingCriteria"                           # :148-160
    _field(m, "max_new_tokens", 1, "uint32")
    _field(m, "min_new_tokens", 2, "uint32")
    _field(m, "time_limit_millis", 3, "uint32")
    _field(m, "stop_sequences", 4, "string", repeated=True)
    _field(m, "include_stop_sequence", 5, "bool", optional=True)

    m = fd.message_type.add(); m.name = "ResponseOptions"                            # :162-183
    _field(m, "input_text", 1, "bool")
    _field(m, "generated_tokens", 2, "bool")
    _field(m, "input_tokens", 3, "bool")
    _field(m, "token_logprobs", 4, "bool")
    _field(m, "token_ranks", 5, "bool")
    _field(m, "top_n_tokens", 6, "uint32")

    m = fd.message_type.add(); m.name = "TokenInfo"                                  # :204-221
    tt = m.nested_type.add(); tt.name = "TopToken"
    _field(tt, "text", 2, "string")
    _field(tt, "logprob", 3, "float")
    _field(m, "text", 2, "string")
    _field(m, "logprob", 3, "float")
    _field(m, "rank", 4, "uint32")
    _field(m, "top_tokens", 5, "message", type_name=".fmaas.TokenInfo.TopToken", repeated=True)

    m = fd.message_type.add(); m.name = "BatchedTokenizeRequest"                     # :227-238
    _field(m, "model_id", 1, "string")
    _field(m, "prefix_id", 6, "string", optional=True)
    _field(m, "adapter_id", 7, "string", optional=True)
    _field(m, "requests", 2, "message", type_name=".fmaas.TokenizeRequest", repeated=True)
    _field(m, "return_tokens", 3, "bool")
    _field(m, "return_offsets", 4, "bool")
    _field(m, "truncate_input_tokens", 5, "uint32")

    m = fd.message_type.add(); m.name = "BatchedTokenizeResponse"                    # :240-242
    _field(m, "responses", 1, "message", type_name=".fmaas.TokenizeResponse", repeated=True)

    m = fd.message_type.add(); m.name = "TokenizeRequest"                            # :244-246
    _field(m, "text", 1, "string")

    m = fd.message_type.add(); m.name = "TokenizeResponse"                           # :248-260
    off = m.nested_type.add(); off.name = "Offset"
    _field(off, "start", 1, "uint32")
    _field(off, "end", 2, "uint32")
    _field(m, "token_count", 1, "uint32")
    _field(m, "tokens", 2, "string", repeated=True)
    _field(m, "offsets", 3, "message", type_name=".fmaas.TokenizeResponse.Offset", repeated=True)

    m = fd.message_type.add(); m.name = "ModelInfoRequest"                           # :266-268
    _field(m, "model_id", 1, "string")

    m = fd.message_type.add(); m.name = "ModelInfoResponse"                          # :270-279
    _enum(m, "ModelKind", [("DECODER_ONLY", 0), ("ENCODER_DECODER", 1)])
    _field(m, "model_kind", 1, "enum", type_name=".fmaas.ModelInfoResponse.ModelKind")
    _field(m, "max_sequence_length", 2, "uint32")
    _field(m, "max_new_tokens", 3, "uint32")

    svc = fd.service.add(); svc.name = "GenerationService"                           # :9-18
    for name, inp, out, stream in [
        ("Generate", "BatchedGenerationRequest", "BatchedGenerationResponse", False),
        ("GenerateStream", "SingleGenerationRequest", "GenerationResponse", True),
        ("Tokenize", "BatchedTokenizeRequest", "BatchedTokenizeResponse", False),
        ("ModelInfo", "ModelInfoRequest", "ModelInfoResponse", False),
    ]:
        meth = svc.method.add()
        meth.name, meth.input_type, meth.output_type = name, ".fmaas." + inp, ".fmaas." + out
        meth.server_streaming = stream
    return fd


def _fix_optional_oneof_order(fd: descriptor_pb2.FileDescriptorProto) -> None:
    """protobuf requires synthetic (proto3_optional) oneofs to come AFTER all real oneofs; _field() appends them in
    field order, which already satisfies that because `guided` is declared first in DecodingParameters."""


_pool = descriptor_pool.DescriptorPool()
_FILE_PROTO = _build_file()
DESCRIPTOR = _pool.Add(_FILE_PROTO)


def _cls(name: str):
    return message_factory.GetMessageClass(_pool.FindMessageTypeByName("fmaas." + name))


BatchedGenerationRequest = _cls("BatchedGenerationRequest")
SingleGenerationRequest = _cls("SingleGenerationRequest")
BatchedGenerationResponse = _cls("BatchedGenerationResponse")
GenerationRequest = _cls("GenerationRequest")
GenerationResponse = _cls("GenerationResponse")
Parameters = _cls("Parameters")
DecodingParameters = _cls("DecodingParameters")
SamplingParameters = _cls("SamplingParameters")
StoppingCriteria = _cls("StoppingCriteria")
ResponseOptions = _cls("ResponseOptions")
TokenInfo = _cls("TokenInfo")
BatchedTokenizeRequest = _cls("BatchedTokenizeRequest")
BatchedTokenizeResponse = _cls("BatchedTokenizeResponse")
TokenizeRequest = _cls("TokenizeRequest")
TokenizeResponse = _cls("TokenizeResponse")
ModelInfoRequest = _cls("ModelInfoRequest")
ModelInfoResponse = _cls("ModelInfoResponse")

DecodingMethod = enum_type_wrapper.EnumTypeWrapper(_pool.FindEnumTypeByName("fmaas.DecodingMethod"))
StopReason = enum_type_wrapper.EnumTypeWrapper(_pool.FindEnumTypeByName("fmaas.StopReason"))

SERVICE_NAME = "fmaas.GenerationService"
FILE_DESCRIPTOR_SERIALIZED = _FILE_PROTO.SerializeToString()
