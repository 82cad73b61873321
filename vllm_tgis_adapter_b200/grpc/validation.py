"""TGIS request validation with the error strings of the TGIS API.

Behavioural mirror of /root/reference/src/vllm_tgis_adapter/grpc/validation.py:1-144 (message texts at :18-61 must
match the TGIS router byte for byte because clients grep them); re-written, not copied: checks are table driven.
"""
from __future__ import annotations

import enum
from typing import NoReturn

from .pb.generation_pb2 import DecodingMethod

MAX_TOP_N_TOKENS = 10       # validation.py:8
MAX_STOP_SEQS = 6           # validation.py:10
MAX_STOP_SEQ_LENGTH = 240   # validation.py:11
STRICT_PARAMETER_VALIDATION = False  # validation.py:15


class TGISValidationError(str, enum.Enum):
    TopP = "top_p must be > 0.0 and <= 1.0"
    TopK = "top_k must be strictly positive"
    TypicalP = "typical_p must be <= 1.0"
    RepetitionPenalty = "repetition_penalty must be > 0.0 and <= 2.0"
    LengthPenalty = "length_penalty.decay_factor must be >= 1.0 and <= 10.0"
    MaxNewTokens = "max_new_tokens must be <= {0}"
    MinNewTokens = "min_new_tokens must be <= max_new_tokens"
    InputLength = "input tokens ({0}) plus prefix length ({1}) plus min_new_tokens ({2}) must be <= {3}"
    InputLength2 = "input tokens ({0}) plus prefix length ({1}) must be < {2}"
    Tokenizer = "tokenizer error {0}"
    StopSequences = "can specify at most {0} non-empty stop sequences, each not more than {1} UTF8 bytes"
    TokenDetail = "must request input and/or generated tokens to request extra token detail"
    PromptPrefix = "can't retrieve prompt prefix with id '{0}': {1}"
    SampleParametersGreedy = "sampling parameters aren't applicable in greedy decoding mode"
    TopN = "top_n_tokens ({0}) must be <= {1}"
    AdapterNotFound = "can't retrieve adapter with id '{0}': {1}"
    AdaptersDisabled = "adapter_id supplied but no adapter store was configured"
    AdapterUnsupported = "adapter type {0} is not currently supported"
    InvalidAdapterID = "Invalid adapter id '{0}', must contain only alphanumeric, _ and - and /"
    # additions of this server (features the reference delegates to vLLM subsystems that are out of scope here)
    GuidedUnsupported = "guided decoding is not supported by this server"

    def error(self, *args: object) -> NoReturn:
        raise ValueError(self.value.format(*args))


def validate_input(min_tokens: int, token_num: int, max_model_len: int) -> None:
    """validation.py:64-77."""
    if token_num >= max_model_len:
        TGISValidationError.InputLength2.error(token_num, 0, max_model_len)
    if token_num + min_tokens > max_model_len:
        TGISValidationError.InputLength.error(token_num, 0, min_tokens, max_model_len)


def validate_params(params, max_max_new_tokens: int) -> None:
    """validation.py:80-144 (same checks, same order => same first error)."""
    resp, sampling, stopping, decoding = params.response, params.sampling, params.stopping, params.decoding
    greedy = params.method == DecodingMethod.GREEDY
    stop_seqs = list(stopping.stop_sequences)
    checks = [
        (decoding.HasField("length_penalty") and not (1.0 <= decoding.length_penalty.decay_factor <= 10.0),
         TGISValidationError.LengthPenalty, ()),
        (not (0 <= decoding.repetition_penalty <= 2), TGISValidationError.RepetitionPenalty, ()),
        (stopping.max_new_tokens > max_max_new_tokens, TGISValidationError.MaxNewTokens, (max_max_new_tokens,)),
        (stopping.min_new_tokens > (stopping.max_new_tokens or max_max_new_tokens),
         TGISValidationError.MinNewTokens, ()),
        (len(stop_seqs) > MAX_STOP_SEQS or not all(0 < len(s) <= MAX_STOP_SEQ_LENGTH for s in stop_seqs),
         TGISValidationError.StopSequences, (MAX_STOP_SEQS, MAX_STOP_SEQ_LENGTH)),
        (resp.top_n_tokens > MAX_TOP_N_TOKENS, TGISValidationError.TopN, (resp.top_n_tokens, MAX_TOP_N_TOKENS)),
        ((resp.token_logprobs or resp.token_ranks or resp.top_n_tokens)
         and not (resp.input_tokens or resp.generated_tokens), TGISValidationError.TokenDetail, ()),
        (STRICT_PARAMETER_VALIDATION and greedy
         and bool(sampling.temperature or sampling.top_k or sampling.top_p or sampling.typical_p),
         TGISValidationError.SampleParametersGreedy, ()),
        (sampling.top_k < 0, TGISValidationError.TopK, ()),
        (not (0 <= sampling.top_p <= 1), TGISValidationError.TopP, ()),
        (sampling.typical_p > 1, TGISValidationError.TypicalP, ()),
    ]
    for failed, err, args in checks:
        if failed:
            err.error(*args)
