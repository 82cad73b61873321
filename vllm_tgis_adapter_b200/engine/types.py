"""Value types crossing the drop-in boundary: the subset of vLLM's SamplingParams / RequestOutput /
CompletionOutput / Logprob that /root/reference/src/vllm_tgis_adapter/grpc/grpc_server.py actually touches
(:205-225 generate kwargs, :282-283,298,374,438-456 RequestOutput, :391-392,477-487,670-690 CompletionOutput,
:727-755 Logprob).  Same attribute names so the adapter code reads identically."""
from __future__ import annotations

import dataclasses
import enum


class RequestOutputKind(enum.Enum):
    CUMULATIVE = 0
    DELTA = 1
    FINAL_ONLY = 2


@dataclasses.dataclass
class StructuredOutputsParams:
    """vllm/sampling_params.py `StructuredOutputsParams`, the members the adapter sets
    (/root/reference/src/vllm_tgis_adapter/tgis_utils/structured_outputs.py:14-38): exactly one of them."""

    json: str | dict | None = None
    regex: str | None = None
    choice: list[str] | None = None
    grammar: str | None = None
    json_object: bool | None = None

    def __post_init__(self) -> None:
        count = sum(v is not None and v is not False for v in
                    (self.json, self.regex, self.choice, self.grammar, self.json_object))
        if count > 1:
            raise ValueError("You can only use one kind of structured outputs constraint "
                             f"but multiple are specified: {dataclasses.asdict(self)}")
        if count < 1:
            raise ValueError("You must use one kind of structured outputs constraint "
                             f"but none are specified: {dataclasses.asdict(self)}")


@dataclasses.dataclass
class SamplingParams:
    """Field names follow vllm/sampling_params.py; normalisation rules follow :398-440."""

    temperature: float = 1.0
    top_k: int = -1
    top_p: float = 1.0
    seed: int | None = None
    logprobs: int | None = None
    prompt_logprobs: int | None = None
    max_tokens: int | None = None
    min_tokens: int = 0
    repetition_penalty: float = 1.0
    stop: list[str] | None = None
    include_stop_str_in_output: bool = False
    skip_special_tokens: bool = True
    output_kind: RequestOutputKind = RequestOutputKind.CUMULATIVE
    # the reference's per-request logits processors (grpc_server.py:560-578), expressed as data for the fused kernel
    typical_p: float = 0.0
    length_penalty: tuple[int, float] | None = None
    eos_token_id: int | None = None
    stop_token_ids: list[int] = dataclasses.field(default_factory=list)
    structured_outputs: StructuredOutputsParams | None = None   # grpc_server.py:580-586

    def __post_init__(self) -> None:
        if self.temperature < 0.0:
            raise ValueError(f"temperature must be non-negative, got {self.temperature}.")
        if 0 < self.temperature < 1e-2:   # vllm sampling_params.py: clamp tiny temperatures
            self.temperature = 1e-2
        if self.seed == -1:
            self.seed = None
        if self.top_k == 0:
            self.top_k = -1
        if self.top_k < -1:
            raise ValueError(f"top_k must be -1 (disable), or at least 1, got {self.top_k}.")
        if not 0.0 < self.top_p <= 1.0:
            raise ValueError(f"top_p must be in (0, 1], got {self.top_p}.")
        if not 0.0 < self.repetition_penalty <= 2.0:
            raise ValueError(f"repetition_penalty must be in (0, 2], got {self.repetition_penalty}.")
        if self.max_tokens is not None and self.max_tokens < 1:
            raise ValueError(f"max_tokens must be at least 1, got {self.max_tokens}.")
        if self.min_tokens < 0:
            raise ValueError(f"min_tokens must be greater than or equal to 0, got {self.min_tokens}.")
        if self.max_tokens is not None and self.min_tokens > self.max_tokens:
            raise ValueError(
                f"min_tokens must be less than or equal to max_tokens={self.max_tokens}, got {self.min_tokens}.")
        if self.logprobs is not None and self.logprobs < 0:
            raise ValueError(f"logprobs must be non-negative, got {self.logprobs}.")
        if self.temperature < 1e-5:       # greedy: top-p/top-k are irrelevant
            self.top_p, self.top_k = 1.0, -1

    @property
    def greedy(self) -> bool:
        return self.temperature < 1e-5


@dataclasses.dataclass
class Logprob:
    logprob: float
    rank: int | None = None
    decoded_token: str | None = None


@dataclasses.dataclass
class CompletionOutput:
    index: int
    text: str
    token_ids: list[int]
    logprobs: list[dict[int, Logprob]] | None
    finish_reason: str | None = None          # None | "length" | "stop" | "abort"
    stop_reason: int | str | None = None


@dataclasses.dataclass
class RequestMetrics:
    arrival_time: float = 0.0
    first_scheduled_time: float | None = None
    first_token_time: float | None = None
    last_token_time: float = 0.0
    time_in_queue: float | None = None
    finished_time: float | None = None


@dataclasses.dataclass
class RequestOutput:
    request_id: str
    prompt: str | None
    prompt_token_ids: list[int]
    prompt_logprobs: list[dict[int, Logprob] | None] | None
    outputs: list[CompletionOutput]
    finished: bool
    metrics: RequestMetrics | None = None


@dataclasses.dataclass
class LoRARequest:
    """vllm/lora/request.py, the members the adapter layer sets (grpc/adapters.py:146-155 via load_lora_adapter)."""

    lora_name: str
    lora_int_id: int
    lora_path: str


@dataclasses.dataclass
class TokensPrompt:
    prompt_token_ids: list[int]
