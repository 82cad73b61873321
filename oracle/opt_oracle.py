"""CPU oracle for the OPT model family ("OPT ... to serve the literal facebook/opt-125m of BASELINE
configs[0]") — TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this module; the product path never does.

The reference's own OPT fixture is `facebook/opt-125m` served through `self.engine.generate(...)`
(/root/reference/tests/conftest.py:79-91 (no --model: vLLM's default, facebook/opt-125m) and tests/test_hub.py:17, tests/test_grpc_server.py:42-49, grpc/grpc_server.py:222); the arithmetic is
the un-vendored dependency vLLM (installed 0.22.0).  Restated here in plain torch ops with the model-dtype rounding
points of that path, each step citing the file:line it follows:

  embeddings         vllm model_executor/models/opt.py:61-70 (learned positions, offset 2), :268-290
                     (inputs_embeds + pos_embeds: a model-dtype add); word_embed_proj_dim == hidden (no project_in/out:
                     facebook/opt-125m, -1.3b ... all have it equal)
  decoder layer      opt.py:170-197 with do_layer_norm_before=True (every released OPT except 350m):
                     LayerNorm -> attention -> residual add -> LayerNorm -> fc1 -> ReLU -> fc2 -> residual add
                     == transformers models/opt/modeling_opt.py OPTDecoderLayer.forward
  LayerNorm          torch.nn.LayerNorm on a model-dtype tensor: fp32 mean / biased variance, (x - mean) * rstd * w + b
                     evaluated in fp32, rounded ONCE to the model dtype
  linear + bias      F.linear(x, W, b) in the model dtype: fp32 accumulation, bias added to the accumulator, one
                     rounding (cuBLASLt bias epilogue on the GPU, oneDNN on the CPU)
  attention          opt.py:73-124: softmax(q k^T * head_dim**-0.5) v, causal, no rotary embedding; fp32 scores and
                     probabilities (flash-attention semantics), output rounded to the model dtype
  logits             opt.py:364-420: final_layer_norm, lm_head tied to embed_tokens, no bias; rounded to the model dtype
                     (vllm layers/logits_processor.py:89-104), the sampler casts to fp32 (v1/sample/sampler.py:91)

PARITY PINNING: tests/test_opt_oracle_cpu.py checks this module against transformers' OPTForCausalLM (random-init, fp32 and
bf16, eager attention) through the fixtures tests/golden/opt_hf_{fp32,bf16}.json written by oracle/gen_opt_golden.py.
The real facebook/opt-125m weights are not obtainable here (no network): the fixtures use seeded random weights of the
same architecture, which pins the arithmetic but not a checkpoint.
"""
from __future__ import annotations

import dataclasses

import torch
import torch.nn.functional as F


@dataclasses.dataclass
class OPTConfig:
    n_layers: int
    hidden: int
    n_heads: int
    ffn: int
    vocab: int
    max_positions: int = 2048  # max_position_embeddings (the table has 2 more rows: opt.py:66)
    ln_eps: float = 1e-5
    max_model_len: int = 2048

    @property
    def head_dim(self) -> int:
        return self.hidden // self.n_heads


OPT_CONFIGS = {
    # 64-dim heads like every released OPT; small enough for the CPU suite
    "opt-tiny": OPTConfig(n_layers=2, hidden=256, n_heads=4, ffn=1024, vocab=1024, max_positions=512, max_model_len=512),
    # facebook/opt-125m (config.json: 12 layers, hidden 768, 12 heads, ffn 3072, vocab 50272, 2048 positions)
    "opt-125m": OPTConfig(n_layers=12, hidden=768, n_heads=12, ffn=3072, vocab=50272),
}

POSITION_OFFSET = 2


def synthetic_opt_weights(cfg: OPTConfig, seed: int = 0, dtype: torch.dtype = torch.bfloat16,
                          std: float = 0.02) -> dict[str, torch.Tensor]:
    """Seeded random weights under HF's OPT parameter names.  Unlike HF's init (biases 0, LayerNorm 1/0) the biases and
    the LayerNorm affine parameters are random too, so that a dropped bias or a swapped weight/bias shows up."""
    g = torch.Generator().manual_seed(seed)

    def rnd(*shape: int, s: float = std) -> torch.Tensor:
        return (torch.randn(*shape, generator=g, dtype=torch.float32) * s).to(dtype)

    def ln_w() -> torch.Tensor:
        return (1.0 + 0.1 * torch.randn(cfg.hidden, generator=g, dtype=torch.float32)).to(dtype)

    H = cfg.hidden
    w = {
        "model.decoder.embed_tokens.weight": rnd(cfg.vocab, H),
        "model.decoder.embed_positions.weight": rnd(cfg.max_positions + POSITION_OFFSET, H),
        "model.decoder.final_layer_norm.weight": ln_w(),
        "model.decoder.final_layer_norm.bias": rnd(H, s=0.05),
    }
    for i in range(cfg.n_layers):
        p = f"model.decoder.layers.{i}."
        for m in ("q_proj", "k_proj", "v_proj", "out_proj"):
            w[p + f"self_attn.{m}.weight"] = rnd(H, H)
            w[p + f"self_attn.{m}.bias"] = rnd(H, s=0.05)
        w[p + "self_attn_layer_norm.weight"] = ln_w()
        w[p + "self_attn_layer_norm.bias"] = rnd(H, s=0.05)
        w[p + "fc1.weight"] = rnd(cfg.ffn, H)
        w[p + "fc1.bias"] = rnd(cfg.ffn, s=0.05)
        w[p + "fc2.weight"] = rnd(H, cfg.ffn)
        w[p + "fc2.bias"] = rnd(H, s=0.05)
        w[p + "final_layer_norm.weight"] = ln_w()
        w[p + "final_layer_norm.bias"] = rnd(H, s=0.05)
    return w


class OPTSeqState:
    def __init__(self, cfg: OPTConfig, dtype: torch.dtype, device: str = "cpu"):
        self.k = [torch.empty(0, cfg.n_heads, cfg.head_dim, dtype=dtype, device=device) for _ in range(cfg.n_layers)]
        self.v = [torch.empty(0, cfg.n_heads, cfg.head_dim, dtype=dtype, device=device) for _ in range(cfg.n_layers)]
        self.n = 0


class OPTOracle:
    """Flat-batch OPT forward; same interface as LlamaOracle (new_seq / step / head)."""

    def __init__(self, cfg: OPTConfig, weights: dict[str, torch.Tensor], dtype: torch.dtype = torch.bfloat16,
                 device: str = "cpu"):
        self.cfg = cfg
        self.dtype = dtype
        self.device = device
        w = {k: v.to(device=device, dtype=dtype) for k, v in weights.items()}
        d = "model.decoder."
        self.embed = w[d + "embed_tokens.weight"]
        self.pos_embed = w[d + "embed_positions.weight"]
        self.lm_head = w.get("lm_head.weight", self.embed)  # tie_word_embeddings
        self.final_ln = (w[d + "final_layer_norm.weight"], w[d + "final_layer_norm.bias"])
        self.layers = []
        for i in range(cfg.n_layers):
            p = f"{d}layers.{i}."
            self.layers.append({
                "qkv_w": torch.cat([w[p + f"self_attn.{m}.weight"] for m in ("q_proj", "k_proj", "v_proj")], 0).contiguous(),
                "qkv_b": torch.cat([w[p + f"self_attn.{m}.bias"] for m in ("q_proj", "k_proj", "v_proj")], 0).contiguous(),
                "o_w": w[p + "self_attn.out_proj.weight"], "o_b": w[p + "self_attn.out_proj.bias"],
                "ln1": (w[p + "self_attn_layer_norm.weight"], w[p + "self_attn_layer_norm.bias"]),
                "fc1_w": w[p + "fc1.weight"], "fc1_b": w[p + "fc1.bias"],
                "fc2_w": w[p + "fc2.weight"], "fc2_b": w[p + "fc2.bias"],
                "ln2": (w[p + "final_layer_norm.weight"], w[p + "final_layer_norm.bias"]),
            })
        self._lm_head_f32 = None
        self._f32: dict[int, tuple[torch.Tensor, torch.Tensor]] = {}

    def _ln(self, x: torch.Tensor, wb: tuple[torch.Tensor, torch.Tensor]) -> torch.Tensor:
        xf = x.float()
        mean = xf.mean(-1, keepdim=True)
        var = (xf - mean).pow(2).mean(-1, keepdim=True)
        y = (xf - mean) * torch.rsqrt(var + self.cfg.ln_eps) * wb[0].float() + wb[1].float()
        return y.to(self.dtype)

    def _linear(self, x: torch.Tensor, w: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
        # fp32 copies of the (static) parameters are cached: same arithmetic, no per-step conversion of the weights
        key = w.data_ptr()
        if key not in self._f32:
            self._f32[key] = (w.float().t().contiguous(), b.float())
        wt, bf = self._f32[key]
        return (x.float() @ wt + bf).to(self.dtype)

    def _attend(self, q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, first_pos: int) -> torch.Tensor:
        s = torch.einsum("qhd,khd->hqk", q.float(), k.float()) * (self.cfg.head_dim ** -0.5)
        tq, tk = q.shape[0], k.shape[0]
        qpos = first_pos + torch.arange(tq, device=q.device)[:, None]
        mask = torch.arange(tk, device=q.device)[None, :] <= qpos
        s = s.masked_fill(~mask[None], float("-inf"))
        return torch.einsum("hqk,khd->qhd", torch.softmax(s, dim=-1), v.float()).to(self.dtype)

    @torch.no_grad()
    def step(self, work: list[tuple[OPTSeqState, list[int]]], want_all_logits: bool = False) -> torch.Tensor:
        cfg = self.cfg
        H, nh, hd = cfg.hidden, cfg.n_heads, cfg.head_dim
        toks = torch.tensor([t for _, ts in work for t in ts], dtype=torch.long, device=self.device)
        pos = torch.tensor([st.n + j for st, ts in work for j in range(len(ts))], dtype=torch.long, device=self.device)
        T = toks.numel()
        h = self.embed[toks] + self.pos_embed[pos + POSITION_OFFSET]
        for li, L in enumerate(self.layers):
            xn = self._ln(h, L["ln1"])
            qkv = self._linear(xn, L["qkv_w"], L["qkv_b"])
            q = qkv[:, :H].reshape(T, nh, hd)
            k = qkv[:, H: 2 * H].reshape(T, nh, hd)
            v = qkv[:, 2 * H:].reshape(T, nh, hd)
            outs, off = [], 0
            for st, ts in work:
                n = len(ts)
                st.k[li] = torch.cat([st.k[li], k[off: off + n]], dim=0)
                st.v[li] = torch.cat([st.v[li], v[off: off + n]], dim=0)
                outs.append(self._attend(q[off: off + n], st.k[li], st.v[li], st.n))
                off += n
            attn = torch.cat(outs, dim=0).reshape(T, H)
            h = h + self._linear(attn, L["o_w"], L["o_b"])
            xn = self._ln(h, L["ln2"])
            act = F.relu(self._linear(xn, L["fc1_w"], L["fc1_b"]))
            h = h + self._linear(act, L["fc2_w"], L["fc2_b"])
        for st, ts in work:
            st.n += len(ts)
        if not want_all_logits:
            last, off = [], 0
            for _, ts in work:
                off += len(ts)
                last.append(off - 1)
            h = h[torch.tensor(last, device=self.device)]
        return self.head(h)

    def head(self, h: torch.Tensor) -> torch.Tensor:
        x = self._ln(h, self.final_ln)
        if self._lm_head_f32 is None:
            self._lm_head_f32 = self.lm_head.float()
        return (x.float() @ self._lm_head_f32.t()).to(self.dtype).float()

    def new_seq(self) -> OPTSeqState:
        return OPTSeqState(self.cfg, self.dtype, self.device)
