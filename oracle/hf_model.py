"""CPU ORACLE (test infrastructure, not product code): the Qwen2 forward the reference's Jacobi functions call.

The arithmetic of the hot path is NOT in /root/reference: modeling/cllm2_..._multiblock_lookahead_unified.py:429-463 calls
`self.model.embed_tokens`, `self.model.rotary_emb`, every `Qwen2DecoderLayer`, `self.model.norm` and `self.lm_head`
of the third-party `transformers` package (pinned transformers==4.53.0 in requirements.txt:97; this image has 5.5.0,
whose Qwen2 math is unchanged: fp32 RMSNorm statistics with a bf16 weight multiply, bf16 rotary tables, GQA, SwiGLU).
This module therefore does not restate the math - it *executes* the same library the reference executes, on CPU,
behind the small `fwd` interface oracle/jacobi_policy.py drives.  No golden logits exist in the reference
(SURVEY.md section 4), so floating-point parity is "unpinned" beyond: same library, same weights, same inputs.

Also here: the seeded random-init checkpoints every test / bench uses (there are no model files offline).
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence

import torch

# Published shapes (HF model cards); the reference ships no config.json.
QWEN_SHAPES: Dict[str, dict] = {
    # small shapes the CUDA path supports (head_dim in {64,128}, hidden % 64 == 0) for fast CPU parity runs
    "tiny": dict(hidden_size=256, num_hidden_layers=2, num_attention_heads=4, num_key_value_heads=2, head_dim=64,
                 intermediate_size=512, vocab_size=1000, tie_word_embeddings=False, rope_theta=1e6,
                 max_position_embeddings=4096),
    "tiny128": dict(hidden_size=512, num_hidden_layers=3, num_attention_heads=4, num_key_value_heads=1,
                    head_dim=128, intermediate_size=1152, vocab_size=2048, tie_word_embeddings=False,
                    rope_theta=1e4, max_position_embeddings=4096),
    # small but with several k-splits per GEMM (QKV 4, O 4, down 8 on 148 SMs): exercises the fused split-K tails
    "mid128": dict(hidden_size=1024, num_hidden_layers=2, num_attention_heads=8, num_key_value_heads=2,
                   head_dim=128, intermediate_size=2048, vocab_size=2048, tie_word_embeddings=False,
                   rope_theta=1e6, max_position_embeddings=4096),
    "qwen2-0.5b": dict(hidden_size=896, num_hidden_layers=24, num_attention_heads=14, num_key_value_heads=2,
                       head_dim=64, intermediate_size=4864, vocab_size=151936, tie_word_embeddings=True,
                       rope_theta=1e6, max_position_embeddings=32768),
    "qwen2.5-coder-7b": dict(hidden_size=3584, num_hidden_layers=28, num_attention_heads=28, num_key_value_heads=4,
                             head_dim=128, intermediate_size=18944, vocab_size=152064, tie_word_embeddings=False,
                             rope_theta=1e6, max_position_embeddings=32768),
    "qwen2.5-math-7b": dict(hidden_size=3584, num_hidden_layers=28, num_attention_heads=28, num_key_value_heads=4,
                            head_dim=128, intermediate_size=18944, vocab_size=152064, tie_word_embeddings=False,
                            rope_theta=1e4, max_position_embeddings=4096),
}


def make_config(name: str, **overrides):
    from transformers import Qwen2Config
    kw = dict(QWEN_SHAPES[name])
    kw.update(overrides)
    theta = kw.pop("rope_theta")
    head_dim = kw.pop("head_dim")
    cfg = Qwen2Config(rms_norm_eps=1e-6, attention_dropout=0.0, use_sliding_window=False,
                      rope_parameters={"rope_type": "default", "rope_theta": float(theta)}, **kw)
    assert cfg.hidden_size // cfg.num_attention_heads == head_dim, "Qwen2 derives head_dim = hidden / heads"
    cfg.rope_theta = float(theta)            # transformers 5.x keeps it in rope_parameters; keep the 4.x attribute too
    return cfg


def build_model(name: str, seed: int = 0, dtype: torch.dtype = torch.float32, attn: str = "eager", **overrides):
    """Seeded HF-`_init_weights` random model (N(0, 0.02) linears/embeddings, unit norms, zero biases -> biases are then
    re-drawn N(0, 0.02) so the qkv-bias path is exercised)."""
    from transformers import Qwen2ForCausalLM
    cfg = make_config(name, **overrides)
    cfg._attn_implementation = attn
    torch.manual_seed(seed)
    model = Qwen2ForCausalLM(cfg)
    g = torch.Generator().manual_seed(seed + 1000)
    with torch.no_grad():
        for layer in model.model.layers:
            for proj in (layer.self_attn.q_proj, layer.self_attn.k_proj, layer.self_attn.v_proj):
                proj.bias.copy_(torch.randn(proj.bias.shape, generator=g) * 0.02)
            for norm in (layer.input_layernorm, layer.post_attention_layernorm):
                norm.weight.copy_(1.0 + 0.1 * torch.randn(norm.weight.shape, generator=g))
        model.model.norm.weight.copy_(1.0 + 0.1 * torch.randn(model.model.norm.weight.shape, generator=g))
    model = model.to(dtype).eval()
    return model


def state_dict_bf16(model) -> Dict[str, torch.Tensor]:
    """HF-named tensors as the reference loader would read them from safetensors (inference_engine/utils/loader.py:27-47)."""
    sd = {}
    for k, v in model.state_dict().items():
        sd[k] = v.detach().to(torch.bfloat16).contiguous()
    if getattr(model.config, "tie_word_embeddings", False):
        sd.pop("lm_head.weight", None)
    return sd


class HFForward:
    """`fwd` interface of oracle/jacobi_policy.py on top of transformers' DynamicCache - the operations the reference
    applies to `past_key_values` (unified.py:93-127 resize, :500-502 row select, :36-59 trim)."""

    def __init__(self, model, record_logits: bool = False):
        from transformers.cache_utils import DynamicCache
        self.model = model
        self.cache = DynamicCache(config=model.config)
        self.record_logits = record_logits
        self.last_logits: Optional[torch.Tensor] = None
        self.n_forwards = 0
        self.min_margin: List[float] = []

    def kv_len(self) -> int:
        return int(self.cache.get_seq_length())

    def _layers(self):
        return [l for l in self.cache.layers if getattr(l, "is_initialized", False) and l.keys is not None]

    def _resize(self, B: int) -> None:
        for l in self._layers():
            b = l.keys.shape[0]
            if b == B:
                continue
            if b == 1:
                l.keys = l.keys.expand(B, -1, -1, -1).contiguous()
                l.values = l.values.expand(B, -1, -1, -1).contiguous()
            else:
                l.keys = l.keys[:1].expand(B, -1, -1, -1).contiguous()
                l.values = l.values[:1].expand(B, -1, -1, -1).contiguous()

    @torch.no_grad()
    def forward(self, rows: Sequence[Sequence[int]]) -> List[List[int]]:
        ids = torch.tensor([list(r) for r in rows], dtype=torch.long)
        self._resize(ids.shape[0])
        out = self.model(input_ids=ids, past_key_values=self.cache, use_cache=True)
        logits = out.logits.float()                      # reference: lm_head(...).float(), unified.py:463
        self.n_forwards += 1
        if self.record_logits:
            self.last_logits = logits
        top2 = logits.topk(2, dim=-1).values
        self.min_margin.append(float((top2[..., 0] - top2[..., 1]).min()))
        return torch.argmax(logits, dim=-1).tolist()     # first index wins ties on CPU

    def select_row(self, r: int) -> None:
        for l in self._layers():
            if l.keys.shape[0] > 1:
                l.keys = l.keys[r:r + 1].contiguous()
                l.values = l.values[r:r + 1].contiguous()

    def trim(self, n: int) -> None:
        if n <= 0:
            return
        for l in self._layers():
            l.keys = l.keys[..., :-n, :]
            l.values = l.values[..., :-n, :]


@torch.no_grad()
def teacher_forced_logits(model, ids: Sequence[int]) -> torch.Tensor:
    """fp32 logits [len, V] of one full-sequence causal forward (the HF forward, no cache)."""
    t = torch.tensor([list(ids)], dtype=torch.long)
    return model(input_ids=t, use_cache=False).logits[0].float()


def bf16_copy(model):
    """A bf16 copy of an fp32 oracle model, the dtype the reference runs in.  A COPY on purpose: `model.to(bfloat16)`
    converts in place and also rounds the rotary embedding's `inv_freq` buffer, which a later `.to(float32)` does not
    restore - the "fp32 oracle" would silently keep bf16-rounded rotary frequencies (error grows with the position).
    The copy keeps the fp32 inv_freq, as a model loaded with torch_dtype=bfloat16 does."""
    import copy
    m = copy.deepcopy(model).to(torch.bfloat16)
    src = dict(model.named_buffers())
    for name, buf in m.named_buffers():
        if name.endswith("inv_freq") and name in src:
            buf.data = src[name].detach().clone().float()
    return m.eval()


@torch.no_grad()
def greedy_ar(model, prompt: Sequence[int], n_new: int) -> List[int]:
    """HF greedy decoding, the property the reference's own tests pin Jacobi against
    (inference_engine/tests/test_jacobi_decoding_greedy.py:180-206)."""
    fwd = HFForward(model)
    g = fwd.forward([list(prompt)])[0]
    fwd.select_row(0)
    out = [int(g[-1])]
    while len(out) < n_new:
        g = fwd.forward([[out[-1]]])[0]
        out.append(int(g[-1]))
    return out


def rope_tables(theta: float, head_dim: int, max_pos: int):
    """cos/sin exactly as transformers' Qwen2RotaryEmbedding computes them (fp32 inv_freq, fp32 outer product, fp32
    cos/sin); the caller casts to the activation dtype like `cos.to(dtype=x.dtype)` does."""
    inv_freq = 1.0 / (theta ** (torch.arange(0, head_dim, 2, dtype=torch.int64).float() / head_dim))
    pos = torch.arange(max_pos, dtype=torch.float32)
    freqs = pos[:, None] * inv_freq[None, :]
    return freqs.cos(), freqs.sin()          # [max_pos, head_dim/2]
