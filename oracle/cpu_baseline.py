"""CPU ORACLE (test / measurement infrastructure): the reference's HF Jacobi path timed on the host cores.

`/root/reference` is Python and does not travel to the GPU box, so the timed code is the pinned restatement
(oracle/jacobi_policy.py, bit-exact against the reference's golden forwards) driving the SAME library forward the
reference calls (transformers Qwen2 on CPU) - cpu_baseline.kind = "port".  Used only by bench.py (cpu_baseline leg and
`--impl reference`).
"""
from __future__ import annotations

import os
import random
import time
from typing import Dict, List, Optional

import torch

from oracle import jacobi_policy as jp
from oracle.hf_model import HFForward, make_config


def build_timing_model(name: str, dtype=torch.bfloat16, seed: int = 0, **overrides):
    """Random-init Qwen2 of the named shape for TIMING: parameters are materialised from one pre-drawn N(0, 0.02)
    pool (distinct memory per tensor, values recycled) because drawing 7.6e9 fresh normals on the host costs minutes.
    Norm weights are 1, biases 0 (HF `_init_weights`)."""
    from transformers import Qwen2ForCausalLM
    cfg = make_config(name, **overrides)
    cfg._attn_implementation = "sdpa"
    with torch.device("meta"):
        model = Qwen2ForCausalLM(cfg)
    model = model.to_empty(device="cpu")
    g = torch.Generator().manual_seed(seed)
    biggest = max(p.numel() for p in model.parameters())
    pool = (torch.randn(biggest + 4096, generator=g, dtype=torch.float32) * 0.02).to(dtype)
    off = 0
    with torch.no_grad():
        for pname, p in model.named_parameters():
            if pname.endswith("norm.weight") or "layernorm" in pname:
                p.data = torch.ones(p.shape, dtype=dtype)
            elif pname.endswith(".bias"):
                p.data = torch.zeros(p.shape, dtype=dtype)
            else:
                off = (off + 1021) % 4096
                p.data = pool[off:off + p.numel()].view(p.shape).clone()
        # non-persistent buffers (rotary inv_freq) are lost by to_empty: rebuild the rotary module
        from transformers.models.qwen2.modeling_qwen2 import Qwen2RotaryEmbedding
        model.model.rotary_emb = Qwen2RotaryEmbedding(config=cfg)
    if cfg.tie_word_embeddings:
        model.tie_weights()
    return model.to(dtype).eval()


def build_model_from_named_tensors(name: str, named_tensors, dtype=torch.bfloat16, **overrides):
    """HF Qwen2 on CPU whose parameters are exactly the given (HF name, tensor) pairs."""
    from transformers import Qwen2ForCausalLM
    from transformers.models.qwen2.modeling_qwen2 import Qwen2RotaryEmbedding
    cfg = make_config(name, **overrides)
    cfg._attn_implementation = "sdpa"
    with torch.device("meta"):
        model = Qwen2ForCausalLM(cfg)
    model = model.to_empty(device="cpu")
    params = dict(model.named_parameters())
    seen = set()
    with torch.no_grad():
        for pname, t in named_tensors:
            p = params[pname]
            assert tuple(p.shape) == tuple(t.shape), pname
            p.data = t.to(dtype).contiguous()
            seen.add(pname)
        model.model.rotary_emb = Qwen2RotaryEmbedding(config=cfg)
    if cfg.tie_word_embeddings:
        model.tie_weights()
        seen.add("lm_head.weight")
    missing = [k for k in params if k not in seen]
    assert not missing, f"tensors missing for {missing[:3]}"
    return model.to(dtype).eval()


def time_reference_jacobi(model, prompt: List[int], n: int, mode: str = "single", max_new_tokens: Optional[int] = None,
                          budget_s: float = 20.0, K: int = 2, r: float = 0.85, pool: int = 4, pad: int = 0,
                          draft_seed: int = 2) -> Dict[str, float]:
    """Prefill (untimed, like the reference's gen_only_time, MR_humaneval.py:217-230) then Jacobi calls until
    `max_new_tokens` or the time budget; returns tokens, iterations and seconds of the timed part."""
    rng = random.Random(draft_seed)
    fwd = HFForward(model)
    t0 = time.perf_counter()
    draft = [rng.choice(prompt) for _ in range(n)]
    first, drafted = jp.hf_prefill(fwd, prompt, draft)
    t_prefill = time.perf_counter() - t0
    generated = list(prompt)
    tokens = iters = calls = 0
    t_gen = 0.0
    max_new = max_new_tokens if max_new_tokens is not None else n
    while tokens < max_new and t_gen < budget_s:
        inp = list(drafted) if calls == 0 else [first] + [rng.choice(generated) for _ in range(n - 1)]
        t1 = time.perf_counter()
        if mode == "multiblock":
            first, acc, itr = jp.hf_multiblock(fwd, inp, n, K=K, r=r, n_gram_pool_size=pool, pad=pad)
        else:
            first, acc, itr = jp.hf_single_block(fwd, inp, n)
        t_gen += time.perf_counter() - t1
        generated += acc
        tokens += len(acc)
        iters += itr
        calls += 1
    return {"tokens": tokens, "iters": iters, "seconds": t_gen, "prefill_seconds": t_prefill, "calls": calls,
            "cores": torch.get_num_threads()}


def host_cores() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1
