"""Shared helpers of the GPU parity tests: build a seeded HF model (oracle side, CPU) and the runtime loaded with the
same bf16 weights; a `fwd` adapter that lets the oracle's policy restatement drive the DEVICE forward."""
from __future__ import annotations

import numpy as np
import torch

from jacobiforcing_b200.runtime import JacobiRuntime
from oracle.hf_model import build_model, rope_tables, state_dict_bf16


def make_pair(name: str, seed: int = 0, model_overrides=None, **rt_kw):
    """(hf_model_fp32_with_bf16_rounded_weights, runtime).  Both sides hold bit-identical weights."""
    model = build_model(name, seed=seed, dtype=torch.float32, **(model_overrides or {}))
    sd = state_dict_bf16(model)
    with torch.no_grad():                       # round the oracle's weights to bf16 too
        for k, v in model.state_dict().items():
            v.copy_(v.to(torch.bfloat16).to(torch.float32))
    kw = dict(max_rows=256, max_seqs=4, num_kv_pages=16, max_pages_per_seq=4, trace_capacity=1 << 16)
    kw.update(rt_kw)
    rt = JacobiRuntime(model.config, **kw)
    rt.load_state_dict(sd)
    cfg = model.config
    hd = cfg.hidden_size // cfg.num_attention_heads
    cos, sin = rope_tables(float(cfg.rope_theta), hd, cfg.max_position_embeddings)
    rt.set_rope_tables(cos.numpy(), sin.numpy())
    return model, rt


def bf16_bits_to_float(a: np.ndarray) -> np.ndarray:
    return (a.astype(np.uint32) << 16).view(np.float32)


class DeviceForward:
    """`fwd` interface of oracle/jacobi_policy.py on top of the runtime's planned forward: the ORACLE's state machine
    consumes the DEVICE's argmaxes, so comparing it with the device-resident loop isolates the policy (integer work,
    bit-exact) from floating point."""

    def __init__(self, rt: JacobiRuntime, seq):
        self.rt, self.seq = rt, seq
        self.kv = seq.ctx            # logical KV length the policy sees
        self.base = seq.ctx          # kv length before the last forward
        self.cached = seq.ctx        # tokens physically in the paged cache
        self.last_L = 0
        self.sel = 0
        self.n_forwards = 0
        self.rows_in_cache = 1

    def kv_len(self):
        return self.kv

    def _flush(self):
        if self.kv > self.cached:    # keep what the policy kept of the last forward's selected row
            commits = [(self.sel * self.last_L + (p - self.base), 0, p) for p in range(self.cached, self.kv)]
            self.rt.commit_rows([self.seq], commits)
            self.cached = self.kv
        elif self.kv < self.cached:
            self.cached = self.kv

    def forward(self, rows):
        self._flush()
        B, L = len(rows), len(rows[0])
        self.rt.ensure_capacity(self.seq, self.kv + L)
        toks = [t for r in rows for t in r]
        groups = [(i * L, L, 0, self.kv) for i in range(B)]
        res = self.rt.forward(toks, groups, [self.seq], 0, B * L)
        self.base, self.last_L, self.sel = self.kv, L, 0
        self.rows_in_cache = B
        self.kv += L
        self.n_forwards += 1
        g = res["greedy"]
        return [[int(t) for t in g[i * L:(i + 1) * L]] for i in range(B)]

    def select_row(self, r):
        # DynamicCache semantics: the first narrowing after a forward picks the row; once the cache has one row,
        # further narrowings (the pseudo spans' best_idx = 0, unified.py:487-502) are no-ops
        if self.rows_in_cache > 1:
            self.sel = r
            self.rows_in_cache = 1

    def trim(self, n):
        if n > 0:
            self.kv -= n


class DeviceEngineForward:
    """forward_step callback of oracle.jacobi_policy.engine_fixed_len on the device forward: draft[0] is the last
    committed token at position S-1; what the policy accepted of the previous draft is committed lazily."""

    def __init__(self, rt: JacobiRuntime, seq, want_logits: bool = False):
        self.rt, self.seq = rt, seq
        self.cached = seq.ctx          # tokens physically in the cache
        self.last_base = None
        self.n_forwards = 0
        self.want_logits = want_logits
        self.last_logits = None        # fp32 [L, V] of the last forward (want_logits)

    def __call__(self, token_ids, draft):
        S = len(token_ids)
        base = S - 1
        if self.last_base is not None and base > self.last_base:
            self.rt.commit_rows([self.seq], [(p - self.last_base, 0, p) for p in range(self.last_base, base)])
        L = len(draft)
        self.rt.ensure_capacity(self.seq, base + L)
        res = self.rt.forward(list(draft), [(0, L, 0, base)], [self.seq], 0, L, want_logits=self.want_logits)
        self.last_base = base
        self.n_forwards += 1
        self.last_logits = res["logits"]
        return [int(t) for t in res["greedy"][:L - 1]]
