"""Fuzz parity of the PRODUCT policy code (csrc/jf_policy.h, host build) against the pinned Python oracle
(oracle/jacobi_policy.py) on synthetic "models": a deterministic next-token function with strong repetition, so that
multi-token accepts, candidate rows from the n-gram pool, spawns, promotions and EOS paths are exercised far more often
than a random-init transformer does.  The oracle drives the fake forward and records every forward (kv, rows, greedy);
the C++ state machine must then replay exactly that trace and produce the same ids / iteration counts."""
import ctypes
import random

import numpy as np
import pytest

from jacobiforcing_b200.policy import MODES, PolicyKnobs, STOP_REASONS, mt_state_words
from oracle import jacobi_policy as jp
from tests.test_policy_host import _arr, _load_harness, _p


class FakeForward:
    """greedy[j] of a row = f(token at position j, absolute position): a toy language with cycles.  KV bookkeeping as
    the reference does to its DynamicCache."""

    def __init__(self, vocab, period, noise, seed):
        self.V, self.period, self.noise, self.seed = vocab, period, noise, seed
        self.kv = 0
        self.records = []

    def _next(self, tok, pos):
        h = (tok * 2654435761 + (pos // self.period) * 40503 + self.seed * 97) & 0xFFFFFFFF
        if (h >> 7) % 100 < self.noise:                      # occasional position-dependent surprise
            return (h >> 11) % self.V
        return (tok * 7 + 3) % self.V if tok % 5 else (tok + 1) % self.V

    def kv_len(self):
        return self.kv

    def forward(self, rows):
        L = len(rows[0])
        g = [[self._next(t, self.kv + j) for j, t in enumerate(r)] for r in rows]
        self.records.append({"kv_before": self.kv, "rows": [list(r) for r in rows], "greedy": g})
        self.kv += L
        return g

    def select_row(self, r):
        pass

    def trim(self, n):
        if n > 0:
            self.kv -= n


def run_case(seed):
    rnd = random.Random(seed)
    n = rnd.choice([4, 8, 16, 32])
    mode = rnd.choice(["single", "multiblock", "multiblock", "multiblock"])
    K = rnd.choice([1, 2, 2])
    r = rnd.choice([0.5, 0.75, 0.85, 1.0])
    pool = rnd.choice([0, 1, 2, 4, 8, 12])
    lsr = rnd.choice([0.0, 0.0, 0.25, 0.5])
    V = rnd.choice([11, 50, 997])
    max_new = rnd.choice([n, 3 * n, 5 * n + 3])
    P = rnd.randint(3, 40)
    prompt = [rnd.randrange(V) for _ in range(P)]
    eos = rnd.choice([None, None, rnd.randrange(V)])
    pad = rnd.choice([0, V - 1])
    fwd = FakeForward(V, rnd.choice([3, 7, 64]), rnd.choice([0, 5, 20, 50]), seed)
    # prefill as hf_driver does it
    rng = random.Random(seed + 1)
    res = jp.hf_driver(fwd, prompt, n, max_new, rng, mode=mode, K=K, r=r, lookahead_start_ratio=lsr,
                       n_gram_pool_size=pool, eos=eos, pad=pad)
    return dict(n=n, mode=mode, K=K, r=r, pool=pool, lsr=lsr, V=V, max_new=max_new, prompt=prompt, eos=eos, pad=pad,
                seed=seed), fwd.records, res


@pytest.mark.parametrize("seed", range(120))
def test_cxx_policy_equals_oracle_on_synthetic_models(seed):
    cfg, recs, res = run_case(seed)
    lib = _load_harness()
    n, prompt = cfg["n"], cfg["prompt"]
    kn = PolicyKnobs(mode=MODES[cfg["mode"]], n=n, K=cfg["K"], r=cfg["r"], lookahead_start_ratio=cfg["lsr"],
                     n_gram_pool_size=cfg["pool"], eos=cfg["eos"], pad=cfg["pad"], max_new_tokens=cfg["max_new"],
                     max_calls=1024)
    rng = random.Random(cfg["seed"] + 1)
    _ = [rng.choice(prompt) for _ in range(n)]
    prefill = recs[0]
    drafted = prefill["greedy"][0][len(prompt) - 1: len(prompt) + n - 1]
    if not res.generated and res.calls <= 1:
        return
    n_calls = res.calls + 2
    mt = _arr(mt_state_words(rng), np.uint32)
    fws = recs[1:]
    if not fws:
        return
    kv = _arr([f["kv_before"] for f in fws]); fB = _arr([len(f["rows"]) for f in fws])
    fL = _arr([len(f["rows"][0]) for f in fws])
    offs, rows, greedy, o = [], [], [], 0
    for f in fws:
        offs.append(o)
        for r_, g_ in zip(f["rows"], f["greedy"]):
            rows.extend(r_); greedy.extend(g_)
        o += len(f["rows"]) * len(f["rows"][0])
    offs = _arr(offs, np.int64); rows = _arr(rows); greedy = _arr(greedy)
    gen_cap = len(prompt) + cfg["max_new"] + 4 * n + 8
    gen = np.zeros(gen_cap, np.int32); iters = np.zeros(n_calls + 4, np.int32); out = np.zeros(8, np.int32)
    knobs = _arr(kn.as_ints()); pr = _arr(prompt); fct = _arr(drafted)
    rc = lib.jf_policy_replay(_p(knobs), _p(pr), len(prompt), len(prompt), drafted[0], _p(fct), _p(mt), 625,
                              None, 0, len(fws), _p(kv), _p(fB), _p(fL), _p(offs), _p(rows), _p(greedy), _p(gen),
                              gen_cap, _p(iters), len(iters), _p(out))
    assert rc == 0, f"{cfg}: diverged at forward {rc - 100 if rc >= 100 else rc}"
    gen_len, calls, stop, total_new, err, used, total_iters, ctx = [int(x) for x in out]
    assert err == 0, cfg
    assert gen[len(prompt):gen_len].tolist() == res.generated, cfg
    assert calls == res.calls and iters[:calls].tolist() == res.iters, cfg
    assert STOP_REASONS[stop] == res.stop_reason, cfg
    assert used == len(fws), cfg
