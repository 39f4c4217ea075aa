"""Pins oracle/jacobi_policy.py (the CPU restatement) against the UNMODIFIED reference:

  * replay:   tests/golden/*.json hold every forward the reference issued (kv length, `out` rows, greedy table) and every
              call's outputs, produced by oracle/make_golden.py running the reference functions themselves.  The
              restatement, driven by a table-lookup forward, must issue the SAME forwards and return the SAME ids and
              iteration counts - no model, no floating point, machine independent.
  * hf:       the same cases re-run with the HF-transformers forward (oracle/hf_model.py) on the seeded model.
  * live:     when /root/reference is present (this container, not the GPU box) the reference is re-run and compared.
"""
import glob
import json
import os
import random

import pytest

from oracle import jacobi_policy as jp

GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "*.json")))
CASES = [os.path.splitext(os.path.basename(p))[0] for p in GOLDEN
         if json.load(open(p)).get("mode") in ("single", "multiblock")]


def load(case):
    with open(os.path.join(os.path.dirname(__file__), "golden", case + ".json")) as f:
        return json.load(f)


class ReplayForward:
    """fwd interface backed by the recorded greedy table; asserts the restatement asks for exactly the forwards the
    reference issued, in order."""

    def __init__(self, forwards):
        self.forwards = forwards
        self.i = 0
        self.kv = 0

    def kv_len(self):
        return self.kv

    def forward(self, rows):
        assert self.i < len(self.forwards), "restatement issued more forwards than the reference"
        rec = self.forwards[self.i]
        assert rec["kv_before"] == self.kv, f"forward {self.i}: kv {self.kv} != reference {rec['kv_before']}"
        assert [list(r) for r in rows] == rec["rows"], f"forward {self.i}: rows differ from the reference's `out`"
        self.i += 1
        self.kv += len(rows[0])
        return rec["greedy"]

    def select_row(self, r):
        pass

    def trim(self, n):
        if n > 0:
            self.kv -= n


def drive(g, fwd):
    return jp.hf_driver(fwd, g["prompt"], g["n"], g["max_new_tokens"], random.Random(g["draft_seed"]),
                        mode=g["mode"], K=g["K"], r=g["r"], lookahead_start_ratio=g["lookahead_start_ratio"],
                        n_gram_pool_size=g["n_gram_pool_size"], eos=g["eos"], pad=g["pad"],
                        max_iteration_count=g["max_iteration_count"])


def check(g, res):
    assert res.generated == g["generated"]
    assert res.iters == g["iters"]
    assert res.calls == g["calls"]
    assert res.stop_reason == g["stop_reason"]
    assert res.total_new_tokens == g["total_new_tokens"]


def test_goldens_exist():
    assert len(CASES) >= 10


@pytest.mark.parametrize("case", CASES)
def test_replay_matches_reference(case):
    g = load(case)
    fwd = ReplayForward(g["forwards"])
    res = drive(g, fwd)
    check(g, res)
    assert fwd.i == len(g["forwards"]), "restatement issued fewer forwards than the reference"
    decode_calls = [c for c in g["call_log"] if c["kind"] == "decode"]
    assert sum(c["iters"] for c in decode_calls) == res.total_iterations


@pytest.mark.parametrize("case", CASES)
def test_reference_property_equals_greedy_ar(case):
    g = load(case)
    if g["eos"] is None:
        assert g["equals_greedy_ar"] is True      # inference_engine/tests/test_jacobi_decoding_greedy.py:180-206


@pytest.mark.parametrize("case", [c for c in CASES if "tiny128" not in c][:6] + [c for c in CASES if "tiny128" in c][:1])
def test_hf_forward_matches_reference(case):
    import torch
    from oracle.hf_model import HFForward, build_model
    g = load(case)
    model = build_model(g["model"], seed=g["model_seed"], dtype=getattr(torch, g["dtype"]))
    fwd = HFForward(model)
    res = drive(g, fwd)
    check(g, res)
    assert fwd.n_forwards == len(g["forwards"])


@pytest.mark.parametrize("case", ["single_n8", "multi_n8_k2"])
def test_live_reference_matches_golden(case):
    from oracle import ref_shim
    if not ref_shim.reference_available():
        pytest.skip("reference tree not present (GPU box)")
    import torch
    from oracle import make_golden
    g = load(case)
    make_golden.GOLDEN_DIR = os.path.join("/tmp", "jf_golden_live")
    out = make_golden.run_case(case, g["model"], g["model_seed"], len(g["prompt"]), g["n"], g["max_new_tokens"],
                               g["mode"], K=g["K"], r=g["r"], lsr=g["lookahead_start_ratio"],
                               pool=g["n_gram_pool_size"], eos=g["eos"], pad=g["pad"],
                               max_iteration_count=g["max_iteration_count"], draft_seed=g["draft_seed"],
                               dtype=getattr(torch, g["dtype"]))
    assert out["generated"] == g["generated"] and out["iters"] == g["iters"]
    assert [f["rows"] for f in out["forwards"]] == [f["rows"] for f in g["forwards"]]


# ---------------------------------------------------------------------------------------------------------------------
# engine convention: goldens recorded from the reference JacobiDecoder itself (oracle/make_golden.py --engine)
# ---------------------------------------------------------------------------------------------------------------------
ENGINE_CASES = [os.path.splitext(os.path.basename(p))[0] for p in GOLDEN if json.load(open(p)).get("mode") == "engine"]


@pytest.mark.parametrize("case", ENGINE_CASES)
def test_engine_policy_replay_matches_reference(case):
    g = load(case)
    fwds = g["forwards"]
    state = {"i": 0}

    def forward_step(token_ids, draft):
        rec = fwds[state["i"]]
        assert rec["kv_before"] == len(token_ids) - 1
        assert rec["rows"][0] == list(draft), f"forward {state['i']}: draft differs from the reference"
        state["i"] += 1
        return rec["greedy"][0][:-1]

    pads = list(g["pad_stream"])

    def randint(k):
        out = pads[:k]
        del pads[:k]
        assert len(out) == k
        return out

    toks = list(g["prompt"])
    acc, iters, per_iter = jp.engine_fixed_len(forward_step, toks, g["prefill_draft"], g["L"], g["max_tokens"],
                                               g["max_iterations"], 0, randint, eos=g["eos"])
    assert acc == g["accepted"]
    assert iters == g["iters"] == len(fwds) == state["i"]
    assert toks == g["prompt"] + g["accepted"]
    assert not pads                                           # consumed the padding stream exactly
    if g["eos"] is None:
        assert g["equals_greedy_ar"] is True


def test_engine_goldens_exist():
    assert len(ENGINE_CASES) >= 3
