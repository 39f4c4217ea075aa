"""The PRODUCT policy code (jacobiforcing_b200/csrc/jf_policy.h - the header the persistent kernel compiles) built for
the host and replayed against the golden forwards recorded from the unmodified reference (tests/golden/*.json):
it must issue the same forwards (kv length + `out` rows) and return the same ids / iteration counts, bit-exact."""
import ctypes
import glob
import json
import os
import random
import subprocess

import numpy as np
import pytest

from jacobiforcing_b200.policy import MODES, PolicyKnobs, STOP_REASONS, mt_state_words

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = sorted(glob.glob(os.path.join(ROOT, "tests", "golden", "*.json")))


def _load_harness():
    so = os.path.join(ROOT, "build", "libjf_policy_host.so")
    src = os.path.join(ROOT, "tests", "host", "policy_harness.cpp")
    hdr = os.path.join(ROOT, "jacobiforcing_b200", "csrc", "jf_policy.h")
    if not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        os.makedirs(os.path.dirname(so), exist_ok=True)
        subprocess.check_call(["g++", "-O2", "-shared", "-fPIC", "-I", os.path.join(ROOT, "jacobiforcing_b200", "csrc"),
                               src, "-o", so])
    return ctypes.CDLL(so)


def _arr(x, dtype=np.int32):
    return np.ascontiguousarray(np.asarray(x, dtype=dtype))


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def replay(g, lib):
    kn = PolicyKnobs(mode=MODES[g["mode"]], n=g["n"], K=g["K"], r=g["r"],
                     lookahead_start_ratio=g["lookahead_start_ratio"], n_gram_pool_size=g["n_gram_pool_size"],
                     eos=g["eos"], pad=g["pad"], max_iteration_count=g["max_iteration_count"],
                     max_new_tokens=g["max_new_tokens"], max_calls=1024)
    assert lib.jf_policy_knob_ints() == len(kn.as_ints())
    prompt = g["prompt"]
    rng = random.Random(g["draft_seed"])
    _ = [rng.choice(prompt) for _ in range(g["n"])]              # the prefill draft draws come first
    prefill = g["call_log"][0]
    n_calls = g["calls"] + 2
    mt = _arr(mt_state_words(rng), np.uint32)
    fws = g["forwards"][1:]                                      # forward 0 is the prefill
    kv = _arr([f["kv_before"] for f in fws]); fB = _arr([len(f["rows"]) for f in fws])
    fL = _arr([len(f["rows"][0]) for f in fws])
    offs, rows, greedy, o = [], [], [], 0
    for f in fws:
        offs.append(o)
        for r_, g_ in zip(f["rows"], f["greedy"]):
            rows.extend(r_); greedy.extend(g_)
        o += len(f["rows"]) * len(f["rows"][0])
    offs = _arr(offs, np.int64); rows = _arr(rows); greedy = _arr(greedy)
    gen_cap = len(prompt) + g["max_new_tokens"] + 4 * g["n"]
    gen = np.zeros(gen_cap, np.int32); iters = np.zeros(n_calls + 4, np.int32); res = np.zeros(8, np.int32)
    knobs = _arr(kn.as_ints()); pr = _arr(prompt); fct = _arr(prefill["prefill_drafted"])
    rc = lib.jf_policy_replay(_p(knobs), _p(pr), len(prompt), prefill["kv_after"], prefill["first_correct"], _p(fct),
                              _p(mt), 625, None, 0, len(fws), _p(kv), _p(fB), _p(fL), _p(offs), _p(rows),
                              _p(greedy), _p(gen), gen_cap, _p(iters), len(iters), _p(res))
    return rc, gen, iters, res, len(fws)


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[:-5] for p in GOLDEN])
def test_device_policy_replays_reference(path):
    g = json.load(open(path))
    if g.get("mode") not in ("single", "multiblock"):
        pytest.skip("not an HF-convention case")
    lib = _load_harness()
    if g["K"] > 2:
        with pytest.raises(NotImplementedError):
            PolicyKnobs(mode=MODES[g["mode"]], n=g["n"], K=g["K"]).validate()
        return
    rc, gen, iters, res, nfw = replay(g, lib)
    assert rc == 0, f"policy diverged from the reference at forward {rc - 100}" if rc >= 100 else "ran out of forwards"
    gen_len, calls, stop, total_new, err, used, total_iters, ctx = [int(x) for x in res]
    assert err == 0
    P = len(g["prompt"])
    assert gen[P:gen_len].tolist() == g["generated"]
    assert calls == g["calls"]
    assert iters[:calls].tolist() == g["iters"]
    assert STOP_REASONS[stop] == g["stop_reason"]
    assert total_new == g["total_new_tokens"]
    assert used == nfw
    assert total_iters == sum(g["iters"])


ENGINE_GOLDEN = [p for p in GOLDEN if json.load(open(p)).get("mode") == "engine"]


@pytest.mark.parametrize("path", ENGINE_GOLDEN, ids=[os.path.basename(p)[:-5] for p in ENGINE_GOLDEN])
def test_device_engine_policy_replays_reference(path):
    """csrc/jf_policy.h eng_advance (mode 2) against forwards recorded from the reference JacobiDecoder."""
    g = json.load(open(path))
    lib = _load_harness()
    kn = PolicyKnobs(mode=MODES["engine"], n=g["L"], eos=g["eos"], pad=0, max_iteration_count=g["max_iterations"],
                     max_new_tokens=g["max_tokens"])
    prompt = g["prompt"]
    fws = g["forwards"]
    kv = _arr([f["kv_before"] for f in fws]); fB = _arr([1] * len(fws)); fL = _arr([g["L"]] * len(fws))
    offs = _arr([i * g["L"] for i in range(len(fws))], np.int64)
    rows = _arr([t for f in fws for t in f["rows"][0]]); greedy = _arr([t for f in fws for t in f["greedy"][0]])
    gen_cap = len(prompt) + g["max_tokens"] + 4 * g["L"]
    gen = np.zeros(gen_cap, np.int32); iters = np.zeros(8, np.int32); res = np.zeros(8, np.int32)
    knobs = _arr(kn.as_ints()); pr = _arr(prompt)
    fct = _arr(g["prefill_draft"][:g["L"] - 1] + [0]) if g["prefill_draft"] is not None else None
    pads = _arr(g["pad_stream"] + [0])
    rc = lib.jf_policy_replay(_p(knobs), _p(pr), len(prompt), len(prompt) - 1, 0, _p(fct) if fct is not None else None,
                              None, 0, _p(pads), len(g["pad_stream"]), len(fws), _p(kv), _p(fB), _p(fL), _p(offs),
                              _p(rows), _p(greedy), _p(gen), gen_cap, _p(iters), len(iters), _p(res))
    assert rc == 0, f"policy diverged from the reference at forward {rc - 100}" if rc >= 100 else "ran out of forwards"
    gen_len, calls, stop, total_new, err, used, total_iters, ctx = [int(x) for x in res]
    assert err == 0
    assert gen[len(prompt):gen_len].tolist() == g["accepted"]
    assert used == len(fws) == g["iters"] == total_iters
    assert total_new == len(g["accepted"])
    assert ctx == gen_len - 1
    expected_stop = "eos" if g["eos"] is not None else ("max_iterations" if g["iters"] >= g["max_iterations"] and len(g["accepted"]) < g["max_tokens"] else "max_new_tokens")
    assert STOP_REASONS[stop] == expected_stop
