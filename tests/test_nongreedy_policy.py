"""CPU parity chain of the non-greedy Jacobi path (rejection-sampling verify):

  unmodified reference class JacobiDecoderNonGreedy  ==golden==>  oracle/jacobi_policy.py::engine_nongreedy
                                                     ==fuzz====>  csrc/jf_policy.h::eng_ng_advance (host build)

The goldens (tests/golden/ng_*.json, oracle/make_golden_nongreedy.py) hold every random draw the reference consumed, in
order, so the restatement is replayed decision by decision: same committed tokens, same next drafts, same iteration
count, same number of draws.  The device policy is then compared with the restatement on random inputs, both reading
their accept uniforms from the same counter-based generator (Philox4x32-10, pinned to the Random123 known answers)."""
import ctypes as C
import glob
import json
import os
import random

import numpy as np
import pytest

from oracle import jacobi_policy as jp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = sorted(glob.glob(os.path.join(ROOT, "tests", "golden", "ng_*.json")))


def test_philox_known_answers():
    """Random123 kat_vectors (philox4x32 10 rounds)."""
    assert jp.philox4x32_10(0, 0, 0, 0, 0, 0) == (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)
    f = 0xffffffff
    assert jp.philox4x32_10(f, f, f, f, f, f) == (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)
    assert jp.philox4x32_10(0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344, 0xa4093822, 0x299f31d0) == \
        (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1)


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[:-5] for p in GOLDEN])
def test_restatement_replays_the_reference_run(path):
    g = json.load(open(path))
    L, fw = g["L"], g["forwards"]
    rand, multi, pads = list(g["rand"]), list(g["multinomial"]), list(g["randint"])
    cur = {"f": -1, "u": 0, "m": 0, "p": 0}

    def forward_step(token_ids, draft):
        cur["f"] += 1
        rec = fw[cur["f"]]
        assert rec["kv_before"] == len(token_ids) - 1
        assert rec["draft"] == list(draft), f"draft of forward {cur['f']} differs"
        assert rec["draws_before"] == [cur["u"], cur["m"]], "the reference had consumed a different number of draws"

        def residual(t):
            def draw():
                cur["m"] += 1
                return multi[cur["m"] - 1]
            return jp.sample_not_equal(draw, draft[t + 1], fallback=lambda: pytest.fail("fallback not recorded"))
        return rec["greedy"], rec["px"], residual

    def uniform(it, t):
        cur["u"] += 1
        return rand[cur["u"] - 1]

    def randint(k):
        out = pads[cur["p"]:cur["p"] + k]
        cur["p"] += k
        assert len(out) == k
        return out

    toks = list(g["prompt"])
    accepted, iters, per_iter, drafts = jp.engine_nongreedy(forward_step, toks, L, g["max_tokens"], g["max_iterations"],
                                                            randint, uniform, eos=g["eos"])
    assert accepted == g["accepted"]
    assert iters == g["iters"] == len(fw)
    assert (cur["u"], cur["m"], cur["p"]) == (len(rand), len(multi), len(pads)), "every recorded draw is consumed"
    assert toks == g["prompt"] + g["accepted"]


def _harness():
    import __graft_entry__ as ge
    ge.build()
    lib = C.CDLL(os.path.join(ROOT, "build", "libjf_policy_host.so"))
    lib.jf_ng_accept_uniform.restype = C.c_float
    return lib


def test_device_uniforms_match_the_oracle_generator():
    from jacobiforcing_b200.policy import MODES, PolicyKnobs
    lib = _harness()
    kn = PolicyKnobs(mode=MODES["engine_nongreedy"], n=8, temperature=0.8, seed=0x1234567890abcdef, rng_offset=7, vocab=1000)
    ints = (C.c_int32 * 16)(*[x if x < 2 ** 31 else x - 2 ** 32 for x in kn.as_ints()])
    assert lib.jf_policy_knob_ints() == 16
    for row, fwd in [(0, 0), (5, 3), (31, 511), (1000, 12345)]:
        got = lib.jf_ng_accept_uniform(ints, row, fwd)
        assert got == np.float32(jp.device_accept_uniform(kn.seed, kn.rng_offset, row, fwd))
    out = (C.c_uint32 * 4)()
    lib.jf_philox(1, 2, 3, 4, 5, 6, out)
    assert tuple(out) == jp.philox4x32_10(1, 2, 3, 4, 5, 6)


@pytest.mark.parametrize("seed", range(40))
def test_device_policy_equals_restatement_on_random_inputs(seed):
    """eng_ng_advance (the code the persistent kernel runs) against engine_nongreedy on synthetic forwards: random
    argmax rows, acceptance probabilities biased towards long accepted runs, random residual samples, EOS sometimes."""
    from jacobiforcing_b200.policy import MODES, PolicyKnobs
    lib = _harness()
    rng = random.Random(seed)
    L = rng.choice([2, 3, 8, 16, 32])
    V = 50
    max_tokens = rng.choice([5, 17, 40])
    max_iters = rng.choice([3, 128])
    eos = rng.choice([None, None, 7])
    kseed = rng.getrandbits(64)
    off = rng.randrange(0, 1000)
    nfw = 80
    px = np.array([[rng.choice([0.0, 0.3, 0.9, 0.999, 1.0]) for _ in range(L - 1)] for _ in range(nfw)], np.float32)
    bonus = np.array([[rng.randrange(V) for _ in range(L - 1)] for _ in range(nfw)], np.int32)
    greedy = np.array([[rng.randrange(V) for _ in range(L - 1)] for _ in range(nfw)], np.int32)
    pads = [rng.randrange(V) for _ in range(nfw * L + L)]
    prompt = [rng.randrange(V) for _ in range(rng.randrange(1, 6))]
    # restatement
    cur = {"f": -1, "p": 0}

    def forward_step(token_ids, draft):
        cur["f"] += 1
        f = cur["f"]
        return greedy[f].tolist(), px[f].tolist(), (lambda t: int(bonus[f][t]))

    def randint(k):
        out = pads[cur["p"]:cur["p"] + k]
        cur["p"] += k
        return out
    toks = list(prompt)
    accepted, iters, per_iter, drafts = jp.engine_nongreedy(
        forward_step, toks, L, max_tokens, max_iters, randint,
        lambda it, t: jp.device_accept_uniform(kseed, off, t, it), eos=eos)
    # device policy, host build
    kn = PolicyKnobs(mode=MODES["engine_nongreedy"], n=L, eos=eos, pad=0, max_iteration_count=max_iters,
                     max_new_tokens=max_tokens, vocab=V, temperature=1.0, seed=kseed, rng_offset=off)
    ints = (C.c_int32 * 16)(*[x if x < 2 ** 31 else x - 2 ** 32 for x in kn.as_ints()])
    cap = len(prompt) + max_tokens + 2 * L + 8
    gen = (C.c_int32 * cap)()
    rows = (C.c_int32 * (nfw * L))()
    ctxs = (C.c_int32 * nfw)()
    res = (C.c_int32 * 8)()
    n = lib.jf_policy_replay_ng(ints, (C.c_int32 * len(prompt))(*prompt), len(prompt), len(prompt) - 1,
                                (C.c_int32 * len(pads))(*pads), len(pads), nfw,
                                px.ctypes.data_as(C.c_void_p), bonus.ctypes.data_as(C.c_void_p),
                                greedy.ctypes.data_as(C.c_void_p), rows, ctxs, gen, cap, res)
    assert n == iters == res[6], (n, iters, res[6])
    assert list(gen[:res[0]]) == toks
    assert [list(rows[f * L:(f + 1) * L]) for f in range(n)] == drafts
    # committed KV before every forward = tokens - 1 (the seed is recomputed)
    lens = [len(prompt)]
    for k in per_iter[:-1]:
        lens.append(lens[-1] + k)
    assert list(ctxs[:n]) == [x - 1 for x in lens[:n]]
    stop = {1: "eos", 2: "max_new_tokens", 4: "max_iterations"}[res[2]]
    if eos is not None and eos in accepted:
        assert stop == "eos"
    elif len(accepted) >= max_tokens:
        assert stop == "max_new_tokens"
    else:
        assert stop == "max_iterations" and iters == max_iters
