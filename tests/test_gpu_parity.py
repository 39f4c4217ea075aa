"""GPU parity tests (-m gpu): the CUDA path, called through the C ABI, against the CPU oracle.

Floating point: the device computes in bf16 with fp32 accumulation at the HF modules' rounding points.  Tolerance on
logits: max |d| <= 2.5e-2 and mean |d| <= 4e-3 against the fp32 HF forward with the same bf16-rounded weights (logits of
magnitude ~1-4, where one bf16 ulp is 0.0078-0.0156: the reference's own bf16 logits are quantised coarser than the
1e-3 the north star quotes, so 1e-3 is only meaningful for fp16/fp32 activations).  The test also requires the device
to be no further from the fp32 oracle than 1.5x the reference-dtype (HF bf16 on CPU) forward is.
Integer work (accepted ids, iteration counts, forwards issued) is bit-exact against the oracle policy driven by the
device forward; against the HF-forward oracle, a token mismatch is accepted only where the oracle's top-2 margin is a
near-tie (hard part 1).
"""
import random

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

LOGIT_MAX_TOL = 2.5e-2
LOGIT_MEAN_TOL = 4e-3
NEAR_TIE = 3e-2


@pytest.fixture(scope="module", params=["tiny", "tiny128"])
def pair(request):
    from tests.gpu_helpers import make_pair
    # rows: n=64 multiblock keeps up to (1 + pool) candidate rows of 2n columns in flight
    model, rt = make_pair(request.param, max_rows=1024)
    yield request.param, model, rt
    rt.close()


def _prompt(model, P, seed=1):
    g = torch.Generator().manual_seed(seed)
    return torch.randint(0, model.config.vocab_size, (P,), generator=g).tolist()


@pytest.mark.parametrize("rows", [72, 200, 256])      # 200 / 256 rows: the wide (whole-TMEM) SwiGLU accumulator mode
def test_forward_logits_within_tolerance(pair, rows):
    from oracle.hf_model import bf16_copy, teacher_forced_logits
    name, model, rt = pair
    ids = _prompt(model, rows)
    seq = rt.new_sequence()
    rt.ensure_capacity(seq, len(ids))
    res = rt.forward(ids, [(0, len(ids), 0, 0)], [seq], 0, len(ids), want_logits=True)
    ref = teacher_forced_logits(model, ids).numpy()
    err = np.abs(res["logits"] - ref)
    assert err.max() <= LOGIT_MAX_TOL, f"max |dlogit| {err.max()}"
    assert err.mean() <= LOGIT_MEAN_TOL, f"mean |dlogit| {err.mean()}"
    # the reference's own dtype (bf16 HF forward) is no closer to fp32 than we are
    ref16 = teacher_forced_logits(bf16_copy(model), ids).numpy()     # a copy: .to() would round inv_freq in place
    err16 = np.abs(ref16 - ref)
    assert err.mean() <= 1.5 * err16.mean() + 1e-4
    # fused argmax == argmax of the bf16-rounded device logits, lowest index on ties
    dev16 = torch.from_numpy(res["logits"]).to(torch.bfloat16).float()
    assert (res["greedy"] == dev16.argmax(-1).numpy()).all()
    rt.release(seq)


def test_committed_kv_equals_inflight_kv_bitwise(pair):
    """Rows verified against committed pages give the same logits, bit for bit, as the same rows verified in flight
    (attention chunks are aligned to absolute key positions) - the property that makes Jacobi == AR on this engine."""
    name, model, rt = pair
    ids = _prompt(model, 80)
    a = rt.new_sequence(); rt.ensure_capacity(a, 80)
    full = rt.forward(ids, [(0, 80, 0, 0)], [a], 0, 80, want_logits=True)["logits"]
    b = rt.new_sequence(); rt.ensure_capacity(b, 80)
    rt.forward(ids[:56], [(0, 56, 0, 0)], [b], 0, 0, commits=[(i, 0, i) for i in range(56)])
    tail = rt.forward(ids[56:], [(0, 24, 0, 56)], [b], 0, 24, want_logits=True)["logits"]
    assert np.array_equal(full[56:], tail)
    rt.release(a); rt.release(b)


def test_prefill_chunked_equals_unchunked(pair):
    from tests.gpu_helpers import make_pair
    name, model, rt = pair
    prompt = _prompt(model, 100)
    draft = [prompt[i] for i in range(8)]
    s = rt.new_sequence()
    fc, drafted, _ = rt.prefill(s, prompt, draft)
    _, rt_small = make_pair(name, max_rows=48)
    s2 = rt_small.new_sequence()
    fc2, drafted2, _ = rt_small.prefill(s2, prompt, draft)
    assert (fc, drafted) == (fc2, drafted2)
    rt.release(s); rt_small.close()


CASES = [("single", 8, 40, dict()), ("multiblock", 8, 40, dict(K=2, r=0.85, pool=4)),
         ("multiblock", 16, 64, dict(K=2, r=0.5, pool=8)), ("multiblock", 8, 32, dict(K=1, r=0.85, pool=1)),
         ("multiblock", 8, 40, dict(K=2, r=0.85, pool=4, lsr=0.5)),
         # the headline block size (BASELINE configs 2-4) and the README-recommended n=64 (config 5)
         ("single", 32, 96, dict()), ("multiblock", 32, 96, dict(K=2, r=0.85, pool=4)),
         ("multiblock", 64, 160, dict(K=2, r=0.85, pool=4)), ("multiblock", 32, 64, dict(K=2, r=0.5, pool=8))]


def _oracle_on_device(rt, prompt, mode, n, max_new, K=2, r=0.85, pool=4, lsr=0.0, eos=None, seed=2):
    """oracle/jacobi_policy.py (pinned against the reference) consuming the DEVICE forward's argmaxes."""
    from jacobiforcing_b200.policy import draw_prefill_draft
    from oracle import jacobi_policy as jp
    from tests.gpu_helpers import DeviceForward
    rng = random.Random(seed)
    seq = rt.new_sequence()
    fc, drafted, _ = rt.prefill(seq, prompt, draw_prefill_draft(rng, prompt, n))
    fwd = DeviceForward(rt, seq)
    trace = jp.Trace()
    generated, iters, calls, total_new, first = list(prompt), [0], 1, 0, fc
    stop = ""
    while True:
        if eos is not None and eos in generated[len(prompt):]:
            stop = "eos"; break
        if total_new >= max_new:
            stop = "max_new_tokens"; break
        inp = list(drafted) if calls == 1 else [first] + [rng.choice(generated) for _ in range(n - 1)]
        if mode == "multiblock":
            first, acc, itr = jp.hf_multiblock(fwd, inp, n, K=K, r=r, lookahead_start_ratio=lsr,
                                               n_gram_pool_size=pool, eos=eos, pad=0, trace=trace)
        else:
            first, acc, itr = jp.hf_single_block(fwd, inp, n, eos=eos, trace=trace)
        generated += acc; total_new += len(acc); iters.append(itr); calls += 1
    rt.release(seq)
    return generated[len(prompt):], iters, stop, trace


@pytest.mark.parametrize("mode,n,max_new,kw", CASES)
def test_device_loop_equals_oracle_policy_bit_exact(pair, mode, n, max_new, kw):
    from jacobiforcing_b200.runtime import hf_generate
    name, model, rt = pair
    prompt = _prompt(model, 24)
    seqs, results, run = hf_generate(rt, [prompt], n=n, max_new_tokens=max_new, mode=mode, K=kw.get("K", 2),
                                     r=kw.get("r", 0.85), n_gram_pool_size=kw.get("pool", 4),
                                     lookahead_start_ratio=kw.get("lsr", 0.0), pad=0, rngs=[random.Random(2)],
                                     want_trace=True)
    toks, iters, stop, trace = _oracle_on_device(rt, prompt, mode, n, max_new, **kw)
    res = results[0]
    assert res.tokens == toks
    assert res.iters == iters
    assert res.stop_reason == stop
    assert [t["rows"] for t in run["trace"]] == trace.rows[0:]
    assert [t["kv_before"] for t in run["trace"]] == trace.kv_before
    assert run["forwards"] == sum(iters)
    for s in seqs:
        rt.release(s)


def test_batch8_multiblock_equals_oracle_policy_bit_exact():
    """BASELINE config 3 shape of work (B=8 multiblock + rejection recycling, n=32): every sequence of the batched
    device-resident loop equals the oracle policy run on that sequence alone (ids, per-call iterations, stop reason)."""
    from jacobiforcing_b200.runtime import hf_generate
    from tests.gpu_helpers import make_pair
    model, rt = make_pair("tiny128", max_rows=2560, max_seqs=8, num_kv_pages=32, max_pages_per_seq=4)
    prompts = [_prompt(model, 20 + 5 * i, seed=30 + i) for i in range(8)]
    mk = lambda: [random.Random(200 + i) for i in range(8)]
    kw = dict(K=2, r=0.85, pool=4)
    seqs, batch, run = hf_generate(rt, prompts, n=32, max_new_tokens=64, mode="multiblock", K=2, r=0.85,
                                   n_gram_pool_size=4, pad=0, rngs=mk())
    for s in seqs:
        rt.release(s)
    assert run["rows"] / run["forwards"] > 32 * 4          # really a batched forward
    for i in range(8):
        toks, iters, stop, _ = _oracle_on_device(rt, prompts[i], "multiblock", 32, 64, seed=200 + i, **kw)
        assert batch[i].tokens == toks, f"sequence {i}"
        assert batch[i].iters == iters and batch[i].stop_reason == stop
    rt.close()


def test_finalize_rejects_partial_checkpoint():
    """a checkpoint that lacks a layer tensor must not finalize (zero weights would decode garbage)."""
    from jacobiforcing_b200.runtime import JacobiRuntime
    from oracle.hf_model import build_model, state_dict_bf16
    model = build_model("tiny", seed=0, dtype=torch.float32)
    sd = state_dict_bf16(model)
    del sd["model.layers.1.mlp.up_proj.weight"]
    rt = JacobiRuntime(model.config, max_rows=64, max_seqs=1, num_kv_pages=2, max_pages_per_seq=2)
    with pytest.raises(ValueError, match="model.layers.1.mlp.up_proj.weight"):
        rt.load_state_dict(sd)
    rt.close()


def test_eos_paths_bit_exact(pair):
    from jacobiforcing_b200.runtime import hf_generate
    name, model, rt = pair
    prompt = _prompt(model, 24)
    base, _, _, _ = _oracle_on_device(rt, prompt, "multiblock", 8, 40)
    eos = next((t for i, t in enumerate(base) if i >= 9 and t not in base[:i] and t not in prompt), None)
    if eos is None:
        pytest.skip("no late-first-occurrence token")
    for mode in ("single", "multiblock"):
        seqs, results, run = hf_generate(rt, [prompt], n=8, max_new_tokens=40, mode=mode, eos=eos, pad=0,
                                         rngs=[random.Random(2)])
        toks, iters, stop, _ = _oracle_on_device(rt, prompt, mode, 8, 40, eos=eos)
        assert results[0].tokens == toks and results[0].iters == iters
        assert results[0].stop_reason == stop == "eos"
        assert results[0].tokens[-1] == eos
        for s in seqs:
            rt.release(s)


@pytest.mark.parametrize("mode,n", [("single", 8), ("multiblock", 16)])
def test_tokens_equal_hf_oracle_up_to_near_ties(pair, mode, n):
    """End to end against the oracle proper (HF fp32 forward + pinned policy): ids equal; a divergence is tolerated only
    at a position where the oracle's own top-2 margin is below NEAR_TIE."""
    from jacobiforcing_b200.runtime import hf_generate
    from oracle import jacobi_policy as jp
    from oracle.hf_model import HFForward, teacher_forced_logits
    name, model, rt = pair
    prompt = _prompt(model, 24, seed=3)
    seqs, results, _ = hf_generate(rt, [prompt], n=n, max_new_tokens=48, mode=mode, pad=0, rngs=[random.Random(2)])
    ref = jp.hf_driver(HFForward(model), prompt, n, 48, random.Random(2), mode=mode, pad=0)
    got = results[0].tokens
    assert len(got) == len(ref.generated) == 48
    same = next((i for i, (a, b) in enumerate(zip(got, ref.generated)) if a != b), len(got))
    if same < len(got):
        lg = teacher_forced_logits(model, prompt + ref.generated[:same])[-1]
        top2 = lg.topk(2).values
        margin = float(top2[0] - top2[1])
        assert margin < NEAR_TIE, f"diverged at {same} although the oracle margin is {margin}"
        assert got[same] in lg.topk(2).indices.tolist()
    else:
        # same ids; iteration counts may differ by a little because REJECTED draft tokens (which never reach the output
        # but steer the next draft) also sit on near-ties.  Exact iteration parity is pinned by
        # test_device_loop_equals_oracle_policy_bit_exact.
        assert abs(results[0].total_iterations - ref.total_iterations) <= max(2, ref.total_iterations // 8)
    for s in seqs:
        rt.release(s)


def test_jacobi_equals_autoregressive_bitwise(pair):
    """Fixed-point property the reference's own tests pin (Jacobi tokens == greedy AR tokens, first 200,
    inference_engine/tests/test_jacobi_decoding_greedy.py:180-206) - here exact, because a row's logits do not depend
    on how many other rows are in flight."""
    from jacobiforcing_b200.runtime import hf_generate
    name, model, rt = pair
    prompt = _prompt(model, 40, seed=5)
    outs = []
    for mode, n in (("single", 1), ("single", 8), ("multiblock", 16)):
        seqs, results, run = hf_generate(rt, [prompt], n=n, max_new_tokens=48, mode=mode, pad=0,
                                         rngs=[random.Random(2)])
        outs.append(results[0].tokens[:48])
        if n == 1:
            assert run["forwards"] == 48
        for s in seqs:
            rt.release(s)
    assert outs[0] == outs[1] == outs[2]


def test_batch_rows_are_independent(pair):
    from jacobiforcing_b200.runtime import hf_generate
    name, model, rt = pair
    prompts = [_prompt(model, 20 + 7 * i, seed=10 + i) for i in range(3)]
    rngs = lambda: [random.Random(100 + i) for i in range(3)]
    seqs, batch, _ = hf_generate(rt, prompts, n=8, max_new_tokens=32, mode="multiblock", pad=0, rngs=rngs())
    for s in seqs:
        rt.release(s)
    for i in range(3):
        seqs, solo, _ = hf_generate(rt, [prompts[i]], n=8, max_new_tokens=32, mode="multiblock", pad=0,
                                    rngs=[rngs()[i]])
        assert solo[0].tokens == batch[i].tokens and solo[0].iters == batch[i].iters
        for s in seqs:
            rt.release(s)


def test_error_mapping(pair):
    from jacobiforcing_b200.policy import MODES, PolicyKnobs
    from jacobiforcing_b200.runtime import hf_generate
    name, model, rt = pair
    prompt = _prompt(model, 24)
    with pytest.raises(NotImplementedError):
        hf_generate(rt, [prompt], n=8, max_new_tokens=16, mode="multiblock", K=3, pad=0)
    free = len(rt.free_pages)
    with pytest.raises(ValueError, match="pad_token_id"):        # unified.py:631-632
        hf_generate(rt, [prompt], n=8, max_new_tokens=16, mode="multiblock", K=2, r=0.2, pad=None)
    rt.free_pages = list(range(rt.runtime_cfg.num_kv_pages - 1, -1, -1))
    with pytest.raises(RuntimeError):
        rt.forward([1] * 1100, [(0, 1100, 0, 0)], [rt.new_sequence()], 0, 8)    # more rows than max_rows


# ---------------------------------------------------------------------------------------------------------------
# fused GEMM tails (QKV + rotary, O / down + reduce + RMSNorm, tagged split-K partial sums) == stand-alone stages
# ---------------------------------------------------------------------------------------------------------------
def _run_mid(monkeypatch, fused: bool, rows_list, gen_kw):
    import random
    from tests.gpu_helpers import make_pair
    from jacobiforcing_b200.runtime import hf_generate
    for k in ("JF_FUSE_NORM", "JF_FUSE_ROPE"):
        monkeypatch.setenv(k, "1" if fused else "0")          # read once, when the engine is created
    # (mid128 also runs the split-K gate/up projection - 4 k-splits, SwiGLU in the tail - in BOTH engines: that is a
    # property of the engine, not of the fusion switches)
    model, rt = make_pair("mid128")
    out = {"logits": [], "greedy": []}
    for rows in rows_list:
        ids = _prompt(model, rows, seed=3)
        seq = rt.new_sequence()
        rt.ensure_capacity(seq, len(ids))
        res = rt.forward(ids, [(0, len(ids), 0, 0)], [seq], 0, len(ids), want_logits=True)
        out["logits"].append(res["logits"].copy())
        out["greedy"].append(np.asarray(res["greedy"]).copy())
        rt.release(seq)
    prompt = _prompt(model, 40, seed=5)
    seqs, results, run = hf_generate(rt, [prompt], rngs=[random.Random(7)], pad=0, **gen_kw)
    out["tokens"] = list(results[0].tokens)
    out["forwards"] = run["forwards"]
    for s_ in seqs:
        rt.release(s_)
    rt.close()
    return model, out


@pytest.mark.parametrize("gen_kw", [dict(n=16, max_new_tokens=48, mode="single"),
                                    dict(n=16, max_new_tokens=48, mode="multiblock", K=2, r=0.85, n_gram_pool_size=4)])
def test_fused_tails_equal_standalone_stages_bitwise(monkeypatch, gen_kw):
    from oracle.hf_model import teacher_forced_logits
    rows_list = [1, 7, 24, 64, 100, 256]          # 1..64: rotary + norm tails; 100, 256: norm tails only; all multi-split
    model, a = _run_mid(monkeypatch, True, rows_list, gen_kw)
    _, b = _run_mid(monkeypatch, False, rows_list, gen_kw)
    for rows, la, lb, ga, gb in zip(rows_list, a["logits"], b["logits"], a["greedy"], b["greedy"]):
        assert np.array_equal(la.view(np.uint32), lb.view(np.uint32)), f"fused != stand-alone logits at {rows} rows"
        assert np.array_equal(ga, gb)
    assert a["tokens"] == b["tokens"] and a["forwards"] == b["forwards"]
    # and both sit within the usual tolerance of the HF oracle
    ids = _prompt(model, 64, seed=3)
    ref = teacher_forced_logits(model, ids).numpy()
    err = np.abs(a["logits"][3] - ref)
    from oracle.hf_model import bf16_copy
    err16 = np.abs(teacher_forced_logits(bf16_copy(model), ids).numpy() - ref)   # the reference's own dtype against fp32
    # hidden 1024: logits reach |6|, one bf16 ulp there is 0.03; tolerance relative to what the HF bf16 forward achieves
    assert err.mean() <= 1.5 * err16.mean() + 1e-4
    assert err.max() <= max(LOGIT_MAX_TOL, 1.5 * err16.max())


# ---------------------------------------------------------------------------------------------------------------
# epilogues through shared memory + bulk stores, fence placement: the same bits as the direct-store / textbook variants
# ---------------------------------------------------------------------------------------------------------------
def _run_switches(monkeypatch, shape, env, rows_list, gen_kw):
    import random
    from tests.gpu_helpers import make_pair
    from jacobiforcing_b200.runtime import hf_generate
    for k, v in env.items():
        monkeypatch.setenv(k, v)                              # read once, when the engine is created
    model, rt = make_pair(shape)
    out = {"logits": []}
    for rows in rows_list:
        ids = _prompt(model, rows, seed=11)
        seq = rt.new_sequence()
        rt.ensure_capacity(seq, len(ids))
        res = rt.forward(ids, [(0, len(ids), 0, 0)], [seq], 0, len(ids), want_logits=True)
        out["logits"].append(res["logits"].copy())
        rt.release(seq)
    prompt = _prompt(model, 40, seed=5)
    seqs, results, run = hf_generate(rt, [prompt], rngs=[random.Random(7)], pad=0, **gen_kw)
    out["tokens"], out["iters"], out["forwards"] = list(results[0].tokens), list(results[0].iters), run["forwards"]
    for s_ in seqs:
        rt.release(s_)
    rt.close()
    return out


@pytest.mark.parametrize("shape,common", [("tiny128", {}), ("mid128", {}), ("mid128", {"JF_GU_SPLITS": "1"})])
def test_bulk_epilogues_and_fence_placement_do_not_change_a_bit(monkeypatch, shape, common):
    """Defaults (SwiGLU tile and O / down partial sums leave through shared memory + bulk copies, acquire fence behind the
    barrier arrival, fence-less activation producer) against JF_SWIGLU_BULK=0 JF_EPI_BULK=0 JF_EARLY_FENCE=0 JF_X_FENCE=1:
    logits of 1 .. 100-row forwards (<= 32 rows: bulk row stores; <= 128: bulk SwiGLU tile) and a generation, bit for bit.
    JF_GU_SPLITS=1 forces the single-accumulator-pair gate/up projection (the 7B layout: SwiGLU in the epilogue, gate/up -> down
    hand-over by arrival counters) on the small shape, whose default is the split-K gate/up + SwiGLU tail."""
    rows_list = [1, 7, 16, 24, 32, 40, 64, 100]
    gen_kw = dict(n=16, max_new_tokens=48, mode="multiblock", K=2, r=0.85, n_gram_pool_size=4)
    a = _run_switches(monkeypatch, shape, dict(common, JF_SWIGLU_BULK="1", JF_EPI_BULK="1", JF_EARLY_FENCE="1", JF_X_FENCE="0"),
                      rows_list, gen_kw)
    b = _run_switches(monkeypatch, shape, dict(common, JF_SWIGLU_BULK="0", JF_EPI_BULK="0", JF_EARLY_FENCE="0", JF_X_FENCE="1"),
                      rows_list, gen_kw)
    for rows, la, lb in zip(rows_list, a["logits"], b["logits"]):
        assert np.array_equal(la.view(np.uint32), lb.view(np.uint32)), f"logits differ at {rows} rows"
    assert a["tokens"] == b["tokens"] and a["iters"] == b["iters"] and a["forwards"] == b["forwards"]
