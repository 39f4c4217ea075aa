"""Non-greedy Jacobi decoding on the device (-m gpu): the rejection-sampling verify of the reference's
JacobiDecoderNonGreedy (inference_engine/engine/jacobi_decoding_nongreedy.py:300-355, loop :385-470) inside the persistent
kernel - LM-head epilogue EPI_SAMPLE (softmax normaliser, drafted-token probability and a Gumbel-max residual sample in one
pass, no logits in HBM) + policy eng_ng_advance.

 1. Decision-level parity: the device-resident loop against the pinned restatement (oracle/jacobi_policy.py::
    engine_nongreedy, itself pinned to the unmodified reference class by tests/test_nongreedy_policy.py) driven by the
    DEVICE forward's logits and the same counter-based random numbers: same committed tokens, drafts and iteration count.
 2. Distribution: tokens committed by the verify follow softmax(logits / T) - the property the reference's own test checks
    (tests/test_jacobi_decoding_nongreedy.py:324-442, JS divergence against autoregressive sampling).
 3. The engine surface: LLM.generate with temperature > 0 and decode_strategy="jacobi"."""
import random

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def pair():
    from tests.gpu_helpers import make_pair
    model, rt = make_pair("tiny")
    yield model, rt
    rt.close()


def _bf16(a: np.ndarray) -> np.ndarray:
    return torch.from_numpy(np.ascontiguousarray(a)).to(torch.bfloat16).float().numpy()


def _run_device(rt, prompt, L, temperature, seed, max_new, max_iters=128, eos=None, pads=None, rng_offset=0):
    from jacobiforcing_b200.policy import MODES, PolicyKnobs
    seq = rt.new_sequence()
    rt.prefill(seq, prompt, [])
    seq.ctx = len(prompt) - 1                              # engine convention: the seed (last token) is recomputed
    knobs = PolicyKnobs(mode=MODES["engine_nongreedy"], n=L, eos=eos, pad=0, max_iteration_count=max_iters,
                        max_new_tokens=max_new, vocab=rt.vocab, temperature=temperature, seed=seed, rng_offset=rng_offset)
    results, run = rt.jacobi_run([seq], knobs, [0], None, pad_stream=[pads], want_trace=True)
    rt.release(seq)
    return results[0], run


@pytest.mark.parametrize("L,temperature,seed", [(8, 1.0, 11), (8, 0.05, 12), (16, 0.3, 13), (2, 0.7, 14)])
def test_device_loop_equals_restatement_on_device_logits(pair, L, temperature, seed):
    from oracle import jacobi_policy as jp
    from tests.gpu_helpers import DeviceEngineForward
    model, rt = pair
    V = model.config.vocab_size
    g = torch.Generator().manual_seed(seed)
    prompt = torch.randint(0, V, (24,), generator=g).tolist()
    prng = random.Random(seed)
    pads = [prng.randrange(V) for _ in range(64 * L)]
    res, run = _run_device(rt, prompt, L, temperature, seed * 7919 + 1, 40, pads=pads, rng_offset=5)
    # the restatement on the device forward
    seq = rt.new_sequence()
    rt.prefill(seq, prompt, [])
    fwd = DeviceEngineForward(rt, seq, want_logits=True)
    cur = {"p": 0, "it": -1}
    kseed, off = seed * 7919 + 1, 5
    inv_t = np.float32(1.0) / np.float32(temperature)

    def forward_step(token_ids, draft):
        greedy = fwd(token_ids, draft)
        cur["it"] += 1
        sc = _bf16(fwd.last_logits[:L - 1]) * inv_t                       # logits / T on bf16-rounded logits
        m = sc.max(axis=1, keepdims=True)
        e = np.exp(sc - m)
        probs = e / e.sum(axis=1, keepdims=True)
        px = [float(probs[t, draft[t + 1]]) for t in range(L - 1)]

        def residual(t):
            u = jp.device_gumbel_uniforms(kseed, off, t, cur["it"], V)
            key = sc[t] + (-np.log(-np.log(u))).astype(np.float32)
            key[draft[t + 1]] = -np.inf
            return int(np.argmax(key))
        return greedy, px, residual

    def randint(k):
        out = pads[cur["p"]:cur["p"] + k]
        cur["p"] += k
        return out
    toks = list(prompt)
    accepted, iters, per_iter, drafts = jp.engine_nongreedy(
        forward_step, toks, L, 40, 128, randint, lambda it, t: jp.device_accept_uniform(kseed, off, t, it))
    rt.release(seq)
    assert res.tokens == accepted, "committed tokens differ"
    assert res.total_iterations == iters == run["forwards"]
    assert [t["rows"][0] for t in run["trace"]] == drafts
    assert [t["kv_before"] for t in run["trace"]] == [len(prompt) - 1 + sum(per_iter[:i]) for i in range(iters)]
    if temperature <= 0.05:
        assert iters < len(accepted), "a sharp distribution must accept drafted runs"


def test_committed_tokens_follow_the_target_distribution(pair):
    """2000 one-token runs (L = 2: one drafted token, accepted with probability p(x), else a residual sample) from the
    same prefix: the empirical distribution of the committed token against softmax(logits / T), compared on the 12 most
    probable tokens + the rest (chi-square, 12 degrees of freedom; 99.9 % quantile 32.9)."""
    from jacobiforcing_b200.policy import MODES, PolicyKnobs
    from oracle.hf_model import teacher_forced_logits
    model, rt = pair
    V = model.config.vocab_size
    g = torch.Generator().manual_seed(5)
    prompt = torch.randint(0, V, (20,), generator=g).tolist()
    T = 0.25
    lg = _bf16(teacher_forced_logits(model, prompt)[-1].numpy())
    p = torch.softmax(torch.from_numpy(lg) / T, -1).numpy().astype(np.float64)
    seq = rt.new_sequence()
    rt.prefill(seq, prompt, [])
    N = 2000
    counts = np.zeros(V, np.int64)
    prng = random.Random(1)
    for i in range(N):
        seq.ctx, seq.tokens = len(prompt) - 1, list(prompt)
        knobs = PolicyKnobs(mode=MODES["engine_nongreedy"], n=2, eos=None, pad=0, max_iteration_count=4, max_new_tokens=1,
                            vocab=V, temperature=T, seed=prng.getrandbits(64), rng_offset=i)
        pads = [int(np.argmax(p)) if i % 2 else prng.randrange(V) for _ in range(8)]     # half the drafts are the mode
        results, _ = rt.jacobi_run([seq], knobs, [0], None, pad_stream=[pads])
        assert len(results[0].tokens) == 1
        counts[results[0].tokens[0]] += 1
    rt.release(seq)
    top = np.argsort(-p)[:12]
    obs = np.append(counts[top], N - counts[top].sum()).astype(np.float64)
    exp = np.append(p[top], 1.0 - p[top].sum()) * N
    chi2 = float(((obs - exp) ** 2 / np.maximum(exp, 1e-9)).sum())
    print(f"chi-square over top-12 + rest: {chi2:.1f}; p(top) {p[top][:4]}, observed {counts[top][:4] / N}")
    assert exp.min() > 5.0, "buckets too thin for a chi-square test"
    assert chi2 < 32.9


def test_llm_generate_with_temperature(pair):
    from jacobiforcing_b200 import LLM, SamplingParams
    from oracle.hf_model import state_dict_bf16
    model, _ = pair
    kw = dict(hf_config=model.config, state_dict=state_dict_bf16(model), max_model_len=2048, num_kvcache_blocks=64,
              max_num_seqs=8, max_num_batched_tokens=2048)
    g = torch.Generator().manual_seed(0)
    prompts = [torch.randint(0, 1000, (16 + 5 * i,), generator=g).tolist() for i in range(3)]

    def gen(seed, temperature, strategy="jacobi"):
        llm = LLM(None, draft_seed=seed, **kw)
        llm.config.eos, llm.config.pad = -1, 0
        llm.scheduler.eos = -1
        try:
            out = llm.generate(prompts, SamplingParams(temperature=temperature, max_tokens=32, decode_strategy=strategy,
                                                       jacobi_block_len=8, ignore_eos=True), use_tqdm=False)
            return [o["token_ids"][:32] for o in out], llm.model_runner.jacobi_decoder.stats
        finally:
            llm.exit()
    a, st = gen(7, 0.8)
    b, _ = gen(7, 0.8)
    c, _ = gen(8, 0.8)
    assert all(len(x) == 32 for x in a)
    assert a == b, "same draft_seed, same samples"
    assert a != c, "another seed, other samples"
    assert st["tokens_accepted"] >= 96 and st["num_jacobi_iterations"] <= st["tokens_accepted"]
    # T -> 0: the verify degenerates to greedy decoding, except where the two best bf16 logits (nearly) tie - there the
    # target distribution itself is split between them
    from oracle.hf_model import teacher_forced_logits
    cold, _ = gen(7, 1e-3)
    greedy, _ = gen(7, 0.0)
    for p_ids, x, y in zip(prompts, cold, greedy):
        same = next((i for i, (u, v) in enumerate(zip(x, y)) if u != v), len(x))
        if same < len(x):
            lg = teacher_forced_logits(model, p_ids + y[:same])[-1]
            t2 = lg.topk(2)
            assert float(t2.values[0] - t2.values[1]) < 3e-2 and x[same] in t2.indices.tolist()
    with pytest.raises(NotImplementedError):
        gen(7, 0.8, strategy="jacobi_multiblock_rejection_recycling")
