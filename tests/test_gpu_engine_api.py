"""GPU tests of the reference-facing engine surface (LLM / SamplingParams / ModelRunner.run), modelled on the reference's
own script tests (inference_engine/tests/test_jacobi_decoding_greedy.py): Jacobi greedy == autoregressive greedy on the
first N tokens (:180-206), batch consistency (:209-302, here 100 % instead of the reference's >= 50 % bar because rows
are bit-independent of the batch), the multiblock strategy error (:97-115), and the decoder statistics (:152-176)."""
import random

import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def llm():
    from jacobiforcing_b200 import LLM
    from oracle.hf_model import build_model, state_dict_bf16
    model = build_model("tiny", seed=0, dtype=torch.float32)
    eng = LLM(None, hf_config=model.config, state_dict=state_dict_bf16(model), max_model_len=2048,
              num_kvcache_blocks=64, max_num_seqs=8, max_num_batched_tokens=2048, draft_seed=7)
    eng.config.eos, eng.config.pad = -1, 0
    eng.scheduler.eos = -1
    yield eng
    eng.exit()


def prompts(n, seed=0, V=1000):
    g = torch.Generator().manual_seed(seed)
    return [torch.randint(0, V, (16 + 5 * i,), generator=g).tolist() for i in range(n)]


def reset(llm):
    llm.scheduler.waiting.clear()
    llm.scheduler.running.clear()
    llm.model_runner.jacobi_decoder.reset()


def test_single_sequence_jacobi_equals_autoregressive(llm):
    from jacobiforcing_b200 import SamplingParams
    p = prompts(1)
    reset(llm)
    ar = llm.generate(p, SamplingParams(temperature=0.0, max_tokens=96, decode_strategy="autoregressive"), use_tqdm=False)
    reset(llm)
    jc = llm.generate(p, SamplingParams(temperature=0.0, max_tokens=96, decode_strategy="jacobi", jacobi_block_len=16,
                                        jacobi_max_iterations=128), use_tqdm=False)
    assert len(ar[0]["token_ids"]) == 96
    assert jc[0]["token_ids"][:96] == ar[0]["token_ids"][:96]
    st = llm.model_runner.jacobi_decoder.stats
    assert set(st) == {"num_chunk_calls", "num_jacobi_iterations", "tokens_accepted", "tokens_per_call",
                       "tokens_per_iteration", "iterations_per_call"}
    assert st["tokens_accepted"] == sum(st["tokens_per_call"]) == len(jc[0]["token_ids"])
    assert st["num_jacobi_iterations"] == sum(st["iterations_per_call"]) == len(st["tokens_per_iteration"])
    assert st["num_jacobi_iterations"] <= st["tokens_accepted"]          # TPF >= 1


def test_batch_jacobi_equals_autoregressive(llm):
    from jacobiforcing_b200 import SamplingParams
    p = prompts(6, seed=3)
    reset(llm)
    ar = llm.generate(p, SamplingParams(temperature=0.0, max_tokens=48, decode_strategy="autoregressive"), use_tqdm=False)
    reset(llm)
    jc = llm.generate(p, SamplingParams(temperature=0.0, max_tokens=48, decode_strategy="jacobi", jacobi_block_len=8),
                      use_tqdm=False)
    for a, j in zip(ar, jc):
        assert j["token_ids"][:48] == a["token_ids"][:48]


def test_multiblock_strategy_error_then_enabled(llm):
    from jacobiforcing_b200 import SamplingParams
    p = prompts(2, seed=5)
    sp = SamplingParams(temperature=0.0, max_tokens=40, decode_strategy="jacobi_multiblock_rejection_recycling",
                        jacobi_block_len=8, jacobi_max_blocks=2, jacobi_spawn_ratio=0.85, jacobi_n_gram_pool_size=4)
    reset(llm)
    with pytest.raises(NotImplementedError, match="jacobi_multiblock_rejection_recycling"):
        llm.generate(p, sp, use_tqdm=False)                    # reference behaviour (test_jacobi_decoding_greedy.py:97-115)
    for s in list(llm.scheduler.running):
        llm.scheduler.block_manager.deallocate(s)
    reset(llm)
    ar = llm.generate(p, SamplingParams(temperature=0.0, max_tokens=40, decode_strategy="autoregressive"), use_tqdm=False)
    llm.config.jacobi_multiblock_enabled = True
    try:
        reset(llm)
        mb = llm.generate(p, sp, use_tqdm=False)
    finally:
        llm.config.jacobi_multiblock_enabled = False
    for a, m in zip(ar, mb):
        assert m["token_ids"][:40] == a["token_ids"][:40]


def test_mixed_strategies_and_sampling_rejected(llm):
    from jacobiforcing_b200 import SamplingParams
    p = prompts(2, seed=9)
    reset(llm)
    with pytest.raises(NotImplementedError, match="Mixed decode strategies"):
        llm.generate(p, [SamplingParams(temperature=0.0, decode_strategy="jacobi", jacobi_block_len=8),
                         SamplingParams(temperature=0.0, decode_strategy="autoregressive")], use_tqdm=False)
    for s in list(llm.scheduler.running):
        llm.scheduler.block_manager.deallocate(s)
    reset(llm)
    # mixed greedy / non-greedy temperatures in one batch (reference: model_runner.py:331-373), and sampling on the plain
    # autoregressive path (served by the non-greedy Jacobi verify instead)
    with pytest.raises(NotImplementedError, match="Mixed greedy"):
        llm.generate(p, [SamplingParams(temperature=0.8, decode_strategy="jacobi", jacobi_block_len=8),
                         SamplingParams(temperature=0.0, decode_strategy="jacobi", jacobi_block_len=8)], use_tqdm=False)
    for s in list(llm.scheduler.running):
        llm.scheduler.block_manager.deallocate(s)
    reset(llm)
    with pytest.raises(NotImplementedError):
        llm.generate(p[:1], SamplingParams(temperature=0.8, decode_strategy="autoregressive"), use_tqdm=False)
    for s in list(llm.scheduler.running):
        llm.scheduler.block_manager.deallocate(s)
    reset(llm)


@pytest.mark.parametrize("L,max_tokens,max_iters", [(8, 40, 128), (16, 64, 128), (8, 200, 10)])
def test_engine_policy_device_loop_bit_exact(L, max_tokens, max_iters):
    """Mode 2 on the device == the oracle's engine_fixed_len (pinned against the reference JacobiDecoder) consuming the
    device forward's argmaxes: same accepted ids, iteration count and padding consumption."""
    from jacobiforcing_b200.policy import MODES, PolicyKnobs
    from oracle import jacobi_policy as jp
    from tests.gpu_helpers import DeviceEngineForward, make_pair
    model, rt = make_pair("tiny")
    V = model.config.vocab_size
    g = torch.Generator().manual_seed(11)
    prompt = torch.randint(0, V, (30,), generator=g).tolist()
    rng = random.Random(5)
    pre = [rng.choice(prompt) for _ in range(L)]
    pads = [rng.randrange(V) for _ in range(max_iters * L + L)]
    # device loop
    s = rt.new_sequence()
    _, drafted, _ = rt.prefill(s, prompt, pre)
    s.ctx = len(prompt) - 1
    kn = PolicyKnobs(mode=MODES["engine"], n=L, eos=None, pad=0, max_iteration_count=max_iters,
                     max_new_tokens=max_tokens, vocab=V)
    res, run = rt.jacobi_run([s], kn, [0], [drafted], pad_stream=[pads])
    # oracle policy on the device forward
    s2 = rt.new_sequence()
    _, drafted2, _ = rt.prefill(s2, prompt, pre)
    assert drafted2 == drafted
    left = list(pads)

    def randint(k):
        out = left[:k]
        del left[:k]
        return out
    toks = list(prompt)
    acc, iters, _ = jp.engine_fixed_len(DeviceEngineForward(rt, s2), toks, list(drafted), L, max_tokens, max_iters, V,
                                        randint)
    assert res[0].tokens == acc
    assert res[0].total_iterations == iters == run["forwards"]
    assert res[0].stop_reason == ("max_iterations" if iters >= max_iters and len(acc) < max_tokens else "max_new_tokens")
    rt.close()


def test_streaming_and_profile_drivers():
    """Block-by-block streaming (applications/jacobi_streaming_driver.py) and the CSV profile rows
    (MR_humaneval.py:259-273) produce the same ids as the one-launch generation."""
    from jacobiforcing_b200.drivers import CSV_COLUMNS, jacobi_stream, profile_rows
    from jacobiforcing_b200.runtime import hf_generate
    from tests.gpu_helpers import make_pair
    model, rt = make_pair("tiny")
    g = torch.Generator().manual_seed(21)
    prompt = torch.randint(0, model.config.vocab_size, (20,), generator=g).tolist()
    seqs, res, _ = hf_generate(rt, [prompt], n=8, max_new_tokens=40, mode="multiblock", K=2, r=0.8, pad=0,
                               rngs=[random.Random(0)])
    for s in seqs:
        rt.release(s)
    blocks = list(jacobi_stream(rt, prompt, n=8, max_new_tokens=40, K=2, r=0.8, pad=0, rng=random.Random(0)))
    assert all(len(b) == 8 for b in blocks) and len(blocks) == 5
    assert [t for b in blocks for t in b] == res[0].tokens[:40]
    rows = profile_rows(rt, [prompt], n=8, max_new_tokens=40, K=2, r=0.8, pad=0, seed=0)
    assert set(CSV_COLUMNS) <= set(rows[0])
    assert rows[0]["token_ids"][:40] == res[0].tokens[:40]
    assert rows[0]["calls"] == res[0].calls and rows[0]["total_iterations"] == res[0].total_iterations
    rt.close()


def test_prefix_cache_reuse_skips_cached_pages_and_keeps_ids(llm):
    """Prefill prefix cache (reference: engine/block_manager.py:135-168): a request whose leading full pages were
    committed by an earlier request shares them, its prefill forwards only the remainder, and its output is identical
    (greedy, Jacobi and autoregressive)."""
    from jacobiforcing_b200 import SamplingParams
    g = torch.Generator().manual_seed(11)
    shared = torch.randint(0, 1000, (600,), generator=g).tolist()            # two full 256-token pages + 88
    p1, p2 = shared + [3, 1, 4], shared[:560] + [2, 7, 1, 8]
    bm = llm.scheduler.block_manager
    for strategy in ("jacobi", "autoregressive"):
        sp = SamplingParams(temperature=0.0, max_tokens=24, decode_strategy=strategy, jacobi_block_len=8, ignore_eos=True)
        reset(llm)
        bm.hash_to_block_id.clear()                                          # cold cache
        for b in bm.blocks:
            b.hash, b.token_ids = -1, []
        rows0 = llm.model_runner.rt.stats.get("prefill_rows", 0)
        cold = llm.generate([p2], sp, use_tqdm=False)[0]["token_ids"]
        rows_cold = llm.model_runner.rt.stats["prefill_rows"] - rows0
        reset(llm)
        bm.hash_to_block_id.clear()
        for b in bm.blocks:
            b.hash, b.token_ids = -1, []
        llm.generate([p1], sp, use_tqdm=False)                               # commits the shared pages
        hits0 = bm.stats["prefix_hit_tokens"]
        rows0 = llm.model_runner.rt.stats["prefill_rows"]
        warm = llm.generate([p2], sp, use_tqdm=False)[0]["token_ids"]
        rows_warm = llm.model_runner.rt.stats["prefill_rows"] - rows0
        assert bm.stats["prefix_hit_tokens"] - hits0 == 512
        assert rows_warm == rows_cold - 512, (rows_warm, rows_cold)
        assert warm == cold, f"{strategy}: prefix-cache hit changed the output"
