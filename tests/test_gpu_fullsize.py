"""Size-independent properties at BASELINE.json's full size (Qwen2.5-7B shape, random-init, n=32) - the sizes the CPU
oracle cannot reach in test time:
  * Jacobi(n=32) ids == autoregressive (n=1) ids, bit for bit (the fixed-point property the reference's tests pin,
    inference_engine/tests/test_jacobi_decoding_greedy.py:180-206, exact here by construction of the attention chunks)
  * multiblock + rejection recycling (K=2, pool=4) == the same ids; batch rows independent of the batch composition
  * prompts crossing KV page boundaries (512-token prompt, 256-token pages)."""
import random

import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def rt7b():
    import bench
    from jacobiforcing_b200.runtime import JacobiRuntime
    free, _ = torch.cuda.mem_get_info(0)
    if free < 40 << 30:
        pytest.skip("needs ~25 GB of free device memory")
    cfg = bench.SHAPES["qwen2.5-coder-7b"]
    rt = JacobiRuntime(cfg, max_rows=1280, max_seqs=4, num_kv_pages=16, max_pages_per_seq=4, gen_capacity=1024,
                       max_calls=128)
    bench.load_random_weights(rt, cfg, 1234, "cuda:0")
    yield cfg, rt
    rt.close()


def _prompts(cfg, B, P, seed=1):
    g = torch.Generator().manual_seed(seed)
    return torch.randint(0, cfg["vocab_size"], (B, P), generator=g).tolist()


def test_7b_jacobi_equals_autoregressive(rt7b):
    from jacobiforcing_b200.runtime import hf_generate
    cfg, rt = rt7b
    prompt = _prompts(cfg, 1, 512)[0]
    outs = {}
    for mode, n in (("single", 1), ("single", 32), ("multiblock", 32)):
        seqs, res, run = hf_generate(rt, [prompt], n=n, max_new_tokens=64, mode=mode, K=2, r=0.85, n_gram_pool_size=4,
                                     pad=0, rngs=[random.Random(2)])
        outs[(mode, n)] = res[0].tokens[:64]
        assert res[0].stop_reason == "max_new_tokens"
        if n == 1:
            assert run["forwards"] == 64
        for s in seqs:
            rt.release(s)
    assert outs[("single", 1)] == outs[("single", 32)] == outs[("multiblock", 32)]
    assert len(set(outs[("single", 1)])) > 8          # a real generation, not a collapsed constant stream


def test_7b_batch_rows_independent(rt7b):
    from jacobiforcing_b200.runtime import hf_generate
    cfg, rt = rt7b
    prompts = _prompts(cfg, 4, 300, seed=7)
    mk = lambda: [random.Random(50 + i) for i in range(4)]
    seqs, batch, run = hf_generate(rt, prompts, n=32, max_new_tokens=64, mode="multiblock", K=2, r=0.85,
                                   n_gram_pool_size=4, pad=0, rngs=mk())
    for s in seqs:
        rt.release(s)
    for i in (0, 3):
        seqs, solo, _ = hf_generate(rt, [prompts[i]], n=32, max_new_tokens=64, mode="multiblock", K=2, r=0.85,
                                    n_gram_pool_size=4, pad=0, rngs=[mk()[i]])
        assert solo[0].tokens == batch[i].tokens and solo[0].iters == batch[i].iters
        for s in seqs:
            rt.release(s)
