"""Tensor parallelism on ONE GPU (-m gpu): tp ranks as host threads of this process, each with its own engine on a
slice of the device's SMs (jf_runtime_cfg.num_ctas), wired with jf_tp_attach_local.  The kernels are the ones that run
one-per-GPU over NVLink - the same push buffers, staging buffers, tags and flag barriers, addressed through plain
pointers instead of IPC mappings - so the sharded forward (reference: inference_engine/layers/linear.py:76-153,
layers/embed_head.py:56-66) is covered on the single-GPU box the driver tests on.  tests/test_gpu_tp.py repeats the
checks with one process per GPU when the box has several.

Checks per (model, tp, attention sharded or replicated):
  * every rank returns the same argmaxes / ids / per-call iteration counts;
  * the gathered vocabulary-parallel logits match the HF fp32 oracle within the bf16 tolerance of test_gpu_parity.py and
    the merged argmax equals the argmax of the bf16-rounded gathered logits (lowest index on ties);
  * the device-resident loop == the pinned policy restatement driven by the (sharded) device forward, bit-exact;
  * Jacobi(n = 8) == multiblock == AR on the sharded engine, bit for bit;
  * ids equal the single-engine ids up to near-ties of the oracle (the cross-rank sum changes the fp32 summation order).
"""
import random
import threading

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

LOGIT_MAX_TOL, LOGIT_MEAN_TOL, NEAR_TIE = 2.5e-2, 4e-3, 3e-2


def _rank_main(rank, tp, group, cfg, sd, shard_attn, num_ctas, out, errors):
    try:
        from jacobiforcing_b200.runtime import JacobiRuntime, hf_generate
        from oracle.hf_model import rope_tables
        from tests.test_gpu_parity import _oracle_on_device
        cos, sin = rope_tables(float(cfg.rope_theta), cfg.hidden_size // cfg.num_attention_heads,
                               cfg.max_position_embeddings)
        rt = JacobiRuntime(cfg, device=0, max_rows=512, max_seqs=4, num_kv_pages=16, max_pages_per_seq=4,
                           trace_capacity=1 << 16, tp_rank=rank, tp_size=tp, tp_group=group,
                           tp_shard_attention=shard_attn, num_ctas=num_ctas)
        rt.load_state_dict(sd)
        rt.set_rope_tables(cos.numpy(), sin.numpy())
        g = torch.Generator().manual_seed(1)
        res = {}
        for rows in (5, 72, 400):                     # fused push tails (<= 320 rows) and the flag-barrier prefill path
            ids = torch.randint(0, cfg.vocab_size, (rows,), generator=g).tolist()
            seq = rt.new_sequence()
            rt.ensure_capacity(seq, rows)
            lm0 = 0 if rows <= 256 else rows - 256
            r = rt.forward(ids, [(0, rows, 0, 0)], [seq], lm0, rows - lm0, want_logits=True, want_hidden=True)
            rt.release(seq)
            res[rows] = (ids, lm0, r["logits"], r["greedy"][lm0:].tolist(), r["hidden"])
        prompt = torch.randint(0, cfg.vocab_size, (24,), generator=g).tolist()
        runs = {}
        for mode, n in (("single", 1), ("single", 8), ("multiblock", 8)):
            seqs, results, run = hf_generate(rt, [prompt], n=n, max_new_tokens=40, mode=mode, pad=0,
                                             rngs=[random.Random(2)], want_trace=True)
            for s in seqs:
                rt.release(s)
            runs[(mode, n)] = (results[0].tokens[:40], results[0].iters, run["forwards"],
                               [t["rows"] for t in run["trace"]], [t["kv_before"] for t in run["trace"]])
        # the oracle's policy restatement on the sharded device forward (every rank runs it: the calls are collective)
        toks, iters, stop, trace = _oracle_on_device(rt, prompt, "multiblock", 8, 40, K=2, r=0.85, pool=4)
        out[rank] = dict(fwd=res, runs=runs, prompt=prompt, oracle=(toks[:40], iters, trace.rows, trace.kv_before))
        group.barrier()
        rt.close()
    except BaseException as e:       # noqa: BLE001 - a dead rank must not leave its peers waiting forever
        errors.append((rank, repr(e)))
        try:
            group._barrier.abort()
        except Exception:
            pass
        raise


CASES = [("tiny", 2, False, True), ("tiny", 2, True, True), ("tiny", 4, False, True), ("tiny", 8, False, True),
         ("mid128", 2, True, True), ("mid128", 4, False, True), ("mid128", 8, False, True),
         # fused tails off: every forward takes the stand-alone reduce stage, whose partial sums the peers PULL after a
         # cross-GPU flag barrier (the path of forwards larger than the fused row capacity)
         ("tiny", 4, False, False), ("mid128", 2, True, False)]


@pytest.mark.parametrize("model_name,tp,shard_attn,fused", CASES)
def test_virtual_ranks_match_oracle_and_single_engine(model_name, tp, shard_attn, fused, monkeypatch):
    if not fused:
        monkeypatch.setenv("JF_FUSE_NORM", "0")
    from jacobiforcing_b200.tp import LocalTPGroup
    from jacobiforcing_b200.runtime import hf_generate
    from oracle.hf_model import teacher_forced_logits
    from tests.gpu_helpers import make_pair
    sms = torch.cuda.get_device_properties(0).multi_processor_count
    num_ctas = (sms - 4) // tp                      # co-resident persistent kernels: leave a few SMs free
    group = LocalTPGroup(tp)
    out, errors = {}, []
    model, rt1 = make_pair(model_name)              # the oracle model (bf16-rounded weights) and the single engine
    from oracle.hf_model import state_dict_bf16
    sd = state_dict_bf16(model)                     # built here: transformers' lazy imports are not thread-safe
    import tests.test_gpu_parity  # noqa: F401
    threads = [threading.Thread(target=_rank_main, args=(r, tp, group, model.config, sd, shard_attn, num_ctas, out, errors))
               for r in range(tp)]
    for t in threads:
        t.start()
    for t in threads:
        t.join(900)
    assert not errors, errors
    assert len(out) == tp
    from oracle.hf_model import bf16_copy
    m16 = bf16_copy(model)
    r0 = out[0]
    for rows, (ids, lm0, _, greedy0, hidden0) in r0["fwd"].items():
        logits = np.concatenate([out[r]["fwd"][rows][2] for r in range(tp)], axis=1)
        for r in range(1, tp):
            assert out[r]["fwd"][rows][3] == greedy0, "ranks disagree on the merged argmax"
            assert (out[r]["fwd"][rows][4] == hidden0).all(), "ranks hold different hidden states"
        ref = teacher_forced_logits(model, ids).numpy()[lm0:]
        err = np.abs(logits - ref)
        err16 = np.abs(teacher_forced_logits(m16, ids).numpy()[lm0:] - ref)   # what the reference's own dtype achieves
        print(f"[tp{tp} {model_name} attn_sharded={shard_attn} rows={rows}] max {err.max():.4g} mean {err.mean():.4g} "
              f"(HF bf16: max {err16.max():.4g} mean {err16.mean():.4g})")
        assert err.max() <= max(LOGIT_MAX_TOL, 1.5 * err16.max())
        assert err.mean() <= max(LOGIT_MEAN_TOL, 1.5 * err16.mean() + 1e-4)
        dev16 = torch.from_numpy(logits).to(torch.bfloat16).float()
        assert greedy0 == dev16.argmax(-1).tolist()
    for r in range(1, tp):
        assert out[r]["runs"] == r0["runs"], "ranks disagree on a generation"
        assert out[r]["oracle"] == r0["oracle"]
    runs = r0["runs"]
    assert runs[("single", 1)][0] == runs[("single", 8)][0] == runs[("multiblock", 8)][0], "Jacobi != AR on the sharded engine"
    mb = runs[("multiblock", 8)]
    assert (mb[0], mb[1], mb[3], mb[4]) == r0["oracle"], "device loop != oracle policy on the sharded device forward"
    # against the single engine: equal, or first different where the oracle's top-2 margin is a near-tie
    seqs, results, _ = hf_generate(rt1, [r0["prompt"]], n=8, max_new_tokens=40, mode="multiblock", pad=0,
                                   rngs=[random.Random(2)])
    one = results[0].tokens[:40]
    rt1.close()
    same = next((i for i, (a, b) in enumerate(zip(mb[0], one)) if a != b), 40)
    if same < 40:
        lg = teacher_forced_logits(model, r0["prompt"] + one[:same])[-1]
        t2 = lg.topk(2).values
        assert float(t2[0] - t2[1]) < NEAR_TIE


@pytest.mark.parametrize("tp", [2, 4])
def test_llm_tensor_parallel_from_a_plain_process(tp):
    """`LLM(..., tensor_parallel_size=N)` without a launcher: the ranks are threads of this process
    (runtime.ThreadedTensorParallel; here they share one GPU).  Jacobi == autoregressive on the sharded engine, and both
    equal the single-engine output up to near-ties of the oracle."""
    from jacobiforcing_b200 import LLM, SamplingParams
    from oracle.hf_model import build_model, state_dict_bf16, teacher_forced_logits
    model = build_model("tiny", seed=0, dtype=torch.float32)
    sd = state_dict_bf16(model)
    with torch.no_grad():
        for k, v in model.state_dict().items():
            v.copy_(v.to(torch.bfloat16).to(torch.float32))
    g = torch.Generator().manual_seed(0)
    prompts = [torch.randint(0, 1000, (16 + 5 * i,), generator=g).tolist() for i in range(3)]
    kw = dict(hf_config=model.config, state_dict=sd, max_model_len=2048, num_kvcache_blocks=64, max_num_seqs=8,
              max_num_batched_tokens=2048, draft_seed=7)

    def run(tp_size):
        extra = dict(tensor_parallel_size=tp_size, tp_devices=[0] * tp_size) if tp_size > 1 else {}
        llm = LLM(None, **kw, **extra)
        llm.config.eos, llm.config.pad = -1, 0
        llm.scheduler.eos = -1
        try:
            ar = llm.generate(prompts, SamplingParams(temperature=0.0, max_tokens=40, decode_strategy="autoregressive"),
                              use_tqdm=False)
            jc = llm.generate(prompts, SamplingParams(temperature=0.0, max_tokens=40, decode_strategy="jacobi",
                                                      jacobi_block_len=8), use_tqdm=False)
            return [a["token_ids"][:40] for a in ar], [j["token_ids"][:40] for j in jc]
        finally:
            llm.exit()
    ar_tp, jc_tp = run(tp)
    assert ar_tp == jc_tp
    ar_1, _ = run(1)
    for p, a, b in zip(prompts, ar_tp, ar_1):
        same = next((i for i, (x, y) in enumerate(zip(a, b)) if x != y), 40)
        if same < 40:
            t2 = teacher_forced_logits(model, p + b[:same])[-1].topk(2).values
            assert float(t2[0] - t2[1]) < NEAR_TIE


def _ng_rank_main(rank, tp, group, cfg, sd, shard_attn, num_ctas, shared, out, errors):
    """Non-greedy generation on the sharded engine, then the restatement driven by the SHARDED device forward: every rank
    runs the same replay (the forwards are collective) and assembles the full logits from the ranks' vocabulary slices."""
    try:
        from jacobiforcing_b200.policy import MODES, PolicyKnobs
        from jacobiforcing_b200.runtime import JacobiRuntime
        from oracle import jacobi_policy as jp
        from oracle.hf_model import rope_tables
        from tests.gpu_helpers import DeviceEngineForward
        cos, sin = rope_tables(float(cfg.rope_theta), cfg.hidden_size // cfg.num_attention_heads,
                               cfg.max_position_embeddings)
        rt = JacobiRuntime(cfg, device=0, max_rows=256, max_seqs=2, num_kv_pages=8, max_pages_per_seq=4,
                           trace_capacity=1 << 16, tp_rank=rank, tp_size=tp, tp_group=group,
                           tp_shard_attention=shard_attn, num_ctas=num_ctas)
        rt.load_state_dict(sd)
        rt.set_rope_tables(cos.numpy(), sin.numpy())
        V, L, T, kseed, off = cfg.vocab_size, 8, 0.7, 0x5eed5eed1234, 3
        g = torch.Generator().manual_seed(21)
        prompt = torch.randint(0, V, (24,), generator=g).tolist()
        prng = random.Random(21)
        pads = [prng.randrange(V) for _ in range(64 * L)]
        seq = rt.new_sequence()
        rt.prefill(seq, prompt, [])
        seq.ctx = len(prompt) - 1
        knobs = PolicyKnobs(mode=MODES["engine_nongreedy"], n=L, eos=None, pad=0, max_iteration_count=128, max_new_tokens=32,
                            vocab=V, temperature=T, seed=kseed, rng_offset=off)
        results, run = rt.jacobi_run([seq], knobs, [0], None, pad_stream=[pads], want_trace=True)
        rt.release(seq)
        dev = (results[0].tokens, results[0].total_iterations, [t["rows"][0] for t in run["trace"]])
        # replay
        seq = rt.new_sequence()
        rt.prefill(seq, prompt, [])
        fwd = DeviceEngineForward(rt, seq, want_logits=True)
        cur = {"p": 0, "it": -1}
        inv_t = np.float32(1.0) / np.float32(T)

        def forward_step(token_ids, draft):
            greedy = fwd(token_ids, draft)
            cur["it"] += 1
            shared[(cur["it"], rank)] = fwd.last_logits[:L - 1]
            group.barrier()
            full = np.concatenate([shared[(cur["it"], r)] for r in range(tp)], axis=1)
            group.barrier()
            sc = torch.from_numpy(np.ascontiguousarray(full)).to(torch.bfloat16).float().numpy() * inv_t
            e = np.exp(sc - sc.max(axis=1, keepdims=True))
            probs = e / e.sum(axis=1, keepdims=True)
            px = [float(probs[t, draft[t + 1]]) for t in range(L - 1)]

            def residual(t):
                u = jp.device_gumbel_uniforms(kseed, off, t, cur["it"], V)
                key = sc[t] + (-np.log(-np.log(u))).astype(np.float32)
                key[draft[t + 1]] = -np.inf
                return int(np.argmax(key))
            return greedy, px, residual

        def randint(k):
            o = pads[cur["p"]:cur["p"] + k]
            cur["p"] += k
            return o
        toks = list(prompt)
        accepted, iters, per_iter, drafts = jp.engine_nongreedy(
            forward_step, toks, L, 32, 128, randint, lambda it, t: jp.device_accept_uniform(kseed, off, t, it))
        rt.release(seq)
        out[rank] = (dev, (accepted, iters, drafts))
        group.barrier()
        rt.close()
    except BaseException as e:       # noqa: BLE001
        errors.append((rank, repr(e)))
        try:
            group._barrier.abort()
        except Exception:
            pass
        raise


@pytest.mark.parametrize("tp,shard_attn", [(2, False), (2, True), (4, False)])
def test_nongreedy_on_the_sharded_engine(tp, shard_attn):
    """Non-greedy Jacobi under tensor parallelism: the vocabulary-parallel sampling epilogue (each rank holds the softmax
    normaliser, the drafted token's logit and a Gumbel-max sample of ITS slice; the policy stage of every rank merges all
    of them) gives every rank the same run, equal to the restatement on the gathered device logits."""
    from jacobiforcing_b200.tp import LocalTPGroup
    from oracle.hf_model import state_dict_bf16
    from tests.gpu_helpers import make_pair
    sms = torch.cuda.get_device_properties(0).multi_processor_count
    model, rt1 = make_pair("tiny")
    rt1.close()
    sd = state_dict_bf16(model)
    group = LocalTPGroup(tp)
    shared, out, errors = {}, {}, []
    threads = [threading.Thread(target=_ng_rank_main, args=(r, tp, group, model.config, sd, shard_attn, (sms - 4) // tp,
                                                            shared, out, errors)) for r in range(tp)]
    for t in threads:
        t.start()
    for t in threads:
        t.join(600)
    assert not errors, errors
    for r in range(1, tp):
        assert out[r] == out[0], "ranks disagree"
    dev, rep = out[0]
    assert dev[0] == rep[0] and dev[1] == rep[1] and dev[2] == rep[2]
    assert len(dev[0]) >= 32
