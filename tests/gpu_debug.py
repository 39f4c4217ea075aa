"""Diagnostic run for a GPU box (not a pytest file): prints error statistics of the device forward against the HF
oracle and a first Jacobi run, so one gpurun call tells where a mismatch starts."""
import random
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from tests.gpu_helpers import DeviceForward, bf16_bits_to_float, make_pair   # noqa: E402
from jacobiforcing_b200.runtime import hf_generate                            # noqa: E402
from oracle import jacobi_policy as jp                                        # noqa: E402
from oracle.hf_model import HFForward, teacher_forced_logits                  # noqa: E402


def forward_check(name, P=40, **over):
    model, rt = make_pair(name)
    V = model.config.vocab_size
    g = torch.Generator().manual_seed(1)
    ids = torch.randint(0, V, (P,), generator=g).tolist()
    seq = rt.new_sequence()
    rt.ensure_capacity(seq, P)
    t0 = time.time()
    res = rt.forward(ids, [(0, P, 0, 0)], [seq], 0, P, want_logits=True, want_hidden=True)
    dt = time.time() - t0
    with torch.no_grad():
        out = model(input_ids=torch.tensor([ids]), output_hidden_states=True, use_cache=False)
    ref_logits = out.logits[0].float().numpy()
    ref_hidden = out.hidden_states[-1]           # after the final norm in HF; compare pre-norm via layer outputs
    hid = bf16_bits_to_float(res["hidden"])
    # HF hidden_states[-1] is post-norm; take the last layer's output through a hook-free route: hidden_states[-2] is
    # the input of the last layer, so recompute the pre-norm residual from the model directly
    with torch.no_grad():
        x = model.model.embed_tokens(torch.tensor([ids]))
        pos = torch.arange(P)[None]
        pe = model.model.rotary_emb(x, pos)
        from transformers.masking_utils import create_causal_mask
        mask = create_causal_mask(config=model.config, inputs_embeds=x, attention_mask=None,
                                  cache_position=torch.arange(P), past_key_values=None, position_ids=pos)
        per_layer = []
        for layer in model.model.layers:
            x = layer(x, attention_mask=mask, position_ids=pos, position_embeddings=pe)
            if isinstance(x, tuple):
                x = x[0]
            per_layer.append(x[0].numpy().copy())
    pre = per_layer[-1]
    err_h = np.abs(hid - pre)
    err_l = np.abs(res["logits"] - ref_logits)
    agree = (res["logits"].argmax(-1) == ref_logits.argmax(-1)).mean()
    g_agree = (res["greedy"] == ref_logits.argmax(-1)).mean()
    print(f"[{name}] forward {P} rows in {res['ms']:.3f} ms (wall {dt:.3f}s): hidden max|d|={err_h.max():.4g} "
          f"(ref max {np.abs(pre).max():.3g}) mean|d|={err_h.mean():.3g}; logits max|d|={err_l.max():.4g} "
          f"(ref max {np.abs(ref_logits).max():.3g}) mean|d|={err_l.mean():.3g}; argmax(logits) agree {agree:.3f}; "
          f"fused argmax agree {g_agree:.3f}; fused==argmax(dev logits) "
          f"{(res['greedy'] == bf16_round(res['logits']).argmax(-1)).mean():.3f}", flush=True)
    bad_rows = np.where(err_h.max(-1) > 0.1 * max(1.0, np.abs(pre).max()))[0]
    if len(bad_rows):
        print("   rows with large hidden error:", bad_rows[:20], flush=True)
    return model, rt, ids


def bf16_round(a):
    t = torch.from_numpy(np.ascontiguousarray(a)).to(torch.bfloat16).to(torch.float32)
    return t.numpy()


def jacobi_check(model, rt, name, mode, n, max_new, K=2, pool=4, r=0.85):
    V = model.config.vocab_size
    g = torch.Generator().manual_seed(1)
    prompt = torch.randint(0, V, (24,), generator=g).tolist()
    t0 = time.time()
    seqs, results, run = hf_generate(rt, [prompt], n=n, max_new_tokens=max_new, mode=mode, K=K, r=r,
                                     n_gram_pool_size=pool, pad=0, rngs=[random.Random(2)], want_trace=True)
    dt = time.time() - t0
    res = results[0]
    # oracle policy + HF forward (fp32, same bf16-rounded weights)
    fwd = HFForward(model)
    ref = jp.hf_driver(fwd, prompt, n, max_new, random.Random(2), mode=mode, K=K, r=r, n_gram_pool_size=pool, pad=0)
    same = 0
    for a, b in zip(res.tokens, ref.generated):
        if a != b:
            break
        same += 1
    print(f"[{name}/{mode} n={n}] device: {len(res.tokens)} tokens, calls={res.calls}, iters={res.total_iterations}, "
          f"stop={res.stop_reason}, forwards={run['forwards']}, {run['device_ms']:.2f} ms device, wall {dt:.2f}s | "
          f"oracle(HF fp32): {len(ref.generated)} tokens, iters={ref.total_iterations}; common prefix {same}", flush=True)
    # oracle policy + DEVICE forward: must match the device loop exactly
    seq2 = rt.new_sequence()
    rng = random.Random(2)

    class Drv:
        pass
    dfw = None

    def run_oracle_on_device():
        nonlocal dfw
        from jacobiforcing_b200.policy import draw_prefill_draft
        rng2 = random.Random(2)
        # replicate hf_driver but with the runtime's prefill for call 0
        draft = draw_prefill_draft(rng2, prompt, n)
        fc, drafted, _ = rt.prefill(seq2, prompt, draft)
        dfw = DeviceForward(rt, seq2)
        generated = list(prompt)
        iters = [0]
        calls = 1
        total_new = 0
        first = fc
        while total_new < max_new:
            inp = list(drafted) if calls == 1 else [first] + [rng2.choice(generated) for _ in range(n - 1)]
            if mode == "multiblock":
                first, acc, itr = jp.hf_multiblock(dfw, inp, n, K=K, r=r, n_gram_pool_size=pool, pad=0)
            else:
                first, acc, itr = jp.hf_single_block(dfw, inp, n)
            generated += acc
            total_new += len(acc)
            iters.append(itr)
            calls += 1
        return generated[len(prompt):], iters
    toks2, iters2 = run_oracle_on_device()
    print(f"    oracle policy on device forward: tokens equal {toks2 == res.tokens}, iters equal {iters2 == res.iters} "
          f"({sum(iters2)} vs {res.total_iterations})", flush=True)
    if toks2 != res.tokens:
        for i, (a, b) in enumerate(zip(res.tokens, toks2)):
            if a != b:
                print("    first token mismatch at", i, a, b)
                break
    for s in seqs + [seq2]:
        rt.release(s)


if __name__ == "__main__":
    which = sys.argv[1:] or ["tiny", "tiny128"]
    for name in which:
        model, rt, ids = forward_check(name)
        jacobi_check(model, rt, name, "single", 8, 32)
        jacobi_check(model, rt, name, "multiblock", 8, 32)
        jacobi_check(model, rt, name, "multiblock", 16, 48, pool=8)
        rt.close()
    print("DEBUG DONE", flush=True)
