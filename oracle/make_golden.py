"""CPU ORACLE tooling: generate tests/golden/*.json by running the UNMODIFIED reference functions (under
oracle/ref_shim.py) on seeded random-init Qwen2 models, through a driver loop that follows
JacobiForcing/jacobi_forcing_inference_MR_humaneval.py:152-240 call for call but draws its random draft tokens from
an explicit random.Random(seed) (the reference uses the unseeded global RNG - SURVEY.md finding 10).

Run here (needs /root/reference):   python -m oracle.make_golden
Each golden file records, for one (model, prompt, knobs) case:
  * every forward the reference issued: kv length before, the `out` rows fed to embed_tokens, and argmax(lm_head(...).float())
    for every row/position (the "greedy table") -> the policy restatement can be replayed against it with NO model
  * every call's input_ids, accepted ids, next token and iteration count
so tests/test_oracle_policy.py pins oracle/jacobi_policy.py bit-exactly on any machine.
"""
from __future__ import annotations

import json
import os
import random
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle import ref_shim                      # noqa: E402
from oracle.hf_model import build_model          # noqa: E402

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


class Recorder:
    """Hooks embed_tokens (rows in) and lm_head (greedy out) of an HF model; reads the KV length from the cache the
    reference passes to create_causal_mask."""

    def __init__(self, model, mb, sb):
        self.forwards = []
        self._cur = None
        self.kv_probe = lambda: 0
        emb = model.model.embed_tokens
        head = model.lm_head
        rec = self

        def emb_hook(_m, args):
            ids = args[0]
            rec._cur = {"kv_before": int(rec.kv_probe()), "rows": ids.tolist()}

        def head_hook(_m, _args, out):
            g = torch.argmax(out.float(), dim=-1).tolist()
            top2 = out.float().topk(2, dim=-1).values
            rec._cur["greedy"] = g
            rec._cur["min_margin"] = float((top2[..., 0] - top2[..., 1]).min())
            rec.forwards.append(rec._cur)
            rec._cur = None

        self.h1 = emb.register_forward_pre_hook(emb_hook)
        self.h2 = head.register_forward_hook(head_hook)

    def close(self):
        self.h1.remove()
        self.h2.remove()


def run_case(name, model_name, seed, prompt_len, n, max_new, mode, K=2, r=0.85, lsr=0.0, pool=4, eos=None, pad=0,
             max_iteration_count=128, draft_seed=2, dtype=torch.float32):
    mb, sb = ref_shim.install()
    model = ref_shim.bind(build_model(model_name, seed=seed, dtype=dtype))
    V = model.config.vocab_size
    g = torch.Generator().manual_seed(1)
    prompt = torch.randint(0, V, (1, prompt_len), generator=g)
    rng = random.Random(draft_seed)
    rec = Recorder(model, mb, sb)

    calls_log = []
    generated = prompt.clone()
    prompt_ids = prompt
    input_ids = prompt
    past = None
    first_correct = None
    prefill_drafted = None
    calls = 0
    total_new = 0
    prev_len = prompt.shape[1]
    iters = []
    prefill_phase = True
    stop_reason = None
    pkv_holder = {"c": None}
    rec.kv_probe = lambda: (pkv_holder["c"].get_seq_length() if pkv_holder["c"] is not None else 0)
    eos_cmp = eos if eos is not None else -1

    while True:
        gen_part = generated[0, prompt_len:]
        if (gen_part == eos_cmp).any().item():
            stop_reason = "eos"
            break
        if total_new >= max_new:
            stop_reason = "max_new_tokens"
            break
        if calls >= 1024:
            stop_reason = "max_calls"
            break
        if prefill_phase:
            draft = torch.tensor([[rng.choice(generated[0].tolist()) for _ in range(n)]], dtype=torch.long)
            pin = torch.cat((prompt_ids, draft), dim=-1)
            pkv_holder["c"] = None
            if mode == "multiblock":
                past, first_correct, prefill_drafted, itr = model.jacobi_forward_greedy_multiblock(
                    input_ids=pin, attention_mask=torch.ones_like(pin), past_key_values=None, use_cache=True,
                    prefill_phase=True, n_token_seq_len=n, eos_token_id=eos, pad_token_id=pad)
            else:
                past, first_correct, prefill_drafted, itr = model.jacobi_forward_greedy(
                    input_ids=pin, attention_mask=torch.ones_like(pin), past_key_values=None, use_cache=True,
                    prefill_phase=True, n_token_seq_len=n, eos_token_id=eos)
            pkv_holder["c"] = past
            prefill_phase = False
            generated = prompt_ids
            itr = 0
            calls_log.append({"kind": "prefill", "draft": draft[0].tolist(),
                              "first_correct": int(first_correct.view(-1)[0]),
                              "prefill_drafted": prefill_drafted[0].tolist(), "kv_after": past.get_seq_length()})
        else:
            if calls == 1:
                input_ids = prefill_drafted
            else:
                q = torch.tensor([[rng.choice(generated[0].tolist()) for _ in range(n - 1)]], dtype=torch.long)
                input_ids = torch.cat((first_correct.view(1, -1), q), dim=-1)
            inp = input_ids[0].tolist()
            if mode == "multiblock":
                past, first_correct, acc, itr = model.jacobi_forward_greedy_multiblock(
                    input_ids=input_ids.clone(), attention_mask=None, past_key_values=past, use_cache=True,
                    prefill_phase=False, n_token_seq_len=n, K=K, r=r, lookahead_start_ratio=lsr,
                    n_gram_pool_size=pool, eos_token_id=eos, pad_token_id=pad,
                    max_iteration_count=max_iteration_count)
            else:
                past, first_correct, acc, itr = model.jacobi_forward_greedy(
                    input_ids=input_ids.clone(), attention_mask=None, past_key_values=past, use_cache=True,
                    prefill_phase=False, n_token_seq_len=n, eos_token_id=eos)
            pkv_holder["c"] = past
            generated = torch.cat((generated, acc), dim=-1)
            calls_log.append({"kind": "decode", "input_ids": inp, "accepted": acc[0].tolist(),
                              "next_token": (int(first_correct.view(-1)[0]) if first_correct.numel() > 0 else None),
                              "iters": int(itr), "kv_after": past.get_seq_length()})
        calls += 1
        iters.append(int(itr))
        added = generated.shape[1] - prev_len
        if added > 0:
            total_new += added
        prev_len = generated.shape[1]
    rec.close()

    # the property the reference's own tests pin (tests/test_jacobi_decoding_greedy.py:180-206): == greedy AR
    from oracle.hf_model import greedy_ar
    gen = generated[0, prompt_len:].tolist()
    ar = greedy_ar(model, prompt[0].tolist(), len(gen)) if eos is None else None
    out = {
        "case": name, "model": model_name, "model_seed": seed, "dtype": str(dtype).replace("torch.", ""),
        "prompt": prompt[0].tolist(), "n": n, "max_new_tokens": max_new, "mode": mode, "K": K, "r": r,
        "lookahead_start_ratio": lsr, "n_gram_pool_size": pool, "eos": eos, "pad": pad,
        "max_iteration_count": max_iteration_count, "draft_seed": draft_seed,
        "generated": gen, "iters": iters, "calls": calls, "stop_reason": stop_reason, "total_new_tokens": total_new,
        "equals_greedy_ar": (None if ar is None else bool(ar == gen)),
        "call_log": calls_log, "forwards": rec.forwards,
        "reference": "hao-ai-lab/JacobiForcing @ 028dbd6, functions run unmodified under oracle/ref_shim.py",
    }
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    path = os.path.join(GOLDEN_DIR, name + ".json")
    with open(path, "w") as f:
        json.dump(out, f, separators=(",", ":"))
    mm = min(fw["min_margin"] for fw in rec.forwards)
    print(f"{name}: {len(gen)} tokens, calls={calls}, iters={sum(iters)}, forwards={len(rec.forwards)}, "
          f"stop={stop_reason}, ==AR {out['equals_greedy_ar']}, min top-2 margin {mm:.4g}, "
          f"max rows {max(len(fw['rows']) for fw in rec.forwards)} -> {os.path.relpath(path, ROOT)} "
          f"({os.path.getsize(path) // 1024} KiB)")
    return out


def pick_eos(golden, after: int):
    """An EOS id that will actually be hit: a token the no-EOS run generated after `after` tokens."""
    gen = golden["generated"]
    for i in range(after, len(gen)):
        if gen[i] not in gen[:i] and gen[i] not in golden["prompt"]:
            return int(gen[i])
    raise RuntimeError("no late-first-occurrence token to use as EOS")


def main():
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    base = run_case("single_n8", "tiny", 0, 24, 8, 48, "single")
    run_case("single_n16_tiny128", "tiny128", 3, 40, 16, 64, "single")
    run_case("single_n8_eos", "tiny", 0, 24, 8, 48, "single", eos=pick_eos(base, 19))
    mbase = run_case("multi_n8_k2", "tiny", 0, 24, 8, 64, "multiblock", K=2, r=0.85, pool=4)
    run_case("multi_n16_k2_r05", "tiny", 5, 32, 16, 96, "multiblock", K=2, r=0.5, pool=4)
    run_case("multi_n16_k2_pool8_tiny128", "tiny128", 7, 40, 16, 96, "multiblock", K=2, r=0.85, pool=8)
    run_case("multi_n8_k1_pool1", "tiny", 0, 24, 8, 48, "multiblock", K=1, r=0.85, pool=1)
    run_case("multi_n8_k2_lsr05", "tiny", 9, 24, 8, 64, "multiblock", K=2, r=0.85, pool=4, lsr=0.5)
    run_case("multi_n8_k2_eos", "tiny", 0, 24, 8, 64, "multiblock", K=2, r=0.85, pool=4, eos=pick_eos(mbase, 21))
    run_case("multi_n8_k2_eos_b", "tiny", 0, 24, 8, 64, "multiblock", K=2, r=0.5, pool=4, eos=pick_eos(mbase, 12))
    run_case("multi_n8_k3_quirk", "tiny", 0, 24, 8, 40, "multiblock", K=3, r=0.5, pool=4, max_iteration_count=40)


def main_r2():
    """Round-2 additions: the headline block size n=32 (BASELINE configs 2-4) and the README-recommended n=64
    (config 5), single block and multiblock + rejection recycling."""
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    run_case("single_n32", "tiny", 11, 40, 32, 64, "single")
    run_case("multi_n32_k2_pool4", "tiny", 12, 40, 32, 96, "multiblock", K=2, r=0.85, pool=4)
    run_case("multi_n64_k2_pool4_tiny128", "tiny128", 13, 48, 64, 128, "multiblock", K=2, r=0.85, pool=4)


if __name__ == "__main__" and "--engine" not in sys.argv and "--r2" not in sys.argv:
    main()
if __name__ == "__main__" and "--r2" in sys.argv:
    main_r2()


# --------------------------------------------------------------------------------------------------------------------
# engine convention: the reference JacobiDecoder itself (inference_engine/engine/jacobi_decoding.py:303-445), run on CPU
# through its forward_step callback with the HF forward; torch.randint is wrapped so the padding stream is recorded.
# --------------------------------------------------------------------------------------------------------------------
def run_engine_case(name, model_name, seed, prompt_len, L, max_tokens, max_iters=128, eos=None, with_prefill_draft=True):
    sys.path.insert(0, ref_shim.REFERENCE_ROOT)
    from inference_engine.engine.jacobi_decoding import JacobiDecoder
    from inference_engine.engine.sequence import Sequence
    from inference_engine.sampling_params import SamplingParams
    model = build_model(model_name, seed=seed, dtype=torch.float32)
    V = model.config.vocab_size
    g = torch.Generator().manual_seed(1)
    prompt = torch.randint(0, V, (prompt_len,), generator=g).tolist()
    forwards = []

    @torch.no_grad()
    def forward_step(seq, q_draft):
        toks = list(seq.token_ids)
        S = len(toks)
        draft = q_draft[0].tolist()
        assert draft[0] == toks[-1]                       # model_runner.py:1157-1162
        ids = torch.tensor([toks[:-1] + draft], dtype=torch.long)
        logits = model(input_ids=ids, use_cache=False).logits[:, S - 1:, :].float()      # L positions
        forwards.append({"kv_before": S - 1, "rows": [draft], "greedy": [torch.argmax(logits[0], -1).tolist()]})
        seq.num_cached_tokens = (S - 1) + len(draft)      # model_runner.py: sets (len-1)+L; trimmed by the decoder
        return logits[:, :-1, :]

    pad_stream = []
    real_randint = torch.randint

    def rec_randint(*a, **k):
        out = real_randint(*a, **k)
        pad_stream.extend(out.flatten().tolist())
        return out

    class FakeBM:                                          # the decoder only needs these three calls
        def may_append_batch(self, seq, n):
            pass

        def may_append(self, seq):
            pass

        def trim_kv_only_fast(self, seq, n):
            seq.num_cached_tokens -= n

    sp = SamplingParams(temperature=0.0, max_tokens=max_tokens, decode_strategy="jacobi", jacobi_block_len=L,
                        jacobi_max_iterations=max_iters)
    seq = Sequence(prompt, sp)
    seq.num_cached_tokens = len(prompt)
    prefill_draft = None
    if with_prefill_draft:
        # model_runner.py:797-918: append L random prompt tokens, greedy draft = argmax(logits[P-1 : P+L-1])
        rng = random.Random(2)
        rnd = [rng.choice(prompt) for _ in range(L)]
        with torch.no_grad():
            lg = model(input_ids=torch.tensor([prompt + rnd]), use_cache=False).logits[0].float()
        prefill_draft = torch.argmax(lg[prompt_len - 1: prompt_len + L - 1], -1).tolist()
        seq._prefill_draft = list(prefill_draft)
    dec = JacobiDecoder(FakeBM(), forward_step=forward_step, eos_token_id=eos, pad_token_id=0, vocab_size=V,
                        device=torch.device("cpu"))
    torch.manual_seed(1234)
    torch.randint = rec_randint
    try:
        accepted = dec.generate_chunk(seq)
    finally:
        torch.randint = real_randint
    from oracle.hf_model import greedy_ar
    ar = greedy_ar(model, prompt, len(accepted)) if eos is None else None
    out = {"case": name, "mode": "engine", "model": model_name, "model_seed": seed, "dtype": "float32",
           "prompt": prompt, "L": L, "max_tokens": max_tokens, "max_iterations": max_iters, "eos": eos,
           "prefill_draft": prefill_draft, "pad_stream": pad_stream, "accepted": accepted,
           "iters": dec.stats["num_jacobi_iterations"], "tokens_per_iteration": dec.stats["tokens_per_iteration"],
           "equals_greedy_ar": (None if ar is None else bool(ar == accepted)), "forwards": forwards,
           "reference": "hao-ai-lab/JacobiForcing @ 028dbd6, inference_engine JacobiDecoder run unmodified on CPU"}
    path = os.path.join(GOLDEN_DIR, name + ".json")
    with open(path, "w") as f:
        json.dump(out, f, separators=(",", ":"))
    print(f"{name}: {len(accepted)} tokens in {out['iters']} iterations, pads used {len(pad_stream)}, "
          f"==AR {out['equals_greedy_ar']} -> {os.path.relpath(path, ROOT)}")
    return out


def main_engine():
    base = run_engine_case("engine_L8", "tiny", 0, 24, 8, 40)
    run_engine_case("engine_L16_nodraft_tiny128", "tiny128", 3, 40, 16, 64, with_prefill_draft=False)
    run_engine_case("engine_L8_maxiter", "tiny", 0, 24, 8, 200, max_iters=12)
    gen = base["accepted"]
    eos = next((t for i, t in enumerate(gen) if i >= 12 and t not in gen[:i] and t not in base["prompt"]), None)
    if eos is not None:
        run_engine_case("engine_L8_eos", "tiny", 0, 24, 8, 40, eos=eos)


if __name__ == "__main__" and "--engine" in sys.argv:
    main_engine()
