"""Golden runs of the UNMODIFIED reference class JacobiDecoderNonGreedy (inference_engine/engine/
jacobi_decoding_nongreedy.py) on CPU, for pinning oracle/jacobi_policy.py::engine_nongreedy.

    python oracle/make_golden_nongreedy.py          (needs /root/reference; writes tests/golden/ng_*.json)

The decoder is driven through its forward_step callback with the HF forward of a tiny seeded Qwen2 (the same callback
contract as the greedy goldens, oracle/make_golden.py::run_engine_case).  Its three sources of randomness - torch.rand
(accept test, :322), torch.multinomial (residual sample, :128-131) and torch.randint (draft padding, :227, :463) - are
wrapped IN THE REFERENCE MODULE'S NAMESPACE so every draw is recorded in call order; nothing else of the module is
touched.  Per forward the golden keeps the draft, the argmax row, the target probability of every drafted token and the
draws consumed, which is exactly what the restatement needs to replay the run decision by decision."""
from __future__ import annotations

import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_shim                                   # noqa: E402
from oracle.hf_model import build_model                       # noqa: E402

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


class _TorchProxy:
    """`torch` as the reference module sees it: rand / multinomial / randint record their results."""

    def __init__(self, log):
        self._log = log

    def __getattr__(self, name):
        return getattr(torch, name)

    def rand(self, *a, **k):
        out = torch.rand(*a, **k)
        self._log["rand"].append(float(out))
        self._log["order"].append("u")
        return out

    def multinomial(self, probs, num_samples=1, **k):
        out = torch.multinomial(probs, num_samples=num_samples, **k)
        self._log["multinomial"].append(int(out.flatten()[0]))
        self._log["order"].append("m")
        return out

    def randint(self, *a, **k):
        out = torch.randint(*a, **k)
        self._log["randint"].extend(out.flatten().tolist())
        return out


def run_case(name, model_name, seed, prompt_len, L, max_tokens, temperature, max_iters=128, eos=None, torch_seed=1234):
    sys.path.insert(0, ref_shim.REFERENCE_ROOT)
    import inference_engine.engine.jacobi_decoding_nongreedy as ng
    from inference_engine.engine.sequence import Sequence
    from inference_engine.sampling_params import SamplingParams
    model = build_model(model_name, seed=seed, dtype=torch.float32)
    V = model.config.vocab_size
    g = torch.Generator().manual_seed(1)
    prompt = torch.randint(0, V, (prompt_len,), generator=g).tolist()
    forwards = []
    log = {"rand": [], "multinomial": [], "randint": [], "order": []}

    @torch.no_grad()
    def forward_step(seq, q_draft):
        toks = list(seq.token_ids)
        S = len(toks)
        draft = q_draft[0].tolist()
        assert draft[0] == toks[-1]
        ids = torch.tensor([toks[:-1] + draft], dtype=torch.long)
        logits = model(input_ids=ids, use_cache=False).logits[:, S - 1:, :].float()
        lg = logits[0, :-1]
        probs = torch.softmax(lg / temperature, dim=-1)
        px = [float(probs[t, draft[t + 1]]) for t in range(len(draft) - 1)]
        forwards.append({"kv_before": S - 1, "draft": draft, "greedy": torch.argmax(lg, -1).tolist(), "px": px,
                         "draws_before": [len(log["rand"]), len(log["multinomial"])]})
        seq.num_cached_tokens = (S - 1) + len(draft)
        return logits[:, :-1, :]

    class FakeBM:
        def may_append_batch(self, seq, n):
            pass

        def trim_kv_only_fast(self, seq, n):
            seq.num_cached_tokens -= n

    sp = SamplingParams(temperature=temperature, max_tokens=max_tokens, decode_strategy="jacobi", jacobi_block_len=L,
                        jacobi_max_iterations=max_iters)
    seq = Sequence(prompt, sp)
    seq.num_cached_tokens = len(prompt)
    dec = ng.JacobiDecoderNonGreedy(FakeBM(), forward_step=forward_step, eos_token_id=eos, pad_token_id=0, vocab_size=V,
                                    device=torch.device("cpu"))
    real = ng.torch
    ng.torch = _TorchProxy(log)
    torch.manual_seed(torch_seed)
    try:
        accepted = dec.generate_chunk(seq)
    finally:
        ng.torch = real
    out = {"case": name, "mode": "engine_nongreedy", "model": model_name, "model_seed": seed, "dtype": "float32",
           "prompt": prompt, "L": L, "max_tokens": max_tokens, "max_iterations": max_iters, "eos": eos,
           "temperature": temperature, "accepted": accepted, "iters": dec.stats["num_jacobi_iterations"],
           "rand": log["rand"], "multinomial": log["multinomial"], "randint": log["randint"], "forwards": forwards,
           "reference": "hao-ai-lab/JacobiForcing @ 028dbd6, inference_engine JacobiDecoderNonGreedy run unmodified on "
                        "CPU, random draws recorded"}
    path = os.path.join(GOLDEN_DIR, name + ".json")
    with open(path, "w") as f:
        json.dump(out, f, separators=(",", ":"))
    print(f"{name}: {len(accepted)} tokens in {out['iters']} iterations, {len(log['rand'])} accept draws, "
          f"{len(log['multinomial'])} multinomial draws, {len(log['randint'])} pad tokens -> {os.path.relpath(path, ROOT)}")
    return out


def main():
    base = run_case("ng_L8_t1", "tiny", 0, 24, 8, 40, 1.0)
    run_case("ng_L16_t07_tiny128", "tiny128", 3, 40, 16, 48, 0.7)
    run_case("ng_L8_t03_sharp", "tiny", 0, 24, 8, 40, 0.05)          # near-greedy: long accepted runs
    run_case("ng_L8_maxiter", "tiny", 0, 24, 8, 200, 1.0, max_iters=9)
    gen = base["accepted"]
    eos = next((t for i, t in enumerate(gen) if i >= 6), None)
    if eos is not None:
        run_case("ng_L8_eos", "tiny", 0, 24, 8, 40, 1.0, eos=eos)


if __name__ == "__main__":
    main()
