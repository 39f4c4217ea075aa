"""Drivers around the device-resident loop that mirror the reference's user-level scripts:

  profile_rows(...)      the per-prompt benchmark loop of JacobiForcing/jacobi_forcing_inference_MR_humaneval.py:152-273
                         with its CSV schema (index, task_id, prompt_tokens, new_tokens, calls, total_iterations,
                         avg_iter_per_call, avg_iter_per_token, time_sec, toks_per_sec, stop_reason)
  jacobi_stream(...)     the block-by-block streaming generator of applications/jacobi_streaming_driver.py:7-193: one
                         launch per n-token block, `on_tokens` called with the newly accepted ids of each block

Both take prompts as token ids (tokenizers are the caller's business) and an explicit RNG for the random drafts.
"""
from __future__ import annotations

import random
import time
from typing import Callable, Dict, Iterator, List, Optional, Sequence

from .policy import MODES, PolicyKnobs, draw_prefill_draft
from .runtime import JacobiRuntime, hf_generate

CSV_COLUMNS = ["index", "task_id", "prompt_tokens", "new_tokens", "calls", "total_iterations", "avg_iter_per_call",
               "avg_iter_per_token", "time_sec", "toks_per_sec", "stop_reason"]


def profile_rows(rt: JacobiRuntime, prompts: Sequence[Sequence[int]], *, n: int = 64, max_new_tokens: int = 1024,
                 K: int = 2, r: float = 0.85, n_gram_pool_size: int = 4, lookahead_start_ratio: float = 0.0,
                 eos: Optional[int] = None, pad: Optional[int] = 0, mode: str = "multiblock", max_calls: int = 1024,
                 task_ids: Optional[Sequence[str]] = None, seed: int = 0) -> List[Dict]:
    """One row per prompt, the quantities and definitions of MR_humaneval.py:242-273: `toks_per_sec` =
    new_tokens / generation-only time (prefill excluded), `new_tokens` excludes the prefill call's token."""
    rows = []
    for idx, prompt in enumerate(prompts):
        t0 = time.time()
        seqs, results, run = hf_generate(rt, [prompt], n=n, max_new_tokens=max_new_tokens, mode=mode, K=K, r=r,
                                         lookahead_start_ratio=lookahead_start_ratio,
                                         n_gram_pool_size=n_gram_pool_size, eos=eos, pad=pad, max_calls=max_calls,
                                         rngs=[random.Random(seed + idx)])
        res = results[0]
        dt = time.time() - t0
        new_tokens = res.total_new_tokens - 1                      # "subtract prefill" (:242-243)
        gen_time = run["device_ms"] / 1e3
        rows.append({
            "index": idx, "task_id": (task_ids[idx] if task_ids else str(idx)), "prompt_tokens": len(prompt),
            "new_tokens": new_tokens, "calls": res.calls, "total_iterations": res.total_iterations,
            "avg_iter_per_call": res.total_iterations / max(1, res.calls),
            "avg_iter_per_token": res.total_iterations / max(1, new_tokens),
            "time_sec": dt, "toks_per_sec": new_tokens / gen_time if gen_time > 0 else 0.0,
            "stop_reason": res.stop_reason, "token_ids": res.tokens,
        })
        for s in seqs:
            rt.release(s)
    return rows


def jacobi_stream(rt: JacobiRuntime, prompt: Sequence[int], *, n: int = 64, max_new_tokens: int = 512, K: int = 2,
                  r: float = 0.8, n_gram_pool_size: int = 4, eos: Optional[int] = None, pad: Optional[int] = 0,
                  on_tokens: Optional[Callable[[List[int]], None]] = None, max_calls: int = 128,
                  rng: Optional[random.Random] = None) -> Iterator[List[int]]:
    """Yields the accepted ids of every n-token block as soon as its fixed point is reached (the reference streams
    after each `jacobi_forward_greedy_multiblock` call, jacobi_streaming_driver.py:61-160)."""
    rng = rng or random.Random(0)
    prompt = [int(t) for t in prompt]
    seq = rt.new_sequence()
    first, drafted, _ = rt.prefill(seq, prompt, draw_prefill_draft(rng, prompt, n))
    generated = list(prompt)
    calls, total = 1, 0
    try:
        while total < max_new_tokens and calls < max_calls:
            if eos is not None and eos in generated[len(prompt):]:
                break
            draft = list(drafted) if calls == 1 else [first] + [rng.choice(generated) for _ in range(n - 1)]
            knobs = PolicyKnobs(mode=MODES["multiblock"], n=n, K=K, r=r, n_gram_pool_size=n_gram_pool_size, eos=eos,
                                pad=pad, max_new_tokens=n, max_calls=2, vocab=rt.vocab)
            results, _ = rt.jacobi_run([seq], knobs, [first], [draft])       # exactly one block: calls 1 -> 2
            toks = results[0].tokens
            if not toks:
                break
            generated += toks
            total += len(toks)
            first = results[0].next_token
            calls += 1
            if on_tokens:
                on_tokens(toks)
            yield toks
    finally:
        rt.release(seq)
