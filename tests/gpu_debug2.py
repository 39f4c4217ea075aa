"""Diagnostic (GPU box): 7B-shaped single/two-layer model, rows > 256, multi-page contexts, vs the HF fp32 forward."""
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from tests.gpu_helpers import bf16_bits_to_float, make_pair   # noqa: E402
from oracle.hf_model import teacher_forced_logits              # noqa: E402


def stats(tag, dev, ref):
    err = np.abs(dev - ref)
    print(f"   {tag}: nan={int(np.isnan(dev).sum())} max|d|={np.nanmax(err):.4g} mean|d|={np.nanmean(err):.3g} "
          f"ref max {np.abs(ref).max():.3g}; argmax agree {(dev.argmax(-1) == ref.argmax(-1)).mean():.3f}", flush=True)
    rows = np.where(np.nanmax(err, -1) > 0.1 * max(1.0, np.abs(ref).max()))[0]
    if len(rows):
        print("      bad rows:", rows[:40].tolist(), "... total", len(rows), flush=True)


def run(name, layers, vocab=None, P_list=(40, 300), **rt_kw):
    over = {"num_hidden_layers": layers}
    if vocab:
        over["vocab_size"] = vocab
    t0 = time.time()
    kw = dict(max_rows=576, max_seqs=2, num_kv_pages=8, max_pages_per_seq=4)
    kw.update(rt_kw)
    model, rt = make_pair(name, model_overrides=over, **kw)
    print(f"[{name} layers={layers} vocab={model.config.vocab_size}] built in {time.time() - t0:.1f}s", flush=True)
    V = model.config.vocab_size
    g = torch.Generator().manual_seed(1)
    for P in P_list:
        ids = torch.randint(0, V, (P,), generator=g).tolist()
        ref = teacher_forced_logits(model, ids).numpy()
        seq = rt.new_sequence(); rt.ensure_capacity(seq, P + 16)
        lm0 = max(0, ((P - 200) // 8) * 8)
        res = rt.forward(ids, [(0, P, 0, 0)], [seq], lm0, P - lm0, want_logits=True)
        print(f"  full forward P={P} ({res['ms']:.2f} ms), logits rows [{lm0},{P})", flush=True)
        stats("in-flight", res["logits"], ref[lm0:])
        # commit the first P-8 rows, verify the last 8 against committed pages (crosses a page when P-8 > 256)
        seq2 = rt.new_sequence(); rt.ensure_capacity(seq2, P + 16)
        C0 = P - 8
        rt.forward(ids[:C0], [(0, C0, 0, 0)], [seq2], 0, 0, commits=[(i, 0, i) for i in range(C0)])
        res2 = rt.forward(ids[C0:], [(0, 8, 0, C0)], [seq2], 0, 8, want_logits=True)
        stats(f"tail 8 rows over {C0} committed keys", res2["logits"], ref[C0:])
        print(f"      bitwise equal to in-flight: {np.array_equal(res2['logits'], res['logits'][C0 - lm0:])}", flush=True)
        rt.release(seq); rt.release(seq2)
    rt.close()


if __name__ == "__main__":
    run("tiny128", 3, P_list=(300, 544))
    run("qwen2.5-coder-7b", 1, vocab=4096)
    run("qwen2.5-coder-7b", 2, vocab=4096, P_list=(544,))
    run("qwen2.5-coder-7b", 1, P_list=(40,))
    print("DEBUG2 DONE", flush=True)
