"""Developer diagnostic: logits error of the 2-layer 7B-width model vs HF fp32 per row count, and prefix consistency."""
import numpy as np, torch, sys
sys.path.insert(0, ".")
from tests.gpu_helpers import make_pair
from oracle.hf_model import teacher_forced_logits
model, rt = make_pair("qwen2.5-coder-7b", model_overrides=dict(num_hidden_layers=2), max_rows=512, max_seqs=2, num_kv_pages=8, max_pages_per_seq=4)
g = torch.Generator().manual_seed(1)
ids_all = torch.randint(0, model.config.vocab_size, (256,), generator=g).tolist()
ref_all = teacher_forced_logits(model, ids_all).numpy()
m16 = model.to(torch.bfloat16); ref16 = teacher_forced_logits(m16, ids_all).numpy(); model.to(torch.float32)
outs = {}
for rows in (32, 72, 128, 136, 200, 256):
    ids = ids_all[:rows]
    seq = rt.new_sequence(); rt.ensure_capacity(seq, rows)
    res = rt.forward(ids, [(0, rows, 0, 0)], [seq], 0, rows, want_logits=True, want_hidden=True)
    rt.release(seq)
    err = np.abs(res["logits"] - ref_all[:rows]); e16 = np.abs(ref16[:rows] - ref_all[:rows])
    outs[rows] = res
    print(f"rows={rows}: dev max {err.max():.4g} mean {err.mean():.4g} | hf16 max {e16.max():.4g} mean {e16.mean():.4g} | per-row-mean last row dev {err[-1].mean():.4g} hf16 {e16[-1].mean():.4g}")
    if rows > 32:
        a = outs[32]["logits"]; b = res["logits"][:32]
        print("   first 32 rows equal to the 32-row forward:", bool((a == b).all()), "max diff", float(np.abs(a-b).max()), "hidden equal", bool((outs[32]["hidden"] == res["hidden"][:32]).all()))
# per-row error profile of the 256 row forward
err = np.abs(outs[256]["logits"] - ref_all); e16 = np.abs(ref16 - ref_all)
for r in (0, 16, 31, 64, 100, 127, 128, 129, 160, 200, 255):
    print(f"row {r}: dev mean {err[r].mean():.4g} hf16 mean {e16[r].mean():.4g}")
