"""Developer aid: one forward of the tiny model with the tcgen05 attention path, compared with the mma.sync path."""
import os, sys, time
sys.path.insert(0, ".")
import numpy as np, torch
from tests.gpu_helpers import make_pair
name = sys.argv[1] if len(sys.argv) > 1 else "tiny"
rows = int(sys.argv[2]) if len(sys.argv) > 2 else 8
os.environ["JF_ATTN_TC"] = "0"
model, rt0 = make_pair(name)
g = torch.Generator().manual_seed(1)
ids = torch.randint(0, model.config.vocab_size, (rows,), generator=g).tolist()
seq = rt0.new_sequence(); rt0.ensure_capacity(seq, rows)
a = rt0.forward(ids, [(0, rows, 0, 0)], [seq], 0, rows, want_logits=True)
rt0.close()
os.environ["JF_ATTN_TC"] = "1"
model, rt1 = make_pair(name)
seq = rt1.new_sequence(); rt1.ensure_capacity(seq, rows)
t0 = time.time()
try:
    b = rt1.forward(ids, [(0, rows, 0, 0)], [seq], 0, rows, want_logits=True)
except Exception as e:
    print("TC forward failed after %.1f s:" % (time.time() - t0), e); sys.exit(1)
d = np.abs(a["logits"] - b["logits"])
print(f"{name} rows={rows}: max |dlogit| tc vs mma.sync {d.max():.4g}, mean {d.mean():.4g}, greedy equal {(a['greedy'] == b['greedy']).all()}")
