"""Developer aid: per-stage timeline from the fine-grained marks (JF_PROF_MARKS): for every grid-barrier release (901), the
offsets of the marks that follow it until the next barrier arrival (900)."""
import sys, collections
rows = [tuple(map(int, l.split())) for l in open(sys.argv[1])]
khz = float(sys.argv[2]) if len(sys.argv) > 2 else 1880.0
seqs = collections.defaultdict(list)
cur, t0 = None, None
for mid, clk in rows:
    if mid == 901:
        cur, t0 = [], clk
        continue
    if cur is None:
        continue
    cur.append((mid, (clk - t0) / khz))
    if mid == 900:
        key = tuple(m for m, _ in cur)
        seqs[key].append([t for _, t in cur])
        cur = None
for key, lst in sorted(seqs.items(), key=lambda kv: -len(kv[1])):
    n = len(lst)
    means = [sum(x[i] for x in lst) / n for i in range(len(key))]
    print(f"n={n:3d} " + "  ".join(f"{m}@{t:.2f}" for m, t in zip(key, means)))
