"""developer aid: per-forward durations from the coarse launch marks (JF_DEBUG_SKIP=2048 JF_PROF_MARKS=file)."""
import sys
rows = [tuple(map(int, l.split())) for l in open(sys.argv[1])]
khz = 1965.0
out = []
for (a, ca), (b, cb) in zip(rows, rows[1:]):
    out.append((a, (cb - ca) / khz))
print(" ".join(f"{a}:{d:.0f}" for a, d in out))
print("sum ms", sum(d for _, d in out) / 1000, "n", len(out))
