import sys, collections
rows=[tuple(map(int,l.split())) for l in open(sys.argv[1])]
khz=1965.0
# print deltas for the first two layers' worth of marks
prev=None; out=[]
for i,(mid,clk) in enumerate(rows[:120]):
    d = (clk-prev)/khz if prev is not None else 0.0
    out.append(f"{mid}:{d:.2f}")
    prev=clk
print(" ".join(out))
# aggregate by transition
agg=collections.defaultdict(list)
for (a,ca),(b,cb) in zip(rows,rows[1:]):
    agg[(a,b)].append((cb-ca)/khz)
for k,v in sorted(agg.items()):
    print(k, f"n={len(v)} mean={sum(v)/len(v):.2f} us")
