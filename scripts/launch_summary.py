"""Aggregate an ncu launch-list CSV (gpu__time_duration.sum) of `bench.py --steps 2 --warmup 1`: the last COMPLETE
train step in the list (a step starts with the statistics memsets just before timestep_sinusoid_kernel)."""
import collections, csv, sys
path = sys.argv[1]
detail = sys.argv[2] if len(sys.argv) > 2 else ""
with open(path) as f:
    lines = [l for l in f if l.startswith('"')]
rows = list(csv.DictReader(lines))
names = [r["Kernel Name"] for r in rows]
idx = [i for i, x in enumerate(names) if "timestep_sinusoid" in x]
step = rows[idx[-2]:idx[-1]] if len(idx) >= 2 else rows[idx[-1]:]
agg = collections.defaultdict(lambda: [0, 0.0])
det = collections.defaultdict(lambda: [0, 0.0])
for r in step:
    nm = r["Kernel Name"]
    short = nm.split("(")[0].replace("void ", "").replace("<unnamed>::", "")[:48]
    v = float(r["Metric Value"].replace(",", ""))
    u = r["Metric Unit"]
    v = v / 1e3 if u in ("ns", "nsecond") else v * 1e3 if u in ("ms", "msecond") else v
    agg[short][0] += 1; agg[short][1] += v
    if detail and detail in nm:
        det[(short, r["Grid Size"])][0] += 1; det[(short, r["Grid Size"])][1] += v
tot = sum(v[1] for v in agg.values())
print(f"total {tot:.1f} us over {len(step)} launches")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{v[1]:10.1f} us {v[0]:5d}  avg {v[1] / v[0]:8.1f}  {100 * v[1] / tot:5.1f}%  {k}")
if detail:
    for k, v in sorted(det.items(), key=lambda kv: -kv[1][1])[:40]:
        print(f"   {v[1]:9.1f} us n={v[0]:3d} avg {v[1] / v[0]:7.1f}  {k}")
