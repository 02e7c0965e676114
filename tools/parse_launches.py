"""Summarise an ncu launch list (gpu__time_duration.sum per launch) by kernel and grid size."""
import collections, csv, re, sys
path = sys.argv[1]
with open(path) as f:
    lines = [l for l in f if not l.startswith('==')]
rows = list(csv.DictReader(lines))
agg = collections.defaultdict(lambda: [0, 0.0])
agg2 = collections.defaultdict(lambda: [0, 0.0])
for r in rows:
    n = re.sub(r'\(.*', '', r['Kernel Name']).replace('attndm::', '').replace('void ', '')
    v = float(r['Metric Value'].replace(',', ''))
    agg[n][0] += 1; agg[n][1] += v
    agg2[(n, r['Grid Size'])][0] += 1; agg2[(n, r['Grid Size'])][1] += v
tot = sum(v[1] for v in agg.values())
print(f"launches {len(rows)}  total {tot/1e6:.3f} ms")
for k, (c, v) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:20]:
    print(f"{k[:60]:60s} n={c:4d} total={v/1000:8.1f}us avg={v/c/1000:7.2f}us share={v/tot*100:5.1f}%")
print("--- by (kernel, grid), top 30")
for k, (c, v) in sorted(agg2.items(), key=lambda kv: -kv[1][1])[:30]:
    print(f"{k[0][:44]:44s} grid={k[1]:16s} n={c:3d} total={v/1000:8.1f}us avg={v/c/1000:7.2f}us")
