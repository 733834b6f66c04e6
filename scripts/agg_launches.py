"""Aggregate an ncu gpu__time_duration launch list (csv) by kernel name."""
import csv, collections, re, sys
fn = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
rows = [r for r in csv.reader(open(fn)) if len(r) > 14 and r[0].isdigit()]
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows:
  name = re.sub(r'\(.*', '', r[4])[:100]
  agg[name][0] += 1; agg[name][1] += float(r[14]) / 1e3
tot = sum(v[1] for v in agg.values())
print(f"{fn}: {len(rows)} launches, total {tot:.1f} us")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
  print(f"{v[1]:10.1f} us {v[0]:5d} {100*v[1]/tot:5.1f}%  {k}")
