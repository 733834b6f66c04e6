"""Summarise an ncu launch list (--metrics gpu__time_duration.sum --csv): the last occurrence of a full step (from the last launch
of the FIRST kernel name given) aggregated per kernel.   python scripts/launch_summary.py file.csv [first-kernel-substring]"""
import collections, csv, sys
rows = list(csv.reader(open(sys.argv[1])))
first = sys.argv[2] if len(sys.argv) > 2 else "conv_small_tc_kernel<3, 5"
hdr, out = None, []
for r in rows:
  if "Kernel Name" in r:
    hdr = r; continue
  if hdr is None or len(r) != len(hdr):
    continue
  d = dict(zip(hdr, r))
  if d.get("Metric Name") != "gpu__time_duration.sum":
    continue
  out.append((d["Kernel Name"], float(d["Metric Value"].replace(",", "")) / 1e3))
idx = [i for i, o in enumerate(out) if first in o[0]]
nper = int(sys.argv[3]) if len(sys.argv) > 3 else 1
s = idx[-nper]
agg = collections.OrderedDict()
for k, t in out[s:]:
  a = agg.setdefault(k[:100], [0, 0.0]); a[0] += 1; a[1] += t
tot = sum(v[1] for v in agg.values())
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
  print(f"{v[1]:9.1f} us {100 * v[1] / tot:5.1f}%  n={v[0]:3d}  {k}")
print(f"{tot:9.1f} us total, {len(out) - s} launches")
