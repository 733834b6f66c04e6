"""Top stalled SASS instructions of each kernel in an `ncu --page source --print-source sass --csv` dump."""
import csv, sys
fn = sys.argv[1]; which = int(sys.argv[2]) if len(sys.argv) > 2 else -1; top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
kernels = []; cur = None
for r in csv.reader(open(fn)):
  if r and r[0] == "Kernel Name": cur = {"name": r[1], "hdr": None, "rows": []}; kernels.append(cur); continue
  if cur is None: continue
  if cur["hdr"] is None: cur["hdr"] = r; continue
  cur["rows"].append(r)
for ki, k in enumerate(kernels):
  if which >= 0 and ki != which: continue
  h = k["hdr"]; iS = h.index("# Samples"); stall_cols = [i for i, n in enumerate(h) if n.startswith("stall_") and "Not Issued" not in n]
  tot = sum(int(r[iS]) for r in k["rows"])
  print(f"== kernel {ki}: {k['name']}  total samples {tot}")
  agg = {}
  for i in stall_cols: agg[h[i]] = sum(int(r[i]) for r in k["rows"])
  print("   by reason:", {n: v for n, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v > tot * 0.01})
  idx = sorted(range(len(k["rows"])), key=lambda i: -int(k["rows"][i][iS]))[:top]
  for i in sorted(idx):
    r = k["rows"][i]
    reasons = {h[c][6:]: int(r[c]) for c in stall_cols if int(r[c]) > 0}
    rs = " ".join(f"{n}={v}" for n, v in sorted(reasons.items(), key=lambda kv: -kv[1])[:3])
    print(f"   #{i:5d} {int(r[iS]):6d} {100*int(r[iS])/tot:5.1f}%  {r[1].strip()[:70]:70s} {rs}")
