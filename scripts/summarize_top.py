"""Summarise gpurun_out/prof_top.ncu-rep (scripts/prof_top.py under `ncu --set full`) into profiles/: a text table of the
metrics quoted in DESIGN.md and the JSON that bench.py reads for `roofline.traffic` / tensor-pipe activity."""
import csv, json, re, subprocess, sys
rep = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/prof_top.ncu-rep"
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines())); h, u, data = rows[0], rows[1], rows[2:]
want = ["Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "sm__cycles_elapsed.max", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "smsp__inst_executed.sum"]
out = [{w: (r[h.index(w)], u[h.index(w)]) for w in want if w in h} for r in data]
with open("profiles/r1_ncu_top_kernels_full.txt", "w") as f:
  f.write("# ncu --set full --clock-control none, scripts/prof_top.py (KITTI sizes, L2 flushed before every launch), third repetition\n")
  for d in out:
    f.write("\n")
    for w in want:
      if w in d: f.write(f"{w:95s} {d[w][0]} {d[w][1]}\n")
def nbytes(d, k):
  v, unit = d[k]
  return float(v) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[unit]
names = ["conv2d_dil1", "conv2d_dil8", "conv3d", "wgrad2d", "wgrad3d", "cost_volume_b1", "cost_volume_b8", "taps27", "tapsum_softargmin", "first_conv5x5s2", "refine_in_conv"]
summ = {}
for n, d in zip(names, out):
  summ[n] = {"kernel": re.sub(r"\(.*", "", d["Kernel Name"][0]), "us": float(d["gpu__time_duration.sum"][0]),
             "dram_bytes": nbytes(d, "dram__bytes_read.sum") + nbytes(d, "dram__bytes_write.sum"),
             "tensor_pipe_active_pct": float(d["sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"][0]),
             "tensor_pipe_elapsed_pct": float(d["sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed"][0]),
             "dram_pct": float(d["gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"][0])}
  print(n, summ[n])
json.dump(summ, open("profiles/r1_ncu_summary.json", "w"), indent=1)
