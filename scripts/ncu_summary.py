"""Summarise an `ncu --set full` report (last launch of every kernel of interest) into profiles/r2_ncu_summary.json.
  ncu -i gpurun_out/r2_prof_final.ncu-rep --page raw --csv > /tmp/raw.csv ; python scripts/ncu_summary.py /tmp/raw.csv profiles/r2_ncu_summary.json
Keys bench.py reads: <kernel>.dram_bytes (read + write, per launch) and <kernel>.tensor_pipe_active_pct."""
import csv, json, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
NAMES = [  # (key, substring of the demangled kernel name, occurrence selector among the launches in capture order)
  ("conv2d_ws_dil1", "conv_c32_ws_kernel<0", "big0"), ("conv2d_ws_dil4", "conv_c32_ws_kernel<0", "big1"),
  ("conv2d_ws_feature_block", "conv_c32_ws_kernel<0", "small"),
  ("conv3d_ws", "conv_c32_ws_kernel<1", "last"), ("conv5x5s2_p4_ws_big", "conv_c32_ws_kernel<2", "first"), ("conv5x5s2_p4_ws_strided", "conv_c32_ws_kernel<2", "last"),
  ("conv3d_tma_h_crosscheck", "conv3d_c32_tma_kernel", "last"),
  ("head_fused", "conv3d_out_softargmin_kernel", "last"),
  ("cost_volume_b1_direct", "cost_volume_fwd_direct_kernel", "last"), ("cost_volume_b8_staged", "cost_volume_fwd_kernel", "last"),
  ("first_conv5x5s2", "conv_small_tc_kernel<3", "last"), ("refine_in_conv_im2col", "conv_small_tc_kernel<4", "last"),
  ("refine_in_pack", "refine_pack_input_kernel", "last"), ("refine_in_conv_c4_ws", "conv_c32_ws_kernel<3", "last"),
  ("head_taps27", "conv_c32_taps_tc_kernel<27", "last"), ("head_tapsum_softargmin", "tapsum_softargmin_kernel", "last"),
  ("refine_out_taps9", "conv_c32_taps_tc_kernel", "last"), ("refine_out_tapsum", "tapsum_refine_out_kernel", "last"),
  ("upsample", "upsample_bilinear_kernel", "last"), ("wgrad_tc", "conv_c32_wgrad_tc_kernel", "wg"),
]
def f(r, name):
  if name not in idx:
    return None
  v = r[idx[name]].replace(",", "")
  try:
    return float(v)
  except ValueError:
    return None
def unit(name):
  return units[idx[name]] if name in idx else ""
def scale_bytes(v, u):
  return None if v is None else v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
out = {"source": "ncu --set full --clock-control none, scripts/prof_r2.py (KITTI sizes, L2 flushed before every launch, third repetition)"}
data = rows[2:]
for key, sub, sel in NAMES:
  ls = [r for r in data if sub in r[idx["Kernel Name"]]]
  if not ls:
    continue
  dur = lambda r: f(r, "gpu__time_duration.sum")
  if sel == "last":
    picks = [ls[-1]]
  elif sel == "first":
    picks = [ls[0]]
  elif sel == "small":
    picks = [min(ls, key=dur)]
  elif sel in ("big0", "big1"):
    big = [r for r in ls if dur(r) > 25.0]        # the full-resolution launches of the last repetition: dil 1 then dil 4
    picks = [big[-2] if sel == "big0" else big[-1]] if len(big) >= 2 else [ls[-1]]
  elif sel == "wg":
    picks = ls[-2:]
  for n, r in enumerate(picks):
    k = key if len(picks) == 1 else f"{key}_{'2d' if n == 0 else '3d'}"
    cyc = f(r, "sm__cycles_elapsed.max")
    hm = f(r, "TPC.TriageCompute.sm__pipe_tensor_subpipe_hmma_cycles_active_realtime.avg")
    e = {
      "kernel": r[idx["Kernel Name"]][:90], "grid": r[idx["Grid Size"]] if "Grid Size" in idx else None,
      "duration_us": dur(r), "sm_cycles": cyc,
      "dram_bytes": (scale_bytes(f(r, "dram__bytes_read.sum"), unit("dram__bytes_read.sum")) or 0) + (scale_bytes(f(r, "dram__bytes_write.sum"), unit("dram__bytes_write.sum")) or 0),
      "dram_read_bytes": scale_bytes(f(r, "dram__bytes_read.sum"), unit("dram__bytes_read.sum")),
      "dram_write_bytes": scale_bytes(f(r, "dram__bytes_write.sum"), unit("dram__bytes_write.sum")),
      "tensor_pipe_active_pct": f(r, "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed"),
      "tensor_pipe_active_pct_of_active_cycles": f(r, "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
      "tensor_pipe_in_flight_pct": None if (hm is None or not cyc) else round(50.0 * hm / cyc, 1),   # realtime counter is per TPC (2 SMs)
      "tensor_mem_active_pct": f(r, "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed"),
      "l1tex_throughput_pct": f(r, "l1tex__throughput.avg.pct_of_peak_sustained_elapsed"),
      "lts_throughput_pct": f(r, "lts__throughput.avg.pct_of_peak_sustained_elapsed"),
      "sm_throughput_pct": f(r, "sm__throughput.avg.pct_of_peak_sustained_elapsed"),
      "smem_lsu_wavefronts": f(r, "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"),
      "smem_bank_conflicts": f(r, "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"),
      "warp_instructions": f(r, "smsp__inst_executed.sum"),
      "registers_per_thread": f(r, "launch__registers_per_thread"),
      "warps_active_pct": f(r, "sm__warps_active.avg.pct_of_peak_sustained_active"),
    }
    out[k] = e
json.dump(out, open(sys.argv[2], "w"), indent=1)
for k, e in out.items():
  if isinstance(e, dict):
    print(f"{k:28s} {e['duration_us']:7.1f} us  dram {e['dram_bytes'] / 1e6:7.1f} MB (r {e['dram_read_bytes'] / 1e6:6.1f} w {e['dram_write_bytes'] / 1e6:6.1f})  "
          f"tensor pipe active {e['tensor_pipe_active_pct']:5.1f} % (in flight {e['tensor_pipe_in_flight_pct']})  l1tex {e['l1tex_throughput_pct']:5.1f} %  "
          f"lts {e['lts_throughput_pct']:5.1f} %  smem wavefronts {e['smem_lsu_wavefronts']:.3g}  warp instr {e['warp_instructions']:.3g}  regs {e['registers_per_thread']:.0f}")
