#!/usr/bin/env python
"""bench.py — StereoNet forward (BASELINE.json configs[1]: KITTI 376x1248, D=192, k=3, batch 1 per GPU).

  python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path (one process per GPU under torchrun)
  python bench.py --impl reference --steps K --warmup W    # the reference algorithm on the host CPU cores (oracle port)

A step = one pass of the hot path over one synthetic stereo pair per GPU: feature_net(L), feature_net(R),
stereo_net(L, fl, fr, 'l').  Prints ONE JSON line on rank 0 (contract in the task statement / DESIGN.md §6).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "adaptive-stereo-icra-2021_b200")
for p in (ROOT, PKG):
  if p not in sys.path:
    sys.path.insert(0, p)

import torch  # noqa: E402

H, W, K_DOWN, MAXDISP = 376, 1248, 3, 192
METRIC = "StereoNet pairs/s @KITTI 376x1248 D=192"
WORKLOAD = "StereoNet k=3 D=192 fp32 forward, KITTI 1x3x376x1248 per GPU (BASELINE.json configs[1])"


def peaks():
  path = os.path.join(ROOT, "MEASURED_PEAKS.json")
  if os.path.exists(path):
    d = json.load(open(path))
    return dict(hbm=d["hbm_gbs"], tensor=d["bf16_tflops"], tensor_sustained=d.get("bf16_tflops_sustained"), src="measured")
  return dict(hbm=6650.0, tensor=1590.0, tensor_sustained=1400.0, src="fallback")


# ------------------------------------------------------------------------------------------------- synthetic data
def synthetic_pair(seed, h=H, w=W, batch=1):
  """Textured left image, ground-plane disparity ramp, right = left warped by the known disparity.
  Plain torch on the host; no oracle import on the product path."""
  import torch.nn.functional as F
  g = torch.Generator().manual_seed(seed)
  noise = torch.randn((batch, 3, h, w), generator=g)
  low = F.avg_pool2d(F.avg_pool2d(noise, 9, 1, 4), 9, 1, 4) * 6.0
  left = (0.5 + 0.25 * low + 0.1 * torch.randn((batch, 3, h, w), generator=g)).clamp(0, 1)
  ramp = 4.0 + 60.0 * torch.arange(h, dtype=torch.float32) / (h - 1)
  rows, cols = torch.meshgrid(torch.arange(h), torch.arange(w), indexing="ij")
  flow = torch.stack([cols, rows], dim=-1).float().expand(batch, -1, -1, -1).clone()
  flow[..., 0] = flow[..., 0] + ramp.view(1, h, 1)
  flow[..., 0] = (2 * flow[..., 0] / w) - 1.0
  flow[..., 1] = (2 * flow[..., 1] / h) - 1.0
  right = F.grid_sample(left, flow, mode="bilinear", padding_mode="border", align_corners=False)
  return left.contiguous(), right.contiguous()


# ------------------------------------------------------------------------------------------------- clocks sampler
class ClockSampler:
  Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
      "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

  def __init__(self, index):
    self.index, self.rows, self.proc = index, [], None

  def start(self):
    try:
      self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                    "--format=csv,noheader,nounits", "-lms", "25"], stdout=subprocess.PIPE, text=True)
      self.thread = threading.Thread(target=self._read, daemon=True)
      self.thread.start()
    except Exception:
      self.proc = None

  def _read(self):
    for line in self.proc.stdout:
      self.rows.append([c.strip() for c in line.split(",")])

  def stop(self):
    if self.proc is None:
      return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
    self.proc.terminate()
    try:
      self.proc.wait(timeout=2)
    except Exception:
      self.proc.kill()
    sm, mx, reasons = [], None, set()
    names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
    for r in self.rows:
      try:
        sm.append(float(r[0])); mx = float(r[1])
        for n, v in zip(names, r[3:7]):
          if v.lower().startswith("active"):
            reasons.add(n)
      except Exception:
        pass
    sm.sort()
    # under load = upper half of the samples (the sampler also sees the idle gaps around the region)
    med = sm[(len(sm) * 3) // 4] if sm else None
    return {"sm_mhz": med, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------- reference arm (CPU)
def cpu_reference(steps, warmup, budget_s):
  """The reference algorithm (oracle port of adaptive_stereo/models/stereo_net.py) on the host cores, eval/no-grad,
  same shapes and seeded weights as the GPU arm.  Each step = one KITTI-sized pair."""
  sys.path.insert(0, os.path.join(ROOT, "oracle"))
  import stereonet_oracle as O
  cores = os.cpu_count() or 1
  try:
    cores = len(os.sched_getaffinity(0))
  except Exception:
    pass
  torch.set_num_threads(cores)
  fsd, ssd = O.make_feature_state(K_DOWN, 11), O.make_stereo_state(22, sharpen=40.0)
  left, right = synthetic_pair(1000)
  done, t_total = 0, 0.0
  with torch.no_grad():
    for _ in range(warmup):
      O.predict_disparity_left(fsd, ssd, left, right, K_DOWN)
    t_begin = time.perf_counter()
    while done < steps:
      t0 = time.perf_counter()
      O.predict_disparity_left(fsd, ssd, left, right, K_DOWN)
      t_total += time.perf_counter() - t0
      done += 1
      if time.perf_counter() - t_begin > budget_s:
        break
  return dict(value=done / t_total, unit="pairs/s", cores=cores, kind="port", steps=done, ms_per_step=1e3 * t_total / done,
              sample=f"{done} full forward passes of the oracle port (torch CPU fp32, {cores} threads) on one synthetic KITTI pair")


def cpu_reference_adapt(steps):
  """One adaptation step of the oracle port on the host cores (adapt.py:313-337,381-394 restated in oracle/)."""
  sys.path.insert(0, os.path.join(ROOT, "oracle"))
  import stereonet_oracle as O
  cores = torch.get_num_threads()
  fsd = O.clone_state(O.make_feature_state(K_DOWN, 11), True)
  ssd = O.clone_state(O.make_stereo_state(22, sharpen=10.0), True)
  left, right = synthetic_pair(1000)
  adam = {}
  O.adapt_step(fsd, ssd, left, right, K_DOWN, adam)
  t0 = time.perf_counter()
  for _ in range(steps):
    O.adapt_step(fsd, ssd, left, right, K_DOWN, adam)
  dt = (time.perf_counter() - t0) / steps
  return {"value": 1.0 / dt, "unit": "steps/s", "cores": cores, "kind": "port", "sample": f"{steps} adaptation steps of the oracle port"}


# ------------------------------------------------------------------------------------------------- GPU arm
def time_kernel(fn, iters, flush, stream):
  """Average duration (ms) of fn() alone, L2 flushed before every launch, CUDA events on the launching stream."""
  evs = []
  for _ in range(3):
    fn()
  for _ in range(iters):
    flush.zero_()                       # > L2 (126 MB): evicts the kernel's inputs/outputs
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(stream); fn(); b.record(stream)
    evs.append((a, b))
  torch.cuda.synchronize()
  ts = sorted(a.elapsed_time(b) for a, b in evs)
  return sum(ts) / len(ts), ts[len(ts) // 2]


def gpu_arm(args):
  import stereonet_b200 as S
  from stereonet_b200 import ops
  from stereonet_b200.autograd import fused
  from stereonet_b200.runtime import StereoEngine
  import torch.distributed as dist

  world = int(os.environ.get("WORLD_SIZE", "1"))
  rank = int(os.environ.get("RANK", "0"))
  local = int(os.environ.get("LOCAL_RANK", "0"))
  if world != args.gpus and world > 1:
    raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
  torch.cuda.set_device(local)
  dev = torch.device("cuda", local)
  if world > 1:
    dist.init_process_group("nccl", device_id=dev)

  torch.manual_seed(123)                                   # the reference's seed (adapt.py:29, train.py:141)
  fnet = S.FeatureExtractorNetwork(K_DOWN).to(dev).eval()  # seeded random init of the reference architecture
  snet = S.StereoNet(K_DOWN, 1, 0, maxdisp=MAXDISP).to(dev).eval()
  engine = StereoEngine(fnet, snet, output_cost_volume=True, use_graph=not args.no_graph)

  left, right = synthetic_pair(1000 + rank)
  shape = tuple(left.shape)
  left_pin, right_pin = left.pin_memory(), right.pin_memory()
  out_pin = torch.empty((1, 1, H, W), dtype=torch.float32).pin_memory()
  e = engine._entry(shape, dev)
  e["left"].copy_(left_pin); e["right"].copy_(right_pin)
  stream = torch.cuda.current_stream(dev)

  def barrier():
    torch.cuda.synchronize()
    if world > 1:
      dist.barrier()
    torch.cuda.synchronize()

  # ---- device-resident throughput: inputs already in HBM, K graph replays
  for _ in range(max(args.warmup, 3)):
    engine.run_static(shape, dev)
  sampler = ClockSampler(local)
  if rank == 0:
    sampler.start()
  barrier()
  ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  ev0.record(stream)
  for _ in range(args.steps):
    _, launches = engine.run_static(shape, dev)
  ev1.record(stream)
  barrier()
  ms_dev = ev0.elapsed_time(ev1)

  # ---- end to end through the public engine API: pinned host images in, host disparity out, every step
  # (a) streaming API: H2D of frame i+1 / forward of frame i / D2H of frame i-1 overlap (throughput = the headline e2e);
  # (b) blocking API: the three phases of one frame back to back (latency).
  for _ in range(3):
    engine.infer_host_async(left_pin, right_pin, out_pin)
  engine.synchronize()
  barrier()
  ev2, ev3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  ev2.record(stream)
  for _ in range(args.steps):
    engine.infer_host_async(left_pin, right_pin, out_pin)
  engine.synchronize()
  ev3.record(stream)
  barrier()
  ms_e2e = ev2.elapsed_time(ev3)
  ev6, ev7 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  ev6.record(stream)
  for _ in range(args.steps):
    engine.infer_host(left_pin, right_pin, out_pin)
  ev7.record(stream)
  barrier()
  ms_e2e_seq = ev6.elapsed_time(ev7)

  # ---- online-adaptation step (BASELINE.json configs[2]/[4]): train-mode fwd + photometric loss + bwd + clip + Adam.
  # N > 1: shared-model data parallel, one stream per rank, ONE flat-bucket NCCL all-reduce of the used gradients.
  ms_adapt, adapt_steps, adapt_launches = float("nan"), 0, 0
  if not args.skip_adapt:
    from stereonet_b200.adapt import AdaptStepper, make_optimizer
    from stereonet_b200 import parallel
    fnet.train(); snet.train()
    stepper = AdaptStepper(fnet, snet, make_optimizer(fnet, snet, lr=5e-5, capturable=not args.no_graph), H, W,
                           clip_grad_norm=True, use_graph=not args.no_graph)
    dl, dr = left.to(dev), right.to(dev)
    used = parallel.used_parameters(snet, fnet)
    dp = used if world > 1 else None
    adapt_steps = max(3, args.steps // 5)
    for _ in range(3):
      stepper.step(dl, dr, dp_params=dp)
    barrier()
    n0 = ops.LAUNCHES
    ev4, ev5 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev4.record(stream)
    for _ in range(adapt_steps):
      stepper.step(dl, dr, dp_params=dp)
    ev5.record(stream)
    barrier()
    ms_adapt = ev4.elapsed_time(ev5)
    adapt_launches = stepper.launches_per_step if stepper.launches_per_step else (ops.LAUNCHES - n0) // adapt_steps
    fnet.eval(); snet.eval()

  clocks = sampler.stop() if rank == 0 else None          # sampled across all three timed regions (device, e2e, adapt)
  t = torch.tensor([ms_dev, ms_e2e, ms_adapt], device=dev, dtype=torch.float64)
  if world > 1:
    dist.all_reduce(t, op=dist.ReduceOp.MAX)              # max over ranks
  ms_dev, ms_e2e, ms_adapt = t.tolist()

  result = None
  if rank == 0:
    pk = peaks()
    # ---- per-kernel roofline numbers, each kernel timed alone with an L2 flush before every launch
    flush = torch.empty(256 * 1024 * 1024 // 4, device=dev, dtype=torch.float32)
    hc, wc, D = (H - 1) // 8 + 1, (W - 1) // 8 + 1, (MAXDISP + 1) // 8
    with torch.no_grad():
      fl = torch.randn(1, hc, wc, 32, device=dev); fr = torch.randn(1, hc, wc, 32, device=dev)
      cv_ms, cv_med = time_kernel(lambda: ops.cost_volume(fl, fr, D), 20, flush, stream)
      cv_bytes = 4 * 32 * hc * wc * (2 + D)
      fl8 = torch.randn(8, hc, wc, 32, device=dev); fr8 = torch.randn(8, hc, wc, 32, device=dev)
      cv8_ms, _ = time_kernel(lambda: ops.cost_volume(fl8, fr8, D), 20, flush, stream)    # 195 MB > L2: fixed costs amortised
      x3 = torch.randn(1, D, hc, wc, 32, device=dev)
      conv3, bn3 = snet.filter[0][0][0], snet.filter[0][0][1]
      f3_ms, _ = time_kernel(lambda: fused.conv_bn_lrelu(x3, conv3, bn3, 1, False, False), 10, flush, stream)
      f3_flops = 2 * 27 * 32 * 32 * D * hc * wc
      x2 = torch.randn(1, H, W, 32, device=dev)
      blk = snet.edge_aware_refinements[0].residual_astrous_blocks[2]
      r2_ms, _ = time_kernel(lambda: blk.forward_cl(x2), 10, flush, stream)
      r2_flops = 2 * 9 * 32 * 32 * H * W
      taps = ops.conv_c32_taps(x3, snet.conv3d_alone.weight, 27)
      sa_ms, _ = time_kernel(lambda: ops.tapsum_softargmin(taps, snet.conv3d_alone.bias, True), 20, flush, stream)
      sa_bytes = 4 * hc * wc * (27 * D + D + 1)
      tp_ms, _ = time_kernel(lambda: ops.conv_c32_taps(x3, snet.conv3d_alone.weight, 27), 20, flush, stream)
      tp_bytes = 4 * D * hc * wc * (32 + 27)
      img2 = torch.rand(2, 3, H, W, device=dev)
      c0 = fnet.downsample[0]
      fc_ms, _ = time_kernel(lambda: ops.conv5x5s2_c3(img2, c0.weight, c0.bias), 20, flush, stream)
      fc_flops = 2 * 2 * 75 * 32 * ((H - 1) // 2 + 1) * ((W - 1) // 2 + 1)
      ref = snet.edge_aware_refinements[0]
      cz = torch.rand(1, hc, wc, device=dev) * 20
      rconv = ref.conv2d_feature[0][0]
      ri_ms, _ = time_kernel(lambda: ops.refine_in_conv(cz, img2[:1], rconv.weight, rconv.bias.detach(), lrelu=True), 20, flush, stream)
      ri_bytes = 4 * H * W * (32 + 3 + 1)
    kernels = {
      "cost_volume": {"bound": "hbm", "ms": cv_ms, "achieved": cv_bytes / cv_ms / 1e6, "peak": pk["hbm"], "unit": "GB/s",
                      "frac": cv_bytes / cv_ms / 1e6 / pk["hbm"], "frac_of_nominal_8TBs": cv_bytes / cv_ms / 1e6 / 8000.0,
                      "algorithmic_bytes": cv_bytes},
      "cost_volume_batch8": {"bound": "hbm", "ms": cv8_ms, "achieved": 8 * cv_bytes / cv8_ms / 1e6, "peak": pk["hbm"], "unit": "GB/s",
                             "frac": 8 * cv_bytes / cv8_ms / 1e6 / pk["hbm"], "frac_of_nominal_8TBs": 8 * cv_bytes / cv8_ms / 1e6 / 8000.0,
                             "algorithmic_bytes": 8 * cv_bytes},
      "filter_conv3d_32x32": {"bound": "tensor", "ms": f3_ms, "achieved": f3_flops / f3_ms / 1e9, "peak": pk["tensor"],
                              "unit": "TFLOP/s", "frac": f3_flops / f3_ms / 1e9 / pk["tensor"], "algorithmic_flops": f3_flops},
      "refine_conv2d_32x32_dil4": {"bound": "tensor", "ms": r2_ms, "achieved": r2_flops / r2_ms / 1e9, "peak": pk["tensor"],
                                   "unit": "TFLOP/s", "frac": r2_flops / r2_ms / 1e9 / pk["tensor"], "algorithmic_flops": r2_flops},
      "head_tap_contraction_27": {"bound": "hbm", "ms": tp_ms, "achieved": tp_bytes / tp_ms / 1e6, "peak": pk["hbm"], "unit": "GB/s",
                                  "frac": tp_bytes / tp_ms / 1e6 / pk["hbm"], "algorithmic_bytes": tp_bytes},
      "first_conv5x5s2_2images": {"bound": "tensor", "ms": fc_ms, "achieved": fc_flops / fc_ms / 1e9, "peak": pk["tensor"],
                                  "unit": "TFLOP/s", "frac": fc_flops / fc_ms / 1e9 / pk["tensor"], "algorithmic_flops": fc_flops},
      "refine_in_conv": {"bound": "hbm", "ms": ri_ms, "achieved": ri_bytes / ri_ms / 1e6, "peak": pk["hbm"], "unit": "GB/s",
                         "frac": ri_bytes / ri_ms / 1e6 / pk["hbm"], "algorithmic_bytes": ri_bytes},
      "tapsum_softargmin": {"bound": "hbm", "ms": sa_ms, "achieved": sa_bytes / sa_ms / 1e6, "peak": pk["hbm"], "unit": "GB/s",
                            "frac": sa_bytes / sa_ms / 1e6 / pk["hbm"], "algorithmic_bytes": sa_bytes},
    }
    # dominant kernel of the step = the 32->32 convolution (4 x 3-D filter + 6 x refinement launches per pair)
    dom_name = "filter_conv3d_32x32" if 4 * f3_ms >= 6 * r2_ms else "refine_conv2d_32x32_dil4"
    dom = kernels[dom_name]
    # DRAM traffic and tensor-pipe activity of the same kernel from the committed `ncu --set full` capture (profiles/)
    ncu = {}
    try:
      ncu = json.load(open(os.path.join(ROOT, "profiles", "r1_ncu_summary.json")))
    except Exception:
      pass
    ncu_k = ncu.get("conv3d" if dom_name == "filter_conv3d_32x32" else "conv2d_dil1", {})
    roofline = {"kernel": dom_name, "bound": dom["bound"], "achieved": dom["achieved"], "peak": dom["peak"], "unit": dom["unit"],
                "frac": dom["frac"], "traffic": ncu_k.get("dram_bytes"),
                "peak_source": pk["src"] + " (bf16 cuBLAS burst; kernel timed alone)",
                "note": "achieved = algorithmic fp32 conv FLOPs / time; the kernel runs 3 TF32 MMA passes per algorithmic product "
                        "(error-compensated split, fp32-grade parity) and TF32 MMAs issue at half the bf16 rate, so 1/6 of the "
                        "bf16 peak (~276 TFLOP/s) is the ceiling of this formulation",
                "frac_of_3xtf32_ceiling": dom["achieved"] / (pk["tensor"] / 6.0),
                "ncu_tensor_pipe_active_pct": ncu_k.get("tensor_pipe_active_pct"),
                "conv_backend": os.environ.get("SNB200_CONV", "default")}
    del flush

    cpu = cpu_reference(steps=args.cpu_steps, warmup=1, budget_s=25.0) if world == 1 and not args.skip_cpu else None

    in_bytes = 2 * left.numel() * 4
    out_bytes = out_pin.numel() * 4
    result = {
      "metric": METRIC, "value": world * args.steps / (ms_dev / 1e3), "unit": "pairs/s", "n_gpus": world,
      "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_dev / args.steps, "higher_is_better": True,
      "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic (textured left, right = left warped by a known disparity ramp; seeded random-init weights)",
      "config": {"workload": WORKLOAD, "pairs_per_gpu_per_step": 1, "k": K_DOWN, "maxdisp": MAXDISP,
                 "l2": "per-step working set ~0.7 GB of intermediates >> 126 MB L2; per-kernel numbers flush L2 before every launch",
                 "cuda_graph": not args.no_graph, "parallelism": f"independent pairs per GPU x{world}, no collective"},
      "e2e": {"value": world * args.steps / (ms_e2e / 1e3), "unit": "pairs/s", "ms_per_step": ms_e2e / args.steps,
              "h2d_bytes_per_step": in_bytes, "d2h_bytes_per_step": out_bytes,
              "api": "stereonet_b200.runtime.StereoEngine.infer_host_async (pinned host images -> host disparity, every frame; "
                     "copies of neighbouring frames overlap the forward on separate streams)",
              "blocking_call_ms": ms_e2e_seq / args.steps},
      "gpu_launches": launches * args.steps,
      "launches_per_step": launches,
      "roofline": roofline, "kernels": kernels, "clocks": clocks,
    }
    if adapt_steps:
      result["adapt"] = {"metric": "online adaptation steps/s @KITTI 376x1248 (train-mode fwd + Monodepth loss + bwd + clip + Adam lr 5e-5)",
                         "value": world * adapt_steps / (ms_adapt / 1e3), "unit": "steps/s", "ms_per_step": ms_adapt / adapt_steps,
                         "steps": adapt_steps, "library_launches_per_step": adapt_launches,
                         "cuda_graph": not args.no_graph,
                         "parallelism": "single stream" if world == 1 else f"shared-model DP x{world}, one NCCL all-reduce of 288066 grads per step",
                         "note": "model fwd+bwd, the photometric loss (+ its gradient) and the feature-contrast score are libsnb200 kernels; "
                                 "clip_grad_norm_ and Adam are the caller's PyTorch ops as in adapt.py"}
    if cpu is not None:
      result["cpu_baseline"] = {k: cpu[k] for k in ("value", "unit", "cores", "kind", "sample")}
      if adapt_steps and not args.skip_cpu_adapt:
        result["adapt"]["cpu_baseline"] = cpu_reference_adapt(steps=2)
  if world > 1:
    dist.barrier()
    dist.destroy_process_group()
  return result


def main():
  if os.environ.get("SNB_BENCH_DUMP_AFTER"):        # debugging aid: dump all Python stacks and exit if the run stalls
    import faulthandler
    faulthandler.dump_traceback_later(int(os.environ["SNB_BENCH_DUMP_AFTER"]), exit=True)
  ap = argparse.ArgumentParser()
  ap.add_argument("--gpus", type=int, default=1)
  ap.add_argument("--steps", type=int, default=50)
  ap.add_argument("--warmup", type=int, default=5)
  ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
  ap.add_argument("--no-graph", action="store_true")
  ap.add_argument("--skip-cpu", action="store_true")
  ap.add_argument("--cpu-steps", type=int, default=12)
  ap.add_argument("--skip-adapt", action="store_true")
  ap.add_argument("--skip-cpu-adapt", action="store_true")
  args = ap.parse_args()

  rank = int(os.environ.get("RANK", "0"))
  if args.impl == "reference":
    if rank != 0:
      return
    c = cpu_reference(steps=args.steps, warmup=min(args.warmup, 2), budget_s=150.0)
    line = {"impl": "reference", "metric": METRIC, "value": c["value"], "unit": "pairs/s", "n_gpus": args.gpus, "steps": c["steps"],
            "warmup": min(args.warmup, 2), "ms_per_step": c["ms_per_step"], "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": {"workload": WORKLOAD},
            "cpu_baseline": {k: c[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": c["value"], "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))
    return
  res = gpu_arm(args)
  if res is not None:
    print(json.dumps(res))


if __name__ == "__main__":
  main()
