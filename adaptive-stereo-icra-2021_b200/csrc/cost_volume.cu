// Difference cost volume (stereo_net.py:173-184) on channels-last features.
//
// Layout: left/right [B,H,W,32] fp32, cost [B,D,H,W,32] fp32.  A CTA = (b, y, segment of XS columns) produces ALL D disparity
// levels of its segment: the left segment (XS*128 B) and the right-feature row WINDOW [x0-(D-1), x0+XS) (one contiguous run of
// channels-last memory) are staged into shared memory with two 1-D TMA bulk copies (cp.async.bulk -> UBLKCP) completing on an
// mbarrier; a disparity shift is then a 128-byte-aligned smem offset, never a materialised copy.  Every thread streams
// float4s: cost[d][x] = L[x] - R[x-d] (x >= d) else 0; a warp writes 512 contiguous bytes per store instruction.
// Staged bytes per CTA are (2 XS + D - 1)/(D XS) of the bytes it stores (~12 % at KITTI size; round 1's (b, y, 2 disparities)
// CTAs re-read both full rows for every 2 levels: as many bytes in as out).
// Algorithmic bytes: 4*B*32*H*W*(2 + D)   — HBM-write bound.
#include "common.cuh"

__global__ void __launch_bounds__(256)
cost_volume_fwd_kernel(const float4* __restrict__ left, const float4* __restrict__ right, float4* __restrict__ cost,
                       int D, int H, int W, int XS) {
  pdl_launch();
  extern __shared__ __align__(128) unsigned char smem_raw[];
  float4* sL = reinterpret_cast<float4*>(smem_raw);                 // [XS][8]
  float4* sR = sL + (size_t)XS * 8;                                 // [XS + D - 1][8] window starting at column wstart
  __shared__ __align__(8) uint64_t bar;

  const int row = blockIdx.x;                 // b*H + y
  const int x0 = blockIdx.y * XS;
  const int xs = min(XS, W - x0);             // columns of this segment
  const int wstart = max(0, x0 - (D - 1));
  const int wlen = x0 + xs - wstart;          // right-feature columns [wstart, x0 + xs)

  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    mbar_fence_init();
  }
  __syncthreads();
  pdl_wait();                                 // the feature maps come from the previous kernel
  if (threadIdx.x == 0) {
    mbar_expect_tx(&bar, (uint32_t)(xs + wlen) * 128u);
    bulk_g2s(sL, left + ((size_t)row * W + x0) * 8, (uint32_t)xs * 128u, &bar);
    bulk_g2s(sR, right + ((size_t)row * W + wstart) * 8, (uint32_t)wlen * 128u, &bar);
  }
  const int b = row / H, y = row - b * H;
  const int seg_f4 = xs * 8;                  // float4s per (d) output segment
  const size_t dstride = (size_t)H * W * 8;   // float4s between disparity levels
  float4* out0 = cost + (((size_t)b * D) * H + y) * (size_t)W * 8 + (size_t)x0 * 8;
  const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
  // zeros of the x < d triangle need no operands: write them while the bulk copies are in flight
  for (int d = x0 + 1; d < D; ++d) {
    const int nz = min(d - x0, xs) * 8;
    float4* out = out0 + d * dstride;
    for (int i = threadIdx.x; i < nz; i += 256) out[i] = zero;
  }
  mbar_wait(&bar, 0);
  for (int d = 0; d < D; ++d) {
    float4* out = out0 + d * dstride;
    const int first = max(d - x0, 0) * 8;     // columns x0 + i/8 < d were zero-filled above
    const int roff = (x0 - d - wstart) * 8;   // sR index of column x - d for x = x0 (>= 0 wherever i >= first)
    for (int i = first + threadIdx.x; i < seg_f4; i += 256) {
      const float4 l = sL[i], r = sR[i + roff];
      out[i] = make_float4(l.x - r.x, l.y - r.y, l.z - r.z, l.w - r.w);
    }
  }
}

// Small volumes (the batch-1 KITTI volume is 22.5 MB and stays in L2): a CTA per (row, disparity level) and no staging at all —
// both operand rows are L1 / L2 hits, every float4 of output is two independent loads, a subtraction and a store, and all
// B*H*D CTAs are resident at once, so nothing waits on a staging round trip.
__global__ void __launch_bounds__(256)
cost_volume_fwd_direct_kernel(const float4* __restrict__ left, const float4* __restrict__ right, float4* __restrict__ cost,
                              int D, int H, int W) {
  pdl_launch(); pdl_wait();
  const int row = blockIdx.x, d = blockIdx.y;        // row = b*H + y
  const int b = row / H, y = row - b * H;
  const int row_f4 = W * 8, first = min(d, W) * 8;
  const float4* lp = left + (size_t)row * row_f4;
  const float4* rp = right + (size_t)row * row_f4 - d * 8;
  float4* out = cost + (((size_t)b * D + d) * H + y) * (size_t)row_f4;
  const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int i = threadIdx.x; i < first; i += 256) out[i] = zero;
  int i = first + threadIdx.x;
  for (; i + 768 < row_f4; i += 1024) {               // four independent load pairs in flight per thread
    float4 l[4], r[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) { l[k] = __ldg(lp + i + 256 * k); r[k] = __ldg(rp + i + 256 * k); }
#pragma unroll
    for (int k = 0; k < 4; ++k) out[i + 256 * k] = make_float4(l[k].x - r[k].x, l[k].y - r[k].y, l[k].z - r[k].z, l[k].w - r[k].w);
  }
  for (; i < row_f4; i += 256) {
    const float4 l = __ldg(lp + i), r = __ldg(rp + i);
    out[i] = make_float4(l.x - r.x, l.y - r.y, l.z - r.z, l.w - r.w);
  }
}

// Adjoint: dL[x] = sum_{d<=x} g[d][x];  dR[x'] = -sum_{d : x'+d < W} g[d][x'+d]   
__global__ void __launch_bounds__(256)
cost_volume_bwd_kernel(const float4* __restrict__ g, float4* __restrict__ dleft, float4* __restrict__ dright,
                       int D, int H, int W) {
  pdl_launch(); pdl_wait();
  const int row = blockIdx.x;
  const int b = row / H, y = row - b * H;
  const int row_f4 = W * 8;
  for (int i = threadIdx.x; i < row_f4; i += 256) {
    const int x = i >> 3;
    float4 aL = make_float4(0.f, 0.f, 0.f, 0.f), aR = aL;
    for (int d = 0; d < D; ++d) {
      const float4* gp = g + (((size_t)b * D + d) * H + y) * (size_t)row_f4;
      if (x >= d) { const float4 v = gp[i]; aL.x += v.x; aL.y += v.y; aL.z += v.z; aL.w += v.w; }
      if (x + d < W) { const float4 v = gp[i + d * 8]; aR.x -= v.x; aR.y -= v.y; aR.z -= v.z; aR.w -= v.w; }
    }
    dleft[(size_t)row * row_f4 + i] = aL;
    dright[(size_t)row * row_f4 + i] = aR;
  }
}

extern "C" int snb_cost_volume_fwd(const float* left, const float* right, float* cost, int B, int D, int H, int W, void* stream) {
  SNB_REQUIRE(B > 0 && D > 0 && H > 0 && W > 0, "snb_cost_volume_fwd: bad dims");
  int dev = 0, sms = 148;
  SNB_CUDA(cudaGetDevice(&dev));
  SNB_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const int rows = B * H;
  // Volumes that stay in L2 (the batch-1 KITTI volume is 22.5 MB) take the un-staged kernel: 7.1 against 11.0 us back to back
  // (3.4 TB/s); from batch 8 on (195 MB) both forms run at the same 4.8 TB/s and the staged one moves a tenth of the operand bytes.
  if ((size_t)B * D * H * W * 128 <= (size_t)64 << 20 && D <= 65535) {
    snb_launch(cost_volume_fwd_direct_kernel, dim3(rows, D), 256, 0, stream, (const float4*)left, (const float4*)right, (float4*)cost, D, H, W);
    SNB_LAUNCH_CHECK("cost_volume_fwd_direct_kernel");
    return 0;
  }
  // >= 2 CTAs per SM so one CTA's bulk-copy latency hides behind the stores of another; segments of at least 8 columns
  int nseg = (2 * sms + rows - 1) / rows;
  if (nseg > (W + 7) / 8) nseg = (W + 7) / 8;
  if (nseg < 1) nseg = 1;
  int XS = (W + nseg - 1) / nseg;
  while ((size_t)(2 * XS + D - 1) * 128 > 200 * 1024) { ++nseg; XS = (W + nseg - 1) / nseg; }     // very wide rows
  nseg = (W + XS - 1) / XS;
  SNB_REQUIRE(nseg <= 65535, "snb_cost_volume_fwd: too many column segments");
  const size_t smem = (size_t)(2 * XS + D - 1) * 128;
  SNB_CUDA(cudaFuncSetAttribute(cost_volume_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  snb_launch(cost_volume_fwd_kernel, dim3(rows, nseg), 256, smem, stream,
      (const float4*)left, (const float4*)right, (float4*)cost, D, H, W, XS);
  SNB_LAUNCH_CHECK("cost_volume_fwd_kernel");
  return 0;
}

extern "C" int snb_cost_volume_bwd(const float* dcost, float* dleft, float* dright, int B, int D, int H, int W, void* stream) {
  SNB_REQUIRE(B > 0 && D > 0 && H > 0 && W > 0, "snb_cost_volume_bwd: bad dims");
  snb_launch(cost_volume_bwd_kernel, B * H, 256, 0, stream, (const float4*)dcost, (float4*)dleft, (float4*)dright, D, H, W);
  SNB_LAUNCH_CHECK("cost_volume_bwd_kernel");
  return 0;
}
