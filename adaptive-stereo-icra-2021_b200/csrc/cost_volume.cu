// Difference cost volume (stereo_net.py:173-184) on channels-last features.
//
// Layout: left/right [B,H,W,32] fp32, cost [B,D,H,W,32] fp32.  One feature row is W*128 contiguous bytes, so the
// right-feature "row window" of a CTA is staged into shared memory with two 1-D TMA bulk copies (cp.async.bulk ->
// UBLKCP) completing on an mbarrier; a disparity shift is then a 128-byte-aligned smem offset, never a copy.
// Each CTA = (b, y, group of disparities); every thread streams float4s: cost[d][x] = L[x] - R[x-d] (x >= d) else 0.
// Algorithmic bytes: 4*B*32*H*W*(2 + D)   — HBM-write bound.
#include "common.cuh"

__global__ void __launch_bounds__(256)
cost_volume_fwd_kernel(const float4* __restrict__ left, const float4* __restrict__ right, float4* __restrict__ cost,
                       int D, int H, int W, int dper) {
  pdl_launch(); pdl_wait();
  extern __shared__ __align__(128) unsigned char smem_raw[];
  float4* sL = reinterpret_cast<float4*>(smem_raw);
  float4* sR = sL + (size_t)W * 8;
  __shared__ __align__(8) uint64_t bar;

  const int row = blockIdx.x;                 // b*H + y
  const int d0 = blockIdx.y * dper;
  const int d1 = min(D, d0 + dper);
  const int row_f4 = W * 8;                   // float4s per feature row
  const uint32_t row_bytes = (uint32_t)row_f4 * 16u;

  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    mbar_fence_init();
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    mbar_expect_tx(&bar, 2u * row_bytes);
    bulk_g2s(sL, left + (size_t)row * row_f4, row_bytes, &bar);
    bulk_g2s(sR, right + (size_t)row * row_f4, row_bytes, &bar);
  }
  mbar_wait(&bar, 0);

  const int b = row / H, y = row - b * H;
  const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int d = d0; d < d1; ++d) {
    float4* out = cost + (((size_t)b * D + d) * H + y) * (size_t)row_f4;
    const int shift = d * 8;                  // d pixels = d*8 float4
    for (int i = threadIdx.x; i < row_f4; i += 256) {
      float4 v = zero;
      if (i >= shift) {
        const float4 l = sL[i], r = sR[i - shift];
        v = make_float4(l.x - r.x, l.y - r.y, l.z - r.z, l.w - r.w);
      }
      out[i] = v;
    }
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
  const size_t smem = (size_t)W * 256;      // two rows of W*128 B
  SNB_REQUIRE(smem <= 200 * 1024, "snb_cost_volume_fwd: W=%d too wide for the smem row window", W);
  int rows = B * H;
  // aim for ~4 CTAs per SM so bulk-copy latency of one CTA hides behind the stores of the others
  int groups = (4 * 148 + rows - 1) / rows;
  if (groups > D) groups = D;
  if (groups < 1) groups = 1;
  int dper = (D + groups - 1) / groups;
  groups = (D + dper - 1) / dper;
  SNB_CUDA(cudaFuncSetAttribute(cost_volume_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  snb_launch(cost_volume_fwd_kernel, dim3(rows, groups), 256, smem, stream, 
      (const float4*)left, (const float4*)right, (float4*)cost, D, H, W, dper);
  SNB_LAUNCH_CHECK("cost_volume_fwd_kernel");
  return 0;
}

extern "C" int snb_cost_volume_bwd(const float* dcost, float* dleft, float* dright, int B, int D, int H, int W, void* stream) {
  SNB_REQUIRE(B > 0 && D > 0 && H > 0 && W > 0, "snb_cost_volume_bwd: bad dims");
  snb_launch(cost_volume_bwd_kernel, B * H, 256, 0, stream, (const float4*)dcost, (float4*)dleft, (float4*)dright, D, H, W);
  SNB_LAUNCH_CHECK("cost_volume_bwd_kernel");
  return 0;
}
