// the replay loss and the evaluation metrics that sit right after the model in the
// adaptation loop, as device-resident reductions (no boolean indexing, no host sync).
//   snb_khamis_loss   khamis_robust_loss, adaptive_stereo/utils/loss_functions.py:6-15 (the ER term, adapt.py:339-349):
//                     sum over gt > 0 of sqrt((gt - pred)^2 + 4) / 2 - 1, divided by max(#valid, 1); value AND d loss / d pred
//   snb_eval_metrics  train.py:98-107 (evaluate): per-sample sums for EPE = |pred - gt|[gt > 0].mean() and D1-all at 2/3/4/5 px
#include "common.cuh"

namespace {

constexpr int KH_BLOCK = 256;

__global__ void __launch_bounds__(KH_BLOCK)
khamis_partial_kernel(const float* __restrict__ pred, const float* __restrict__ gt, float* __restrict__ dpred,
                      float* __restrict__ part, long long n) {
  pdl_launch(); pdl_wait();
  __shared__ float sS[KH_BLOCK / 32], sC[KH_BLOCK / 32];
  float s = 0.f, c = 0.f;
  for (long long i = (long long)blockIdx.x * KH_BLOCK + threadIdx.x; i < n; i += (long long)gridDim.x * KH_BLOCK) {
    const float g = gt[i];
    float gr = 0.f;
    if (g > 0.f) {
      const float d = g - pred[i];
      const float r = sqrtf(fmaf(d, d, 4.f));
      s += 0.5f * r - 1.f;
      c += 1.f;
      gr = -0.5f * d / r;                       // d/dpred of sqrt((gt - pred)^2 + 4) / 2
    }
    dpred[i] = gr;                              // scaled by 1 / max(#valid, 1) in the finalize kernel
  }
  s = warp_sum(s); c = warp_sum(c);
  if ((threadIdx.x & 31) == 0) { sS[threadIdx.x >> 5] = s; sC[threadIdx.x >> 5] = c; }
  __syncthreads();
  if (threadIdx.x == 0) {
    float a = 0.f, b = 0.f;
#pragma unroll
    for (int w = 0; w < KH_BLOCK / 32; ++w) { a += sS[w]; b += sC[w]; }
    part[2 * blockIdx.x] = a; part[2 * blockIdx.x + 1] = b;
  }
}

// Every block re-reduces the partials in the same fixed order (deterministic, no extra launch), then scales its share of dpred.
__global__ void __launch_bounds__(KH_BLOCK)
khamis_finalize_kernel(const float* __restrict__ part, int nblk, float* __restrict__ dpred, float* __restrict__ loss_out, long long n) {
  pdl_launch(); pdl_wait();
  __shared__ double rs[KH_BLOCK], rc[KH_BLOCK];
  double s = 0.0, c = 0.0;
  for (int i = threadIdx.x; i < nblk; i += KH_BLOCK) { s += (double)part[2 * i]; c += (double)part[2 * i + 1]; }
  rs[threadIdx.x] = s; rc[threadIdx.x] = c;
  __syncthreads();
  for (int off = KH_BLOCK / 2; off > 0; off >>= 1) {
    if (threadIdx.x < off) { rs[threadIdx.x] += rs[threadIdx.x + off]; rc[threadIdx.x] += rc[threadIdx.x + off]; }
    __syncthreads();
  }
  const double nv = rc[0] > 1.0 ? rc[0] : 1.0;  // num_valid = max(mask.sum(), 1)
  const float inv = (float)(1.0 / nv);
  if (blockIdx.x == 0 && threadIdx.x == 0) loss_out[0] = (float)(rs[0] / nv);
  for (long long i = (long long)blockIdx.x * KH_BLOCK + threadIdx.x; i < n; i += (long long)gridDim.x * KH_BLOCK) dpred[i] *= inv;
}

// one CTA per sample: {sum |e|, #valid, #(|e| > 2), #(|e| > 3), #(|e| > 4), #(|e| > 5)} over gt > 0
__global__ void __launch_bounds__(1024)
eval_metrics_kernel(const float* __restrict__ pred, const float* __restrict__ gt, float* __restrict__ out, long long hw) {
  pdl_launch(); pdl_wait();
  __shared__ double sE[32];
  __shared__ int sN[32][5];
  const float* p = pred + (size_t)blockIdx.x * hw;
  const float* g = gt + (size_t)blockIdx.x * hw;
  double e = 0.0;
  int cnt[5] = {0, 0, 0, 0, 0};
  for (long long i = threadIdx.x; i < hw; i += 1024) {
    const float gv = g[i];
    if (gv > 0.f) {
      const float a = fabsf(p[i] - gv);
      e += (double)a;
      cnt[0] += 1; cnt[1] += a > 2.f; cnt[2] += a > 3.f; cnt[3] += a > 4.f; cnt[4] += a > 5.f;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    e += __shfl_xor_sync(0xffffffffu, e, o);
#pragma unroll
    for (int k = 0; k < 5; ++k) cnt[k] += __shfl_xor_sync(0xffffffffu, cnt[k], o);
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) {
    sE[warp] = e;
#pragma unroll
    for (int k = 0; k < 5; ++k) sN[warp][k] = cnt[k];
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double te = 0.0; long long tn[5] = {0, 0, 0, 0, 0};
    for (int w = 0; w < 32; ++w) { te += sE[w]; for (int k = 0; k < 5; ++k) tn[k] += sN[w][k]; }
    float* o = out + (size_t)blockIdx.x * 6;
    o[0] = (float)te; o[1] = (float)tn[0]; o[2] = (float)tn[1]; o[3] = (float)tn[2]; o[4] = (float)tn[3]; o[5] = (float)tn[4];
  }
}

int khamis_blocks(long long n) {
  const long long b = (n + 4 * KH_BLOCK - 1) / (4 * KH_BLOCK);
  return (int)(b < 1 ? 1 : (b > 592 ? 592 : b));
}

}  // namespace

extern "C" int snb_khamis_loss_workspace_floats(long long n) { return 2 * khamis_blocks(n); }

extern "C" int snb_khamis_loss(const float* pred, const float* gt, float* loss_out, float* dpred, float* workspace, long long n,
                               void* stream) {
  SNB_REQUIRE(pred && gt && loss_out && dpred && workspace && n > 0, "snb_khamis_loss: bad args");
  const int nblk = khamis_blocks(n);
  snb_launch(khamis_partial_kernel, nblk, KH_BLOCK, 0, stream, pred, gt, dpred, workspace, n);
  SNB_LAUNCH_CHECK("khamis_partial_kernel");
  snb_launch(khamis_finalize_kernel, nblk, KH_BLOCK, 0, stream, (const float*)workspace, nblk, dpred, loss_out, n);
  SNB_LAUNCH_CHECK("khamis_finalize_kernel");
  return 0;
}

extern "C" int snb_eval_metrics(const float* pred, const float* gt, float* out, int B, long long hw, void* stream) {
  SNB_REQUIRE(pred && gt && out && B > 0 && hw > 0, "snb_eval_metrics: bad args");
  snb_launch(eval_metrics_kernel, B, 1024, 0, stream, pred, gt, out, hw);
  SNB_LAUNCH_CHECK("eval_metrics_kernel");
  return 0;
}
