// Optimizer side of the adaptation step (adapt.py:391-393): clip_grad_norm_(stereo_net.parameters(), 1.0) + torch.optim.Adam
// (defaults: betas 0.9 / 0.999, eps 1e-8, no weight decay) over ONE flat gradient bucket, in three launches — PyTorch's
// clip + fused Adam are ~8 multi-tensor launches, and shared-model data parallelism wants the gradients in one flat buffer for
// its single NCCL all-reduce anyway.
//
//   snb_multi_gather     the per-parameter gradient tensors autograd produced -> the flat bucket (one launch per 64 tensors)
//   snb_adam_clip_step   (1) sum of squares of the clipped group (= the first n_clip elements: stereo_net) in fixed-order
//                            block partials; (2) finalise: total norm, clip coefficient max_norm / (norm + 1e-6) capped at 1
//                            (torch.nn.utils.clip_grad_norm_), step += 1, bias corrections; (3) the Adam update of every
//                            parameter chunk.  grad_scale folds the 1 / world of the data-parallel average into the same pass.
#include "common.cuh"
#include <math.h>

namespace {

constexpr int MG_MAX = 64;
struct GatherArgs { const float* src[MG_MAX]; long long off[MG_MAX]; long long cnt[MG_MAX]; int n; };

__global__ void __launch_bounds__(256) multi_gather_kernel(const GatherArgs a, float* __restrict__ dst) {
  pdl_launch(); pdl_wait();
  const int t = blockIdx.y;
  if (t >= a.n) return;
  const float* __restrict__ s = a.src[t];
  float* __restrict__ d = dst + a.off[t];
  const long long n = a.cnt[t];
  for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n; i += (long long)gridDim.x * 256) d[i] = s ? s[i] : 0.f;
}

constexpr int NORM_BLOCKS = 128;

__global__ void __launch_bounds__(256) sumsq_partial_kernel(const float* __restrict__ g, long long n, float scale, double* __restrict__ part) {
  pdl_launch(); pdl_wait();
  __shared__ double red[256];
  double s = 0.0;
  for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n; i += (long long)NORM_BLOCKS * 256) {
    const double v = (double)(g[i] * scale);
    s += v * v;
  }
  red[threadIdx.x] = s;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) { if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o]; __syncthreads(); }
  if (threadIdx.x == 0) part[blockIdx.x] = red[0];
}

// ws layout (floats): [0] clip coefficient, [1] lr / (1 - beta1^t), [2] 1 / sqrt(1 - beta2^t), [3] total norm of the clipped group
__global__ void adam_finalize_kernel(const double* __restrict__ part, float* __restrict__ step, float max_norm, double lr,
                                     double beta1, double beta2, float* __restrict__ ws) {
  pdl_launch(); pdl_wait();
  if (threadIdx.x != 0) return;
  double s = 0.0;
  for (int i = 0; i < NORM_BLOCKS; ++i) s += part[i];
  const float norm = (float)sqrt(s);
  float coef = 1.f;
  if (max_norm > 0.f) { coef = max_norm / (norm + 1e-6f); if (coef > 1.f) coef = 1.f; }
  const float t = step[0] + 1.f;
  step[0] = t;
  ws[0] = coef;
  ws[1] = (float)(lr / (1.0 - pow(beta1, (double)t)));
  ws[2] = (float)(1.0 / sqrt(1.0 - pow(beta2, (double)t)));
  ws[3] = norm;
}

// chunk table: [nchunks][3] int64 = {parameter pointer (already offset to the chunk), flat offset, count <= 1024}
__global__ void __launch_bounds__(256) adam_update_kernel(const long long* __restrict__ table, const float* __restrict__ grad,
                                                          float* __restrict__ exp_avg, float* __restrict__ exp_avg_sq,
                                                          const float* __restrict__ ws, long long n_clip, float grad_scale,
                                                          float beta1, float beta2, float omb1, float omb2, float eps) {
  pdl_launch(); pdl_wait();
  const long long* e = table + 3 * (long long)blockIdx.x;
  float* __restrict__ p = reinterpret_cast<float*>(e[0]);
  const long long off = e[1];
  const int cnt = (int)e[2];
  const float coef = ws[0], step_size = ws[1], inv_bc2 = ws[2];
  for (int i = threadIdx.x; i < cnt; i += 256) {
    const long long j = off + i;
    const float g = grad[j] * grad_scale * (j < n_clip ? coef : 1.f);
    const float m = beta1 * exp_avg[j] + omb1 * g;
    const float v = beta2 * exp_avg_sq[j] + omb2 * g * g;
    exp_avg[j] = m; exp_avg_sq[j] = v;
    const float denom = sqrtf(v) * inv_bc2 + eps;
    p[i] -= step_size * (m / denom);
  }
}

}  // namespace

extern "C" int snb_multi_gather(const float* const* src, const long long* dst_off, const long long* count, int n, float* dst, void* stream) {
  SNB_REQUIRE(src && dst_off && count && dst && n > 0, "snb_multi_gather: bad args");
  for (int base = 0; base < n; base += MG_MAX) {
    GatherArgs a;
    a.n = n - base < MG_MAX ? n - base : MG_MAX;
    long long mx = 1;
    for (int i = 0; i < a.n; ++i) {
      a.src[i] = src[base + i]; a.off[i] = dst_off[base + i]; a.cnt[i] = count[base + i];
      SNB_REQUIRE(a.cnt[i] >= 0 && a.off[i] >= 0, "snb_multi_gather: negative extent");
      if (a.cnt[i] > mx) mx = a.cnt[i];
    }
    int bx = (int)((mx + 2047) / 2048); if (bx > 32) bx = 32; if (bx < 1) bx = 1;
    snb_launch(multi_gather_kernel, dim3(bx, a.n), 256, 0, stream, a, dst);
    SNB_LAUNCH_CHECK("multi_gather_kernel");
  }
  return 0;
}

extern "C" int snb_adam_clip_workspace_bytes(void) { return NORM_BLOCKS * (int)sizeof(double) + 4 * (int)sizeof(float); }

extern "C" int snb_adam_clip_step(const long long* chunk_table, int nchunks, const float* flat_grad, float* exp_avg, float* exp_avg_sq,
                                  float* step, long long n_total, long long n_clip, float max_norm, float grad_scale, double lr,
                                  double beta1, double beta2, double eps, void* workspace, void* stream) {
  SNB_REQUIRE(chunk_table && nchunks > 0 && flat_grad && exp_avg && exp_avg_sq && step && workspace, "snb_adam_clip_step: bad args");
  SNB_REQUIRE(n_clip >= 0 && n_clip <= n_total, "snb_adam_clip_step: clipped group exceeds the bucket");
  SNB_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 7) == 0, "snb_adam_clip_step: workspace must be 8-byte aligned");
  double* part = reinterpret_cast<double*>(workspace);
  float* ws = reinterpret_cast<float*>(part + NORM_BLOCKS);
  snb_launch(sumsq_partial_kernel, NORM_BLOCKS, 256, 0, stream, flat_grad, max_norm > 0.f ? n_clip : 0ll, grad_scale, part);
  SNB_LAUNCH_CHECK("sumsq_partial_kernel");
  snb_launch(adam_finalize_kernel, 1, 32, 0, stream, (const double*)part, step, max_norm, lr, beta1, beta2, ws);
  SNB_LAUNCH_CHECK("adam_finalize_kernel");
  snb_launch(adam_update_kernel, nchunks, 256, 0, stream, chunk_table, flat_grad, exp_avg, exp_avg_sq, (const float*)ws, n_clip, grad_scale,
             (float)beta1, (float)beta2, (float)(1.0 - beta1), (float)(1.0 - beta2), (float)eps);   // 1 - beta in double, as torch.optim.Adam forms it
  SNB_LAUNCH_CHECK("adam_update_kernel");
  return 0;
}
