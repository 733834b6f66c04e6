// Disparity head: the two 32->1 convolutions, soft-argmin, bilinear upsampling.
//
// A 32->1 convolution is split into (1) a per-position channel contraction  taps[p][t] = sum_c w[t][c] * x[p][c]
// (every activation is read exactly once, no halo) and (2) a gather-sum of the ntaps shifted tap planes, which is fused
// with what follows it in the reference:
//   conv3d_alone + softmax + DisparityRegression (stereo_net.py:187-192,124-134) -> snb_tapsum_softargmin
//   conv2d_out + residual add + ReLU           (stereo_net.py:121)              -> snb_tapsum_refine_out
// Tap planes are stored [slice][tap][H][W] so both the scatter (thread = position) and the gather (thread = x) are coalesced.
#include "common.cuh"
#include <math.h>

namespace {

template <int NT>
__global__ void __launch_bounds__(128)
conv_c32_taps_kernel(const float* __restrict__ x, const float* __restrict__ w, float* __restrict__ taps,
                     long long npos, int plane) {
  pdl_launch(); pdl_wait();
  constexpr int NTP = (NT + 3) & ~3;
  __shared__ __align__(16) float sA[128][32];
  __shared__ __align__(16) float sW[32][NTP];
  const int t = threadIdx.x;
  const long long pos0 = (long long)blockIdx.x * 128;
  {
    const int chunk = t & 7;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int r = (t >> 3) + 16 * j;
      const bool ok = pos0 + r < npos;
      cp_async16(&sA[r][(chunk ^ (r & 7)) * 4], x + (ok ? (pos0 + r) * 32 + chunk * 4 : 0), ok);
    }
    cp_async_commit();
  }
  for (int i = t; i < 32 * NTP; i += 128) {
    const int ci = i / NTP, tp = i - ci * NTP;
    sW[ci][tp] = tp < NT ? w[ci * NT + tp] : 0.f;        // w is [1][32][NT]
  }
  cp_async_wait<0>();
  __syncthreads();

  unsigned long long pacc[NTP / 2];              // packed pairs of taps: FFMA2 halves the FMA instruction count
#pragma unroll
  for (int j = 0; j < NTP / 2; ++j) pacc[j] = 0ull;
#pragma unroll 1
  for (int c = 0; c < 8; ++c) {     // not unrolled: a full unroll hoists all 32xNT weights into registers and spills
    const float4 xa = *reinterpret_cast<const float4*>(&sA[t][(c ^ (t & 7)) * 4]);
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      const float xv = kk == 0 ? xa.x : kk == 1 ? xa.y : kk == 2 ? xa.z : xa.w;
      const unsigned long long xp = pack2(xv, xv);
      const ulonglong2* wp = reinterpret_cast<const ulonglong2*>(&sW[c * 4 + kk][0]);
#pragma unroll
      for (int j = 0; j < NTP / 4; ++j) {
        const ulonglong2 wv = wp[j];
        ffma2(pacc[2 * j], xp, wv.x); ffma2(pacc[2 * j + 1], xp, wv.y);
      }
    }
  }
  float acc[NTP];
#pragma unroll
  for (int j = 0; j < NTP / 2; ++j) { acc[2 * j] = lo2(pacc[j]); acc[2 * j + 1] = hi2(pacc[j]); }
  const long long p = pos0 + t;
  if (p < npos) {
    const long long slice = p / plane;
    const int q = (int)(p - slice * plane);
    float* out = taps + slice * (long long)NT * plane + q;
#pragma unroll
    for (int j = 0; j < NT; ++j) out[(size_t)j * plane] = acc[j];
  }
}

// grid (B*H, ceil(W/32)); block (32, D): thread (lane, d) gathers the 27 tap planes of ONE cost value (all 27 loads are
// independent and coalesced over the 32 x-positions of the warp), the D warps drop their costs into smem and warp 0
// finishes softmax_d + expectation for the 32 columns.  D <= 32 (maxdisp 192 -> D' <= 24).
__global__ void __launch_bounds__(1024)
tapsum_softargmin_kernel(const float* __restrict__ taps, const float* __restrict__ bias, float* __restrict__ cost_out,
                         float* __restrict__ pred, int D, int H, int W) {
  pdl_launch(); pdl_wait();
  __shared__ float sC[32][33];
  const int lane = threadIdx.x, d = threadIdx.y;
  const int b = blockIdx.x / H, y = blockIdx.x - b * H;
  const int x = blockIdx.y * 32 + lane;
  const size_t plane = (size_t)H * W;
  float c = bias[0];
  if (x < W) {
#pragma unroll
    for (int kd = 0; kd < 3; ++kd) {
      const int dd = d + kd - 1;
      if ((unsigned)dd >= (unsigned)D) continue;
      const float* tp = taps + (((size_t)b * D + dd) * 27 + kd * 9) * plane;
#pragma unroll
      for (int kh = 0; kh < 3; ++kh) {
        const int yy = y + kh - 1;
        if ((unsigned)yy >= (unsigned)H) continue;
#pragma unroll
        for (int kw = 0; kw < 3; ++kw) {
          const int xx = x + kw - 1;
          if ((unsigned)xx < (unsigned)W) c += __ldg(tp + (size_t)(kh * 3 + kw) * plane + (size_t)yy * W + xx);
        }
      }
    }
    if (cost_out) cost_out[(((size_t)b * D + d) * H + y) * W + x] = c;
  }
  sC[d][lane] = c;
  __syncthreads();
  if (d == 0 && x < W) {
    float m = -INFINITY;
    for (int i = 0; i < D; ++i) m = fmaxf(m, sC[i][lane]);
    float s = 0.f, ws = 0.f;
    for (int i = 0; i < D; ++i) {
      const float e = __expf(sC[i][lane] - m);
      s += e; ws = fmaf(e, (float)i, ws);
    }
    pred[((size_t)b * H + y) * W + x] = ws / s;
  }
}

__global__ void __launch_bounds__(256)
tapsum_refine_out_kernel(const float* __restrict__ taps, const float* __restrict__ bias, const float* __restrict__ up,
                         float* __restrict__ out, int H, int W, int relu) {
  pdl_launch(); pdl_wait();
  const int x = blockIdx.x * 256 + threadIdx.x;
  const int y = blockIdx.y, b = blockIdx.z;
  if (x >= W) return;
  const size_t plane = (size_t)H * W;
  const float* tp = taps + (size_t)b * 9 * plane;
  float c = bias ? bias[0] : 0.f;
#pragma unroll
  for (int kh = 0; kh < 3; ++kh) {
    const int yy = y + kh - 1;
    if ((unsigned)yy >= (unsigned)H) continue;
#pragma unroll
    for (int kw = 0; kw < 3; ++kw) {
      const int xx = x + kw - 1;
      if ((unsigned)xx < (unsigned)W) c += tp[(size_t)(kh * 3 + kw) * plane + (size_t)yy * W + xx];
    }
  }
  const size_t o = (size_t)b * plane + (size_t)y * W + x;
  const float v = (up ? up[o] : 0.f) + c;
  out[o] = (relu && v < 0.f) ? 0.f : v;
}

__global__ void __launch_bounds__(256)
upsample_bilinear_kernel(const float* __restrict__ in, float* __restrict__ out, int h, int w, int H, int W, float mul) {
  pdl_launch(); pdl_wait();
  const int x = blockIdx.x * 256 + threadIdx.x;
  const int y = blockIdx.y, b = blockIdx.z;
  if (x >= W) return;
  const float v = bilinear_sample(in + (size_t)b * h * w, h, w, y, x, (float)h / (float)H, (float)w / (float)W);
  out[((size_t)b * H + y) * W + x] = mul * v;
}

// Deterministic gather form of the adjoint: each coarse pixel visits the fine pixels whose 2x2 footprint contains it.
__global__ void __launch_bounds__(128)
upsample_bilinear_bwd_kernel(const float* __restrict__ dout, float* __restrict__ din, int h, int w, int H, int W, float mul) {
  pdl_launch(); pdl_wait();
  const int j = blockIdx.x * 128 + threadIdx.x;
  const int i = blockIdx.y, b = blockIdx.z;
  if (j >= w) return;
  const float sh = (float)h / (float)H, sw = (float)w / (float)W;
  // fine rows whose source coordinate falls in (i-1, i+1): widen by one pixel on both sides, then test exactly
  int ylo = (int)floorf(((float)i - 1.f + 0.5f) / sh - 0.5f) - 1, yhi = (int)ceilf(((float)i + 1.f + 0.5f) / sh - 0.5f) + 1;
  int xlo = (int)floorf(((float)j - 1.f + 0.5f) / sw - 0.5f) - 1, xhi = (int)ceilf(((float)j + 1.f + 0.5f) / sw - 0.5f) + 1;
  ylo = max(ylo, 0); yhi = min(yhi, H - 1); xlo = max(xlo, 0); xhi = min(xhi, W - 1);
  if (i == 0) ylo = 0;
  if (i == h - 1) yhi = H - 1;
  if (j == 0) xlo = 0;
  if (j == w - 1) xhi = W - 1;
  float acc = 0.f;
  for (int y = ylo; y <= yhi; ++y) {
    int y0, y1; float ly;
    bilinear_src(y, sh, h, y0, y1, ly);
    float wy = 0.f;
    if (y0 == i) wy += 1.f - ly;
    if (y1 == i) wy += ly;
    if (wy == 0.f) continue;
    const float* dp = dout + ((size_t)b * H + y) * W;
    for (int x = xlo; x <= xhi; ++x) {
      int x0, x1; float lx;
      bilinear_src(x, sw, w, x0, x1, lx);
      float wx = 0.f;
      if (x0 == j) wx += 1.f - lx;
      if (x1 == j) wx += lx;
      if (wx != 0.f) acc = fmaf(wy * wx, dp[x], acc);
    }
  }
  din[((size_t)b * h + i) * w + j] = mul * acc;
}

// Feature-contrast score (adaptive_stereo/utils/feature_contrast.py:12-23):
//   fcs[b,y,x] = max_d cost - mean of all but the two largest costs  ( = sorted_desc[0] - mean(sorted_desc[2:]) )
// One pass over D with a running top-2 and sum instead of a torch.sort over [B,D,H,W].
__global__ void __launch_bounds__(256)
feature_contrast_kernel(const float* __restrict__ cost, float* __restrict__ out, int D, long long plane, long long total) {
  pdl_launch(); pdl_wait();
  const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
  if (i >= total) return;
  const long long b = i / plane, q = i - b * plane;
  const float* c = cost + b * D * plane + q;
  float m1 = -INFINITY, m2 = -INFINITY, s = 0.f;
  for (int d = 0; d < D; ++d) {
    const float v = c[(size_t)d * plane];
    s += v;
    if (v > m1) { m2 = m1; m1 = v; } else if (v > m2) m2 = v;
  }
  out[i] = m1 - (s - m1 - m2) / (float)(D - 2);
}

}  // namespace

extern "C" int snb_feature_contrast(const float* cost, float* out, int B, int D, int H, int W, void* stream) {
  SNB_REQUIRE(cost && out && B > 0 && D > 2 && H > 0 && W > 0, "snb_feature_contrast: bad args (needs D > 2)");
  const long long plane = (long long)H * W, total = plane * B;
  snb_launch(feature_contrast_kernel, snb_ceil_div(total, 256), 256, 0, stream, cost, out, D, plane, total);
  SNB_LAUNCH_CHECK("feature_contrast_kernel");
  return 0;
}

extern "C" int snb_conv_c32_taps(const float* x, const float* w, float* taps, long long nslices, int plane, int ntaps, void* stream) {
  SNB_REQUIRE(x && w && taps && nslices > 0 && plane > 0, "snb_conv_c32_taps: bad args");
  SNB_REQUIRE(ntaps == 27 || ntaps == 9, "snb_conv_c32_taps: ntaps must be 27 or 9");
  const long long npos = nslices * plane;
  const int grid = snb_ceil_div(npos, 128);
  if (ntaps == 27) snb_launch(conv_c32_taps_kernel<27>, grid, 128, 0, stream, x, w, taps, npos, plane);
  else             snb_launch(conv_c32_taps_kernel<9>, grid, 128, 0, stream, x, w, taps, npos, plane);
  SNB_LAUNCH_CHECK("conv_c32_taps_kernel");
  return 0;
}

extern "C" int snb_tapsum_softargmin(const float* taps, const float* bias, float* cost_out, float* pred,
                                     int B, int D, int H, int W, void* stream) {
  SNB_REQUIRE(taps && bias && pred && B > 0 && D > 0 && H > 0 && W > 0, "snb_tapsum_softargmin: bad args");
  SNB_REQUIRE(D <= 32, "snb_tapsum_softargmin: at most 32 coarse disparity levels (got %d)", D);
  snb_launch(tapsum_softargmin_kernel, dim3(B * H, snb_ceil_div(W, 32)), dim3(32, D), 0, stream, 
      taps, bias, cost_out, pred, D, H, W);
  SNB_LAUNCH_CHECK("tapsum_softargmin_kernel");
  return 0;
}

extern "C" int snb_tapsum_refine_out(const float* taps, const float* bias, const float* up, float* out,
                                     int B, int H, int W, int relu, void* stream) {
  SNB_REQUIRE(taps && out && B > 0 && H > 0 && W > 0, "snb_tapsum_refine_out: bad args");
  snb_launch(tapsum_refine_out_kernel, dim3(snb_ceil_div(W, 256), H, B), 256, 0, stream, taps, bias, up, out, H, W, relu);
  SNB_LAUNCH_CHECK("tapsum_refine_out_kernel");
  return 0;
}

extern "C" int snb_upsample_bilinear(const float* in, float* out, int B, int h, int w, int H, int W, float mul, void* stream) {
  SNB_REQUIRE(in && out && B > 0 && h > 0 && w > 0 && H > 0 && W > 0, "snb_upsample_bilinear: bad args");
  snb_launch(upsample_bilinear_kernel, dim3(snb_ceil_div(W, 256), H, B), 256, 0, stream, in, out, h, w, H, W, mul);
  SNB_LAUNCH_CHECK("upsample_bilinear_kernel");
  return 0;
}

extern "C" int snb_upsample_bilinear_bwd(const float* dout, float* din, int B, int h, int w, int H, int W, float mul, void* stream) {
  SNB_REQUIRE(dout && din && B > 0 && h > 0 && w > 0 && H > 0 && W > 0, "snb_upsample_bilinear_bwd: bad args");
  snb_launch(upsample_bilinear_bwd_kernel, dim3(snb_ceil_div(w, 128), h, B), 128, 0, stream, dout, din, h, w, H, W, mul);
  SNB_LAUNCH_CHECK("upsample_bilinear_bwd_kernel");
  return 0;
}
