// Disparity head: the two 32->1 convolutions, soft-argmin, bilinear upsampling.
//
// A 32->1 convolution is split into (1) a per-position channel contraction  taps[p][t] = sum_c w[t][c] * x[p][c]
// (every activation is read exactly once, no halo) and (2) a gather-sum of the ntaps shifted tap planes, which is fused
// with what follows it in the reference:
//   conv3d_alone + softmax + DisparityRegression (stereo_net.py:187-192,124-134) -> snb_tapsum_softargmin
//   conv2d_out + residual add + ReLU           (stereo_net.py:121)              -> snb_tapsum_refine_out
// Tap planes are stored [slice][tap][H][W] so both the scatter (thread = position) and the gather (thread = x) are coalesced.
#include <stdlib.h>
#include <string.h>
#include "tc_common.cuh"
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

// ---- the same contraction on the tensor cores: taps[p][t] = sum_c x[p][c] w[t][c] is a GEMM with M = positions, N = 32
// (27 or 9 tap rows used), K = 32 channels; 3xTF32 split, A operand from TMEM.  Two groups of 4 warps alternate over the
// CTA's 128-position tiles: cp.async the tile (16-B chunks XOR-swizzled) -> each thread reads its position's row, splits
// hi/lo and tcgen05.st's it into the group's A slot -> the MMA warp issues 12 MMAs -> the group reads the 27 taps of its
// position back (tcgen05.ld) and stores them plane-major (coalesced over positions).  The FFMA kernel above was bound by
// its 27 x 32 FMAs per position (20 us at KITTI size against a 6 us memory floor); it stays selectable with
// SNB200_TAPS=ffma as the independent cross-check.
constexpr int TAPS_TC_THREADS = 9 * 32;
constexpr int TAPS_TC_SMEM = 80 * 1024;           // 8 KB weights + 4 x 16 KB tile buffers + barriers; >= 78 KB keeps it at 2 CTAs / SM (2 x 256 TMEM columns)

template <int NT>
__global__ void __launch_bounds__(TAPS_TC_THREADS, 2)
conv_c32_taps_tc_kernel(const float* __restrict__ x, const float* __restrict__ w, float* __restrict__ taps,
                        long long npos, int plane, int ntiles) {
  extern __shared__ unsigned char smem_dyn[];
  const uint32_t base_u32 = (smem_u32(smem_dyn) + 1023u) & ~1023u;
  unsigned char* base = smem_dyn + (base_u32 - smem_u32(smem_dyn));
  float* sB = reinterpret_cast<float*>(base);                      // [hi|lo][32 tap rows][32 channels] swizzled
  unsigned char* sTile = base + 8192;                              // [group][2][128 x 128 B]
  uint64_t* a_full = reinterpret_cast<uint64_t*>(sTile + 4 * 16384);   // [2] group -> MMA warp
  uint64_t* done = a_full + 2;                                     // [2] MMA commit -> group
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(done + 2);
  const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
  constexpr int ACC_COL = 0, A_COL = 64;                           // TMEM: 2 accumulators x 32 | 2 A slots x (hi 32 | lo 32)

  if (warp == 8) {
    if (lane == 0) { mbar_init(&a_full[0], 4); mbar_init(&a_full[1], 4); mbar_init(&done[0], 1); mbar_init(&done[1], 1); mbar_fence_init(); }
    __syncwarp();
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;\n" :: "r"(smem_u32(tmem_slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
  }
  pdl_launch();                                    // dependents only once this CTA owns its TMEM columns (no alloc dead-lock with an early dependent)
  pdl_wait();                                                      // nothing above touched global memory
  for (int i = t; i < 32 * 32; i += TAPS_TC_THREADS) {             // B row = tap, K = channel; w is [1][32][NT]
    const int tp = i >> 5, c = i & 31;
    const float v = tp < NT ? __ldg(w + c * NT + tp) : 0.f;
    const float hi = __uint_as_float(tc::tf32_hi_bits(v));
    const int o = tp * 32 + (((c >> 2) ^ (tp & 7)) << 2) + (c & 3);
    sB[o] = hi;
    sB[1024 + o] = v - hi;
  }
  tc::fence_async_smem();
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int nmine = (ntiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;     // tiles blockIdx.x + i * gridDim.x

  if (warp == 8) {
    // =============================================================== MMA issuer
    const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0), smem_u = __shfl_sync(0xffffffffu, base_u32, 0);
    for (int i = 0; i < nmine; ++i) {
      const int g = i & 1, j = i >> 1;
      tc::mbar_wait_warp(&a_full[g], j & 1);
      tc::tc_fence_after();
      if (tc::elect_one()) {
        const uint32_t tmem_d = tmem_u + ACC_COL + g * 32, ta = tmem_u + A_COL + g * 64;
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
          const uint64_t bh = tc::make_desc(smem_u + ks * 32), bl = tc::make_desc(smem_u + 4096 + ks * 32);
          tc::mma_n32_ts(tmem_d, ta + ks * 8, bh, ks > 0);
          tc::mma_n32_ts(tmem_d, ta + 32 + ks * 8, bh, 1);
          tc::mma_n32_ts(tmem_d, ta + ks * 8, bl, 1);
        }
        tc::mma_commit_raw(&done[g]);
      }
      __syncwarp();
    }
  } else {
    // =============================================================== two groups of 4 warps: load, convert, store
    const int g = warp >> 2, quad = warp & 3;
    const int gt = t & 127;                                         // thread within the group = position row = TMEM lane
    const uint32_t tlane = tmem_base + ((uint32_t)(quad * 32) << 16);
    unsigned char* buf = sTile + g * 2 * 16384;
    const int ngrp = (nmine - g + 1) / 2;                           // this group's tiles: i = g, g + 2, ...
    auto prefetch = [&](int j) {
      const long long pos0 = ((long long)blockIdx.x + (long long)(2 * j + g) * gridDim.x) * 128;
      unsigned char* dst = buf + (j & 1) * 16384;
      const int chunk = gt & 7;
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const int r = (gt >> 3) + 16 * k;
        const bool ok = pos0 + r < npos;
        cp_async16(dst + r * 128 + ((chunk ^ (r & 7)) << 4), x + (ok ? (pos0 + r) * 32 + chunk * 4 : 0), ok);
      }
      cp_async_commit();
    };
    if (ngrp > 0) prefetch(0);
    for (int j = 0; j < ngrp; ++j) {
      cp_async_wait<0>();
      asm volatile("bar.sync %0, 128;\n" :: "r"(1 + g) : "memory");   // the whole tile has landed; buffer (j+1)&1 is no longer read
      if (j + 1 < ngrp) prefetch(j + 1);
      const unsigned char* rowp = buf + (j & 1) * 16384 + gt * 128;
      float4 v[8];
#pragma unroll
      for (int c = 0; c < 8; ++c) v[c] = *reinterpret_cast<const float4*>(rowp + ((c ^ (gt & 7)) << 4));
      // the A slot and the accumulator of this group are free: done[g] of tile j-1 was waited for below
      const uint32_t ta = tlane + A_COL + g * 64;
      {
        uint32_t h[32];
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          h[4 * c] = tc::tf32_hi_bits(v[c].x); h[4 * c + 1] = tc::tf32_hi_bits(v[c].y);
          h[4 * c + 2] = tc::tf32_hi_bits(v[c].z); h[4 * c + 3] = tc::tf32_hi_bits(v[c].w);
        }
        tc::tmem_st32(ta, h);
        uint32_t l[32];
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          l[4 * c] = __float_as_uint(v[c].x - __uint_as_float(h[4 * c])); l[4 * c + 1] = __float_as_uint(v[c].y - __uint_as_float(h[4 * c + 1]));
          l[4 * c + 2] = __float_as_uint(v[c].z - __uint_as_float(h[4 * c + 2])); l[4 * c + 3] = __float_as_uint(v[c].w - __uint_as_float(h[4 * c + 3]));
        }
        tc::tmem_st32(ta + 32, l);
      }
      tc::tmem_wait_st();
      tc::tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&a_full[g]);
      tc::mbar_wait(&done[g], j & 1);
      tc::tc_fence_after();
      float acc[32];
      tc::tmem_ld32(tlane + ACC_COL + g * 32, acc);
      tc::tc_fence_before();
      const long long p = ((long long)blockIdx.x + (long long)(2 * j + g) * gridDim.x) * 128 + gt;
      if (p < npos) {
        const long long slice = p / plane;
        const int q = (int)(p - slice * plane);
        float* out = taps + slice * (long long)NT * plane + q;
#pragma unroll
        for (int k = 0; k < NT; ++k) __stcg(out + (size_t)k * plane, acc[k]);
      }
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 8) {
    tc::tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;\n" :: "r"(tmem_base) : "memory");
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

// Input image of the refinement's first layer, packed for snb_conv_c4_ws: x4[b][y][x] = (mul * bilinear(coarse), r, g, b)
// (stereo_net.py:105-113: upsample, scale by W / w, concat with the guidance image), plus the upsampled plane itself.
__global__ void __launch_bounds__(256)
refine_pack_input_kernel(const float* __restrict__ coarse, const float* __restrict__ rgb, float4* __restrict__ x4,
                         float* __restrict__ up, int h, int w, int H, int W, float mul) {
  pdl_launch(); pdl_wait();
  const int x = blockIdx.x * 256 + threadIdx.x;
  const int y = blockIdx.y, b = blockIdx.z;
  if (x >= W) return;
  const float v = mul * bilinear_sample(coarse + (size_t)b * h * w, h, w, y, x, (float)h / (float)H, (float)w / (float)W);
  const size_t plane = (size_t)H * W, o = (size_t)y * W + x;
  const float* img = rgb + (size_t)b * 3 * plane;
  x4[(size_t)b * plane + o] = make_float4(v, __ldg(img + o), __ldg(img + plane + o), __ldg(img + 2 * plane + o));
  up[(size_t)b * plane + o] = v;
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
  static const bool use_ffma = []() { const char* s = getenv("SNB200_TAPS"); return s != nullptr && strcmp(s, "ffma") == 0; }();
  if (!use_ffma) {
    SNB_REQUIRE((reinterpret_cast<uintptr_t>(x) & 15) == 0, "snb_conv_c32_taps: x must be 16-byte aligned");
    int dev = 0, sms = 148;
    SNB_CUDA(cudaGetDevice(&dev));
    SNB_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const int ctas = grid < 2 * sms ? grid : 2 * sms;
    if (ntaps == 27) {
      SNB_CUDA(cudaFuncSetAttribute(conv_c32_taps_tc_kernel<27>, cudaFuncAttributeMaxDynamicSharedMemorySize, TAPS_TC_SMEM));
      snb_launch(conv_c32_taps_tc_kernel<27>, ctas, TAPS_TC_THREADS, TAPS_TC_SMEM, stream, x, w, taps, npos, plane, grid);
    } else {
      SNB_CUDA(cudaFuncSetAttribute(conv_c32_taps_tc_kernel<9>, cudaFuncAttributeMaxDynamicSharedMemorySize, TAPS_TC_SMEM));
      snb_launch(conv_c32_taps_tc_kernel<9>, ctas, TAPS_TC_THREADS, TAPS_TC_SMEM, stream, x, w, taps, npos, plane, grid);
    }
    SNB_LAUNCH_CHECK("conv_c32_taps_tc_kernel");
    return 0;
  }
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

extern "C" int snb_refine_pack_input(const float* coarse, const float* rgb, float* x4, float* up, int B, int h, int w, int H, int W,
                                     float mul, void* stream) {
  SNB_REQUIRE(coarse && rgb && x4 && up && B > 0 && h > 0 && w > 0 && H > 0 && W > 0, "snb_refine_pack_input: bad args");
  SNB_REQUIRE((reinterpret_cast<uintptr_t>(x4) & 15) == 0, "snb_refine_pack_input: x4 must be 16-byte aligned");
  snb_launch(refine_pack_input_kernel, dim3(snb_ceil_div(W, 256), H, B), 256, 0, stream, coarse, rgb, reinterpret_cast<float4*>(x4), up,
             h, w, H, W, mul);
  SNB_LAUNCH_CHECK("refine_pack_input_kernel");
  return 0;
}

extern "C" int snb_upsample_bilinear_bwd(const float* dout, float* din, int B, int h, int w, int H, int W, float mul, void* stream) {
  SNB_REQUIRE(dout && din && B > 0 && h > 0 && w > 0 && H > 0 && W > 0, "snb_upsample_bilinear_bwd: bad args");
  snb_launch(upsample_bilinear_bwd_kernel, dim3(snb_ceil_div(w, 128), h, B), 128, 0, stream, dout, din, h, w, H, W, mul);
  SNB_LAUNCH_CHECK("upsample_bilinear_bwd_kernel");
  return 0;
}
