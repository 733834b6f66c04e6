// Generic 32->32 convolution over channels-last tensors, fp32 CUDA-core implicit GEMM.
// Replaces nn.Conv2d/nn.Conv3d (+ folded BN + LeakyReLU + residual) of stereo_net.py:10-17,23-29,44-51,64-70,77,185.
// Handles any tap set / stride / dilation (5x5 s2 feature convs, 3x3 dilated, 3x3x3) and also serves every
// stride-1 data-gradient (same kernel, flipped/transposed weights from snb_prep_conv_weights mode 1).
//
// Tiling: one CTA (128 threads) = 128 consecutive output positions x 32 output channels; K loop = taps, one tap per
// pipeline stage (A tile 128x32 fp32 = 16 KB via zero-filling cp.async, B tile 32x32 = 4 KB), double buffered.
// smem A is XOR-swizzled at 16-B granularity so that both the cp.async writes and the per-thread float4 reads are
// bank-conflict free.  Thread tile: 4 positions (p, p+32, p+64, p+96) x 8 couts -> 12 LDS.128 per 128 FFMA.
#include "common.cuh"

namespace {

constexpr int TILE_M = 128;
constexpr int NTHREADS = 128;

struct SmemC32 {
  float a[2][TILE_M][32];   // swizzled: chunk c of row r lives at chunk slot (c ^ (r & 7))
  float w[2][32][32];       // [cin][cout]
  int4 coord[TILE_M];       // (b, od, oh, ow) per tile row; b = -1 when the row is past the end
};

__device__ __forceinline__ void load_tap(SmemC32& s, int stage, const float* __restrict__ x, const float* __restrict__ wprep,
                                         const snb_conv_geom& g, int tap) {
  const int t = threadIdx.x;
  const int kw = tap % g.KW;
  const int kh = (tap / g.KW) % g.KH;
  const int kd = tap / (g.KW * g.KH);
  const int chunk = t & 7;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int r = (t >> 3) + 16 * j;
    const int4 c = s.coord[r];
    int id, ih, iw;
    bool ok = c.x >= 0;
    if (!g.transposed) {
      id = c.y + kd - g.pd;
      ih = c.z * g.stride + kh * g.dil - g.ph;
      iw = c.w * g.stride + kw * g.dil - g.pw;
    } else {          // data gradient: written position o reads (o + p - k*dil)/stride when divisible
      id = c.y + g.pd - kd;
      const int th = c.z + g.ph - kh * g.dil, tw = c.w + g.pw - kw * g.dil;
      ok = ok && th >= 0 && tw >= 0 && (th % g.stride) == 0 && (tw % g.stride) == 0;
      ih = th / g.stride; iw = tw / g.stride;
    }
    ok = ok && (unsigned)id < (unsigned)g.D && (unsigned)ih < (unsigned)g.H && (unsigned)iw < (unsigned)g.W;
    const size_t off = ok ? ((((size_t)c.x * g.D + id) * g.H + ih) * g.W + iw) * 32 + chunk * 4 : 0;
    cp_async16(&s.a[stage][r][(chunk ^ (r & 7)) * 4], x + off, ok);
  }
  const float* wt = wprep + (size_t)tap * 1024;
  cp_async16(&s.w[stage][0][0] + t * 4, wt + t * 4, true);
  cp_async16(&s.w[stage][0][0] + (t + 128) * 4, wt + (t + 128) * 4, true);
}

__global__ void __launch_bounds__(NTHREADS)
conv_c32_ffma_kernel(const float* __restrict__ x, const float* __restrict__ wprep, float* __restrict__ y,
                     snb_conv_geom g, snb_conv_epilogue e, long long npos) {
  pdl_launch(); pdl_wait();
  extern __shared__ __align__(128) unsigned char smem_raw[];
  SmemC32& s = *reinterpret_cast<SmemC32*>(smem_raw);
  const int t = threadIdx.x;
  const long long pos0 = (long long)blockIdx.x * TILE_M;

  {  // decode the tile's output coordinates once
    long long p = pos0 + t;
    int4 c;
    if (p < npos) {
      c.w = (int)(p % g.OW); p /= g.OW;
      c.z = (int)(p % g.OH); p /= g.OH;
      c.y = (int)(p % g.OD); c.x = (int)(p / g.OD);
    } else {
      c = make_int4(-1, 0, 0, 0);
    }
    s.coord[t] = c;
  }
  __syncthreads();

  const int ntaps = g.KD * g.KH * g.KW;
  const int pg = t & 31;          // positions pg + 32*i
  const int cg = t >> 5;          // couts cg*8 .. cg*8+7
  float acc[4][8];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

  load_tap(s, 0, x, wprep, g, 0);
  cp_async_commit();
  for (int tap = 0; tap < ntaps; ++tap) {
    const int st = tap & 1;
    if (tap + 1 < ntaps) {
      load_tap(s, st ^ 1, x, wprep, g, tap + 1);
      cp_async_commit();
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    const int sw = pg & 7;
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      float4 xa[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) xa[i] = *reinterpret_cast<const float4*>(&s.a[st][pg + 32 * i][(c ^ sw) * 4]);
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        const float4 w0 = *reinterpret_cast<const float4*>(&s.w[st][c * 4 + kk][cg * 8]);
        const float4 w1 = *reinterpret_cast<const float4*>(&s.w[st][c * 4 + kk][cg * 8 + 4]);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float xv = kk == 0 ? xa[i].x : kk == 1 ? xa[i].y : kk == 2 ? xa[i].z : xa[i].w;
          acc[i][0] = fmaf(xv, w0.x, acc[i][0]); acc[i][1] = fmaf(xv, w0.y, acc[i][1]);
          acc[i][2] = fmaf(xv, w0.z, acc[i][2]); acc[i][3] = fmaf(xv, w0.w, acc[i][3]);
          acc[i][4] = fmaf(xv, w1.x, acc[i][4]); acc[i][5] = fmaf(xv, w1.y, acc[i][5]);
          acc[i][6] = fmaf(xv, w1.z, acc[i][6]); acc[i][7] = fmaf(xv, w1.w, acc[i][7]);
        }
      }
    }
    __syncthreads();
  }

  // ---- epilogue
  float bias[8], sc[8], sh[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    bias[j] = e.bias ? e.bias[cg * 8 + j] : 0.f;
    sc[j] = e.scale ? e.scale[cg * 8 + j] : 1.f;
    sh[j] = e.scale ? e.shift[cg * 8 + j] : 0.f;
  }
  float s1[8], s2[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) { s1[j] = 0.f; s2[j] = 0.f; }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const long long p = pos0 + pg + 32 * i;
    const bool ok = p < npos;
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      v[j] = acc[i][j] + bias[j];
      if (ok) { s1[j] += v[j]; s2[j] = fmaf(v[j], v[j], s2[j]); }
      if (e.scale) v[j] = fmaf(v[j], sc[j], sh[j]);
      if (e.lrelu) v[j] = lrelu(v[j]);
    }
    if (ok) {
      float* yp = y + p * 32 + cg * 8;
      if (e.residual) {
        const float4 r0 = *reinterpret_cast<const float4*>(e.residual + p * 32 + cg * 8);
        const float4 r1 = *reinterpret_cast<const float4*>(e.residual + p * 32 + cg * 8 + 4);
        v[0] += r0.x; v[1] += r0.y; v[2] += r0.z; v[3] += r0.w;
        v[4] += r1.x; v[5] += r1.y; v[6] += r1.z; v[7] += r1.w;
      }
      *reinterpret_cast<float4*>(yp) = make_float4(v[0], v[1], v[2], v[3]);
      *reinterpret_cast<float4*>(yp + 4) = make_float4(v[4], v[5], v[6], v[7]);
    }
  }
  if (e.stats) {   // per-tile partial (sum, sum of squares) per channel; each warp owns 8 distinct channels
#pragma unroll
    for (int j = 0; j < 8; ++j) { s1[j] = warp_sum(s1[j]); s2[j] = warp_sum(s2[j]); }
    if (pg == 0) {
      float* sp = e.stats + (size_t)blockIdx.x * 64;
#pragma unroll
      for (int j = 0; j < 8; ++j) { sp[cg * 8 + j] = s1[j]; sp[32 + cg * 8 + j] = s2[j]; }
    }
  }
}

__global__ void prep_weights_kernel(const float* __restrict__ w, float* __restrict__ out, int cout, int cin, int taps, int mode) {
  pdl_launch(); pdl_wait();
  const int n = cout * cin * taps;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const int t = i % taps;
    const int ci = (i / taps) % cin;
    const int co = i / (taps * cin);
    if (mode == 0)      out[((size_t)t * cin + ci) * cout + co] = w[i];                 // [t][cin][cout]
    else if (mode == 1) out[((size_t)(taps - 1 - t) * cout + co) * cin + ci] = w[i];    // [t'][cout as K][cin as N]
    else                out[((size_t)t * cout + co) * cin + ci] = w[i];                 // [t][cout as K][cin as N]
  }
}

}  // namespace

static int check_geom(const snb_conv_geom* g, const char* who) {
  SNB_REQUIRE(g != nullptr, "%s: null geometry", who);
  SNB_REQUIRE(g->B > 0 && g->D > 0 && g->H > 0 && g->W > 0 && g->OD > 0 && g->OH > 0 && g->OW > 0, "%s: bad dims", who);
  SNB_REQUIRE(g->KD > 0 && g->KH > 0 && g->KW > 0 && g->stride > 0 && g->dil > 0, "%s: bad kernel", who);
  return 0;
}

extern "C" int snb_conv_c32_num_tiles(const snb_conv_geom* g) {
  if (!g) return -1;
  const long long npos = (long long)g->B * g->OD * g->OH * g->OW;
  return snb_ceil_div(npos, TILE_M);
}

extern "C" int snb_conv_c32(const float* x, const float* wprep, float* y, const snb_conv_geom* g, const snb_conv_epilogue* e, void* stream) {
  if (int rc = check_geom(g, "snb_conv_c32")) return rc;
  SNB_REQUIRE(x && wprep && y && e, "snb_conv_c32: null pointer");
  SNB_REQUIRE(!e->scale || e->shift, "snb_conv_c32: scale without shift");
  const long long npos = (long long)g->B * g->OD * g->OH * g->OW;
  const int ntiles = snb_ceil_div(npos, TILE_M);
  const int smem = (int)sizeof(SmemC32);
  SNB_CUDA(cudaFuncSetAttribute(conv_c32_ffma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  snb_launch(conv_c32_ffma_kernel, ntiles, NTHREADS, smem, stream, x, wprep, y, *g, *e, npos);
  SNB_LAUNCH_CHECK("conv_c32_ffma_kernel");
  return 0;
}

extern "C" int snb_prep_conv_weights(const float* w, float* out, int cout, int cin, int taps, int mode, void* stream) {
  SNB_REQUIRE(w && out && cout > 0 && cin > 0 && taps > 0 && mode >= 0 && mode <= 2, "snb_prep_conv_weights: bad args");
  const int n = cout * cin * taps;
  snb_launch(prep_weights_kernel, snb_ceil_div(n, 256), 256, 0, stream, w, out, cout, cin, taps, mode);
  SNB_LAUNCH_CHECK("prep_weights_kernel");
  return 0;
}
