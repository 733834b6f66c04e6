// Backward kernels of the adaptation step (adapt.py:381-394 triggers autograd through stereo_net.py:79-85,168-207).
// fp32 CUDA-core kernels; data gradients of the 3x3 / 3x3x3 convolutions reuse the forward convolution kernels
// (snb_conv_c32 / snb_conv_c32_tc with re-packed weights), everything else lives here:
//   * BatchNorm(train|eval) + LeakyReLU backward (two-pass: per-channel reductions, then dz)
//   * weight gradient of the 32->32 convolutions (per-CTA partials + deterministic tree-free reduction)
//   * bias gradients, soft-argmin backward, 32->1 convolution backward (tap form), ReLU mask
#include "common.cuh"

namespace {

// ------------------------------------------------------------------------------------------------ BN + LeakyReLU backward
// u = z*scale + shift ; du = dy * (u > 0 ? 1 : 0.2) ; zhat = (z - mean) * invstd
// partial[blk][0..31] = sum du, partial[blk][32..63] = sum du*zhat        (per channel)
__global__ void __launch_bounds__(256)
bn_lrelu_bwd_reduce_kernel(const float4* __restrict__ z, const float4* __restrict__ dy, const float* __restrict__ scale,
                           const float* __restrict__ shift, const float* __restrict__ mean, const float* __restrict__ invstd,
                           float* __restrict__ partial, long long n4, int do_lrelu) {
  pdl_launch(); pdl_wait();
  __shared__ float red[2][256][4];
  const int t = threadIdx.x, ch = t & 7;                    // this thread always sees channels 4*ch .. 4*ch+3
  const float4 sc = reinterpret_cast<const float4*>(scale)[ch], sh = reinterpret_cast<const float4*>(shift)[ch];
  const float4 mu = reinterpret_cast<const float4*>(mean)[ch], is = reinterpret_cast<const float4*>(invstd)[ch];
  float a[4] = {0.f, 0.f, 0.f, 0.f}, b[4] = {0.f, 0.f, 0.f, 0.f};
  for (long long i = (long long)blockIdx.x * 256 + t; i < n4; i += (long long)gridDim.x * 256) {
    const float4 zv = z[i], g = dy[i];
    const float zz[4] = {zv.x, zv.y, zv.z, zv.w}, gg[4] = {g.x, g.y, g.z, g.w};
    const float s4[4] = {sc.x, sc.y, sc.z, sc.w}, h4[4] = {sh.x, sh.y, sh.z, sh.w};
    const float m4[4] = {mu.x, mu.y, mu.z, mu.w}, i4[4] = {is.x, is.y, is.z, is.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float u = fmaf(zz[k], s4[k], h4[k]);
      const float du = (do_lrelu && u <= 0.f) ? gg[k] * SNB_LRELU_SLOPE : gg[k];
      a[k] += du;
      b[k] = fmaf(du, (zz[k] - m4[k]) * i4[k], b[k]);
    }
  }
#pragma unroll
  for (int k = 0; k < 4; ++k) { red[0][t][k] = a[k]; red[1][t][k] = b[k]; }
  __syncthreads();
  if (t < 64) {                                 // t -> (which, channel): sum the 32 threads that share chunk ch
    const int which = t >> 5, c = t & 31, chunk = c >> 2, k = c & 3;
    float s = 0.f;
    for (int j = 0; j < 32; ++j) s += red[which][j * 8 + chunk][k];
    partial[(size_t)blockIdx.x * 64 + t] = s;
  }
}

// dz = scale * (du - train * (sum_du / N + zhat * sum_duz / N));  dzpart[blk][c] = per-block sum of dz (conv bias grad)
__global__ void __launch_bounds__(256)
bn_lrelu_bwd_apply_kernel(const float4* __restrict__ z, const float4* __restrict__ dy, const float* __restrict__ scale,
                          const float* __restrict__ shift, const float* __restrict__ mean, const float* __restrict__ invstd,
                          const float* __restrict__ sums, float inv_count, int train, int do_lrelu,
                          float4* __restrict__ dz, float* __restrict__ dzpart, long long n4) {
  pdl_launch(); pdl_wait();
  __shared__ float red[256][4];
  const int t = threadIdx.x, ch = t & 7;
  const float4 sc = reinterpret_cast<const float4*>(scale)[ch], sh = reinterpret_cast<const float4*>(shift)[ch];
  const float4 mu = reinterpret_cast<const float4*>(mean)[ch], is = reinterpret_cast<const float4*>(invstd)[ch];
  const float4 sa = reinterpret_cast<const float4*>(sums)[ch], sb = reinterpret_cast<const float4*>(sums + 32)[ch];
  const float s4[4] = {sc.x, sc.y, sc.z, sc.w}, h4[4] = {sh.x, sh.y, sh.z, sh.w};
  const float m4[4] = {mu.x, mu.y, mu.z, mu.w}, i4[4] = {is.x, is.y, is.z, is.w};
  const float a4[4] = {sa.x * inv_count, sa.y * inv_count, sa.z * inv_count, sa.w * inv_count};
  const float b4[4] = {sb.x * inv_count, sb.y * inv_count, sb.z * inv_count, sb.w * inv_count};
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  for (long long i = (long long)blockIdx.x * 256 + t; i < n4; i += (long long)gridDim.x * 256) {
    const float4 zv = z[i], g = dy[i];
    const float zz[4] = {zv.x, zv.y, zv.z, zv.w}, gg[4] = {g.x, g.y, g.z, g.w};
    float o[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float u = fmaf(zz[k], s4[k], h4[k]);
      float du = (do_lrelu && u <= 0.f) ? gg[k] * SNB_LRELU_SLOPE : gg[k];
      if (train) du = du - a4[k] - (zz[k] - m4[k]) * i4[k] * b4[k];
      o[k] = du * s4[k];
      acc[k] += o[k];
    }
    dz[i] = make_float4(o[0], o[1], o[2], o[3]);
  }
#pragma unroll
  for (int k = 0; k < 4; ++k) red[t][k] = acc[k];
  __syncthreads();
  if (t < 32) {
    const int chunk = t >> 2, k = t & 3;
    float s = 0.f;
    for (int j = 0; j < 32; ++j) s += red[j * 8 + chunk][k];
    dzpart[(size_t)blockIdx.x * 32 + t] = s;
  }
}

// out[j] = mul * sum_i partial[i][j]: block = 32 columns x 8 row groups, double accumulation, fixed order -> deterministic
// taps > 0: the row is a [taps][32 ci][32 co] weight-gradient block and is written in PyTorch's [co][ci][taps] layout.
__global__ void __launch_bounds__(256)
reduce_partials_kernel(const float* __restrict__ partial, int n, int len, float* __restrict__ out, float mul, int taps) {
  pdl_launch(); pdl_wait();
  __shared__ double red[8][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int j = blockIdx.x * 32 + tx;
  double s = 0.0;
  if (j < len)
    for (int i = ty; i < n; i += 8) s += (double)partial[(size_t)i * len + j];
  red[ty][tx] = s;
  __syncthreads();
  if (ty == 0 && j < len) {
    double a = ((red[0][tx] + red[1][tx]) + (red[2][tx] + red[3][tx])) + ((red[4][tx] + red[5][tx]) + (red[6][tx] + red[7][tx]));
    int o = j;
    if (taps > 0) { const int co = j & 31, ci = (j >> 5) & 31, tap = j >> 10; o = (co * 32 + ci) * taps + tap; }
    out[o] = (float)(a * (double)mul);
  }
}

// Narrow outputs (len <= 512: BN sums, bias gradients, 32->1 conv gradients) over many partial rows: the kernel above would
// run 2-16 blocks whose threads each walk n/8 rows serially (8-14 us of pure latency, 80 of them per adaptation step).
// Here a block owns 4 columns and 64 row lanes (4 independent accumulators each), then a fixed-order tree -> deterministic.
__global__ void __launch_bounds__(256)
reduce_partials_narrow_kernel(const float* __restrict__ partial, int n, int len, float* __restrict__ out, float mul) {
  pdl_launch(); pdl_wait();
  __shared__ double red[64][4];
  const int c = threadIdx.x & 3, r = threadIdx.x >> 2;
  const int j = blockIdx.x * 4 + c;
  double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
  if (j < len) {
    int i = r;
    for (; i + 192 < n; i += 256) {
      s0 += (double)partial[(size_t)i * len + j];         s1 += (double)partial[(size_t)(i + 64) * len + j];
      s2 += (double)partial[(size_t)(i + 128) * len + j]; s3 += (double)partial[(size_t)(i + 192) * len + j];
    }
    for (; i < n; i += 64) s0 += (double)partial[(size_t)i * len + j];
  }
  red[r][c] = (s0 + s1) + (s2 + s3);
  __syncthreads();
  for (int off = 32; off > 0; off >>= 1) {
    if (r < off) red[r][c] += red[r + off][c];
    __syncthreads();
  }
  if (r == 0 && j < len) out[j] = (float)(red[0][c] * (double)mul);
}

// per-channel column sums of a [npos][32] tensor -> partial[blk][32]
__global__ void __launch_bounds__(256)
channel_sum_kernel(const float4* __restrict__ x, float* __restrict__ partial, long long n4) {
  pdl_launch(); pdl_wait();
  __shared__ float red[256][4];
  const int t = threadIdx.x;
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  for (long long i = (long long)blockIdx.x * 256 + t; i < n4; i += (long long)gridDim.x * 256) {
    const float4 v = x[i];
    acc[0] += v.x; acc[1] += v.y; acc[2] += v.z; acc[3] += v.w;
  }
#pragma unroll
  for (int k = 0; k < 4; ++k) red[t][k] = acc[k];
  __syncthreads();
  if (t < 32) {
    const int chunk = t >> 2, k = t & 3;
    float s = 0.f;
    for (int j = 0; j < 32; ++j) s += red[j * 8 + chunk][k];
    partial[(size_t)blockIdx.x * 32 + t] = s;
  }
}

// ------------------------------------------------------------------------------------------------ 32->32 weight gradient
// dW[tap][ci][co] = sum_p x[p + off(tap)][ci] * dz[p][co].  One CTA walks a contiguous range of 128-position chunks; per
// chunk the dz tile is staged once, each tap's shifted x tile is staged with zero-filling cp.async, and 128 threads
// (4 position groups x 32 threads, each 4 ci x 8 co) accumulate the 32x32 block, reduce the 4 groups through smem and add
// into a smem accumulator [taps][32][32].  Partials [cta][taps][32][32] are summed by reduce_partials_kernel.
struct SmemWgrad {
  float x[2][128][32];   // swizzled rows, double buffered over taps
  float dz[128][32];     // swizzled rows
  float red[8][32][33];  // cross-group reduction scratch (padded)
  int4 coord[128];
};

__device__ __forceinline__ void wgrad_load_x(SmemWgrad& s, int buf, const float* __restrict__ x, const snb_conv_geom& g, int tap) {
  const int t = threadIdx.x;
  const int kw = tap % g.KW, kh = (tap / g.KW) % g.KH, kd = tap / (g.KW * g.KH);
  const int chunkc = t & 7;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int r = (t >> 3) + 32 * j;
    const int4 c = s.coord[r];
    const int id = c.y + kd - g.pd, ih = c.z * g.stride + kh * g.dil - g.ph, iw = c.w * g.stride + kw * g.dil - g.pw;
    const bool ok = (c.x >= 0) && (unsigned)id < (unsigned)g.D && (unsigned)ih < (unsigned)g.H && (unsigned)iw < (unsigned)g.W;
    const size_t off = ok ? ((((size_t)c.x * g.D + id) * g.H + ih) * g.W + iw) * 32 + chunkc * 4 : 0;
    cp_async16(&s.x[buf][r][(chunkc ^ (r & 7)) * 4], x + off, ok);
  }
}

// 256 threads = 8 position groups (16 positions each) x 32 threads (4 ci x 8 co each).
__global__ void __launch_bounds__(256)
conv_c32_wgrad_kernel(const float* __restrict__ x, const float* __restrict__ dz, float* __restrict__ partial,
                      snb_conv_geom g, long long npos, int chunks_per_cta) {
  pdl_launch(); pdl_wait();
  extern __shared__ __align__(128) unsigned char smem_raw[];
  SmemWgrad& s = *reinterpret_cast<SmemWgrad*>(smem_raw);
  float* sAcc = reinterpret_cast<float*>(smem_raw + sizeof(SmemWgrad));     // [taps][32 ci][32 co]
  const int t = threadIdx.x;
  const int ntaps = g.KD * g.KH * g.KW;
  for (int i = t; i < ntaps * 1024; i += 256) sAcc[i] = 0.f;
  const int grp = t >> 5, l = t & 31;
  const int cig = l >> 2, cog = l & 3;             // ci = cig*4 .. +3 ; co = cog*8 .. +7
  const long long chunk0 = (long long)blockIdx.x * chunks_per_cta;
  const long long nchunks = (npos + 127) / 128;

  for (int cc = 0; cc < chunks_per_cta; ++cc) {
    const long long chunk = chunk0 + cc;
    if (chunk >= nchunks) break;
    const long long pos0 = chunk * 128;
    __syncthreads();                       // previous chunk fully consumed (x, dz, coord, red)
    if (t < 128) {
      long long p = pos0 + t;
      int4 c;
      if (p < npos) {
        c.w = (int)(p % g.OW); p /= g.OW;
        c.z = (int)(p % g.OH); p /= g.OH;
        c.y = (int)(p % g.OD); c.x = (int)(p / g.OD);
      } else c = make_int4(-1, 0, 0, 0);
      s.coord[t] = c;
    }
    {   // dz tile (output positions are contiguous in memory)
      const int chunkc = t & 7;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int r = (t >> 3) + 32 * j;
        const bool ok = pos0 + r < npos;
        cp_async16(&s.dz[r][(chunkc ^ (r & 7)) * 4], dz + (ok ? (pos0 + r) * 32 + chunkc * 4 : 0), ok);
      }
    }
    __syncthreads();                       // coord visible
    wgrad_load_x(s, 0, x, g, 0);
    cp_async_commit();
    for (int tap = 0; tap < ntaps; ++tap) {
      const int buf = tap & 1;
      if (tap + 1 < ntaps) {
        wgrad_load_x(s, buf ^ 1, x, g, tap + 1);      // buffer buf^1 was last read two barriers ago
        cp_async_commit();
        cp_async_wait<1>();
      } else {
        cp_async_wait<0>();
      }
      __syncthreads();
      unsigned long long pacc[4][4];           // 4 ci x 8 co as packed pairs (FFMA2)
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) pacc[i][j] = 0ull;
#pragma unroll 4
      for (int pp = 0; pp < 16; ++pp) {
        const int r = grp * 16 + pp;
        const float4 xv = *reinterpret_cast<const float4*>(&s.x[buf][r][(cig ^ (r & 7)) * 4]);
        const ulonglong2 d0 = *reinterpret_cast<const ulonglong2*>(&s.dz[r][((cog * 2) ^ (r & 7)) * 4]);
        const ulonglong2 d1 = *reinterpret_cast<const ulonglong2*>(&s.dz[r][((cog * 2 + 1) ^ (r & 7)) * 4]);
        const float xx[4] = {xv.x, xv.y, xv.z, xv.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const unsigned long long xp = pack2(xx[i], xx[i]);
          ffma2(pacc[i][0], xp, d0.x); ffma2(pacc[i][1], xp, d0.y);
          ffma2(pacc[i][2], xp, d1.x); ffma2(pacc[i][3], xp, d1.y);
        }
      }
      float acc[4][8];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) { acc[i][2 * j] = lo2(pacc[i][j]); acc[i][2 * j + 1] = hi2(pacc[i][j]); }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) s.red[grp][cig * 4 + i][cog * 8 + j] = acc[i][j];
      __syncthreads();
      float* accp = sAcc + (size_t)tap * 1024;
      for (int i = t; i < 1024; i += 256) {
        const int ci = i >> 5, co = i & 31;
        accp[i] += ((s.red[0][ci][co] + s.red[1][ci][co]) + (s.red[2][ci][co] + s.red[3][ci][co])) +
                   ((s.red[4][ci][co] + s.red[5][ci][co]) + (s.red[6][ci][co] + s.red[7][ci][co]));
      }
      // s.red is rewritten after the next tap's barrier; s.x[buf] is refilled only at tap+2, after that barrier too
    }
  }
  __syncthreads();
  float* out = partial + (size_t)blockIdx.x * ntaps * 1024;
  for (int i = t; i < ntaps * 1024; i += 256) out[i] = sAcc[i];
}

// ------------------------------------------------------------------------------------------------ soft-argmin backward
// pred = sum_d d * softmax(cost)_d  ->  dcost_d = p_d * (d - pred) * dpred  (+ dcost_extra)
__global__ void __launch_bounds__(256)
softargmin_bwd_kernel(const float* __restrict__ cost, const float* __restrict__ pred, const float* __restrict__ dpred,
                      const float* __restrict__ extra, float* __restrict__ dcost, int D, long long plane, long long total) {
  pdl_launch(); pdl_wait();
  const long long i = (long long)blockIdx.x * 256 + threadIdx.x;     // over B*H*W
  if (i >= total) return;
  const long long b = i / plane, q = i - b * plane;
  const float* c = cost + b * D * plane + q;
  float m = -INFINITY;
  for (int d = 0; d < D; ++d) m = fmaxf(m, c[(size_t)d * plane]);
  float s = 0.f;
  for (int d = 0; d < D; ++d) s += __expf(c[(size_t)d * plane] - m);
  const float inv = 1.f / s, pr = pred[i], gp = dpred ? dpred[i] : 0.f;
  for (int d = 0; d < D; ++d) {
    const float p = __expf(c[(size_t)d * plane] - m) * inv;
    float g = p * ((float)d - pr) * gp;
    if (extra) g += extra[b * D * plane + (size_t)d * plane + q];
    dcost[b * D * plane + (size_t)d * plane + q] = g;
  }
}

// dres = dout * (out > 0)    (ReLU of stereo_net.py:121)
__global__ void __launch_bounds__(256)
relu_bwd_kernel(const float* __restrict__ out, const float* __restrict__ dout, float* __restrict__ dres, long long n) {
  pdl_launch(); pdl_wait();
  const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
  if (i < n) dres[i] = out[i] > 0.f ? dout[i] : 0.f;
}

// ------------------------------------------------------------------------------------------------ 32->1 conv backward (tap form)
// forward: cost[v] = b + sum_tap sum_c w[tap][c] * x[v + off(tap)][c].   Given g = dcost (one plane per slice):
//   dx[u][c]   = sum_tap w[tap][c] * g[u - off(tap)]
//   dw[tap][c] = sum_u x[u][c] * g[u - off(tap)]          db = sum g
// One CTA = 128 positions u: gathers G[u][tap] = g[u - off(tap)] into smem, every thread forms its dx row, then 128 threads
// reduce the CTA's dw partial.  partial[blk] = [NT][32] dw + 1 float db.
template <int NT>
__global__ void __launch_bounds__(128)
conv_c32_taps_bwd_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ g,
                         float* __restrict__ dx, float* __restrict__ partial, long long npos, int D, int H, int W) {
  pdl_launch(); pdl_wait();
  constexpr int NTP = NT + 1;
  __shared__ __align__(16) float sX[128][32];
  __shared__ float sG[128][NTP];
  __shared__ __align__(16) float sW[NT][32];
  __shared__ float sB[4];
  const int t = threadIdx.x;
  const long long pos0 = (long long)blockIdx.x * 128;
  {
    const int chunk = t & 7;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int r = (t >> 3) + 16 * j;
      const bool ok = pos0 + r < npos;
      cp_async16(&sX[r][(chunk ^ (r & 7)) * 4], x + (ok ? (pos0 + r) * 32 + chunk * 4 : 0), ok);
    }
    cp_async_commit();
  }
  for (int i = t; i < NT * 32; i += 128) { const int tap = i >> 5, c = i & 31; sW[tap][c] = w[c * NT + tap]; }   // w is [1][32][NT]
  const long long u = pos0 + t;
  float gc = 0.f;
  if (u < npos) {
    long long r = u;
    const int xw = (int)(r % W); r /= W;
    const int yh = (int)(r % H); r /= H;
    const int d = (int)(r % D); const long long b = r / D;
    constexpr int KD = NT / 9;
#pragma unroll
    for (int tap = 0; tap < NT; ++tap) {
      const int kw = tap % 3, kh = (tap / 3) % 3, kd = tap / 9;
      const int dd = d - (kd - (KD == 3 ? 1 : 0)), yy = yh - (kh - 1), xx = xw - (kw - 1);
      float v = 0.f;
      if ((unsigned)dd < (unsigned)D && (unsigned)yy < (unsigned)H && (unsigned)xx < (unsigned)W)
        v = g[((b * D + dd) * H + yy) * (long long)W + xx];
      sG[t][tap] = v;
    }
    gc = g[u];
  } else {
#pragma unroll
    for (int tap = 0; tap < NT; ++tap) sG[t][tap] = 0.f;
  }
  cp_async_wait<0>();
  __syncthreads();
  if (u < npos) {       // dx row of this position
    unsigned long long pacc[16];             // 32 channels as packed pairs (FFMA2)
#pragma unroll
    for (int c = 0; c < 16; ++c) pacc[c] = 0ull;
#pragma unroll 3
    for (int tap = 0; tap < NT; ++tap) {
      const float gv = sG[t][tap];
      const unsigned long long gp = pack2(gv, gv);
#pragma unroll
      for (int c4 = 0; c4 < 8; ++c4) {
        const ulonglong2 wv = *reinterpret_cast<const ulonglong2*>(&sW[tap][c4 * 4]);
        ffma2(pacc[2 * c4], gp, wv.x); ffma2(pacc[2 * c4 + 1], gp, wv.y);
      }
    }
    float acc[32];
#pragma unroll
    for (int c = 0; c < 16; ++c) { acc[2 * c] = lo2(pacc[c]); acc[2 * c + 1] = hi2(pacc[c]); }
    float4* o = reinterpret_cast<float4*>(dx + u * 32);
#pragma unroll
    for (int c4 = 0; c4 < 8; ++c4) o[c4] = make_float4(acc[c4 * 4], acc[c4 * 4 + 1], acc[c4 * 4 + 2], acc[c4 * 4 + 3]);
  }
  // dw partial: thread (c = t & 31, tg = t >> 5) handles taps tg, tg+4, ...
  {
    const int c = t & 31, tg = t >> 5;
    float* out = partial + (size_t)blockIdx.x * (NT * 32 + 1);
    for (int tap = tg; tap < NT; tap += 4) {
      float a = 0.f;
      for (int p = 0; p < 128; ++p) a = fmaf(sX[p][(((c >> 2) ^ (p & 7)) << 2) + (c & 3)], sG[p][tap], a);
      out[tap * 32 + c] = a;
    }
    gc = warp_sum(gc);
    if (c == 0) sB[tg] = gc;
    __syncthreads();
    if (t == 0) out[NT * 32] = (sB[0] + sB[1]) + (sB[2] + sB[3]);
  }
}

}  // namespace

static int grid_for(long long n4) {
  long long blocks = (n4 + 255) / 256;
  if (blocks > 148 * 4) blocks = 148 * 4;
  if (blocks < 1) blocks = 1;
  return (int)blocks;
}

extern "C" int snb_bwd_num_blocks(long long npos) { return grid_for(npos * 8); }

extern "C" int snb_bn_lrelu_bwd_reduce(const float* z, const float* dy, const float* scale, const float* shift, const float* mean,
                                       const float* invstd, float* partial, long long npos, int lrelu_flag, void* stream) {
  SNB_REQUIRE(z && dy && scale && shift && mean && invstd && partial && npos > 0, "snb_bn_lrelu_bwd_reduce: bad args");
  const long long n4 = npos * 8;
  snb_launch(bn_lrelu_bwd_reduce_kernel, grid_for(n4), 256, 0, stream, (const float4*)z, (const float4*)dy, scale, shift, mean,
                                                                            invstd, partial, n4, lrelu_flag);
  SNB_LAUNCH_CHECK("bn_lrelu_bwd_reduce_kernel");
  return 0;
}

extern "C" int snb_bn_lrelu_bwd_apply(const float* z, const float* dy, const float* scale, const float* shift, const float* mean,
                                      const float* invstd, const float* sums, long long npos, int train, int lrelu_flag,
                                      float* dz, float* dzpart, void* stream) {
  SNB_REQUIRE(z && dy && scale && shift && mean && invstd && sums && dz && dzpart && npos > 0, "snb_bn_lrelu_bwd_apply: bad args");
  const long long n4 = npos * 8;
  snb_launch(bn_lrelu_bwd_apply_kernel, grid_for(n4), 256, 0, stream, (const float4*)z, (const float4*)dy, scale, shift, mean,
                                                                           invstd, sums, 1.0f / (float)npos, train, lrelu_flag,
                                                                           (float4*)dz, dzpart, n4);
  SNB_LAUNCH_CHECK("bn_lrelu_bwd_apply_kernel");
  return 0;
}

extern "C" int snb_reduce_partials(const float* partial, int n, int len, float* out, float mul, void* stream) {
  SNB_REQUIRE(partial && out && n > 0 && len > 0, "snb_reduce_partials: bad args");
  if (len <= 512 && n >= 64)
    snb_launch(reduce_partials_narrow_kernel, snb_ceil_div(len, 4), 256, 0, stream, partial, n, len, out, mul);
  else
    snb_launch(reduce_partials_kernel, snb_ceil_div(len, 32), 256, 0, stream, partial, n, len, out, mul, 0);
  SNB_LAUNCH_CHECK("reduce_partials_kernel");
  return 0;
}

extern "C" int snb_reduce_wgrad_partials(const float* partial, int n, int taps, float* out, void* stream) {
  SNB_REQUIRE(partial && out && n > 0 && taps > 0, "snb_reduce_wgrad_partials: bad args");
  snb_launch(reduce_partials_kernel, taps * 32, 256, 0, stream, partial, n, taps * 1024, out, 1.0f, taps);
  SNB_LAUNCH_CHECK("reduce_partials_kernel");
  return 0;
}

extern "C" int snb_channel_sum(const float* x, float* partial, long long npos, void* stream) {
  SNB_REQUIRE(x && partial && npos > 0, "snb_channel_sum: bad args");
  const long long n4 = npos * 8;
  snb_launch(channel_sum_kernel, grid_for(n4), 256, 0, stream, (const float4*)x, partial, n4);
  SNB_LAUNCH_CHECK("channel_sum_kernel");
  return 0;
}

static int wgrad_plan(const snb_conv_geom* g, int* ctas, int* chunks_per_cta) {
  const long long npos = (long long)g->B * g->OD * g->OH * g->OW;
  const long long nchunks = (npos + 127) / 128;
  long long c = nchunks < 148 * 2 ? nchunks : 148 * 2;
  if (c < 1) c = 1;
  const long long per = (nchunks + c - 1) / c;
  *ctas = (int)((nchunks + per - 1) / per);
  *chunks_per_cta = (int)per;
  return 0;
}

extern "C" int snb_conv_c32_wgrad_num_partials(const snb_conv_geom* g) {
  if (!g) return -1;
  int ctas, per;
  wgrad_plan(g, &ctas, &per);
  return ctas;
}

extern "C" int snb_conv_c32_wgrad(const float* x, const float* dz, float* partial, const snb_conv_geom* g, void* stream) {
  SNB_REQUIRE(x && dz && partial && g, "snb_conv_c32_wgrad: bad args");
  SNB_REQUIRE(g->transposed == 0, "snb_conv_c32_wgrad: forward geometry expected");
  const int ntaps = g->KD * g->KH * g->KW;
  const long long npos = (long long)g->B * g->OD * g->OH * g->OW;
  int ctas, per;
  wgrad_plan(g, &ctas, &per);
  const int smem = (int)(sizeof(SmemWgrad) + (size_t)ntaps * 4096);
  SNB_REQUIRE(smem <= 227 * 1024, "snb_conv_c32_wgrad: too many taps for shared memory");
  SNB_CUDA(cudaFuncSetAttribute(conv_c32_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  snb_launch(conv_c32_wgrad_kernel, ctas, 256, smem, stream, x, dz, partial, *g, npos, per);
  SNB_LAUNCH_CHECK("conv_c32_wgrad_kernel");
  return 0;
}

extern "C" int snb_softargmin_bwd(const float* cost, const float* pred, const float* dpred, const float* dcost_extra, float* dcost,
                                  int B, int D, int H, int W, void* stream) {
  SNB_REQUIRE(cost && pred && dcost && B > 0 && D > 0 && H > 0 && W > 0, "snb_softargmin_bwd: bad args");
  const long long plane = (long long)H * W, total = plane * B;
  snb_launch(softargmin_bwd_kernel, snb_ceil_div(total, 256), 256, 0, stream, cost, pred, dpred, dcost_extra, dcost, D, plane, total);
  SNB_LAUNCH_CHECK("softargmin_bwd_kernel");
  return 0;
}

extern "C" int snb_relu_bwd(const float* out, const float* dout, float* dres, long long n, void* stream) {
  SNB_REQUIRE(out && dout && dres && n > 0, "snb_relu_bwd: bad args");
  snb_launch(relu_bwd_kernel, snb_ceil_div(n, 256), 256, 0, stream, out, dout, dres, n);
  SNB_LAUNCH_CHECK("relu_bwd_kernel");
  return 0;
}

extern "C" int snb_conv_c32_taps_bwd(const float* x, const float* w, const float* g, float* dx, float* partial,
                                     int B, int D, int H, int W, int ntaps, void* stream) {
  SNB_REQUIRE(x && w && g && dx && partial && B > 0 && D > 0 && H > 0 && W > 0, "snb_conv_c32_taps_bwd: bad args");
  SNB_REQUIRE((ntaps == 27) || (ntaps == 9 && D == 1), "snb_conv_c32_taps_bwd: ntaps must be 27 (3-D) or 9 (2-D, D = 1)");
  const long long npos = (long long)B * D * H * W;
  const int grid = snb_ceil_div(npos, 128);
  if (ntaps == 27) snb_launch(conv_c32_taps_bwd_kernel<27>, grid, 128, 0, stream, x, w, g, dx, partial, npos, D, H, W);
  else             snb_launch(conv_c32_taps_bwd_kernel<9>, grid, 128, 0, stream, x, w, g, dx, partial, npos, D, H, W);
  SNB_LAUNCH_CHECK("conv_c32_taps_bwd_kernel");
  return 0;
}
