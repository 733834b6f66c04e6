// Small-Cin first-layer convolutions (Cin = 3 or 4 -> 32), direct fp32, NCHW image in, channels-last out.
//  * snb_conv5x5s2_c3   : FeatureExtractorNetwork.downsample[0], stereo_net.py:64-70,81 (K = 75, too thin for MMA)
//  * snb_refine_in_conv : EdgeAwareRefinement head, stereo_net.py:105-117: bilinear upsample + scale + concat + 3x3 conv
// CTA = 256 threads = 8 x 64 output pixels; every thread owns 2 pixels x 32 couts (64 accumulators); the input halo tile
// sits in smem (column-parity de-interleaved for stride 2 so reads are conflict free), weights are broadcast LDS.128.
#include "common.cuh"

namespace {

constexpr int TH = 8, TW = 64;

template <int CIN, int KS, int S>
struct TileDims {
  static constexpr int IH = (TH - 1) * S + KS;
  static constexpr int IW = (TW - 1) * S + KS;
  static constexpr int IWP = (S == 2) ? ((IW + 1) / 2 + 1) : IW + 1;       // per-parity row pitch (+1 pad)
  static constexpr int ROW = (S == 2) ? 2 * IWP : IWP;
  static constexpr int IN_FLOATS = (CIN * IH * ROW + 3) & ~3;   // keep the weight block 16-B aligned
  static constexpr int W_FLOATS = CIN * KS * KS * 32;
  static constexpr int MAIN_FLOATS = (IN_FLOATS + W_FLOATS) > 8 * 1024 ? (IN_FLOATS + W_FLOATS) : 8 * 1024;   // >= output staging (8 warps x 4 KB)
  static constexpr int SMEM_BYTES = (MAIN_FLOATS + 64 * 8) * 4;
};

template <int S, int IWP>
__device__ __forceinline__ int col_index(int c) { return (S == 2) ? (c & 1) * IWP + (c >> 1) : c; }

struct ImgLoader {          // plain NCHW image
  const float* img; int C, H, W;
  __device__ __forceinline__ float operator()(int b, int ci, int iy, int ix) const {
    if ((unsigned)iy >= (unsigned)H || (unsigned)ix >= (unsigned)W) return 0.f;
    return img[(((size_t)b * C + ci) * H + iy) * W + ix];
  }
};

struct RefineLoader {       // channel 0 = scale * bilinear(coarse), channels 1..3 = rgb (stereo_net.py:105-117)
  const float* coarse; const float* rgb; int h, w, H, W; float sh, sw, mul;
  __device__ __forceinline__ float operator()(int b, int ci, int iy, int ix) const {
    if ((unsigned)iy >= (unsigned)H || (unsigned)ix >= (unsigned)W) return 0.f;
    if (ci == 0) return bilinear_sample(coarse + (size_t)b * h * w, h, w, iy, ix, sh, sw) * mul;
    return rgb[(((size_t)b * 3 + (ci - 1)) * H + iy) * W + ix];
  }
};

template <int CIN, int KS, int S, class Loader>
__global__ void __launch_bounds__(256)
conv_small_kernel(Loader ld, const float* __restrict__ w, float* __restrict__ y, float* __restrict__ up_out,
                  int OH, int OW, int pad, snb_conv_epilogue e, int phaseB) {
  using T = TileDims<CIN, KS, S>;
  extern __shared__ __align__(16) float smem[];
  float* sIn = smem;
  float* sW = smem + T::IN_FLOATS;             // [k][32], k = (ci*KS + kh)*KS + kw
  float* sRed = smem + T::MAIN_FLOATS;         // [8 warps][64]
  const int t = threadIdx.x;
  const int tiles_x = (OW + TW - 1) / TW, tiles_y = (OH + TH - 1) / TH;
  const int tile = blockIdx.x;
  const int b = tile / (tiles_x * tiles_y);
  const int trem = tile - b * tiles_x * tiles_y;
  const int oy0 = (trem / tiles_x) * TH, ox0 = (trem % tiles_x) * TW;
  const int iy0 = oy0 * S - pad, ix0 = ox0 * S - pad;

  {  // input halo tile: U global loads in flight per thread before the dependent smem stores (a load->store chain per
     // element exposed one full memory latency per iteration: 20 % of all stall samples in the ncu source view)
    constexpr int NIN = CIN * T::IH * T::IW, U = 8;
    for (int i0 = t; i0 < NIN; i0 += 256 * U) {
      float vals[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int i = i0 + u * 256;
        const int c = i % T::IW;
        const int r = (i / T::IW) % T::IH;
        const int ci = i / (T::IW * T::IH);
        vals[u] = i < NIN ? ld(b, ci, iy0 + r, ix0 + c) : 0.f;
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int i = i0 + u * 256;
        const int c = i % T::IW;
        const int r = (i / T::IW) % T::IH;
        const int ci = i / (T::IW * T::IH);
        if (i < NIN) sIn[(ci * T::IH + r) * T::ROW + col_index<S, T::IWP>(c)] = vals[u];
      }
    }
  }
  for (int i = t; i < T::W_FLOATS; i += 256) {   // w is [32][CIN][KS][KS]
    const int co = i / (CIN * KS * KS);
    const int k = i - co * (CIN * KS * KS);
    sW[k * 32 + co] = w[i];
  }
  __syncthreads();

  const int ty = t >> 5, tx = t & 31;
  // 2 pixels x 32 couts per thread as 32 packed accumulators: FFMA2 (fma.rn.f32x2) halves the FMA instruction count
  unsigned long long pa0[16], pa1[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) { pa0[j] = 0ull; pa1[j] = 0ull; }

  for (int ci = 0; ci < CIN; ++ci) {
#pragma unroll
    for (int kh = 0; kh < KS; ++kh) {
      const float* rowp = sIn + (ci * T::IH + ty * S + kh) * T::ROW;
#pragma unroll
      for (int kw = 0; kw < KS; ++kw) {
        const float x0 = rowp[col_index<S, T::IWP>(tx * S + kw)];
        const float x1 = rowp[col_index<S, T::IWP>((tx + 32) * S + kw)];
        const unsigned long long x0p = pack2(x0, x0), x1p = pack2(x1, x1);
        const ulonglong2* wp = reinterpret_cast<const ulonglong2*>(sW + ((ci * KS + kh) * KS + kw) * 32);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const ulonglong2 wv = wp[j];                     // (w[4j], w[4j+1]) | (w[4j+2], w[4j+3])
          ffma2(pa0[2 * j], x0p, wv.x); ffma2(pa1[2 * j], x1p, wv.x);
          ffma2(pa0[2 * j + 1], x0p, wv.y); ffma2(pa1[2 * j + 1], x1p, wv.y);
        }
      }
    }
  }
  float acc0[32], acc1[32];
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    acc0[2 * j] = lo2(pa0[j]); acc0[2 * j + 1] = hi2(pa0[j]);
    acc1[2 * j] = lo2(pa1[j]); acc1[2 * j + 1] = hi2(pa1[j]);
  }

  const int oy = oy0 + ty;
  const int oxa = ox0 + tx, oxb = ox0 + tx + 32;
  const bool oka = oy < OH && oxa < OW, okb = oy < OH && oxb < OW;

  if (up_out != nullptr) {   // side output: the upsampled+scaled disparity plane (centre tap of input channel 0)
    const float* rowp = sIn + (0 * T::IH + ty * S + pad) * T::ROW;
    if (oka) up_out[((size_t)b * OH + oy) * OW + oxa] = rowp[col_index<S, T::IWP>(tx * S + pad)];
    if (okb) up_out[((size_t)b * OH + oy) * OW + oxb] = rowp[col_index<S, T::IWP>((tx + 32) * S + pad)];
  }

  // z = conv + bias ; per-channel partial stats over valid pixels
  float s1 = 0.f, s2 = 0.f;   // lane j of a warp ends up owning channel j
#pragma unroll
  for (int j = 0; j < 32; ++j) {
    const float bj = e.bias ? e.bias[j] : 0.f;
    acc0[j] += bj; acc1[j] += bj;
    if (e.stats) {
      float a = (oka ? acc0[j] : 0.f) + (okb ? acc1[j] : 0.f);
      float q = (oka ? acc0[j] * acc0[j] : 0.f) + (okb ? acc1[j] * acc1[j] : 0.f);
      a = warp_sum(a); q = warp_sum(q);
      if (tx == j) { s1 = a; s2 = q; }
    }
  }
  if (e.stats) {
    sRed[ty * 64 + tx] = s1; sRed[ty * 64 + 32 + tx] = s2;
    __syncthreads();
    if (t < 64) {
      float a = 0.f;
#pragma unroll
      for (int wv = 0; wv < 8; ++wv) a += sRed[wv * 64 + t];
      e.stats[(size_t)blockIdx.x * 64 + t] = a;
    }
  }
#pragma unroll
  for (int j = 0; j < 32; ++j) {
    if (e.scale) { const float sc = e.scale[j], sh = e.shift[j]; acc0[j] = fmaf(acc0[j], sc, sh); acc1[j] = fmaf(acc1[j], sc, sh); }
    if (e.lrelu) { acc0[j] = lrelu(acc0[j]); acc1[j] = lrelu(acc1[j]); }
  }
  // Coalesced output: a thread holds whole pixels (32 channels = 128 B); writing them directly would touch 32 different
  // lines per STG.  Each warp transposes its 32-pixel run through a private 4 KB smem tile (16-B chunks XOR-swizzled by
  // pixel) and stores 4 KB of consecutive channels-last memory with fully coalesced STG.128.
  __syncthreads();                             // every warp is done with the input tile / weights: reuse as staging
  float* stg = smem + ty * 1024;
#pragma unroll
  for (int half = 0; half < 2; ++half) {
    const float* acc = half ? acc1 : acc0;
#pragma unroll
    for (int j = 0; j < 8; ++j)
      *reinterpret_cast<float4*>(stg + tx * 32 + ((j ^ (tx & 7)) << 2)) = make_float4(acc[4 * j], acc[4 * j + 1], acc[4 * j + 2], acc[4 * j + 3]);
    __syncwarp();
    const int oxs = ox0 + half * 32;
    float* yrow = y + (((size_t)b * OH + oy) * OW + oxs) * 32;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int idx = i * 32 + tx, pp = idx >> 3, c = idx & 7;
      const float4 v = *reinterpret_cast<const float4*>(stg + pp * 32 + ((c ^ (pp & 7)) << 2));
      if (oy < OH && oxs + pp < OW) {
        if (phaseB > 0) {      // polyphase output [4][B][OH/2][OW/2][32] for the stride-2 layer that follows (csrc/phase.cu)
          const int ox = oxs + pp, ph = (oy & 1) * 2 + (ox & 1);
          *reinterpret_cast<float4*>(y + ((((size_t)ph * phaseB + b) * (OH >> 1) + (oy >> 1)) * (OW >> 1) + (ox >> 1)) * 32 + c * 4) = v;
        } else {
          *reinterpret_cast<float4*>(yrow + (size_t)pp * 32 + c * 4) = v;
        }
      }
    }
    __syncwarp();
  }
}

// Weight gradient of the small-Cin convolutions: dW[k][co] = sum_p dy[p][co] * in[k @ p], k = (ci,kh,kw).
// Same tiling / input-tile loader as the forward kernel; the dy tile (512 positions x 32) is staged with cp.async.
// 256 threads = 32 couts x 8 groups, group g owns k = g, g+8, ...  Writes partial[tile][K][32].
template <int CIN, int KS, int S, class Loader>
__global__ void __launch_bounds__(256)
conv_small_wgrad_kernel(Loader ld, const float* __restrict__ dy, float* __restrict__ partial, int OH, int OW, int pad) {
  using T = TileDims<CIN, KS, S>;
  constexpr int K = CIN * KS * KS;
  constexpr int JMAX = (K + 7) / 8;
  extern __shared__ __align__(16) float smem[];
  float* sIn = smem;
  float* sDy = smem + T::IN_FLOATS;            // [512][32]
  const int t = threadIdx.x;
  const int tiles_x = (OW + TW - 1) / TW, tiles_y = (OH + TH - 1) / TH;
  const int tile = blockIdx.x;
  const int b = tile / (tiles_x * tiles_y);
  const int trem = tile - b * tiles_x * tiles_y;
  const int oy0 = (trem / tiles_x) * TH, ox0 = (trem % tiles_x) * TW;
  const int iy0 = oy0 * S - pad, ix0 = ox0 * S - pad;
  {
    const int chunk = t & 7;
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      const int p = (t >> 3) + 32 * j;          // 0..511
      const int oy = oy0 + (p >> 6), ox = ox0 + (p & 63);
      const bool ok = oy < OH && ox < OW;
      cp_async16(sDy + p * 32 + chunk * 4, dy + (ok ? (((size_t)b * OH + oy) * OW + ox) * 32 + chunk * 4 : 0), ok);
    }
    cp_async_commit();
  }
  {  // input halo tile, U loads in flight per thread (see conv_small_kernel)
    constexpr int NIN = CIN * T::IH * T::IW, U = 8;
    for (int i0 = t; i0 < NIN; i0 += 256 * U) {
      float vals[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int i = i0 + u * 256;
        const int c = i % T::IW;
        const int r = (i / T::IW) % T::IH;
        const int ci = i / (T::IW * T::IH);
        vals[u] = i < NIN ? ld(b, ci, iy0 + r, ix0 + c) : 0.f;
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int i = i0 + u * 256;
        const int c = i % T::IW;
        const int r = (i / T::IW) % T::IH;
        const int ci = i / (T::IW * T::IH);
        if (i < NIN) sIn[(ci * T::IH + r) * T::ROW + col_index<S, T::IWP>(c)] = vals[u];
      }
    }
  }
  cp_async_wait<0>();
  __syncthreads();
  const int co = t & 31, grp = t >> 5;
  int basej[JMAX];
  float acc[JMAX];
#pragma unroll
  for (int j = 0; j < JMAX; ++j) {
    int k = grp + 8 * j;
    if (k >= K) k = K - 1;                      // clamped duplicates are discarded at the end
    const int kw = k % KS, kh = (k / KS) % KS, ci = k / (KS * KS);
    basej[j] = (ci * T::IH + kh) * T::ROW + ((S == 2) ? (kw & 1) * T::IWP + (kw >> 1) : kw);
    acc[j] = 0.f;
  }
  for (int p = 0; p < TH * TW; ++p) {
    const float g = sDy[p * 32 + co];
    const int o = (p >> 6) * S * T::ROW + (p & 63);
#pragma unroll
    for (int j = 0; j < JMAX; ++j) acc[j] = fmaf(g, sIn[basej[j] + o], acc[j]);
  }
  float* out = partial + (size_t)blockIdx.x * K * 32;
#pragma unroll
  for (int j = 0; j < JMAX; ++j) {
    const int k = grp + 8 * j;
    if (k < K) out[k * 32 + co] = acc[j];
  }
}

}  // namespace

extern "C" int snb_conv5x5s2_c3(const float* img, const float* w, const float* bias, float* y, int B, int H, int W, void* stream) {
  SNB_REQUIRE(img && w && y && B > 0 && H > 0 && W > 0, "snb_conv5x5s2_c3: bad args");
  const int OH = (H - 1) / 2 + 1, OW = (W - 1) / 2 + 1;      // floor((n + 2*2 - 5)/2) + 1
  using T = TileDims<3, 5, 2>;
  ImgLoader ld{img, 3, H, W};
  snb_conv_epilogue e{bias, nullptr, nullptr, nullptr, nullptr, 0};
  auto kern = conv_small_kernel<3, 5, 2, ImgLoader>;
  SNB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, T::SMEM_BYTES));
  const int tiles = B * ((OH + TH - 1) / TH) * ((OW + TW - 1) / TW);
  kern<<<tiles, 256, T::SMEM_BYTES, (cudaStream_t)stream>>>(ld, w, y, nullptr, OH, OW, 2, e, 0);
  SNB_LAUNCH_CHECK("conv5x5s2_c3");
  return 0;
}

extern "C" int snb_conv5x5s2_c3_phases(const float* img, const float* w, const float* bias, float* yph, int B, int H, int W, void* stream) {
  SNB_REQUIRE(img && w && yph && B > 0 && H > 0 && W > 0, "snb_conv5x5s2_c3_phases: bad args");
  const int OH = (H - 1) / 2 + 1, OW = (W - 1) / 2 + 1;
  SNB_REQUIRE((OH % 2) == 0 && (OW % 2) == 0, "snb_conv5x5s2_c3_phases: output %dx%d must be even (use snb_conv5x5s2_c3 + snb_phase_split)", OH, OW);
  using T = TileDims<3, 5, 2>;
  ImgLoader ld{img, 3, H, W};
  snb_conv_epilogue e{bias, nullptr, nullptr, nullptr, nullptr, 0};
  auto kern = conv_small_kernel<3, 5, 2, ImgLoader>;
  SNB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, T::SMEM_BYTES));
  const int tiles = B * ((OH + TH - 1) / TH) * ((OW + TW - 1) / TW);
  kern<<<tiles, 256, T::SMEM_BYTES, (cudaStream_t)stream>>>(ld, w, yph, nullptr, OH, OW, 2, e, B);
  SNB_LAUNCH_CHECK("conv5x5s2_c3_phases");
  return 0;
}

extern "C" int snb_refine_in_conv_num_tiles(int B, int H, int W) {
  return B * ((H + TH - 1) / TH) * ((W + TW - 1) / TW);
}

extern "C" int snb_refine_in_conv(const float* coarse, const float* rgb, const float* w, float* up, float* z,
                                  int B, int h, int w_, int H, int W, float disp_scale, const snb_conv_epilogue* e, void* stream) {
  SNB_REQUIRE(coarse && rgb && w && up && z && e && B > 0 && h > 0 && w_ > 0 && H > 0 && W > 0, "snb_refine_in_conv: bad args");
  using T = TileDims<4, 3, 1>;
  RefineLoader ld{coarse, rgb, h, w_, H, W, (float)h / (float)H, (float)w_ / (float)W, disp_scale};
  auto kern = conv_small_kernel<4, 3, 1, RefineLoader>;
  SNB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, T::SMEM_BYTES));
  const int tiles = snb_refine_in_conv_num_tiles(B, H, W);
  kern<<<tiles, 256, T::SMEM_BYTES, (cudaStream_t)stream>>>(ld, w, z, up, H, W, 1, *e, 0);
  SNB_LAUNCH_CHECK("refine_in_conv");
  return 0;
}

extern "C" int snb_conv5x5s2_c3_num_tiles(int B, int H, int W) {
  const int OH = (H - 1) / 2 + 1, OW = (W - 1) / 2 + 1;
  return B * ((OH + TH - 1) / TH) * ((OW + TW - 1) / TW);
}

extern "C" int snb_conv5x5s2_c3_wgrad(const float* img, const float* dy, float* partial, int B, int H, int W, void* stream) {
  SNB_REQUIRE(img && dy && partial && B > 0 && H > 0 && W > 0, "snb_conv5x5s2_c3_wgrad: bad args");
  const int OH = (H - 1) / 2 + 1, OW = (W - 1) / 2 + 1;
  using T = TileDims<3, 5, 2>;
  ImgLoader ld{img, 3, H, W};
  auto kern = conv_small_wgrad_kernel<3, 5, 2, ImgLoader>;
  const int smem = (T::IN_FLOATS + 512 * 32) * 4;
  SNB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  kern<<<snb_conv5x5s2_c3_num_tiles(B, H, W), 256, smem, (cudaStream_t)stream>>>(ld, dy, partial, OH, OW, 2);
  SNB_LAUNCH_CHECK("conv5x5s2_c3_wgrad");
  return 0;
}

extern "C" int snb_refine_in_wgrad(const float* coarse, const float* rgb, const float* dz, float* partial,
                                   int B, int h, int w_, int H, int W, float disp_scale, void* stream) {
  SNB_REQUIRE(coarse && rgb && dz && partial && B > 0 && h > 0 && w_ > 0 && H > 0 && W > 0, "snb_refine_in_wgrad: bad args");
  using T = TileDims<4, 3, 1>;
  RefineLoader ld{coarse, rgb, h, w_, H, W, (float)h / (float)H, (float)w_ / (float)W, disp_scale};
  auto kern = conv_small_wgrad_kernel<4, 3, 1, RefineLoader>;
  const int smem = (T::IN_FLOATS + 512 * 32) * 4;
  SNB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  kern<<<snb_refine_in_conv_num_tiles(B, H, W), 256, smem, (cudaStream_t)stream>>>(ld, dz, partial, H, W, 1);
  SNB_LAUNCH_CHECK("refine_in_wgrad");
  return 0;
}
