// Small-Cin first-layer convolutions (Cin = 3 or 4 -> 32), direct fp32, NCHW image in, channels-last out.
//  * snb_conv5x5s2_c3   : FeatureExtractorNetwork.downsample[0], stereo_net.py:64-70,81 (K = 75, too thin for MMA)
//  * snb_refine_in_conv : EdgeAwareRefinement head, stereo_net.py:105-117: bilinear upsample + scale + concat + 3x3 conv
// CTA = 256 threads = 8 x 64 output pixels; every thread owns 2 pixels x 32 couts (64 accumulators); the input halo tile
// sits in smem (column-parity de-interleaved for stride 2 so reads are conflict free), weights are broadcast LDS.128.
// Default path: conv_small_tc_kernel — the same tiles as an im2col GEMM on the tcgen05 tensor cores (M = 128 pixels,
// N = 32 couts, K = Cin*KS*KS padded to a multiple of 8, 3xTF32 split).  A thread builds the im2col row of its pixel in
// registers from the smem halo tile, splits hi/lo and tcgen05.st's it into TMEM (A operand from TMEM, so an MMA only
// fetches its 1 KB weight slice from smem); the weight image (hi|lo, SWIZZLE_128B K-major) is built in smem per CTA.
// The FFMA kernel below stays selectable (SNB200_SMALL_CONV=ffma) as the independent cross-check.
#include <stdlib.h>
#include <string.h>
#include "tc_common.cuh"

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
  static constexpr int NCOMP = 0;        // number of leading channels that are computed rather than copied
  const float* img; int C, H, W;
  __device__ __forceinline__ const float* base() const { return img; }
  __device__ __forceinline__ const float* rowptr(int b, int ci, int iy) const { return img + (((size_t)b * C + ci) * H + iy) * W; }
  __device__ __forceinline__ float operator()(int b, int ci, int iy, int ix) const {
    if ((unsigned)iy >= (unsigned)H || (unsigned)ix >= (unsigned)W) return 0.f;
    return img[(((size_t)b * C + ci) * H + iy) * W + ix];
  }
};

struct RefineLoader {       // channel 0 = scale * bilinear(coarse), channels 1..3 = rgb (stereo_net.py:105-117)
  static constexpr int NCOMP = 1;
  const float* coarse; const float* rgb; int h, w, H, W; float sh, sw, mul;
  __device__ __forceinline__ const float* base() const { return rgb; }
  __device__ __forceinline__ const float* rowptr(int b, int ci, int iy) const { return rgb + (((size_t)b * 3 + (ci - 1)) * H + iy) * W; }
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
  pdl_launch(); pdl_wait();
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

// ------------------------------------------------------------------------------------------------ tensor-core variant
// registers -> TMEM, N consecutive 32-bit columns of the thread's lane
template <int N> __device__ __forceinline__ void tmem_st_n(uint32_t taddr, const uint32_t* r);
template <> __device__ __forceinline__ void tmem_st_n<4>(uint32_t taddr, const uint32_t* r) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};\n"
               :: "r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]) : "memory");
}
template <> __device__ __forceinline__ void tmem_st_n<8>(uint32_t taddr, const uint32_t* r) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};\n"
               :: "r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]) : "memory");
}
template <> __device__ __forceinline__ void tmem_st_n<16>(uint32_t taddr, const uint32_t* r) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};\n"
               :: "r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
                  "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]) : "memory");
}
template <> __device__ __forceinline__ void tmem_st_n<32>(uint32_t taddr, const uint32_t* r) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};\n"
      :: "r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
         "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]),
         "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]),
         "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
}

template <int CIN, int KS, int S>
struct TcDims {
  using T = TileDims<CIN, KS, S>;
  static constexpr int K = CIN * KS * KS;
  static constexpr int KP = (K + 7) / 8 * 8;          // MMA K extent (80 / 40)
  static constexpr int KH = KP / 2;                   // im2col columns built by one warp half (40 / 20)
  static constexpr int KH_A = (KH >= 32) ? 32 : 16;   // tcgen05.st chunking: KH = KH_A + KH_B
  static constexpr int KH_B = KH - KH_A;              // 8 / 4
  static constexpr int KPB = (KP + 31) / 32 * 32;     // smem weight image K extent (whole 32-float swizzle atoms)
  static constexpr int B_FLOATS = 2 * KPB * 32;       // hi | lo
  static constexpr int STAGE_FLOATS = 2 * 128 * 32;   // double-buffered output staging
  static constexpr int OFF_STAGE = B_FLOATS;
  static constexpr int OFF_IN = OFF_STAGE + STAGE_FLOATS;
  static constexpr int OFF_RED = OFF_IN + T::IN_FLOATS;
  static constexpr int OFF_BAR = OFF_RED + 8 * 64;
  static constexpr int USED_BYTES = (OFF_BAR + 16) * 4 + 1024 /*alignment slack*/;
  // two CTAs per SM (2 x 256 TMEM columns): never let a third one become resident and spin in tcgen05.alloc
  static constexpr int SMEM_BYTES = USED_BYTES > 78 * 1024 ? USED_BYTES : 78 * 1024;
  static constexpr int ACC_COL = 0, A_COL = 64;       // TMEM: 2 accumulators x 32 | A_hi[KP] | A_lo[KP]
  static_assert(A_COL + 2 * KP <= 256, "TMEM budget");
  __host__ __device__ static constexpr int in_off(int k) {
    return ((k / (KS * KS)) * T::IH + (k / KS) % KS) * T::ROW + ((S == 2) ? ((k % KS) & 1) * T::IWP + ((k % KS) >> 1) : (k % KS));
  }
};

// im2col columns [K0, K0 + N) of the thread's pixel: hi -> TMEM column ta + K0, lo -> ta + KP + K0 (N = 16, 8 or 4: short
// chunks keep the live register set small)
template <int CIN, int KS, int S, int K0, int N>
__device__ __forceinline__ void build_a_chunk(const float* __restrict__ px, uint32_t ta) {
  using D = TcDims<CIN, KS, S>;
  uint32_t h[N], l[N];
#pragma unroll
  for (int j = 0; j < N; ++j) {
    const int k = K0 + j;
    const float x = (k < D::K) ? px[D::in_off(k < D::K ? k : 0)] : 0.f;
    h[j] = tc::tf32_hi_bits(x);
    l[j] = __float_as_uint(x - __uint_as_float(h[j]));
  }
  tmem_st_n<N>(ta + K0, h);
  tmem_st_n<N>(ta + D::KP + K0, l);
}

template <int CIN, int KS, int S, int HALF>
__device__ __forceinline__ void build_a_half(const float* __restrict__ px, uint32_t ta) {
  using D = TcDims<CIN, KS, S>;
  constexpr int K0 = HALF * D::KH;
  static_assert(D::KH == 40 || D::KH == 20, "chunking below assumes KH = 40 or 20");
  build_a_chunk<CIN, KS, S, K0, 16>(px, ta);
  if constexpr (D::KH == 40) {
    build_a_chunk<CIN, KS, S, K0 + 16, 16>(px, ta);
    build_a_chunk<CIN, KS, S, K0 + 32, 8>(px, ta);
  } else {
    build_a_chunk<CIN, KS, S, K0 + 16, 4>(px, ta);
  }
}

// 4-byte cp.async with zero fill when !valid
__device__ __forceinline__ void cp_async4(void* smem_dst, const void* gsrc, bool valid) {
  const int sz = valid ? 4 : 0;
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;\n" :: "r"(smem_u32(smem_dst)), "l"(gsrc), "r"(sz) : "memory");
}

// Warps 0-7: im2col builders + epilogue (warp % 4 = TMEM lane quadrant, warp / 4 = K half / channel half); warp 8: MMA issuer.
constexpr int TC_THREADS = 9 * 32;

template <int CIN, int KS, int S, class Loader>
__global__ void __launch_bounds__(TC_THREADS, 2)
conv_small_tc_kernel(Loader ld, const float* __restrict__ w, float* __restrict__ y, float* __restrict__ up_out,
                     int OH, int OW, int pad, snb_conv_epilogue e, int phaseB) {
  using T = TileDims<CIN, KS, S>;
  using D = TcDims<CIN, KS, S>;
  constexpr int NMT = TH * TW / 128;                 // 4 M-tiles of 2 rows x 64 pixels
  extern __shared__ unsigned char smem_dyn[];
  const uint32_t base_u32 = (smem_u32(smem_dyn) + 1023u) & ~1023u;
  float* smem = reinterpret_cast<float*>(smem_dyn + (base_u32 - smem_u32(smem_dyn)));
  float* sB = smem;                                  // [hi|lo][atom][32 couts][32 k] swizzled
  float* sStage = smem + D::OFF_STAGE;
  float* sIn = smem + D::OFF_IN;
  float* sRed = smem + D::OFF_RED;
  uint64_t* mma_done = reinterpret_cast<uint64_t*>(smem + D::OFF_BAR);   // MMA commit -> builders / epilogue
  uint64_t* a_full = mma_done + 1;                                       // 8 builder warps -> MMA warp
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(a_full + 1);

  const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
  const int tiles_x = (OW + TW - 1) / TW, tiles_y = (OH + TH - 1) / TH;
  const int tile = blockIdx.x;
  const int b = tile / (tiles_x * tiles_y);
  const int trem = tile - b * tiles_x * tiles_y;
  const int oy0 = (trem / tiles_x) * TH, ox0 = (trem % tiles_x) * TW;
  const int iy0 = oy0 * S - pad, ix0 = ox0 * S - pad;

  if (warp == 8) {
    if (lane == 0) { mbar_init(mma_done, 1); mbar_init(a_full, 8); mbar_fence_init(); }
    __syncwarp();
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;\n" :: "r"(smem_u32(tmem_slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
  }
  pdl_launch();                                    // dependents only once this CTA owns its TMEM columns (no alloc dead-lock with an early dependent)
  pdl_wait();                                        // nothing above touched global memory
  {  // input halo tile, one row per warp at a time (no per-element index arithmetic: this kernel is instruction-issue bound).
     // Plain image channels go global -> smem with 4-byte cp.async (everything in flight at once, zero fill outside the
     // image); computed channels (the bilinear disparity plane of the refinement) go through registers.
    constexpr int NROWS = CIN * T::IH, NJ = (T::IW + 31) / 32;
    for (int row = warp; row < NROWS; row += TC_THREADS / 32) {
      const int ci = row / T::IH, r = row - ci * T::IH;
      const int iy = iy0 + r;
      float* drow = sIn + row * T::ROW;
      if (ci >= Loader::NCOMP) {
        const bool rowok = (unsigned)iy < (unsigned)ld.H;
        const float* srow = ld.rowptr(b, ci, rowok ? iy : 0);
#pragma unroll
        for (int j = 0; j < NJ; ++j) {
          const int c = lane + 32 * j;
          if (c < T::IW) {
            const int ix = ix0 + c;
            const bool ok = rowok && (unsigned)ix < (unsigned)ld.W;
            cp_async4(drow + col_index<S, T::IWP>(c), ok ? srow + ix : ld.base(), ok);
          }
        }
      } else {
#pragma unroll
        for (int j = 0; j < NJ; ++j) {
          const int c = lane + 32 * j;
          if (c < T::IW) drow[col_index<S, T::IWP>(c)] = ld(b, ci, iy, ix0 + c);
        }
      }
    }
    cp_async_commit();
  }
  {  // weight image: w is [32][K] (PyTorch [co][ci][kh][kw]) = K-major rows of the B operand; zero padded to KPB
    constexpr int NJ = D::KPB / 32;
    for (int co = warp; co < 32; co += TC_THREADS / 32) {
      float v[NJ];
#pragma unroll
      for (int j = 0; j < NJ; ++j) { const int k = lane + 32 * j; v[j] = k < D::K ? __ldg(w + co * D::K + k) : 0.f; }
#pragma unroll
      for (int j = 0; j < NJ; ++j) {
        const float hi = __uint_as_float(tc::tf32_hi_bits(v[j]));
        const int o = j * 1024 + co * 32 + (((lane >> 2) ^ (co & 7)) << 2) + (lane & 3);
        sB[o] = hi;
        sB[D::KPB * 32 + o] = v[j] - hi;
      }
    }
  }
  cp_async_wait<0>();
  tc::fence_async_smem();                            // generic-proxy smem writes -> visible to the MMA's async-proxy reads
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 8) {
    // =============================================================== MMA issuer (converged warp, one election per M-tile)
    const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0), smem_u = __shfl_sync(0xffffffffu, base_u32, 0);
    for (int mt = 0; mt < NMT; ++mt) {
      tc::mbar_wait_warp(a_full, mt & 1);
      tc::tc_fence_after();
      if (tc::elect_one()) {
        const uint32_t tmem_d = tmem_u + D::ACC_COL + (mt & 1) * 32, ta = tmem_u + D::A_COL;
#pragma unroll
        for (int ks = 0; ks < D::KP / 8; ++ks) {
          const uint32_t bo = smem_u + (ks >> 2) * 4096 + (ks & 3) * 32;
          const uint64_t bh = tc::make_desc(bo), bl = tc::make_desc(bo + D::KPB * 128);
          tc::mma_n32_ts(tmem_d, ta + ks * 8, bh, ks > 0);
          tc::mma_n32_ts(tmem_d, ta + D::KP + ks * 8, bh, 1);
          tc::mma_n32_ts(tmem_d, ta + ks * 8, bl, 1);
        }
        tc::mma_commit_raw(mma_done);
      }
      __syncwarp();
    }
  } else {
    // =============================================================== im2col builders + epilogue (256 threads)
    const int quad = warp & 3, half = warp >> 2;
    const int m = quad * 32 + lane;                    // pixel row of this thread inside an M-tile (TMEM lane)
    const uint32_t tlane = tmem_base + ((uint32_t)(quad * 32) << 16);
    const int chunk = t & 7, rg = t >> 3;              // store phase: rows rg + 32 j, 16-B chunk `chunk`
    float4 bias4 = make_float4(0.f, 0.f, 0.f, 0.f), sc4 = make_float4(1.f, 1.f, 1.f, 1.f), sh4 = bias4;
    if (e.bias) bias4 = reinterpret_cast<const float4*>(e.bias)[chunk];
    if (e.scale) { sc4 = reinterpret_cast<const float4*>(e.scale)[chunk]; sh4 = reinterpret_cast<const float4*>(e.shift)[chunk]; }
    const bool has_stats = e.stats != nullptr;
    float s1[4] = {0.f, 0.f, 0.f, 0.f}, s2[4] = {0.f, 0.f, 0.f, 0.f};
    // output addressing of the store phase: row r = rg + 32 j of M-tile mt is pixel (oy0 + 2 mt + (r >> 6), ox0 + (r & 63))
    size_t goff[4]; bool okx[4];
    const size_t mt_stride = phaseB > 0 ? (size_t)(OW >> 1) * 32 : (size_t)2 * OW * 32;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int r = rg + 32 * j, rr = r >> 6, ox = ox0 + (r & 63), oy = oy0 + rr;
      okx[j] = ox < OW;
      if (phaseB > 0) {            // polyphase output [4][B][OH/2][OW/2][32] for the stride-2 layer that follows (csrc/phase.cu)
        const int ph = (oy & 1) * 2 + (ox & 1);
        goff[j] = ((((size_t)ph * phaseB + b) * (OH >> 1) + (oy >> 1)) * (OW >> 1) + (ox >> 1)) * 32 + chunk * 4;
      } else {
        goff[j] = (((size_t)b * OH + oy) * OW + ox) * 32 + chunk * 4;
      }
    }

    auto epilogue = [&](int mt) {
      float v[16];
      tc::tmem_ld16(tlane + D::ACC_COL + (mt & 1) * 32 + half * 16, v);
      float* stg = sStage + (mt & 1) * 128 * 32;
      {
        float* row = stg + m * 32;
#pragma unroll
        for (int c = 0; c < 4; ++c)
          *reinterpret_cast<float4*>(row + (((half * 4 + c) ^ (m & 7)) << 2)) = make_float4(v[4 * c], v[4 * c + 1], v[4 * c + 2], v[4 * c + 3]);
      }
      tc::epi_bar();
      float* ymt = y + mt * mt_stride;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int r = rg + 32 * j;
        const bool ok = okx[j] && (oy0 + 2 * mt + (r >> 6)) < OH;
        float4 o = *reinterpret_cast<const float4*>(stg + r * 32 + ((chunk ^ (r & 7)) << 2));
        o.x += bias4.x; o.y += bias4.y; o.z += bias4.z; o.w += bias4.w;
        if (has_stats && ok) {
          s1[0] += o.x; s1[1] += o.y; s1[2] += o.z; s1[3] += o.w;
          s2[0] = fmaf(o.x, o.x, s2[0]); s2[1] = fmaf(o.y, o.y, s2[1]); s2[2] = fmaf(o.z, o.z, s2[2]); s2[3] = fmaf(o.w, o.w, s2[3]);
        }
        if (e.scale) { o.x = fmaf(o.x, sc4.x, sh4.x); o.y = fmaf(o.y, sc4.y, sh4.y); o.z = fmaf(o.z, sc4.z, sh4.z); o.w = fmaf(o.w, sc4.w, sh4.w); }
        if (e.lrelu) { o.x = lrelu(o.x); o.y = lrelu(o.y); o.z = lrelu(o.z); o.w = lrelu(o.w); }
        if (ok) __stcg(reinterpret_cast<float4*>(ymt + goff[j]), o);
      }
    };

    for (int mt = 0; mt < NMT; ++mt) {
      if (mt > 0) { tc::mbar_wait(mma_done, (mt - 1) & 1); tc::tc_fence_after(); }   // MMAs of mt-1 done: A slot free, accumulator ready
      {
        const int ty = 2 * mt + (m >> 6), tx = m & 63;
        const float* px = sIn + ty * S * T::ROW + tx;
        if (half == 0) {
          build_a_half<CIN, KS, S, 0>(px, tlane + D::A_COL);
          if (up_out != nullptr) {   // side output: the upsampled+scaled disparity plane (centre tap of input channel 0)
            const int oy = oy0 + ty, ox = ox0 + tx;
            if (oy < OH && ox < OW) up_out[((size_t)b * OH + oy) * OW + ox] = px[D::in_off(pad * KS + pad)];
          }
        } else {
          build_a_half<CIN, KS, S, 1>(px, tlane + D::A_COL);
        }
      }
      tc::tmem_wait_st();
      tc::tc_fence_before();                           // orders this warp's tcgen05.st (and the tcgen05.ld of epilogue mt-2) before the arrive
      __syncwarp();
      if (lane == 0) mbar_arrive(a_full);
      if (mt > 0) epilogue(mt - 1);                    // overlaps the MMAs of mt
    }
    tc::mbar_wait(mma_done, (NMT - 1) & 1);
    tc::tc_fence_after();
    epilogue(NMT - 1);

    if (has_stats) {
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        s1[c] += __shfl_xor_sync(0xffffffffu, s1[c], 8); s1[c] += __shfl_xor_sync(0xffffffffu, s1[c], 16);
        s2[c] += __shfl_xor_sync(0xffffffffu, s2[c], 8); s2[c] += __shfl_xor_sync(0xffffffffu, s2[c], 16);
      }
      if (lane < 8) {
#pragma unroll
        for (int c = 0; c < 4; ++c) { sRed[warp * 64 + lane * 4 + c] = s1[c]; sRed[warp * 64 + 32 + lane * 4 + c] = s2[c]; }
      }
      tc::epi_bar();
      if (t < 64) {
        float a = 0.f;
#pragma unroll
        for (int wv = 0; wv < 8; ++wv) a += sRed[wv * 64 + t];
        e.stats[(size_t)blockIdx.x * 64 + t] = a;
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

// Weight gradient of the small-Cin convolutions: dW[k][co] = sum_p dy[p][co] * in[k @ p], k = (ci,kh,kw).
// Same tiling / input-tile loader as the forward kernel; the dy tile (512 positions x 32) is staged with cp.async.
// 256 threads = 32 couts x 8 groups, group g owns k = g, g+8, ...  Writes partial[tile][K][32].
template <int CIN, int KS, int S, class Loader>
__global__ void __launch_bounds__(256)
conv_small_wgrad_kernel(Loader ld, const float* __restrict__ dy, float* __restrict__ partial, int OH, int OW, int pad) {
  pdl_launch(); pdl_wait();
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

static bool snb_small_conv_ffma() {
  static const bool v = []() { const char* s = getenv("SNB200_SMALL_CONV"); return s != nullptr && strcmp(s, "ffma") == 0; }();
  return v;
}

template <int CIN, int KS, int S, class Loader>
static int snb_launch_small(const char* name, Loader ld, const float* w, float* y, float* up, int tiles, int OH, int OW, int pad,
                            const snb_conv_epilogue& e, int phaseB, void* stream) {
  if (snb_small_conv_ffma()) {
    using T = TileDims<CIN, KS, S>;
    auto kern = conv_small_kernel<CIN, KS, S, Loader>;
    SNB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, T::SMEM_BYTES));
    snb_launch(kern, tiles, 256, T::SMEM_BYTES, stream, ld, w, y, up, OH, OW, pad, e, phaseB);
  } else {
    using D = TcDims<CIN, KS, S>;
    SNB_REQUIRE((reinterpret_cast<uintptr_t>(y) & 15) == 0, "%s: output must be 16-byte aligned", name);
    SNB_REQUIRE(!e.bias || (reinterpret_cast<uintptr_t>(e.bias) & 15) == 0, "%s: bias must be 16-byte aligned", name);
    SNB_REQUIRE(!e.scale || ((reinterpret_cast<uintptr_t>(e.scale) | reinterpret_cast<uintptr_t>(e.shift)) & 15) == 0, "%s: scale/shift must be 16-byte aligned", name);
    auto kern = conv_small_tc_kernel<CIN, KS, S, Loader>;
    SNB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, D::SMEM_BYTES));
    snb_launch(kern, tiles, TC_THREADS, D::SMEM_BYTES, stream, ld, w, y, up, OH, OW, pad, e, phaseB);
  }
  SNB_LAUNCH_CHECK(name);
  return 0;
}

extern "C" int snb_conv5x5s2_c3(const float* img, const float* w, const float* bias, float* y, int B, int H, int W, void* stream) {
  SNB_REQUIRE(img && w && y && B > 0 && H > 0 && W > 0, "snb_conv5x5s2_c3: bad args");
  const int OH = (H - 1) / 2 + 1, OW = (W - 1) / 2 + 1;      // floor((n + 2*2 - 5)/2) + 1
  ImgLoader ld{img, 3, H, W};
  snb_conv_epilogue e{bias, nullptr, nullptr, nullptr, nullptr, 0};
  const int tiles = B * ((OH + TH - 1) / TH) * ((OW + TW - 1) / TW);
  return snb_launch_small<3, 5, 2>("conv5x5s2_c3", ld, w, y, nullptr, tiles, OH, OW, 2, e, 0, stream);
}

extern "C" int snb_conv5x5s2_c3_phases(const float* img, const float* w, const float* bias, float* yph, int B, int H, int W, void* stream) {
  SNB_REQUIRE(img && w && yph && B > 0 && H > 0 && W > 0, "snb_conv5x5s2_c3_phases: bad args");
  const int OH = (H - 1) / 2 + 1, OW = (W - 1) / 2 + 1;
  SNB_REQUIRE((OH % 2) == 0 && (OW % 2) == 0, "snb_conv5x5s2_c3_phases: output %dx%d must be even (use snb_conv5x5s2_c3 + snb_phase_split)", OH, OW);
  ImgLoader ld{img, 3, H, W};
  snb_conv_epilogue e{bias, nullptr, nullptr, nullptr, nullptr, 0};
  const int tiles = B * ((OH + TH - 1) / TH) * ((OW + TW - 1) / TW);
  return snb_launch_small<3, 5, 2>("conv5x5s2_c3_phases", ld, w, yph, nullptr, tiles, OH, OW, 2, e, B, stream);
}

extern "C" int snb_refine_in_conv_num_tiles(int B, int H, int W) {
  return B * ((H + TH - 1) / TH) * ((W + TW - 1) / TW);
}

extern "C" int snb_refine_in_conv(const float* coarse, const float* rgb, const float* w, float* up, float* z,
                                  int B, int h, int w_, int H, int W, float disp_scale, const snb_conv_epilogue* e, void* stream) {
  SNB_REQUIRE(coarse && rgb && w && up && z && e && B > 0 && h > 0 && w_ > 0 && H > 0 && W > 0, "snb_refine_in_conv: bad args");
  RefineLoader ld{coarse, rgb, h, w_, H, W, (float)h / (float)H, (float)w_ / (float)W, disp_scale};
  const int tiles = snb_refine_in_conv_num_tiles(B, H, W);
  return snb_launch_small<4, 3, 1>("refine_in_conv", ld, w, z, up, tiles, H, W, 1, *e, 0, stream);
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
  snb_launch(kern, snb_conv5x5s2_c3_num_tiles(B, H, W), 256, smem, stream, ld, dy, partial, OH, OW, 2);
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
  snb_launch(kern, snb_refine_in_conv_num_tiles(B, H, W), 256, smem, stream, ld, dz, partial, H, W, 1);
  SNB_LAUNCH_CHECK("refine_in_wgrad");
  return 0;
}
