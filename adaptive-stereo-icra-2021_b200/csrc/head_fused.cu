// conv3d_alone (Conv3d 32->1, 3x3x3, pad 1) + softmax over D' + DisparityRegression (+ the pre-softmax cost output and the
// feature-contrast score) as ONE kernel — stereo_net.py:187-198, :124-134; feature_contrast.py:12-23.
//
// Round 1 ran this as two kernels with a 19 MB tap-plane round trip (head.cu: conv_c32_taps_tc + tapsum_softargmin).  Here a
// CTA owns a band of R output rows x XB output columns x all D' disparities and streams the R + 2 input rows it needs:
//
//   stage   cp.async (16-B chunks, XOR-swizzled, zero fill outside the image) of one input row segment
//           [D'][XB + 2 columns][32 ch] into a double-buffered smem row (the next row loads while this one is contracted);
//   FMA     thread = up to P positions (d, x) of that row; for every channel the 9 (kd,kw) taps of each kernel row kh that
//           this input row feeds (kh = 0 -> output row y+1, kh = 1 -> y, kh = 2 -> y-1; rows outside the band are skipped, so
//           no tap is computed twice) are accumulated with packed FFMA2 into three rotating register sets
//           U_row[d][x][kd,kw] = sum_kh sum_c w[kd,kh,kw,c] * x[d, row+kh-1, x, c];
//   gather  when an output row has received its three kernel rows its 9 tap planes go to smem once and
//           cost[d][x] = bias + sum_{kd,kw} U[d+kd-1][x+kw-1][kd,kw]   is gathered with lane = disparity, so
//   head    softmax max / sum / expectation and the FCS top-2 / sum are warp-shuffle butterflies over D' (no serial loop);
//           pred, fcs and the cost rows leave through a small smem transpose as coalesced row segments.
//
// Algorithmic bytes: 4 * B*H'W' * (32 D' + 1 [+ D' cost] [+ 1 fcs])   — the 22.5 MB activation is read
// from HBM once ((R+2)/R of it from L2), nothing but the results is written.  fp32 FMA throughout (no operand split needed).
#include "common.cuh"
#include <math.h>

namespace {

constexpr int HF_THREADS = 224;          // 7 warps
constexpr int HF_P = 3;                  // positions per thread
constexpr int HF_NPOS = HF_THREADS * HF_P;   // 672 positions (d, x) of one input row segment
constexpr int HF_WROW = 36;              // floats per channel in the smem weight table: 3 kh x 12 (9 taps (kd,kw) + 3 zero pad)

struct HeadParams {
  const float* x; const float* w; const float* bias;
  float* cost; float* pred; float* fcs;
  int B, D, H, W;
  int R, XB, XC, XCP;                    // band rows, output columns per CTA, XC = XB + 2 staged columns, XCP = odd pitch of the tap planes
  int nxb, nbands;
};

__device__ __forceinline__ void ffma2_acc(unsigned long long& acc, unsigned long long a, unsigned long long b) {
  asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc) : "l"(a), "l"(b));
}

// the 9 (kd,kw) weights of one kernel row kh for one channel, as 5 packed pairs (the 10th lane is a zero weight)
struct WRow { ulonglong2 a, b; unsigned long long c; };
__device__ __forceinline__ WRow load_wrow(const float* __restrict__ wrow, int kh) {
  WRow w;
  w.a = *reinterpret_cast<const ulonglong2*>(wrow + kh * 12);
  w.b = *reinterpret_cast<const ulonglong2*>(wrow + kh * 12 + 4);
  w.c = *reinterpret_cast<const unsigned long long*>(wrow + kh * 12 + 8);
  return w;
}
__device__ __forceinline__ void fma_khrow(unsigned long long (&acc)[5], const WRow& w, unsigned long long xp) {
  ffma2_acc(acc[0], xp, w.a.x); ffma2_acc(acc[1], xp, w.a.y);
  ffma2_acc(acc[2], xp, w.b.x); ffma2_acc(acc[3], xp, w.b.y);
  ffma2_acc(acc[4], xp, w.c);
}

// contraction of one staged input row: MASK bit kh set = this row feeds kernel row kh of an output row inside the band.
// The weights of a channel are loaded ONCE per thread (broadcast LDS) and serve its HF_P positions; positions past the end of
// the segment read as zeros (no per-position branch, which would also keep the compiler from sharing the weight loads).
template <int MASK>
__device__ __forceinline__ void contract_row(const unsigned char* __restrict__ row, const float* __restrict__ sW, int npos,
                                             unsigned long long (&accN)[HF_P][5], unsigned long long (&accC)[HF_P][5],
                                             unsigned long long (&accP)[HF_P][5]) {
  const int t = threadIdx.x;
#pragma unroll 2
  for (int c4 = 0; c4 < 8; ++c4) {
    float4 xv[HF_P];
#pragma unroll
    for (int i = 0; i < HF_P; ++i) {
      const int q = t + i * HF_THREADS;
      xv[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (q < npos) xv[i] = *reinterpret_cast<const float4*>(row + (size_t)q * 128 + ((c4 ^ (q & 7)) << 4));
    }
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      const float* wrow = sW + (c4 * 4 + kk) * HF_WROW;
      WRow w0, w1, w2;
      if (MASK & 1) w0 = load_wrow(wrow, 0);
      if (MASK & 2) w1 = load_wrow(wrow, 1);
      if (MASK & 4) w2 = load_wrow(wrow, 2);
#pragma unroll
      for (int i = 0; i < HF_P; ++i) {
        const float xs = kk == 0 ? xv[i].x : kk == 1 ? xv[i].y : kk == 2 ? xv[i].z : xv[i].w;
        const unsigned long long xp = pack2(xs, xs);
        if (MASK & 1) fma_khrow(accN[i], w0, xp);
        if (MASK & 2) fma_khrow(accC[i], w1, xp);
        if (MASK & 4) fma_khrow(accP[i], w2, xp);
      }
    }
  }
}

__global__ void __launch_bounds__(HF_THREADS, 1)
conv3d_out_softargmin_kernel(const HeadParams p) {
  pdl_launch(); pdl_wait();
  extern __shared__ __align__(128) unsigned char smem[];
  unsigned char* sRow0 = smem;                                         // 2 x HF_NPOS x 128 B staged input rows
  unsigned char* sRow1 = smem + (size_t)HF_NPOS * 128;
  float* sU = reinterpret_cast<float*>(smem + (size_t)2 * HF_NPOS * 128);   // [9][D * XCP] tap planes of one output row
  float* sW = sU + ((9 * p.D * p.XCP + 3) & ~3);                       // [32][HF_WROW], 16-byte aligned
  float* sC = sW + 32 * HF_WROW;                                       // [D][XB + 1] cost transpose | [2][XB] pred, fcs
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;

  int bid = blockIdx.x;
  const int xb = bid % p.nxb; bid /= p.nxb;
  const int band = bid % p.nbands;
  const int b = bid / p.nbands;
  const int x0 = xb * p.XB;                                            // first output column
  const int xbw = min(p.XB, p.W - x0);                                 // output columns of this CTA
  const int xc = xbw + 2;                                              // staged columns x0-1 .. x0+xbw
  const int y0 = band * p.R;
  const int nrows = min(p.R, p.H - y0);                                // output rows of this CTA
  const int npos = p.D * xc;

  // weight table: sW[c][kh*12 + kd*3 + kw] = w[c][kd][kh][kw]   (w is [1][32][3][3][3])
  for (int i = t; i < 32 * HF_WROW; i += HF_THREADS) {
    const int c = i / HF_WROW, r = i - c * HF_WROW, kh = r / 12, j = r - kh * 12;
    sW[i] = j < 9 ? __ldg(p.w + c * 27 + (j / 3) * 9 + kh * 3 + (j % 3)) : 0.f;
  }

  auto stage = [&](int yy, unsigned char* dst) {                       // input row yy (inside the image) -> dst
    const float* src = p.x + (((size_t)b * p.D) * p.H + yy) * (size_t)p.W * 32;
    const size_t dstride = (size_t)p.H * p.W * 32;
    for (int e = t; e < npos * 8; e += HF_THREADS) {
      const int q = e >> 3, c = e & 7;
      const int d = q / xc, xi = q - d * xc;
      const int xx = x0 - 1 + xi;
      const bool ok = (unsigned)xx < (unsigned)p.W;
      cp_async16(dst + (size_t)q * 128 + ((c ^ (q & 7)) << 4), src + d * dstride + (size_t)(ok ? xx : 0) * 32 + c * 4, ok);
    }
    cp_async_commit();
  };

  unsigned long long accN[HF_P][5], accC[HF_P][5], accP[HF_P][5];       // output rows yy+1 (kh=0), yy (kh=1), yy-1 (kh=2)
#pragma unroll
  for (int i = 0; i < HF_P; ++i)
#pragma unroll
    for (int j = 0; j < 5; ++j) { accN[i][j] = 0ull; accC[i][j] = 0ull; accP[i][j] = 0ull; }

  // input rows y0-1 .. y0+nrows, clipped to the image
  const int ya = max(y0 - 1, 0), yb = min(y0 + nrows, p.H - 1);
  stage(ya, sRow0);
  const float bias = __ldg(p.bias);
  const int yend = y0 + nrows;                                         // one past the last output row of this CTA

  for (int yy = ya; yy <= yb; ++yy) {
    unsigned char* cur = ((yy - ya) & 1) ? sRow1 : sRow0;
    unsigned char* nxt = ((yy - ya) & 1) ? sRow0 : sRow1;
    cp_async_wait<0>();
    __syncthreads();                                                   // row yy has landed; buffer `nxt` (row yy-1) is no longer read
    if (yy < yb) stage(yy + 1, nxt);
    // which kernel rows does input row yy feed?  kh=0 -> output yy+1, kh=1 -> yy, kh=2 -> yy-1 (each must lie in [y0, yend))
    const int mask = ((yy + 1 >= y0 && yy + 1 < yend) ? 1 : 0) | ((yy >= y0 && yy < yend) ? 2 : 0) | ((yy - 1 >= y0 && yy - 1 < yend) ? 4 : 0);
    switch (mask) {
      case 1: contract_row<1>(cur, sW, npos, accN, accC, accP); break;
      case 2: contract_row<2>(cur, sW, npos, accN, accC, accP); break;
      case 3: contract_row<3>(cur, sW, npos, accN, accC, accP); break;
      case 4: contract_row<4>(cur, sW, npos, accN, accC, accP); break;
      case 5: contract_row<5>(cur, sW, npos, accN, accC, accP); break;
      case 6: contract_row<6>(cur, sW, npos, accN, accC, accP); break;
      case 7: contract_row<7>(cur, sW, npos, accN, accC, accP); break;
      default: break;
    }
    // An output row is complete when its last contributing input row has been contracted: the row below it (kh = 2), or — at
    // the bottom of the image, where that row does not exist — the row itself (kh = 1).
    const bool emit_prev = (mask & 4) != 0;                            // accP holds output row yy-1, complete
    const bool emit_cur_last = (yy == p.H - 1) && (mask & 2) != 0;     // bottom image row: accC (output row yy) is complete too
    for (int pass = 0; pass < 2; ++pass) {
      const bool do_it = pass == 0 ? emit_prev : emit_cur_last;
      if (!do_it) continue;
      const int yout = pass == 0 ? yy - 1 : yy;
      __syncthreads();                                                 // previous gather finished reading sU / sC
#pragma unroll
      for (int i = 0; i < HF_P; ++i) {
        const int q = t + i * HF_THREADS;
        if (q < npos) {
          const int d = q / xc, xi = q - d * xc;
          float* u = sU + d * p.XCP + xi;
          const int plane = p.D * p.XCP;
          const unsigned long long* a = pass == 0 ? accP[i] : accC[i];
          u[0] = lo2(a[0]); u[plane] = hi2(a[0]); u[2 * plane] = lo2(a[1]); u[3 * plane] = hi2(a[1]);
          u[4 * plane] = lo2(a[2]); u[5 * plane] = hi2(a[2]); u[6 * plane] = lo2(a[3]); u[7 * plane] = hi2(a[3]);
          u[8 * plane] = lo2(a[4]);
        }
      }
      __syncthreads();
      // gather + head: warp per output column, lane (and lane + 32) = disparity
      for (int xo = warp; xo < xbw; xo += HF_THREADS / 32) {
        float cv[2];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int d = lane + 32 * h;
          float c = -INFINITY;
          if (d < p.D) {
            c = bias;
#pragma unroll
            for (int kd = 0; kd < 3; ++kd) {
              const int dd = d + kd - 1;
              if ((unsigned)dd >= (unsigned)p.D) continue;
              const float* up = sU + (kd * 3) * (p.D * p.XCP) + dd * p.XCP + xo;     // xi = xo + kw
              c += up[0] + up[p.D * p.XCP + 1] + up[2 * p.D * p.XCP + 2];
            }
            sC[d * (p.XB + 1) + xo] = c;
          }
          cv[h] = c;
        }
        // warp-shuffle softmax + expectation + top-2 / sum over the D' lanes (values of lanes >= D' are -inf / 0)
        float m1 = fmaxf(cv[0], cv[1]), m2 = fminf(cv[0], cv[1]);
        float ssum = (lane < p.D ? cv[0] : 0.f) + (lane + 32 < p.D ? cv[1] : 0.f);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          const float q1 = __shfl_xor_sync(0xffffffffu, m1, o), q2 = __shfl_xor_sync(0xffffffffu, m2, o);
          m2 = fmaxf(fminf(m1, q1), fmaxf(m2, q2));
          m1 = fmaxf(m1, q1);
          ssum += __shfl_xor_sync(0xffffffffu, ssum, o);
        }
        float es = 0.f, ws = 0.f;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int d = lane + 32 * h;
          if (d < p.D) { const float e = expf(cv[h] - m1); es += e; ws = fmaf(e, (float)d, ws); }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          es += __shfl_xor_sync(0xffffffffu, es, o);
          ws += __shfl_xor_sync(0xffffffffu, ws, o);
        }
        if (lane == 0) {
          float* sPF = sC + p.D * (p.XB + 1);
          sPF[xo] = ws / es;
          sPF[p.XB + xo] = m1 - (ssum - m1 - m2) / (float)(p.D - 2);
        }
      }
      __syncthreads();
      // coalesced row-segment stores
      const size_t rowoff = ((size_t)b * p.H + yout) * p.W + x0;
      const float* sPF = sC + p.D * (p.XB + 1);
      for (int i = t; i < xbw; i += HF_THREADS) {
        p.pred[rowoff + i] = sPF[i];
        if (p.fcs) p.fcs[rowoff + i] = sPF[p.XB + i];
      }
      if (p.cost) {
        for (int i = t; i < p.D * xbw; i += HF_THREADS) {
          const int d = i / xbw, xo = i - d * xbw;
          p.cost[(((size_t)b * p.D + d) * p.H + yout) * p.W + x0 + xo] = sC[d * (p.XB + 1) + xo];
        }
      }
    }
    // rotate the register sets: row yy+1 sees today's "next" as "current", today's "current" as "previous"
#pragma unroll
    for (int i = 0; i < HF_P; ++i)
#pragma unroll
      for (int j = 0; j < 5; ++j) { accP[i][j] = accC[i][j]; accC[i][j] = accN[i][j]; accN[i][j] = 0ull; }
  }
}

}  // namespace

extern "C" int snb_conv3d_out_softargmin(const float* x, const float* w, const float* bias, float* cost_out, float* pred,
                                         float* fcs_out, int B, int D, int H, int W, void* stream) {
  SNB_REQUIRE(x && w && bias && pred && B > 0 && D > 0 && H > 0 && W > 0, "snb_conv3d_out_softargmin: bad args");
  SNB_REQUIRE(D <= 64, "snb_conv3d_out_softargmin: at most 64 coarse disparity levels (got %d)", D);
  SNB_REQUIRE(fcs_out == nullptr || D > 2, "snb_conv3d_out_softargmin: the feature-contrast score needs D > 2");
  SNB_REQUIRE((reinterpret_cast<uintptr_t>(x) & 15) == 0, "snb_conv3d_out_softargmin: x must be 16-byte aligned");
  int dev = 0, sms = 148;
  SNB_CUDA(cudaGetDevice(&dev));
  SNB_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  HeadParams p;
  p.x = x; p.w = w; p.bias = bias; p.cost = cost_out; p.pred = pred; p.fcs = fcs_out;
  p.B = B; p.D = D; p.H = H; p.W = W;
  const int xbmax = HF_NPOS / D - 2;                        // D * (XB + 2) <= HF_NPOS
  SNB_REQUIRE(xbmax >= 1, "snb_conv3d_out_softargmin: D too large for the row buffer");
  // (R, XB): fewest waves over the SMs, then least halo traffic (R+2)/R * (XB+2)/XB
  double best = 1e30; int bestR = 1, bestXB = 1;
  for (int R = 1; R <= 4 && R <= H; ++R) {
    const int nb = (H + R - 1) / R;
    for (int nxb = (W + xbmax - 1) / xbmax; nxb <= W; ++nxb) {
      const int XB = (W + nxb - 1) / nxb;
      const long long ctas = (long long)B * nb * ((W + XB - 1) / XB);
      const long long waves = (ctas + sms - 1) / sms;
      // time ~ waves * (rows streamed per CTA * rounds of HF_THREADS positions per row + a fixed cost per emitted row)
      const int rounds = (D * (XB + 2) + HF_THREADS - 1) / HF_THREADS;
      const double cost = (double)waves * ((R + 2) * (double)rounds + 0.35 * R) + 1e-3 * (R + 2) * (XB + 2) / (double)(R * XB);
      if (cost < best - 1e-9) { best = cost; bestR = R; bestXB = XB; }
      if (ctas >= 4ll * sms) break;
    }
  }
  p.R = bestR; p.XB = bestXB; p.XC = p.XB + 2; p.XCP = p.XC | 1;
  p.nxb = (W + p.XB - 1) / p.XB; p.nbands = (H + p.R - 1) / p.R;
  const size_t smem = (size_t)2 * HF_NPOS * 128 + sizeof(float) * ((((size_t)9 * D * p.XCP + 3) & ~(size_t)3) + 32 * HF_WROW + (size_t)D * (p.XB + 1) + 2 * p.XB + 8);
  SNB_REQUIRE(smem <= 227 * 1024, "snb_conv3d_out_softargmin: shared-memory plan does not fit (%zu bytes)", smem);
  SNB_CUDA(cudaFuncSetAttribute(conv3d_out_softargmin_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const long long grid = (long long)B * p.nbands * p.nxb;
  SNB_REQUIRE(grid < (1ll << 31), "snb_conv3d_out_softargmin: grid too large");
  snb_launch(conv3d_out_softargmin_kernel, (int)grid, HF_THREADS, smem, stream, p);
  SNB_LAUNCH_CHECK("conv3d_out_softargmin_kernel");
  return 0;
}
