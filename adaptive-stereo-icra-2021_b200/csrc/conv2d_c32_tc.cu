// 2-D 32->32 stride-1 dilated 3x3 convolution on the tcgen05 tensor cores, "vertical walk" schedule.
// Same GEMM formulation as conv_c32_tc.cu (M = 128 pixels of one image row, N = 96 = 3 kw x 32 cout folded, K = kh x 32
// cin, TF32 operands with the 3xTF32 hi/lo split, fp32 accumulators in TMEM) but the tiles of one CTA walk DOWN a column
// block in steps of `dil` rows, so consecutive tiles share two of their three input-row windows:
//
//   window u  = A tile of input row r_u = rho + (j0 + u - 1)*dil, columns [x0, x0+128)          (loaded ONCE)
//   tile j    = output row rho + (j0 + j)*dil, needs windows j (kh=0), j+1 (kh=1), j+2 (kh=2)
//
// The MMA warp is window-major: when window u lands it issues  acc[tile u] += A_u*B[0], acc[tile u-1] += A_u*B[1],
// acc[tile u-2] += A_u*B[2]  and releases the A buffer at once; tile u-2 is then complete and goes to the epilogue.  Per
// tile the loaders therefore move 1 window (16 KB) instead of 3, and smem store traffic / L2 traffic drop 3x — the
// flat-tiled kernel was shared-memory-bandwidth bound (MMA operand fetch + loader stores + epilogue staging).
// The three B images (one per kh, 24 KB each incl. hi/lo) stay resident in smem for the whole kernel.
#include "tc_common.cuh"

namespace tc2d {

using namespace tc;

constexpr int NA = 3;                                    // A-window ring depth (hi + lo = 32 KB each)
constexpr int AWIN_BYTES = 2 * A_BYTES;
constexpr int BWIN_BYTES = 2 * B_BYTES;
constexpr int SY_BYTES = 3 * 128 * 128;
constexpr int SMEM_BYTES2 = NA * AWIN_BYTES + 3 * BWIN_BYTES + SY_BYTES + 3072 + 1024;

struct Params2 {
  const float* x; const float* wimg; float* y;
  int B, H, W, dil;
  int step, ncb;            // output columns per tile (128 - 2*dil), column blocks per row
  int cmax, L, nseg;        // longest row chain ceil(H/dil), tiles per strip, segments per chain
  int nstrips;              // B * ncb * dil * nseg strip slots (some are empty)
  int passes;
  snb_conv_epilogue e;
};

struct Strip { int b, cb, row0, ntiles; };   // first output row, tiles (rows row0, row0+dil, ...)

__device__ __forceinline__ Strip decode_strip(const Params2& p, int sid) {
  Strip s;
  const int seg = sid % p.nseg; sid /= p.nseg;
  const int rho = sid % p.dil; sid /= p.dil;
  s.cb = sid % p.ncb; s.b = sid / p.ncb;
  const int chain = (rho < p.H) ? (p.H - rho + p.dil - 1) / p.dil : 0;
  const int j0 = seg * p.L;
  s.ntiles = chain - j0; if (s.ntiles > p.L) s.ntiles = p.L; if (s.ntiles < 0) s.ntiles = 0;
  s.row0 = rho + j0 * p.dil;
  return s;
}

__global__ void __launch_bounds__(NTHREADS, 1)
conv2d_c32_tc_kernel(const Params2 p) {
  extern __shared__ unsigned char smem_dyn[];
  const uint32_t base_u32 = (smem_u32(smem_dyn) + 1023u) & ~1023u;
  unsigned char* base = smem_dyn + (base_u32 - smem_u32(smem_dyn));
  unsigned char* sB = base + NA * AWIN_BYTES;
  unsigned char* sYB = sB + 3 * BWIN_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sYB + SY_BYTES);
  uint64_t* full = bars;                 // [NA]   loaders -> MMA
  uint64_t* empty = bars + NA;           // [NA]   MMA commit -> loaders
  uint64_t* tfull = bars + 2 * NA;       // [NACC] MMA commit -> epilogue
  uint64_t* tempty = tfull + NACC;       // [NACC] epilogue -> MMA
  uint64_t* wbar = tempty + NACC;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(wbar + 1);
  float* sRed = reinterpret_cast<float*>(tmem_slot + 2);   // [8 warps][64]

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  if (warp == NUM_LOADER_WARPS) {
    if (lane == 0) {
      for (int i = 0; i < NA; ++i) { mbar_init(&full[i], NUM_LOADER_WARPS); mbar_init(&empty[i], 1); }
      for (int i = 0; i < NACC; ++i) { mbar_init(&tfull[i], 1); mbar_init(&tempty[i], NUM_EPI_WARPS); }
      mbar_init(wbar, 1);
      mbar_fence_init();
    }
    __syncwarp();
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;\n" :: "r"(smem_u32(tmem_slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp < NUM_LOADER_WARPS) {
    // =============================================================== loaders (window items, register prefetch PF deep)
    constexpr int PF = 3;
    const int chunk = tid & 7, rgrp = tid >> 3;            // 8 lanes = one 128-B pixel; rows rgrp + LSTRIDE*j (< 128)
    // load cursor
    int l_sid = blockIdx.x, l_u = 0;
    Strip ls = decode_strip(p, l_sid < p.nstrips ? l_sid : 0);
    auto l_skip_empty = [&]() {
      while (l_sid < p.nstrips && ls.ntiles == 0) { l_sid += gridDim.x; if (l_sid < p.nstrips) ls = decode_strip(p, l_sid); }
    };
    l_skip_empty();
    auto issue_loads = [&](float4 (&v)[LROWS]) {
      const int r = ls.row0 + (l_u - 1) * p.dil;
      const int x0 = ls.cb * p.step - p.dil;
      const bool row_ok = (unsigned)r < (unsigned)p.H;
      const float* rowp = p.x + ((size_t)ls.b * p.H + (row_ok ? r : 0)) * p.W * 32 + chunk * 4;
#pragma unroll
      for (int j = 0; j < LROWS; ++j) {
        const int xx = x0 + rgrp + LSTRIDE * j;
        v[j] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (row_ok && rgrp + LSTRIDE * j < 128 && (unsigned)xx < (unsigned)p.W) v[j] = __ldcg(reinterpret_cast<const float4*>(rowp + (size_t)xx * 32));
      }
      if (++l_u == ls.ntiles + 2) {
        l_u = 0; l_sid += gridDim.x;
        if (l_sid < p.nstrips) { ls = decode_strip(p, l_sid); l_skip_empty(); }
      }
    };
    float4 v[PF][LROWS];
#pragma unroll
    for (int k = 0; k < PF; ++k)
      if (l_sid < p.nstrips) issue_loads(v[k]);

    // store cursor: only needs to know how many windows this CTA produces
    long long n_items = 0;
    for (int sid = blockIdx.x; sid < p.nstrips; sid += gridDim.x) {
      const Strip s = decode_strip(p, sid);
      if (s.ntiles > 0) n_items += s.ntiles + 2;
    }
    uint32_t buf = 0, phase = 0;
    for (long long item0 = 0; item0 < n_items; item0 += PF) {
#pragma unroll
      for (int k = 0; k < PF; ++k) {
        if (item0 + k >= n_items) break;
        unsigned char* st = base + buf * AWIN_BYTES;
        tc::mbar_wait(&empty[buf], phase ^ 1);
#pragma unroll
        for (int j = 0; j < LROWS; ++j) {
          const int r = rgrp + LSTRIDE * j;
          if (r < 128) {
            const uint32_t off = r * 128 + ((chunk ^ (r & 7)) << 4);
            float4 hi, lo;
            split_tf32(v[k][j], hi, lo);
            *reinterpret_cast<float4*>(st + off) = hi;
            if (p.passes == 3) *reinterpret_cast<float4*>(st + A_BYTES + off) = lo;
          }
        }
        fence_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(&full[buf]);
        if (++buf == NA) { buf = 0; phase ^= 1; }
        if (l_sid < p.nstrips) issue_loads(v[k]);
      }
    }
  } else if (warp == NUM_LOADER_WARPS) {
    // =============================================================== MMA issuer (one thread), window-major
    if (lane == 0) {
      mbar_expect_tx(wbar, 3 * BWIN_BYTES);
      for (int w = 0; w < 3; ++w)
        bulk_g2s(sB + w * BWIN_BYTES, p.wimg + (size_t)w * WIMG_FLOATS_PER_WINDOW, BWIN_BYTES, wbar);
      tc::mbar_wait_spin(wbar, 0);
      const uint32_t sb_u32 = base_u32 + NA * AWIN_BYTES;
      uint32_t buf = 0, phase = 0;
      long long tile_base = 0;                 // tiles issued so far by this CTA (TMEM slot = counter % NACC)
      for (int sid = blockIdx.x; sid < p.nstrips; sid += gridDim.x) {
        const Strip s = decode_strip(p, sid);
        if (s.ntiles == 0) continue;
        for (int u = 0; u < s.ntiles + 2; ++u) {
          tc::mbar_wait_spin(&full[buf], phase);
          tc_fence_after();
          const uint32_t sa = base_u32 + buf * AWIN_BYTES;
#pragma unroll
          for (int kk = 0; kk < 3; ++kk) {
            const int kh = 2 - kk;             // finish the oldest tile first so the epilogue can start on it
            const int j = u - kh;
            if (j < 0 || j >= s.ntiles) continue;
            const long long tcount = tile_base + j;
            const int slot = (int)(tcount & (NACC - 1));
            if (kh == 0) {                     // first touch of this tile's accumulator: the epilogue must have drained it
              tc::mbar_wait_spin(&tempty[slot], (uint32_t)(((tcount / NACC) & 1) ^ 1));
              tc_fence_after();
            }
            const uint32_t tmem_d = tmem_base + slot * 128;
            const uint32_t sbw = sb_u32 + kh * BWIN_BYTES;
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
              const uint64_t ah = make_desc(sa + ks * 32), bh = make_desc(sbw + ks * 32);
              mma_tf32(tmem_d, ah, bh, (kh | ks) != 0);
              if (p.passes == 3) {
                mma_tf32(tmem_d, make_desc(sa + A_BYTES + ks * 32), bh, 1);
                mma_tf32(tmem_d, ah, make_desc(sbw + B_BYTES + ks * 32), 1);
              }
            }
            if (kh == 2) mma_commit(&tfull[slot]);      // tile j = u-2 has received all three kh contributions
          }
          mma_commit(&empty[buf]);                      // window u fully consumed
          if (++buf == NA) { buf = 0; phase ^= 1; }
        }
        tile_base += s.ntiles;
      }
    }
    __syncwarp();
  } else {
    // =============================================================== epilogue (8 warps; TMEM lane quadrant = warp % 4)
    const int ew = warp - (NUM_LOADER_WARPS + 1);
    const int quad = warp & 3;
    const int half = ew >> 2;
    const int m = quad * 32 + lane;
    const int et = tid - (NUM_LOADER_WARPS + 1) * 32;
    const int chunk = et & 7, rg = et >> 3;          // rows rg + 32*j, 16-B chunk `chunk`
    float* sY = reinterpret_cast<float*>(sYB);
    const snb_conv_epilogue& e = p.e;
    const bool has_res = e.residual != nullptr, has_stats = e.stats != nullptr;
    float4 bias4 = make_float4(0.f, 0.f, 0.f, 0.f), sc4 = make_float4(1.f, 1.f, 1.f, 1.f), sh4 = bias4;
    if (e.bias) bias4 = reinterpret_cast<const float4*>(e.bias)[chunk];
    if (e.scale) { sc4 = reinterpret_cast<const float4*>(e.scale)[chunk]; sh4 = reinterpret_cast<const float4*>(e.shift)[chunk]; }
    long long tcount = 0;
    for (int sid = blockIdx.x; sid < p.nstrips; sid += gridDim.x) {
      const Strip s = decode_strip(p, sid);
      const int x0 = s.cb * p.step - p.dil;
      for (int j = 0; j < s.ntiles; ++j, ++tcount) {
        const int slot = (int)(tcount & (NACC - 1));
        const uint32_t accphase = (uint32_t)((tcount / NACC) & 1);
        const int h = s.row0 + j * p.dil;
        const size_t rowbase = ((size_t)s.b * p.H + h) * p.W;
        bool okr[4]; int xs[4];
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) {
          const int r = rg + 32 * jj;
          xs[jj] = x0 + r;
          okr[jj] = r >= p.dil && r < 128 - p.dil && (unsigned)xs[jj] < (unsigned)p.W;
        }
        float4 res[4];
        if (has_res) {
#pragma unroll
          for (int jj = 0; jj < 4; ++jj) {
            res[jj] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (okr[jj]) res[jj] = __ldcg(reinterpret_cast<const float4*>(e.residual + (rowbase + xs[jj]) * 32 + chunk * 4));
          }
        }
        tc::mbar_wait(&tfull[slot], accphase);
        tc_fence_after();
        const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + slot * 128 + half * 16;
        {
          float v0[16], v1[16], v2[16];
          tmem_ld16x3(taddr, taddr + 32, taddr + 64, v0, v1, v2);
          float* row = sY + m * 32;
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            const int o = ((half * 4 + c) ^ (m & 7)) << 2;
            *reinterpret_cast<float4*>(row + o) = make_float4(v0[4 * c], v0[4 * c + 1], v0[4 * c + 2], v0[4 * c + 3]);
            *reinterpret_cast<float4*>(row + 128 * 32 + o) = make_float4(v1[4 * c], v1[4 * c + 1], v1[4 * c + 2], v1[4 * c + 3]);
            *reinterpret_cast<float4*>(row + 2 * 128 * 32 + o) = make_float4(v2[4 * c], v2[4 * c + 1], v2[4 * c + 2], v2[4 * c + 3]);
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tempty[slot]);
        epi_bar();
        float s1[4] = {0.f, 0.f, 0.f, 0.f}, s2[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) {       // branch-free: out-of-range rows read clamped smem rows and skip only the store
          const int r = rg + 32 * jj;
          const int r0 = max(r - p.dil, 0), r2 = min(r + p.dil, 127);
          const float4 a = *reinterpret_cast<const float4*>(sY + r0 * 32 + ((chunk ^ (r0 & 7)) << 2));
          const float4 b = *reinterpret_cast<const float4*>(sY + 128 * 32 + r * 32 + ((chunk ^ (r & 7)) << 2));
          const float4 c = *reinterpret_cast<const float4*>(sY + 2 * 128 * 32 + r2 * 32 + ((chunk ^ (r2 & 7)) << 2));
          float4 o;
          o.x = (a.x + b.x) + c.x + bias4.x; o.y = (a.y + b.y) + c.y + bias4.y;
          o.z = (a.z + b.z) + c.z + bias4.z; o.w = (a.w + b.w) + c.w + bias4.w;
          if (has_stats && okr[jj]) {
            s1[0] += o.x; s1[1] += o.y; s1[2] += o.z; s1[3] += o.w;
            s2[0] = fmaf(o.x, o.x, s2[0]); s2[1] = fmaf(o.y, o.y, s2[1]); s2[2] = fmaf(o.z, o.z, s2[2]); s2[3] = fmaf(o.w, o.w, s2[3]);
          }
          if (e.scale) { o.x = fmaf(o.x, sc4.x, sh4.x); o.y = fmaf(o.y, sc4.y, sh4.y); o.z = fmaf(o.z, sc4.z, sh4.z); o.w = fmaf(o.w, sc4.w, sh4.w); }
          if (e.lrelu) { o.x = lrelu(o.x); o.y = lrelu(o.y); o.z = lrelu(o.z); o.w = lrelu(o.w); }
          if (has_res) { o.x += res[jj].x; o.y += res[jj].y; o.z += res[jj].z; o.w += res[jj].w; }
          if (okr[jj]) __stcg(reinterpret_cast<float4*>(p.y + (rowbase + xs[jj]) * 32 + chunk * 4), o);
        }
        if (has_stats) {      // stats row = (image row, column block): [(b*H + h)*ncb + cb][2][32]
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            s1[c] += __shfl_xor_sync(0xffffffffu, s1[c], 8); s1[c] += __shfl_xor_sync(0xffffffffu, s1[c], 16);
            s2[c] += __shfl_xor_sync(0xffffffffu, s2[c], 8); s2[c] += __shfl_xor_sync(0xffffffffu, s2[c], 16);
          }
          if (lane < 8) {
#pragma unroll
            for (int c = 0; c < 4; ++c) { sRed[ew * 64 + lane * 4 + c] = s1[c]; sRed[ew * 64 + 32 + lane * 4 + c] = s2[c]; }
          }
          epi_bar();
          if (et < 64) {
            float a = 0.f;
#pragma unroll
            for (int wq = 0; wq < NUM_EPI_WARPS; ++wq) a += sRed[wq * 64 + et];
            e.stats[(((size_t)s.b * p.H + h) * p.ncb + s.cb) * 64 + et] = a;
          }
        }
        epi_bar();
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == NUM_LOADER_WARPS) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;\n" :: "r"(tmem_base) : "memory");
  }
}

}  // namespace tc2d

static int tc2d_setup(const snb_conv_geom* g, tc2d::Params2& p, const char* who) {
  SNB_REQUIRE(g != nullptr, "%s: null geometry", who);
  SNB_REQUIRE(g->transposed == 0 && g->stride == 1 && g->KD == 1 && g->KH == 3 && g->KW == 3 && g->D == 1 && g->OD == 1,
              "%s: needs a 2-D stride-1 3x3 conv", who);
  SNB_REQUIRE(g->OH == g->H && g->OW == g->W && g->ph == g->dil && g->pw == g->dil, "%s: needs 'same' padding", who);
  SNB_REQUIRE(g->dil >= 1 && g->dil <= 16, "%s: dilation out of range", who);
  p.B = g->B; p.H = g->H; p.W = g->W; p.dil = g->dil;
  p.step = 128 - 2 * g->dil;
  p.ncb = snb_ceil_div(g->W, p.step);
  p.cmax = snb_ceil_div(g->H, g->dil);
  const long long total_tiles = (long long)g->B * p.ncb * g->H;
  int L = (int)((total_tiles + 3 * 148 - 1) / (3 * 148));       // ~3 strips per SM
  if (L < 4) L = 4;
  if (L > p.cmax) L = p.cmax;
  p.L = L;
  p.nseg = snb_ceil_div(p.cmax, L);
  const long long ns = (long long)g->B * p.ncb * g->dil * p.nseg;
  SNB_REQUIRE(ns < (1ll << 30), "%s: too many strips", who);
  p.nstrips = (int)ns;
  return 0;
}

extern "C" int snb_conv2d_c32_tc_num_tiles(const snb_conv_geom* g) {
  tc2d::Params2 p;
  if (tc2d_setup(g, p, "snb_conv2d_c32_tc_num_tiles")) return -1;
  return p.B * p.H * p.ncb;          // one stats row per (image row, column block)
}

extern "C" int snb_conv2d_c32_tc(const float* x, const float* wimg, float* y, const snb_conv_geom* g, const snb_conv_epilogue* e,
                                 int passes, void* stream) {
  tc2d::Params2 p;
  if (int rc = tc2d_setup(g, p, "snb_conv2d_c32_tc")) return rc;
  SNB_REQUIRE(x && wimg && y && e, "snb_conv2d_c32_tc: null pointer");
  SNB_REQUIRE(passes == 1 || passes == 3, "snb_conv2d_c32_tc: passes must be 1 or 3");
  SNB_REQUIRE(!e->scale || e->shift, "snb_conv2d_c32_tc: scale without shift");
  p.x = x; p.wimg = wimg; p.y = y; p.passes = passes; p.e = *e;
  int dev = 0, sms = 148;
  SNB_CUDA(cudaGetDevice(&dev));
  SNB_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const int grid = p.nstrips < sms ? p.nstrips : sms;
  SNB_CUDA(cudaFuncSetAttribute(tc2d::conv2d_c32_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, tc2d::SMEM_BYTES2));
  tc2d::conv2d_c32_tc_kernel<<<grid, tc::NTHREADS, tc2d::SMEM_BYTES2, (cudaStream_t)stream>>>(p);
  SNB_LAUNCH_CHECK("conv2d_c32_tc_kernel");
  return 0;
}
