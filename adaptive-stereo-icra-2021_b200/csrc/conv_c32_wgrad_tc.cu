// Weight gradient of the stride-1 'same' 3x3 (dilated, 2-D) / 3x3x3 (3-D) 32->32 convolutions on the tcgen05 tensor cores.
// (autograd of nn.Conv2d / nn.Conv3d at stereo_net.py:10-17,23-29, triggered by adapt.py:390.)
//
//   dW[kd][kh][kw][ci][co] = sum_{b,d,y,q} X[b, d+kd-1, y+(kh-1)dil, q][ci] * dZ[b, d, y, q-(kw-1)dil][co]
//
// GEMM view: the reduction (K) runs over positions q of one image row, so both operands are "MN-major" — a position is
// one 128-byte channels-last line, i.e. exactly one row of an MN-major (SWIZZLE_128B_BASE32B) smem image, and NO transpose is needed:
//   A (M = 128 = 4 row windows x 32 ci, K = q)  M-block s = one X row window; blocks sit `LBO` bytes apart
//   B (N =  96 = 3 kw x 32 co,        K = q)  ONE dZ row window with a 2*dil halo; the kw shift of N-block i is just a start
//                                               offset of i*dil rows, so LBO = dil*128 B and nothing is stored three times
//   D (128 x 96 fp32 in TMEM) = all (kh,kw) taps of one kd at once; one M-block in four is unused (kh = 3 does not exist).
// TF32 operands with the same error-compensated split as the forward kernels (passes = 3: xh*zh + xl*zh + xh*zl).
//
// Schedules ("walks", every X row window is fetched from global memory ONCE per pass over its 3 uses):
//   2-D: a CTA walks DOWN a column block in steps of dil rows.  X windows live in a 4-slot ring (slot = window counter
//        mod 4); tile j uses windows j, j+1, j+2, so the A descriptor always starts at the ring base and the tile
//        accumulates into TMEM accumulator (j mod 4), in which M-block s holds kh = (s - j) mod 4.  The four rotated
//        accumulators are un-rotated and summed through smem at the end.
//   3-D: a CTA walks along the disparity axis for one (y, column block).  A slot = slice d' with its three rows
//        y-1, y, y+1 (the kh M-blocks, LBO = Kc*128); tile d uses slots d-1, d, d+1 with accumulator kd.
// Per-CTA partial results [taps][ci][co] are summed by snb_reduce_partials (fixed order -> deterministic).
//
// Warp roles (512 threads): warps 0-14 loaders (LDG.128 -> hi/lo split -> STS), warp 15 lane 0 issues the MMAs; all 16
// warps drain TMEM at the end.
#include "tc_common.cuh"

namespace wg {

using namespace tc;

constexpr int R = 4;                       // A ring slots
constexpr int LW = 15;                     // loader warps
constexpr int LT = LW * 32;                // loader threads
constexpr int RG = LT / 8;                 // row groups (8 lanes move one 128-B row)
constexpr int LR = 3;                      // rows per loader thread per item -> items of up to 180 rows
constexpr int PF = 4;                      // items prefetched into registers
constexpr uint32_t IDESC_MN = tc::IDESC | (1u << 15) | (1u << 16);     // A and B MN-major

struct WParams {
  const float* x; const float* dz; float* partial;
  int B, D, H, W, dil, three_d;
  int Kc, ncb;                             // K chunk (positions per tile, multiple of 8), column blocks per row
  int L, nseg, nstrips;                    // tiles per strip, segments per chain
  int ximg, slot_stride, xrows;            // bytes of one slot image (hi), hi+lo, rows per slot (NWIN*Kc)
  int zr, zimg;                            // dZ window rows (multiple of 8), bytes of one dZ image
  int passes;
  float* dbg;                              // optional [grid][64] diagnostics, or NULL
};

struct Strip { int b, cb, y0, d0, nt; };   // 2-D: y0 = first dZ row; 3-D: y0 = row, d0 = first slice

__device__ __forceinline__ Strip decode_strip(const WParams& p, int sid) {
  Strip s;
  const int seg = sid % p.nseg; sid /= p.nseg;
  const int j0 = seg * p.L;
  int chain;
  if (p.three_d) {
    s.cb = sid % p.ncb; sid /= p.ncb;
    s.y0 = sid % p.H; s.b = sid / p.H;
    s.d0 = j0; chain = p.D;
  } else {
    const int rho = sid % p.dil; sid /= p.dil;
    s.cb = sid % p.ncb; s.b = sid / p.ncb;
    chain = (rho < p.H) ? (p.H - rho + p.dil - 1) / p.dil : 0;
    s.y0 = rho + j0 * p.dil; s.d0 = 0;
  }
  s.nt = chain - j0; if (s.nt > p.L) s.nt = p.L; if (s.nt < 0) s.nt = 0;
  return s;
}

// MN-major descriptor.  32-bit MN-major operands only exist in the SWIZZLE_128B_BASE32B layout (layout type 1): a row is
// 128 B (32 elements along M/N), the swizzle atom is 4 K-rows (512 B) and permutes 32-BYTE chunks: chunk c of row r is
// stored at c ^ (r & 3)  (Swizzle<2,5,2> on the byte address).
// start>>4 | LBO (byte distance between 32-element M/N blocks)>>4 <<16 | SBO (distance between 4-row K groups = 512 B)>>4 <<32
// | version 1 <<46 | layout type 1 <<61
__device__ __forceinline__ uint64_t make_desc_mn(uint32_t smem_addr, uint32_t lbo_bytes) {
  return (uint64_t)((smem_addr & 0x3FFFF) >> 4) | ((uint64_t)(lbo_bytes >> 4) << 16) | ((uint64_t)(512 >> 4) << 32) |
         ((uint64_t)1 << 46) | ((uint64_t)1 << 61);
}
// byte offset of 16-B chunk `chunk` of row `rr` inside an MN-major image (1024-B aligned base)
__device__ __forceinline__ uint32_t mn_off(int rr, int chunk) {
  return rr * 128 + (((((chunk >> 1) ^ (rr & 3)) << 1) | (chunk & 1)) << 4);
}

__device__ __forceinline__ void mma_tf32_mn(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t accumulate) {   // un-predicated: call inside `if (elect_one())`
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
      "}\n" :: "r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(IDESC_MN), "r"(accumulate) : "memory");
}

__global__ void __launch_bounds__(NTHREADS, 1)
conv_c32_wgrad_tc_kernel(const WParams p) {
  extern __shared__ unsigned char smem_dyn[];
  const uint32_t base_u32 = (smem_u32(smem_dyn) + 1023u) & ~1023u;
  unsigned char* base = smem_dyn + (base_u32 - smem_u32(smem_dyn));
  const int ring_bytes = R * p.slot_stride;
  unsigned char* sZ = base + ring_bytes;                          // two dZ buffers (hi | lo each)
  const int tail = ring_bytes + 4 * p.zimg;
  const int red_bytes = p.three_d ? 0 : 4 * 9 * 1024 * 4;         // 2-D un-rotation scratch aliases the ring
  unsigned char* misc = base + (tail > red_bytes ? tail : red_bytes);
  uint64_t* xfull = reinterpret_cast<uint64_t*>(misc);            // [R]  loaders -> MMA
  uint64_t* xempty = xfull + R;                                   // [R]  MMA commit -> loaders
  uint64_t* zfull = xempty + R;                                   // [2]
  uint64_t* zempty = zfull + 2;                                   // [2]
  uint64_t* done = zempty + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(done + 1);
  uint32_t* used_slot = tmem_slot + 1;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  if (warp == LW) {
    if (lane == 0) {
      for (int i = 0; i < R; ++i) { mbar_init(&xfull[i], LW); mbar_init(&xempty[i], 1); }
      for (int i = 0; i < 2; ++i) { mbar_init(&zfull[i], LW); mbar_init(&zempty[i], 1); }
      mbar_init(done, 1);
      *used_slot = 0;
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
  pdl_launch();                                    // dependents only once this CTA owns its TMEM columns (no alloc dead-lock with an early dependent)
  pdl_wait();                                      // everything above touched no global memory

  if (warp < LW) {
    // =============================================================== loaders
    // item stream of a strip with nt tiles:  X0 X1 | X2 Z0 | X3 Z1 | ... | X(nt+1) Z(nt-1)     (2*nt + 2 items)
    const int chunk = tid & 7, rgrp = tid >> 3;
    int xw[LR], xk[LR];                     // window / position of this thread's rows inside an X slot
#pragma unroll
    for (int j = 0; j < LR; ++j) {
      const int rr = rgrp + RG * j;
      xw[j] = p.three_d ? rr / p.Kc : 0;
      xk[j] = rr - xw[j] * p.Kc;
    }
    int l_sid = blockIdx.x, l_i = 0;
    Strip ls = decode_strip(p, l_sid < p.nstrips ? l_sid : 0);
    auto l_skip_empty = [&]() {
      while (l_sid < p.nstrips && ls.nt == 0) { l_sid += gridDim.x; if (l_sid < p.nstrips) ls = decode_strip(p, l_sid); }
    };
    l_skip_empty();
    // returns 1 when the item is a dZ window, 0 for an X slot
    auto issue_loads = [&](float4 (&v)[LR]) -> int {
      const int i = l_i;
      const bool isz = i >= 3 && (i & 1);
      const int q0 = ls.cb * p.Kc;
      if (isz) {
        const int j = (i - 3) >> 1;
        const int row = p.three_d ? ls.y0 : ls.y0 + j * p.dil;
        const int dd = p.three_d ? ls.d0 + j : 0;
        const float* rowp = p.dz + ((((size_t)ls.b * p.D + dd) * p.H + row) * p.W) * 32 + chunk * 4;
#pragma unroll
        for (int jj = 0; jj < LR; ++jj) {
          const int rr = rgrp + RG * jj;
          const int q = q0 - p.dil + rr;
          v[jj] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (rr < p.zr && (unsigned)q < (unsigned)p.W) v[jj] = __ldcg(reinterpret_cast<const float4*>(rowp + (ptrdiff_t)q * 32));
        }
      } else {
        const int u = i < 2 ? i : 2 + ((i - 2) >> 1);
#pragma unroll
        for (int jj = 0; jj < LR; ++jj) {
          const int rr = rgrp + RG * jj;
          int row, dd;
          if (p.three_d) { row = ls.y0 + xw[jj] - 1; dd = ls.d0 + u - 1; }
          else           { row = ls.y0 + (u - 1) * p.dil; dd = 0; }
          const int q = q0 + xk[jj];
          v[jj] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (rr < p.xrows && (unsigned)row < (unsigned)p.H && (unsigned)dd < (unsigned)p.D && q < p.W)
            v[jj] = __ldcg(reinterpret_cast<const float4*>(p.x + ((((size_t)ls.b * p.D + dd) * p.H + row) * p.W + q) * 32 + chunk * 4));
        }
      }
      if (++l_i == 2 * ls.nt + 2) {
        l_i = 0; l_sid += gridDim.x;
        if (l_sid < p.nstrips) { ls = decode_strip(p, l_sid); l_skip_empty(); }
      }
      return isz ? 1 : 0;
    };
    float4 v[PF][LR];
    int kind[PF];
#pragma unroll
    for (int k = 0; k < PF; ++k) { kind[k] = 0; if (l_sid < p.nstrips) kind[k] = issue_loads(v[k]); }

    long long n_items = 0;
    for (int sid = blockIdx.x; sid < p.nstrips; sid += gridDim.x) {
      const Strip s = decode_strip(p, sid);
      if (s.nt > 0) n_items += 2 * s.nt + 2;
    }
    uint32_t xc = 0, zc = 0;
    for (long long item0 = 0; item0 < n_items; item0 += PF) {
#pragma unroll
      for (int k = 0; k < PF; ++k) {
        if (item0 + k >= n_items) break;
        unsigned char* dst; int img, nrows; uint64_t* fullbar;
        if (kind[k]) {
          const uint32_t zb = zc & 1;
          tc::mbar_wait(&zempty[zb], ((zc >> 1) & 1) ^ 1);
          dst = sZ + zb * 2 * p.zimg; img = p.zimg; nrows = p.zr; fullbar = &zfull[zb]; ++zc;
        } else {
          const uint32_t slot = xc % R;
          tc::mbar_wait(&xempty[slot], ((xc / R) & 1) ^ 1);
          dst = base + slot * p.slot_stride; img = p.ximg; nrows = p.xrows; fullbar = &xfull[slot]; ++xc;
        }
#pragma unroll
        for (int j = 0; j < LR; ++j) {
          const int rr = rgrp + RG * j;
          if (rr < nrows) {
            const uint32_t off = mn_off(rr, chunk);
            float4 hi, lo;
            split_tf32(v[k][j], hi, lo);
            *reinterpret_cast<float4*>(dst + off) = hi;
            if (p.passes == 3) *reinterpret_cast<float4*>(dst + img + off) = lo;
          }
        }
        fence_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(fullbar);
        if (l_sid < p.nstrips) kind[k] = issue_loads(v[k]);
      }
    }
  } else {
    // =============================================================== MMA issuer (converged warp, elected lane)
    {
      uint32_t xc = 0, zc = 0, used = 0;
      bool any = false;
      const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0), smem_u = __shfl_sync(0xffffffffu, base_u32, 0);
      const uint32_t zb_u32 = smem_u + ring_bytes;
      for (int sid = blockIdx.x; sid < p.nstrips; sid += gridDim.x) {
        const Strip s = decode_strip(p, sid);
        if (s.nt == 0) continue;
        int valid = p.W - s.cb * p.Kc; if (valid > p.Kc) valid = p.Kc;
        const int nks = (valid + 7) >> 3;
        mbar_wait_warp(&xfull[xc % R], (xc / R) & 1);
        mbar_wait_warp(&xfull[(xc + 1) % R], ((xc + 1) / R) & 1);
        for (int j = 0; j < s.nt; ++j) {
          const uint32_t c0 = xc + j;
          mbar_wait_warp(&xfull[(c0 + 2) % R], ((c0 + 2) / R) & 1);
          mbar_wait_warp(&zfull[zc & 1], (zc >> 1) & 1);
          tc_fence_after();
          const uint32_t zb = zb_u32 + (zc & 1) * 2 * p.zimg;
          const uint32_t blbo = p.dil * 128;
          const uint32_t used_before = used;
          used |= p.three_d ? 7u : (1u << (c0 & 3));
          if (elect_one()) {                         // one election for the tile's MMAs + commits (see tc_common.cuh)
            if (p.three_d) {
#pragma unroll
              for (int kd = 0; kd < 3; ++kd) {
                const uint32_t sa = smem_u + ((c0 + kd) % R) * p.slot_stride;
                const uint32_t albo = p.Kc * 128;
                const bool fresh = !((used_before >> kd) & 1);
                for (int ks = 0; ks < nks; ++ks) {
                  const uint64_t ah = make_desc_mn(sa + ks * 1024, albo), bh = make_desc_mn(zb + ks * 1024, blbo);
                  mma_tf32_mn(tmem_u + kd * 128, ah, bh, !(fresh && ks == 0));
                  if (p.passes == 3) {
                    mma_tf32_mn(tmem_u + kd * 128, make_desc_mn(sa + p.ximg + ks * 1024, albo), bh, 1);
                    mma_tf32_mn(tmem_u + kd * 128, ah, make_desc_mn(zb + p.zimg + ks * 1024, blbo), 1);
                  }
                }
              }
            } else {
              const uint32_t acc = c0 & 3;
              const bool fresh = !((used_before >> acc) & 1);
              for (int ks = 0; ks < nks; ++ks) {
                const uint64_t ah = make_desc_mn(smem_u + ks * 1024, p.slot_stride), bh = make_desc_mn(zb + ks * 1024, blbo);
                mma_tf32_mn(tmem_u + acc * 128, ah, bh, !(fresh && ks == 0));
                if (p.passes == 3) {
                  mma_tf32_mn(tmem_u + acc * 128, make_desc_mn(smem_u + p.ximg + ks * 1024, p.slot_stride), bh, 1);
                  mma_tf32_mn(tmem_u + acc * 128, ah, make_desc_mn(zb + p.zimg + ks * 1024, blbo), 1);
                }
              }
            }
            mma_commit_raw(&xempty[c0 % R]);           // window/slot j is not needed by later tiles
            mma_commit_raw(&zempty[zc & 1]);
          }
          __syncwarp();
          any = true;
          ++zc;
        }
        mma_commit(&xempty[(xc + s.nt) % R]);        // the two trailing halo slots of this strip
        mma_commit(&xempty[(xc + s.nt + 1) % R]);
        xc += s.nt + 2;
      }
      if (lane == 0) *used_slot = used;
      if (p.dbg && lane == 0) { float* d = p.dbg + blockIdx.x * 64; d[0] = (float)used; d[1] = (float)xc; d[2] = (float)zc; d[3] = any ? 1.f : 0.f; }
      if (any) mma_commit(done); else if (lane == 0) mbar_arrive(done);
    }
    __syncwarp();
  }

  // =================================================================== drain TMEM -> per-CTA partial [taps][ci][co]
  __syncthreads();
  tc::mbar_wait(done, 0);
  tc_fence_after();
  const uint32_t used = *used_slot;
  if (p.dbg && warp < 4) {               // raw accumulator 0 / 1, columns 0-31 of lane `tid`
    float v[32];
    tmem_ld32(tmem_base + ((uint32_t)(warp * 32) << 16), v);
    if (lane < 4) { float* d = p.dbg + blockIdx.x * 64 + 8 + (warp * 4 + lane) * 2; d[0] = v[0]; d[1] = v[7]; }
  }
  const int quad = warp & 3, sub = warp >> 2;
  float* outp = p.partial + (size_t)blockIdx.x * (p.three_d ? 27 : 9) * 1024;
  if (p.three_d) {
    // accumulator kd, M-block (= TMEM lane quadrant) kh, N-block i holds kw = 2 - i; lane = ci, columns = co
    if (quad < 3) {
      for (int it = sub; it < 9; it += 4) {
        const int kd = it / 3, i = it % 3;
        float v[32];
        if ((used >> kd) & 1) tmem_ld32(tmem_base + ((uint32_t)(quad * 32) << 16) + kd * 128 + i * 32, v);
        else {
#pragma unroll
          for (int c = 0; c < 32; ++c) v[c] = 0.f;
        }
        float4* o = reinterpret_cast<float4*>(outp + ((size_t)((kd * 3 + quad) * 3 + (2 - i)) * 32 + lane) * 32);
#pragma unroll
        for (int c = 0; c < 8; ++c) o[c] = make_float4(v[4 * c], v[4 * c + 1], v[4 * c + 2], v[4 * c + 3]);
      }
    }
  } else {
    // accumulator a, M-block s (= quadrant) holds kh = (s - a) mod 4: un-rotate through smem, sum the four quadrants' shares
    float* red = reinterpret_cast<float*>(base);                 // [4 s][9 taps][32 ci][32 co], 16-B chunks swizzled by ci
    for (int it = sub; it < 9; it += 4) {
      const int kh = it / 3, i = it % 3;
      const int a = (quad - kh) & 3;
      float v[32];
      if ((used >> a) & 1) tmem_ld32(tmem_base + ((uint32_t)(quad * 32) << 16) + a * 128 + i * 32, v);
      else {
#pragma unroll
        for (int c = 0; c < 32; ++c) v[c] = 0.f;
      }
      float* row = red + ((size_t)(quad * 9 + kh * 3 + (2 - i)) * 32 + lane) * 32;
#pragma unroll
      for (int c = 0; c < 8; ++c)
        *reinterpret_cast<float4*>(row + ((c ^ (lane & 7)) << 2)) = make_float4(v[4 * c], v[4 * c + 1], v[4 * c + 2], v[4 * c + 3]);
    }
    __syncthreads();
    for (int idx = tid; idx < 9 * 32 * 8; idx += NTHREADS) {
      const int c = idx & 7, ci = (idx >> 3) & 31, tap = idx >> 8;
      const size_t o = ((size_t)tap * 32 + ci) * 32 + ((c ^ (ci & 7)) << 2);
      const float4 a0 = *reinterpret_cast<const float4*>(red + o);
      const float4 a1 = *reinterpret_cast<const float4*>(red + 9 * 1024 + o);
      const float4 a2 = *reinterpret_cast<const float4*>(red + 2 * 9 * 1024 + o);
      const float4 a3 = *reinterpret_cast<const float4*>(red + 3 * 9 * 1024 + o);
      float4 r;
      r.x = (a0.x + a1.x) + (a2.x + a3.x); r.y = (a0.y + a1.y) + (a2.y + a3.y);
      r.z = (a0.z + a1.z) + (a2.z + a3.z); r.w = (a0.w + a1.w) + (a2.w + a3.w);
      *reinterpret_cast<float4*>(outp + ((size_t)tap * 32 + ci) * 32 + c * 4) = r;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == LW) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;\n" :: "r"(tmem_base) : "memory");
  }
}

}  // namespace wg

static int wgtc_setup(const snb_conv_geom* g, wg::WParams& p, int* smem_bytes, int* grid, const char* who) {
  SNB_REQUIRE(g != nullptr, "%s: null geometry", who);
  SNB_REQUIRE(g->transposed == 0 && g->stride == 1 && g->KH == 3 && g->KW == 3 && (g->KD == 1 || g->KD == 3),
              "%s: needs a stride-1 3x3(x3) conv", who);
  SNB_REQUIRE(g->OD == g->D && g->OH == g->H && g->OW == g->W && g->ph == g->dil && g->pw == g->dil &&
              g->pd == (g->KD == 3 ? 1 : 0), "%s: needs 'same' padding", who);
  SNB_REQUIRE(g->dil >= 1 && g->dil <= 16 && (g->KD == 1 || g->dil == 1), "%s: dilation out of range", who);
  SNB_REQUIRE(g->KD == 3 || g->D == 1, "%s: 2-D conv with D != 1", who);
  p.B = g->B; p.D = g->D; p.H = g->H; p.W = g->W; p.dil = g->dil; p.three_d = g->KD == 3;
  const int kmax = p.three_d ? 56 : 128;
  p.ncb = snb_ceil_div(g->W, kmax);
  p.Kc = ((snb_ceil_div(g->W, p.ncb) + 7) / 8) * 8;
  p.xrows = (p.three_d ? 3 : 1) * p.Kc;
  p.ximg = p.xrows * 128;
  p.slot_stride = 2 * p.ximg;
  p.zr = ((p.Kc + 2 * p.dil + 7) / 8) * 8;
  p.zimg = p.zr * 128;
  SNB_REQUIRE(p.xrows <= wg::RG * wg::LR && p.zr <= wg::RG * wg::LR, "%s: window too tall", who);
  // strips: chains (2-D: B x column blocks x dil row residues of up to cmax rows; 3-D: B x H x column blocks of D slices)
  const long long nchains = p.three_d ? (long long)g->B * g->H * p.ncb : (long long)g->B * p.ncb * g->dil;
  const int cmax = p.three_d ? g->D : snb_ceil_div(g->H, g->dil);
  int dev = 0, sms = 148;
  SNB_CUDA(cudaGetDevice(&dev));
  SNB_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  // pick the number of segments per chain that minimises the busiest CTA's window count (tiles + 2 halo windows per strip)
  long long best = -1; int best_nseg = 1;
  for (int nseg = 1; nseg <= cmax && nseg <= 64; ++nseg) {
    const int L = snb_ceil_div(cmax, nseg);
    if (L < 3 && nseg > 1) break;
    const long long ns = nchains * nseg;
    const long long cost = ((ns + sms - 1) / sms) * (L + 2);
    if (best < 0 || cost < best) { best = cost; best_nseg = nseg; }
  }
  p.nseg = best_nseg;
  p.L = snb_ceil_div(cmax, p.nseg);
  const long long ns = nchains * p.nseg;
  SNB_REQUIRE(ns < (1ll << 30), "%s: too many strips", who);
  p.nstrips = (int)ns;
  *grid = p.nstrips < sms ? p.nstrips : sms;
  const int tail = wg::R * p.slot_stride + 4 * p.zimg;
  const int red = p.three_d ? 0 : 4 * 9 * 1024 * 4;
  *smem_bytes = (tail > red ? tail : red) + 256 /*barriers*/ + 1024 /*alignment slack*/;
  SNB_REQUIRE(*smem_bytes <= 227 * 1024, "%s: shared memory budget exceeded", who);
  return 0;
}

extern "C" int snb_conv_c32_wgrad_tc_num_partials(const snb_conv_geom* g) {
  wg::WParams p; int smem, grid;
  if (wgtc_setup(g, p, &smem, &grid, "snb_conv_c32_wgrad_tc_num_partials")) return -1;
  return grid;
}

static int wgtc_launch(const float* x, const float* dz, float* partial, const snb_conv_geom* g, int passes, float* dbg, void* stream);

extern "C" int snb_conv_c32_wgrad_tc(const float* x, const float* dz, float* partial, const snb_conv_geom* g, int passes, void* stream) {
  return wgtc_launch(x, dz, partial, g, passes, nullptr, stream);
}

extern "C" int snb_conv_c32_wgrad_tc_debug(const float* x, const float* dz, float* partial, const snb_conv_geom* g, int passes,
                                           float* dbg, void* stream) {
  SNB_REQUIRE(dbg != nullptr, "snb_conv_c32_wgrad_tc_debug: null dbg");
  return wgtc_launch(x, dz, partial, g, passes, dbg, stream);
}

static int wgtc_launch(const float* x, const float* dz, float* partial, const snb_conv_geom* g, int passes, float* dbg, void* stream) {
  wg::WParams p; int smem, grid;
  if (int rc = wgtc_setup(g, p, &smem, &grid, "snb_conv_c32_wgrad_tc")) return rc;
  SNB_REQUIRE(x && dz && partial, "snb_conv_c32_wgrad_tc: null pointer");
  SNB_REQUIRE(passes == 1 || passes == 3, "snb_conv_c32_wgrad_tc: passes must be 1 or 3");
  p.x = x; p.dz = dz; p.partial = partial; p.passes = passes; p.dbg = dbg;
  SNB_CUDA(cudaFuncSetAttribute(wg::conv_c32_wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  snb_launch(wg::conv_c32_wgrad_tc_kernel, grid, tc::NTHREADS, smem, stream, p);
  SNB_LAUNCH_CHECK("conv_c32_wgrad_tc_kernel");
  return 0;
}
