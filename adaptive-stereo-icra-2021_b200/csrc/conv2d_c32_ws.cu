// 2-D 32->32 stride-1 dilated 3x3 convolution on the tcgen05 tensor cores — "walk + shifted operand" kernel (round 2).
//
// conv2d_c32_tc.cu folds the three kw taps into N = 96 and un-shifts them in the epilogue (out[m] = Y0[m-dil] + Y1[m] + Y2[m+dil]):
// 48 KB of TMEM -> smem -> registers traffic per 128-pixel tile, and that epilogue (~2400+ cycles per tile, measured with the
// per-role counters) bounds the kernel whatever the MMA format.  Here the shift happens on the way INTO the tensor core:
//
//   raw window  one TMA tile load of the input row segment [x0 - dil, x0 + 128 + dil) x 32 ch (fp32, SWIZZLE_128B, zero fill)
//   converters  thread = output pixel m: for kw = 0,1,2 read row m + kw*dil of the raw window, split it into the fp16 pair
//               (xh | xl') and tcgen05.st it into A slot [kw]  ->  three M128 x K32 operands that are already shifted
//   MMA         acc[tile][128 x 32] += A_kw * W(kh,kw)  (M128 N32 K16 kind::f16, 3 products per k-step, see conv3d_c32_tma.cu)
//               window-major vertical walk as before: window u feeds tiles u (kh=0), u-1 (kh=1), u-2 (kh=2)
//   epilogue    the accumulator IS the convolution output: tcgen05.ld 32 columns -> * 2^-s + bias -> BN scale/shift -> LeakyReLU
//               -> + residual -> 16 KB smem transpose -> fully coalesced 16-byte stores.  4 warps, one barrier per tile.
//   residual    BasicBlock's `x + ...` (stereo_net.py:50) adds the layer's own input: the converter of the tile's centre row
//               already holds x[m] in registers and parks it in TMEM (32 columns) next to the accumulator — no second read of
//               the 60 MB activation.  A residual that is NOT the input (data gradients, polyphase chains) is read from global
//               memory in the coalesced pass.
// Every tile yields 128 valid outputs (the halo is in the raw window, not in the tile): 3 % fewer tiles than the N = 96 kernel
// at dilation 1, 14 % fewer at dilation 8.
// Warps: 0 producer, 1 MMA issuer (+ resident weight images), 2-9 converters (two groups of four, alternating windows), 10-17
// epilogue (two groups of four, alternating tiles; a staged tile leaves through one TMA tile store).
// smem: 4 x 20 KB raw ring | 4 x 16 KB staging | 72 KB weights.
// TMEM: 4 accumulators x 32 | 3 residual slots x 32 | 3 A slots x 96 (= 3 kw x (xh 16 | xl' 16)).
#include <cuda.h>
#include "tc_common.cuh"

namespace ws2 {

using namespace tc;

constexpr int NR = 4;                                    // raw window ring depth
constexpr int RAW_BYTES = 160 * 128;                     // up to 128 + 2*16 pixel rows of 128 B
constexpr int BWIN_BYTES = 2 * B_BYTES;                  // one kh window: image P [wh | wl] + image Q [2^-11 wh | -], 96 rows each
constexpr int STG_BYTES = 128 * 128;                     // one output tile, fp32
// raw ring | weights | two staging tiles per epilogue group | barriers, params
constexpr int SMEM_BYTES = NR * RAW_BYTES + 3 * BWIN_BYTES + 4 * STG_BYTES + 4096 + 1024;
constexpr int NTHREADS_WS = 18 * 32;
constexpr int CONV_WARP0 = 2, EPI_WARP0 = 10;             // 2 converter groups x 4 warps (alternating windows), 2 epilogue groups x 4 warps (alternating tiles)
constexpr int NACCW = 4, NRES = 3, NA = 3;
constexpr int ACC_BASE = 0, RES_BASE = NACCW * 32, A_BASE = RES_BASE + NRES * 32, A_COLS = 96;
static_assert(A_BASE + NA * A_COLS <= 512, "TMEM budget");

struct Params {
  const float* wimg; float* y;
  int B, H, W, dil;
  int ncb;                  // column blocks of 128 output pixels per row
  int cmax, L, nseg;        // longest row chain ceil(H/dil), tiles per strip, segments per chain
  int nstrips;
  int res_mode;             // 0: none, 1: residual == input (on-chip), 2: residual from global memory
  snb_conv_epilogue e;
  long long* dbg;
};

#define WSWAIT(acc, call) do { const long long _t0 = p.dbg ? clock64() : 0; call; if (p.dbg) acc += clock64() - _t0; } while (0)

struct Strip { int b, cb, row0, ntiles; };

__device__ __forceinline__ Strip decode_strip(const Params& p, int sid) {
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

__device__ __forceinline__ void tma_load_4d(void* smem_dst, const CUtensorMap* tmap, uint64_t* bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];\n"
      :: "r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void group_bar(int id) { asm volatile("bar.sync %0, 128;\n" :: "r"(id) : "memory"); }     // 4 warps of one group

// M128 N32 K16 kind::f16 (A from TMEM)
constexpr uint32_t IDESC_F16_N32 = (1u << 4) | ((32u >> 3) << 17) | ((128u >> 4) << 24);
__device__ __forceinline__ void mma_f16_n32(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n"
      "}\n" :: "r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(IDESC_F16_N32), "r"(accumulate) : "memory");
}

__global__ void __launch_bounds__(NTHREADS_WS, 1)
conv2d_c32_ws_kernel(const __grid_constant__ CUtensorMap tmap, const __grid_constant__ CUtensorMap tmap_out, const Params p) {
  extern __shared__ unsigned char smem_dyn[];
  const uint32_t base_u32 = (smem_u32(smem_dyn) + 1023u) & ~1023u;
  unsigned char* base = smem_dyn + (base_u32 - smem_u32(smem_dyn));
  unsigned char* sStg = base + NR * RAW_BYTES;              // [2 epilogue groups][2] output tiles (1024-B aligned: TMA store source)
  unsigned char* sB = sStg + 4 * STG_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sB + 3 * BWIN_BYTES);
  uint64_t* rfull = bars;                // [NR]    TMA (raw window) -> converters
  uint64_t* rempty = rfull + NR;         // [NR]    converters -> producer
  uint64_t* afull = rempty + NR;         // [NA]    converters -> MMA (A slot in TMEM)
  uint64_t* aempty = afull + NA;         // [NA]    MMA commit -> converters
  uint64_t* tfull = aempty + NA;         // [NACCW] MMA commit -> epilogue
  uint64_t* tempty = tfull + NACCW;      // [NACCW] epilogue -> MMA
  uint64_t* rsempty = tempty + NACCW;    // [NRES]  epilogue -> converters (residual slot drained)
  uint64_t* wbar = rsempty + NRES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(wbar + 1);
  float* sPar = reinterpret_cast<float*>(reinterpret_cast<unsigned char*>(bars) + 512);    // [64] alpha | beta (16-B aligned)
  float* sRed = sPar + 96;                                  // [2 groups][4 warps][64]

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int RW = 128 + 2 * p.dil;                           // rows of a raw window

  if (warp == 1) {
    if (lane == 0) {
      for (int i = 0; i < NR; ++i) { mbar_init(&rfull[i], 1); mbar_init(&rempty[i], 4); }
      for (int i = 0; i < NA; ++i) { mbar_init(&afull[i], 4); mbar_init(&aempty[i], 1); }
      for (int i = 0; i < NACCW; ++i) { mbar_init(&tfull[i], 1); mbar_init(&tempty[i], 4); }      // a tile is drained by ONE epilogue group
      for (int i = 0; i < NRES; ++i) mbar_init(&rsempty[i], 4);
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
  pdl_launch();                                    // only after this CTA owns its TMEM columns
  pdl_wait();                                      // everything above touched no global memory
  if (tid >= EPI_WARP0 * 32 && tid < EPI_WARP0 * 32 + 32) {
    // y = lrelu(scale * (acc * 2^-s + bias) + shift) = lrelu(acc * alpha + beta): two broadcast reads per channel quad instead of three
    const int c = tid - EPI_WARP0 * 32;
    const snb_conv_epilogue& e = p.e;
    const float bias = e.bias ? e.bias[c] : 0.f, sc = e.scale ? e.scale[c] : 1.f, sh = e.scale ? e.shift[c] : 0.f;
    sPar[c] = sc * __ldg(p.wimg + WIMG_SCALE_SLOT);
    sPar[32 + c] = fmaf(sc, bias, sh);
  }

  if (warp == 0) {
    // =============================================================== producer: one TMA tile load per window
    if (lane == 0) {
      uint32_t ac = 0;
      long long w_r = 0; const long long t0 = clock64();
      const uint32_t bytes = (uint32_t)RW * 128u;
      for (int sid = blockIdx.x; sid < p.nstrips; sid += gridDim.x) {
        const Strip s = decode_strip(p, sid);
        if (s.ntiles == 0) continue;
        const int xw = s.cb * 128 - p.dil;
        for (int u = 0; u < s.ntiles + 2; ++u) {
          const uint32_t sa = ac % NR;
          WSWAIT(w_r, tc::mbar_wait(&rempty[sa], ((ac / NR) & 1) ^ 1));
          mbar_expect_tx(&rfull[sa], bytes);
          tma_load_4d(base + sa * RAW_BYTES, &tmap, &rfull[sa], 0, xw, s.row0 + (u - 1) * p.dil, s.b);
          ++ac;
        }
      }
      if (p.dbg) { long long* dd = p.dbg + blockIdx.x * 16; dd[0] = w_r; dd[1] = clock64() - t0; }
    }
    __syncwarp();
  } else if (warp == 1) {
    // =============================================================== MMA issuer (converged warp, elected lane), window-major
    if (lane == 0) {
      mbar_expect_tx(wbar, 3 * BWIN_BYTES);
      for (int w = 0; w < 3; ++w)
        bulk_g2s(sB + w * BWIN_BYTES, p.wimg + (size_t)w * WIMG_FLOATS_PER_WINDOW, BWIN_BYTES, wbar);
    }
    __syncwarp();
    tc::mbar_wait_spin(wbar, 0);
    const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
    const uint32_t sb_u32 = __shfl_sync(0xffffffffu, base_u32, 0) + NR * RAW_BYTES + 4 * STG_BYTES;
    long long tile_base = 0;
    uint32_t win_count = 0;
    long long t_full = 0, t_tempty = 0; const long long t_mbegin = clock64();
    for (int sid = blockIdx.x; sid < p.nstrips; sid += gridDim.x) {
      const Strip s = decode_strip(p, sid);
      if (s.ntiles == 0) continue;
      for (int u = 0; u < s.ntiles + 2; ++u) {
        const uint32_t aslot = win_count % NA;
        WSWAIT(t_full, mbar_wait_warp(&afull[aslot], (win_count / NA) & 1));
        tc_fence_after();
        const uint32_t ta = tmem_u + A_BASE + aslot * A_COLS;
        uint32_t tmem_d[3]; bool act[3]; int slot_done = -1;
#pragma unroll
        for (int kk = 0; kk < 3; ++kk) {
          const int kh = 2 - kk;             // finish the oldest tile first so the epilogue can start on it
          const int j = u - kh;
          act[kk] = j >= 0 && j < s.ntiles;
          const long long tcount = tile_base + (act[kk] ? j : 0);
          const int slot = (int)(tcount & (NACCW - 1));
          if (act[kk] && kh == 0) {          // first touch of this tile's accumulator: the epilogue must have drained it
            WSWAIT(t_tempty, mbar_wait_warp(&tempty[slot], (uint32_t)(((tcount / NACCW) & 1) ^ 1)));
            tc_fence_after();
          }
          if (act[kk] && kh == 2) slot_done = slot;
          tmem_d[kk] = tmem_u + ACC_BASE + slot * 32;
        }
        if (elect_one()) {
#pragma unroll
          for (int kk = 0; kk < 3; ++kk) {
            if (!act[kk]) continue;
            const int kh = 2 - kk;
            const uint32_t sbw = sb_u32 + kh * BWIN_BYTES;
#pragma unroll
            for (int kw = 0; kw < 3; ++kw) {
              const uint32_t pimg = sbw + kw * 4096, qimg = pimg + B_BYTES, a = ta + kw * 32;
#pragma unroll
              for (int ks = 0; ks < 2; ++ks) {
                mma_f16_n32(tmem_d[kk], a + ks * 8, make_desc(pimg + ks * 32), (kh | kw | ks) != 0);        // xh . wh
                mma_f16_n32(tmem_d[kk], a + 16 + ks * 8, make_desc(qimg + ks * 32), 1);                     // xl' . 2^-11 wh
                mma_f16_n32(tmem_d[kk], a + ks * 8, make_desc(pimg + 64 + ks * 32), 1);                     // xh . wl
              }
            }
            if (kh == 2) mma_commit_raw(&tfull[slot_done]);      // tile u-2 has received all three kh contributions
          }
          mma_commit_raw(&aempty[aslot]);
        }
        __syncwarp();
        ++win_count;
      }
      tile_base += s.ntiles;
    }
    if (p.dbg && lane == 0) { long long* dd = p.dbg + blockIdx.x * 16; dd[2] = t_full; dd[3] = t_tempty; dd[4] = clock64() - t_mbegin; }
    __syncwarp();
  } else if (warp < EPI_WARP0) {
    // =============================================================== converters: 2 groups of 4 warps, alternating windows (TMEM lane
    // quadrant = warp % 4, thread = output pixel m).  For kw = 0,1,2: row m + kw*dil of the raw window -> (xh | xl') -> A slot [kw].
    // The kernel is bound by shared-memory wavefronts (B-operand fetch 432 + TMA write 144 + these reads 384 + epilogue per tile):
    // converting every row once into a packed smem copy was measured and costs 800 wavefronts instead of 384.
    const int quad = warp & 3;
    const uint32_t grp = (uint32_t)(warp - CONV_WARP0) >> 2;
    const int m = quad * 32 + lane;
    const uint32_t tlane = tmem_base + ((uint32_t)(quad * 32) << 16);
    long long w_rf = 0, w_ae = 0, w_rs = 0; const long long t0 = clock64();
    uint32_t cnt = 0;                                              // windows so far (all strips)
    long long tile_base = 0;
    for (int sid = blockIdx.x; sid < p.nstrips; sid += gridDim.x) {
      const Strip s = decode_strip(p, sid);
      if (s.ntiles == 0) continue;
      for (int u = 0; u < s.ntiles + 2; ++u, ++cnt) {
        if ((cnt & 1) != grp) continue;
        const uint32_t sr = cnt % NR, aslot = cnt % NA;
        WSWAIT(w_rf, tc::mbar_wait(&rfull[sr], (cnt / NR) & 1));
        const unsigned char* rawp = base + sr * RAW_BYTES;
        const uint32_t ta = tlane + A_BASE + aslot * A_COLS;
        // this window is the centre row (kh = 1) of tile j = u - 1: its un-shifted pixels are that tile's residual
        const bool centre = p.res_mode == 1 && u >= 1 && u <= s.ntiles;
        const long long tcount = tile_base + (u - 1);
        bool a_free = false;
#pragma unroll
        for (int kw = 0; kw < 3; ++kw) {
          const int row = m + kw * p.dil;
          const unsigned char* rp = rawp + row * 128;
          float4 v[8];
#pragma unroll
          for (int c = 0; c < 8; ++c) v[c] = *reinterpret_cast<const float4*>(rp + ((c ^ (row & 7)) << 4));
          uint32_t hl[32];
          split_f16(v, hl);
          if (kw == 2) {                                   // last read of the raw window: release it once the loads have returned
            const uint32_t dep = hl[0] ^ hl[5] ^ hl[10] ^ hl[15] ^ hl[3] ^ hl[6] ^ hl[9] ^ hl[12];     // one word of each of the 8 loads
            __syncwarp();
            if (lane == 0) mbar_arrive_after(&rempty[sr], dep);
          }
          if (!a_free) {
            WSWAIT(w_ae, tc::mbar_wait(&aempty[aslot], ((cnt / NA) & 1) ^ 1));
            tc_fence_after();
            a_free = true;
          }
          tmem_st32(ta + kw * 32, hl);
          if (kw == 1 && centre) {
            const int rslot = (int)(tcount % NRES);
            WSWAIT(w_rs, tc::mbar_wait(&rsempty[rslot], (uint32_t)(((tcount / NRES) & 1) ^ 1)));
            tc_fence_after();
            uint32_t raw[32];
#pragma unroll
            for (int c = 0; c < 8; ++c) {
              raw[4 * c] = __float_as_uint(v[c].x); raw[4 * c + 1] = __float_as_uint(v[c].y);
              raw[4 * c + 2] = __float_as_uint(v[c].z); raw[4 * c + 3] = __float_as_uint(v[c].w);
            }
            tmem_st32(tlane + RES_BASE + rslot * 32, raw);
          }
        }
        tmem_wait_st();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&afull[aslot]);
      }
      tile_base += s.ntiles;
    }
    if (p.dbg && warp == CONV_WARP0 && lane == 0) { long long* dd = p.dbg + blockIdx.x * 16; dd[11] = w_rf; dd[12] = w_ae; dd[13] = clock64() - t0; dd[14] = w_rs; }
  } else {
    // =============================================================== epilogue: 2 groups of 4 warps, alternating tiles (TMEM lane
    // quadrant = warp % 4, thread = pixel).  One tile per group at a time: its latency chain (tcgen05.ld -> pointwise -> smem
    // transpose -> coalesced stores) is ~2500 cycles, more than the MMAs of a tile.
    const int egrp = (warp - EPI_WARP0) >> 2;
    const int ew = (warp - EPI_WARP0) & 3;
    const int quad = warp & 3;
    const int m = quad * 32 + lane;
    const int et = (tid - EPI_WARP0 * 32) & 127;
    const int chunk = et & 7, rg = et >> 3;          // coalesced pass: rows rg + 16*i, 16-B chunk `chunk`
    const uint32_t tlane = tmem_base + ((uint32_t)(quad * 32) << 16);
    const snb_conv_epilogue& e = p.e;
    const bool has_stats = e.stats != nullptr;
    const float slope = e.lrelu ? SNB_LRELU_SLOPE : 1.f;
    float* red = sRed + egrp * 256;
    // Output paths.  res_mode 0 / 1: the staged tile (SWIZZLE_128B image of 128 pixels x 32 ch) leaves through ONE TMA tile store
    // (columns >= W are clipped by the tensor map) — no read-back pass.  res_mode 2 (a residual that is not the input): the coalesced
    // pass below adds it from global memory and stores.  Train-mode statistics are a read-only column pass over the staged tile.
    const bool tma_out = p.res_mode != 2;
    uint32_t ntile_g = 0;                                     // tiles this group has staged so far (staging buffer = parity)
    const int bar_id = 4 + egrp;
    asm volatile("bar.sync 6, 256;\n" ::: "memory");          // sPar (written by the first epilogue warp) is visible
    long long tcount = 0;
    long long t_tfull = 0, t_px = 0, t_out = 0, t_bar = 0, t_ld = 0, t_bar2 = 0; const long long t_ebegin = clock64();
    for (int sid = blockIdx.x; sid < p.nstrips; sid += gridDim.x) {
      const Strip s = decode_strip(p, sid);
      const int x0 = s.cb * 128;
      for (int j = 0; j < s.ntiles; ++j, ++tcount) {
        if ((int)(tcount & 1) != egrp) continue;
        const int slot = (int)(tcount & (NACCW - 1));
        const uint32_t accphase = (uint32_t)((tcount / NACCW) & 1);
        const int rslot = (int)(tcount % NRES);
        const int h = s.row0 + j * p.dil;
        const size_t rowbase = ((size_t)s.b * p.H + h) * p.W;
        float* stg = reinterpret_cast<float*>(sStg + (egrp * 2 + (ntile_g & 1)) * STG_BYTES);
        if (tma_out && ntile_g >= 2) {                // the TMA store that read this buffer two tiles ago must have drained it
          if (et == 0) asm volatile("cp.async.bulk.wait_group.read 1;\n" ::: "memory");
          group_bar(bar_id);
        }
        ++ntile_g;
        WSWAIT(t_tfull, tc::mbar_wait(&tfull[slot], accphase));
        tc_fence_after();
        const long long tB = p.dbg ? clock64() : 0;
        {
          float acc[32], res[32];
          if (p.res_mode == 1) tmem_ld32x2(tlane + ACC_BASE + slot * 32, tlane + RES_BASE + rslot * 32, acc, res);
          else tmem_ld32(tlane + ACC_BASE + slot * 32, acc);
          if (p.dbg) t_ld += clock64() - tB;
          tc_fence_before();
          __syncwarp();
          if (lane == 0) { mbar_arrive(&tempty[slot]); if (p.res_mode == 1) mbar_arrive(&rsempty[rslot]); }
          float* row = stg + m * 32;
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            const float4 a4 = *reinterpret_cast<const float4*>(sPar + 4 * c);
            const float4 b4 = *reinterpret_cast<const float4*>(sPar + 32 + 4 * c);
            float4 o;
            o.x = fmaf(acc[4 * c], a4.x, b4.x); o.y = fmaf(acc[4 * c + 1], a4.y, b4.y);
            o.z = fmaf(acc[4 * c + 2], a4.z, b4.z); o.w = fmaf(acc[4 * c + 3], a4.w, b4.w);
            o.x = o.x > 0.f ? o.x : o.x * slope; o.y = o.y > 0.f ? o.y : o.y * slope;
            o.z = o.z > 0.f ? o.z : o.z * slope; o.w = o.w > 0.f ? o.w : o.w * slope;
            if (p.res_mode == 1) { o.x += res[4 * c]; o.y += res[4 * c + 1]; o.z += res[4 * c + 2]; o.w += res[4 * c + 3]; }
            *reinterpret_cast<float4*>(row + ((c ^ (m & 7)) << 2)) = o;
          }
        }
        if (p.dbg) t_px += clock64() - tB;
        if (tma_out) fence_async_smem();              // generic-proxy writes of the tile -> visible to the TMA (async proxy)
        WSWAIT(t_bar, group_bar(bar_id));             // tile staged
        const long long tC = p.dbg ? clock64() : 0;
        if (tma_out && et == 0) {
          asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.tile.bulk_group [%0, {%2, %3, %4, %5}], [%1];\n"
                       :: "l"(reinterpret_cast<uint64_t>(&tmap_out)), "r"(smem_u32(stg)), "r"(0), "r"(x0), "r"(h), "r"(s.b) : "memory");
          asm volatile("cp.async.bulk.commit_group;\n" ::: "memory");
        }
        float s1[4] = {0.f, 0.f, 0.f, 0.f}, s2[4] = {0.f, 0.f, 0.f, 0.f};
        if (!tma_out || has_stats)
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int r = rg + 16 * i;
          const int xx = x0 + r;
          float4 o = *reinterpret_cast<const float4*>(stg + r * 32 + ((chunk ^ (r & 7)) << 2));
          if (xx < p.W) {
            const size_t off = (rowbase + xx) * 32 + chunk * 4;
            if (p.res_mode == 2) {
              const float4 rr = __ldcg(reinterpret_cast<const float4*>(e.residual + off));
              o.x += rr.x; o.y += rr.y; o.z += rr.z; o.w += rr.w;
            }
            if (has_stats) {
              s1[0] += o.x; s1[1] += o.y; s1[2] += o.z; s1[3] += o.w;
              s2[0] = fmaf(o.x, o.x, s2[0]); s2[1] = fmaf(o.y, o.y, s2[1]); s2[2] = fmaf(o.z, o.z, s2[2]); s2[3] = fmaf(o.w, o.w, s2[3]);
            }
            if (!tma_out) __stcg(reinterpret_cast<float4*>(p.y + off), o);
          }
        }
        if (has_stats) {      // stats row = (image row, column block): [(b*H + h)*ncb + cb][2][32]
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            s1[c] += __shfl_xor_sync(0xffffffffu, s1[c], 8); s1[c] += __shfl_xor_sync(0xffffffffu, s1[c], 16);
            s2[c] += __shfl_xor_sync(0xffffffffu, s2[c], 8); s2[c] += __shfl_xor_sync(0xffffffffu, s2[c], 16);
          }
          if (lane < 8) {
#pragma unroll
            for (int c = 0; c < 4; ++c) { red[ew * 64 + lane * 4 + c] = s1[c]; red[ew * 64 + 32 + lane * 4 + c] = s2[c]; }
          }
          group_bar(bar_id);
          if (et < 64) {
            float a = 0.f;
#pragma unroll
            for (int wq = 0; wq < 4; ++wq) a += red[wq * 64 + et];
            e.stats[(((size_t)s.b * p.H + h) * p.ncb + s.cb) * 64 + et] = a;
          }
        }
        if (p.dbg) t_out += clock64() - tC;
        if (!tma_out || has_stats) WSWAIT(t_bar2, group_bar(bar_id));       // staging tile / stats scratch may be rewritten
      }
    }
    if (tma_out && et == 0) asm volatile("cp.async.bulk.wait_group 0;\n" ::: "memory");    // all stores complete before the CTA exits
    if (p.dbg && et == 0 && egrp == 0) { long long* d = p.dbg + blockIdx.x * 16; d[5] = t_tfull; d[6] = clock64() - t_ebegin; d[7] = t_bar; d[9] = t_px; d[10] = t_out; d[8] = t_ld; d[15] = t_bar2; }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;\n" :: "r"(tmem_base) : "memory");
  }
}

}  // namespace ws2

static int ws2_setup(const snb_conv_geom* g, ws2::Params& p, const char* who) {
  SNB_REQUIRE(g != nullptr, "%s: null geometry", who);
  SNB_REQUIRE(g->transposed == 0 && g->stride == 1 && g->KD == 1 && g->KH == 3 && g->KW == 3 && g->D == 1 && g->OD == 1,
              "%s: needs a 2-D stride-1 3x3 conv", who);
  SNB_REQUIRE(g->OH == g->H && g->OW == g->W && g->ph == g->dil && g->pw == g->dil, "%s: needs 'same' padding", who);
  SNB_REQUIRE(g->dil >= 1 && g->dil <= 16, "%s: dilation out of range", who);
  p.B = g->B; p.H = g->H; p.W = g->W; p.dil = g->dil;
  p.ncb = snb_ceil_div(g->W, 128);
  p.cmax = snb_ceil_div(g->H, g->dil);
  // Strips = (b, column block, row residue mod dil) chains of up to cmax tiles, cut into nseg segments of L tiles; a strip
  // costs L + 2 windows.  Pick nseg so that the busiest CTA (ceil(strips / SMs) strips) loads the fewest windows.
  int dev = 0, sms = 148;
  SNB_CUDA(cudaGetDevice(&dev));
  SNB_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const long long nchains = (long long)g->B * p.ncb * g->dil;
  long long best = -1; int best_nseg = 1;
  for (int nseg = 1; nseg <= p.cmax && nseg <= 256; ++nseg) {
    const int L = snb_ceil_div(p.cmax, nseg);
    if (L < 2 && nseg > 1) break;
    const long long ns = nchains * nseg;
    const long long cost = ((ns + sms - 1) / sms) * (L + 2);
    if (best < 0 || cost < best) { best = cost; best_nseg = nseg; }
  }
  p.nseg = best_nseg;
  p.L = snb_ceil_div(p.cmax, p.nseg);
  const long long ns = nchains * p.nseg;
  SNB_REQUIRE(ns < (1ll << 30), "%s: too many strips", who);
  p.nstrips = (int)ns;
  return 0;
}

extern "C" int snb_conv2d_c32_ws_num_tiles(const snb_conv_geom* g) {
  ws2::Params p;
  if (ws2_setup(g, p, "snb_conv2d_c32_ws_num_tiles")) return -1;
  return p.B * p.H * p.ncb;          // one stats row per (image row, column block)
}

static int ws2_launch(const float* x, const float* wimg, float* y, const snb_conv_geom* g, const snb_conv_epilogue* e,
                      long long* dbg, void* stream) {
  ws2::Params p;
  if (int rc = ws2_setup(g, p, "snb_conv2d_c32_ws")) return rc;
  SNB_REQUIRE(x && wimg && y && e, "snb_conv2d_c32_ws: null pointer");
  SNB_REQUIRE((reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(y) & 15) == 0, "snb_conv2d_c32_ws: x / y must be 16-byte aligned");
  SNB_REQUIRE(!e->scale || e->shift, "snb_conv2d_c32_ws: scale without shift");
  SNB_REQUIRE(!e->stats || (!e->scale && !e->lrelu && !e->residual), "snb_conv2d_c32_ws: statistics are taken of the plain conv + bias output");
  p.wimg = wimg; p.y = y; p.e = *e; p.dbg = dbg;
  p.res_mode = e->residual == nullptr ? 0 : (e->residual == x ? 1 : 2);
  snb_encode_tiled_fn enc = snb_get_encode_tiled();
  SNB_REQUIRE(enc != nullptr, "snb_conv2d_c32_ws: cuTensorMapEncodeTiled is not available from the driver");
  CUtensorMap tmap;
  const cuuint64_t dims[4] = {32, (cuuint64_t)g->W, (cuuint64_t)g->H, (cuuint64_t)g->B};
  const cuuint64_t strides[3] = {128, (cuuint64_t)g->W * 128, (cuuint64_t)g->W * g->H * 128};      // bytes, dims 1..3
  const cuuint32_t box[4] = {32, (cuuint32_t)(128 + 2 * g->dil), 1, 1};
  const cuuint32_t estr[4] = {1, 1, 1, 1};
  const CUresult cr = enc(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(x), dims, strides, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  SNB_REQUIRE(cr == CUDA_SUCCESS, "snb_conv2d_c32_ws: cuTensorMapEncodeTiled failed (%d)", (int)cr);
  CUtensorMap tmap_out;
  const cuuint32_t box_out[4] = {32, 128, 1, 1};
  const CUresult cr2 = enc(&tmap_out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, y, dims, strides, box_out, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  SNB_REQUIRE(cr2 == CUDA_SUCCESS, "snb_conv2d_c32_ws: cuTensorMapEncodeTiled (output) failed (%d)", (int)cr2);
  int dev = 0, sms = 148;
  SNB_CUDA(cudaGetDevice(&dev));
  SNB_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const int grid = p.nstrips < sms ? p.nstrips : sms;
  SNB_CUDA(cudaFuncSetAttribute(ws2::conv2d_c32_ws_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, ws2::SMEM_BYTES));
  snb_launch(ws2::conv2d_c32_ws_kernel, grid, ws2::NTHREADS_WS, ws2::SMEM_BYTES, stream, tmap, tmap_out, p);
  SNB_LAUNCH_CHECK("conv2d_c32_ws_kernel");
  return 0;
}

extern "C" int snb_conv2d_c32_ws(const float* x, const float* wimg, float* y, const snb_conv_geom* g, const snb_conv_epilogue* e,
                                 void* stream) {
  return ws2_launch(x, wimg, y, g, e, nullptr, stream);
}

extern "C" int snb_conv2d_c32_ws_profile(const float* x, const float* wimg, float* y, const snb_conv_geom* g,
                                         const snb_conv_epilogue* e, long long* counters, void* stream) {
  SNB_REQUIRE(counters != nullptr, "snb_conv2d_c32_ws_profile: null counters");
  return ws2_launch(x, wimg, y, g, e, counters, stream);
}
