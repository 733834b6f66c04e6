// 3-D 32->32 3x3x3 'same' convolution (cost filtering, stereo_net.py:155-160,185-186) on the tcgen05 tensor cores with a TMA
// producer and the A operand in TMEM.  Same GEMM formulation as conv_c32_tc.cu (M = 128 positions, N = 96 = 3 kw x 32 cout,
// K = 9 (kd,kh) windows x 32 cin, 3xTF32 split, epilogue out[m] = Y0[m-1] + Y1[m] + Y2[m+1]) but:
//   * tiles are 128 consecutive positions of the UNPADDED flat index q = h*W + w of one (b,d) slice, so a (kd,kh) window is
//     16 KB of contiguous channels-last memory: one 4-D TMA tile load (box 32 ch x 128 positions, SWIZZLE_128B, zero fill
//     for q < 0, q >= H*W and d' outside [0,D)) — no loader warps, no per-element address arithmetic.  The wrap of the +-1
//     kw shift across image rows is undone in the epilogue (Y0 is dropped at w = 0, Y2 at w = W-1);
//   * 4 converter warps (one per TMEM lane quadrant, thread = position) read the raw fp32 window (conflict-free LDS.128),
//     split hi/lo in registers and tcgen05.st it into one of two A slots in TMEM; the MMAs read A from TMEM, so the
//     per-MMA shared-memory fetch is the 3 KB B operand only (the smem-operand kernel is bound by its 7 KB A+B fetch);
//   * windows of slices outside the volume (d' = -1, D) are skipped instead of multiplied as zeros.
// Warps: 0 producer (TMA A + bulk B), 1 MMA issuer, 2-9 converters (two per TMEM lane quadrant, alternating windows),
// 10-17 epilogue (two per quadrant).  smem: 6 x 16 KB raw A ring | 3 x 24 KB B ring (hi|lo) | 48 KB epilogue tiles.
// TMEM: 3 accumulators x 96 | 3 A slots x 64.
// A CTA works on SUPER-TILES of two consecutive tiles of one slice: every (kd,kh) weight image is fetched once per pair, which
// halves the dominant L2->smem stream (24 KB of weights against 16 KB of activations per window) and doubles the MMA work
// behind each in-flight weight stage (the single-tile version was bound by exactly that stream / its latency).
// F16 = true (the product path): the error-compensated split uses fp16 operands (tcgen05.mma.kind::f16, K = 16) instead of TF32
// (K = 8): x = xh + 2^-11 xl', xh = fp16(x), xl' = fp16((x - xh) * 2^11); w' = w * 2^s (per-layer power of two, exact), wh = fp16(w'),
// wl = fp16(w' - wh); out = 2^-s (xh.wh + xl'.(2^-11 wh) + xh.wl).  Same three products and the same ~2^-22 relative error as
// 3xTF32 (both formats carry an 11-bit significand; the 2^11 pre-scaling of xl' keeps it out of the fp16 subnormals), but an
// M128 N96 K16 fp16 MMA does twice the work of the K8 TF32 one in the same 48 cycles: 6 instead of 12 MMAs per window, half
// the B-operand fetch, half the tcgen05.st traffic (A slot = 32 TMEM columns, 6 slots).  Range: |x| < 65504 (saturating
// conversion; activations behind BatchNorm are O(1..100)).  (A cta_group::2 variant of the TF32 kernel was built in round 1,
// measured 20 % slower and removed.)
#include <cuda.h>
#include <stdlib.h>
#include <string.h>
#include "tc_common.cuh"

namespace tc3 {

using namespace tc;

constexpr int NR = 6;                                    // raw A ring depth (TMA latency / NR bounds the window rate)
constexpr int NB = 3;                                    // B ring depth (one weight image serves both tiles of a pair)
constexpr int NTHREADS3 = 18 * 32;
constexpr int CONV_WARP0 = 2, EPI_WARP0 = 10;                // 8 converter warps: two per quadrant, alternating windows
constexpr int NACC3 = 3;                                 // TMEM accumulators x 96 columns
// A slots in TMEM behind the accumulators.  TF32: 3 x 64 columns (hi 32 | lo 32).  F16: 6 x 32 columns (hi 16 | lo' 16): the
// slot turnaround commit -> convert -> tcgen05.st -> arrive is longer than one MMA batch, and an fp16 batch is half as long.
template <bool F16> struct ASlots { static constexpr int N = F16 ? 6 : 3, COLS = F16 ? 32 : 64; };
constexpr int ACC_STRIDE = 96, TA_BASE = NACC3 * 96;
constexpr int SMEM_BYTES3 = NR * A_BYTES + NB * 2 * B_BYTES + OUT_BYTES + 3072 /*barriers, stats scratch*/ + 1024 /*alignment slack*/;

struct Params3 {
  const float* wimg; float* y;
  int B, D, H, W;
  int tiles_per_slice, ntiles, step;       // step = 126 output positions per tile
  int spt, nsuper;                         // super-tiles (pairs of tiles) per slice, in total
  int passes;
  snb_conv_epilogue e;
  long long* dbg;                          // optional [grid][16] cycle counters (diagnostics), or NULL
};

#define T3WAIT(acc, call) do { const long long _t0 = p.dbg ? clock64() : 0; call; if (p.dbg) acc += clock64() - _t0; } while (0)

__device__ __forceinline__ void tma_load_4d(void* smem_dst, const CUtensorMap* tmap, uint64_t* bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];\n"
      :: "r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}

// work decomposition shared by all roles: unit = one super-tile (two consecutive tiles of one slice, the last may be single)
struct Unit { int slice, sst, nh; };
__device__ __forceinline__ Unit unit_of(int u, int spt, int tiles_per_slice) {
  Unit r;
  r.slice = u / spt; r.sst = u - r.slice * spt; r.nh = (2 * r.sst + 1 < tiles_per_slice) ? 2 : 1;
  return r;
}

__device__ __forceinline__ void epi_bar3() { asm volatile("bar.sync 1, 256;\n" ::: "memory"); }

template <bool F16>
__global__ void __launch_bounds__(NTHREADS3, 1)
conv3d_c32_tma_kernel(const __grid_constant__ CUtensorMap tmap, const Params3 p) {
  constexpr int NA = ASlots<F16>::N, ACOLS = ASlots<F16>::COLS;
  extern __shared__ unsigned char smem_dyn[];
  const uint32_t base_u32 = (smem_u32(smem_dyn) + 1023u) & ~1023u;
  unsigned char* base = smem_dyn + (base_u32 - smem_u32(smem_dyn));
  unsigned char* sBring = base + NR * A_BYTES;
  unsigned char* sOutB = sBring + NB * 2 * B_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sOutB + OUT_BYTES);
  uint64_t* rfull = bars;                // [NR]   TMA (A window) -> converters
  uint64_t* rempty = rfull + NR;         // [NR]   converters -> producer
  uint64_t* bfull = rempty + NR;         // [NB]   bulk copy (B window) -> MMA
  uint64_t* bempty = bfull + NB;         // [NB]   MMA commit -> producer
  uint64_t* afull = bempty + NB;         // [NA]    converters -> MMA (A slot in TMEM)
  uint64_t* aempty = afull + NA;         // [NA]    MMA commit -> converters
  uint64_t* tfull = aempty + NA;         // [NACC3] MMA commit -> epilogue
  uint64_t* tempty = tfull + NACC3;      // [NACC3] epilogue -> MMA
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + NACC3);
  float* sRed = reinterpret_cast<float*>(tmem_slot + 2);   // [8 warps][64]

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int nunits = p.nsuper;

  if (warp == 1) {
    if (lane == 0) {
      for (int i = 0; i < NR; ++i) { mbar_init(&rfull[i], 1); mbar_init(&rempty[i], 4); }
      for (int i = 0; i < NB; ++i) { mbar_init(&bfull[i], 1); mbar_init(&bempty[i], 1); }
      for (int i = 0; i < NA; ++i) { mbar_init(&afull[i], 4); mbar_init(&aempty[i], 1); }
      for (int i = 0; i < NACC3; ++i) { mbar_init(&tfull[i], 1); mbar_init(&tempty[i], NUM_EPI_WARPS); }
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
  // Dependents may be scheduled only now that this CTA owns its TMEM columns: a dependent CTA that became co-resident and
  // allocated first would spin in griddepcontrol.wait while this one blocks in tcgen05.alloc.
  pdl_launch();
  pdl_wait();                                      // everything above touched no global memory
  const int HW = p.H * p.W;

  if (warp == 0) {
    // =============================================================== producer: one TMA tile load (A) + one bulk copy (B) per window
    // Two independent loops on two lanes: lane 0 streams the A windows, lane 1 the weight stages.  In one thread the wait for a
    // free weight stage also held back the A loads behind it in program order, and the converters starved (cv_wait_rfull).
    if (lane == 0) {
      uint32_t ac = 0;
      long long w_r = 0; const long long t0 = clock64();
      for (int u = blockIdx.x; u < nunits; u += gridDim.x) {
        const Unit un = unit_of(u, p.spt, p.tiles_per_slice);
        const int slice = un.slice, sst = un.sst, nh = un.nh;
        const int b = slice / p.D, d = slice - b * p.D;
        for (int widx = 0; widx < 9; ++widx) {
          const int kd = widx / 3, kh = widx - kd * 3;
          const int dd = d + kd - 1;
          if ((unsigned)dd >= (unsigned)p.D) continue;
          for (int h = 0; h < nh; ++h) {
            const uint32_t sa = ac % NR;
            T3WAIT(w_r, tc::mbar_wait(&rempty[sa], ((ac / NR) & 1) ^ 1));
            mbar_expect_tx(&rfull[sa], A_BYTES);
            const int q = (2 * sst + h) * p.step - 1 + (kh - 1) * p.W;
            tma_load_4d(base + sa * A_BYTES, &tmap, &rfull[sa], 0, q, dd, b);
            ++ac;
          }
        }
      }
      if (p.dbg) { long long* dd = p.dbg + blockIdx.x * 16; dd[0] = w_r; dd[2] = clock64() - t0; }
    } else if (lane == 1) {
      uint32_t bc = 0;
      long long w_b = 0;
      for (int u = blockIdx.x; u < nunits; u += gridDim.x) {
        const Unit un = unit_of(u, p.spt, p.tiles_per_slice);
        const int d = un.slice % p.D;
        for (int widx = 0; widx < 9; ++widx) {
          const int dd = d + widx / 3 - 1;
          if ((unsigned)dd >= (unsigned)p.D) continue;
          const uint32_t sb = bc % NB;
          T3WAIT(w_b, tc::mbar_wait(&bempty[sb], ((bc / NB) & 1) ^ 1));
          mbar_expect_tx(&bfull[sb], 2 * B_BYTES);
          bulk_g2s(sBring + sb * 2 * B_BYTES, p.wimg + (size_t)widx * WIMG_FLOATS_PER_WINDOW, 2 * B_BYTES, &bfull[sb]);
          ++bc;
        }
      }
      if (p.dbg) { long long* dd = p.dbg + blockIdx.x * 16; dd[1] = w_b; }
    }
    __syncwarp();
  } else if (warp == 1) {
    // =============================================================== MMA issuer (converged warp, elected lane)
    // warp-uniform copies (a shuffle from lane 0 tells nvcc the value is uniform: addresses then stay in uniform registers)
    const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0), smem_u = __shfl_sync(0xffffffffu, base_u32, 0);
    uint32_t ac = 0, bc = 0, accpar = 0;         // accpar bit a = (number of earlier uses of accumulator a) & 1
    uint32_t a_ready = 0, b_ready = 0;           // the next A slot / weight stage was already seen full by a look-ahead poll
    long long w_a = 0, w_bf = 0, w_t = 0; const long long t0 = clock64();
    int it = 0;
    for (int u = blockIdx.x; u < nunits; u += gridDim.x, ++it) {
      const Unit un = unit_of(u, p.spt, p.tiles_per_slice);
      const int d = un.slice % p.D;
      const int nh = un.nh;
      const int accs[2] = {(2 * it) % NACC3, (2 * it + 1) % NACC3};
      for (int h = 0; h < nh; ++h)
        T3WAIT(w_t, mbar_wait_warp(&tempty[accs[h]], ((accpar >> accs[h]) & 1) ^ 1));
      tc_fence_after();
      bool first = true;
      for (int widx = 0; widx < 9; ++widx) {
        const int dd = d + widx / 3 - 1;
        if ((unsigned)dd >= (unsigned)p.D) continue;
        const uint32_t sbi = bc % NB;
        if (!b_ready) T3WAIT(w_bf, mbar_wait_warp(&bfull[sbi], (bc / NB) & 1));
        b_ready = 0;
        const uint32_t sb = smem_u + NR * A_BYTES + sbi * 2 * B_BYTES;
        for (int h = 0; h < nh; ++h) {
          const uint32_t aslot = ac % NA;
          if (!a_ready) T3WAIT(w_a, mbar_wait_warp(&afull[aslot], (ac / NA) & 1));
          a_ready = 0;
          tc_fence_after();
          uint32_t peek_a = 0, peek_b = 0;       // look-ahead polls of the elected lane (see below)
          const uint32_t ta = tmem_u + TA_BASE + aslot * ACOLS;
          const uint32_t tmem_d = tmem_u + accs[h] * ACC_STRIDE;
          if (elect_one()) {                     // one election per batch of MMAs + commit (see tc_common.cuh)
            if (F16) {
              // B stage: image P rows = [wh 64 B | wl 64 B], image Q rows = [2^-11 wh 64 B | unused]; A slot = [xh 16 cols | xl' 16 cols]
#pragma unroll
              for (int ks = 0; ks < 2; ++ks) {
                if (ks == 1) {
                  // Look ahead while the pipe still has the MMAs above queued (the tensor pipe's queue is shallow: it idled during
                  // the ~100-cycle barrier polls between batches); a poll that already sees the NEXT batch's operands ready lets
                  // that batch skip its wait.
                  peek_a = mbar_try_wait(&afull[(ac + 1) % NA], ((ac + 1) / NA) & 1) ? 1u : 0u;
                  if (h == nh - 1) peek_b = mbar_try_wait(&bfull[(bc + 1) % NB], ((bc + 1) / NB) & 1) ? 1u : 0u;
                }
                const uint64_t bh = make_desc(sb + ks * 32);
                mma_f16_ts_raw(tmem_d, ta + ks * 8, bh, !(first && ks == 0));
                mma_f16_ts_raw(tmem_d, ta + 16 + ks * 8, make_desc(sb + B_BYTES + ks * 32), 1);
                mma_f16_ts_raw(tmem_d, ta + ks * 8, make_desc(sb + 64 + ks * 32), 1);
              }
            } else {
#pragma unroll
              for (int ks = 0; ks < 4; ++ks) {
                if (ks == 3) {
                  peek_a = mbar_try_wait(&afull[(ac + 1) % NA], ((ac + 1) / NA) & 1) ? 1u : 0u;
                  if (h == nh - 1) peek_b = mbar_try_wait(&bfull[(bc + 1) % NB], ((bc + 1) / NB) & 1) ? 1u : 0u;
                }
                const uint64_t bh = make_desc(sb + ks * 32);
                mma_tf32_ts_raw(tmem_d, ta + ks * 8, bh, !(first && ks == 0));
                if (p.passes == 3) {
                  mma_tf32_ts_raw(tmem_d, ta + 32 + ks * 8, bh, 1);
                  mma_tf32_ts_raw(tmem_d, ta + ks * 8, make_desc(sb + B_BYTES + ks * 32), 1);
                }
              }
            }
            mma_commit_raw(&aempty[aslot]);
          }
          a_ready = __any_sync(0xffffffffu, peek_a != 0);
          if (h == nh - 1) b_ready = __any_sync(0xffffffffu, peek_b != 0);
          ++ac;
        }
        first = false;
        if (elect_one()) mma_commit_raw(&bempty[sbi]);
        __syncwarp();
        ++bc;
      }
      for (int h = 0; h < nh; ++h) {
        if (elect_one()) mma_commit_raw(&tfull[accs[h]]);
        __syncwarp();
        accpar ^= 1u << accs[h];
      }
    }
    if (p.dbg && lane == 0) { long long* dd = p.dbg + blockIdx.x * 16; dd[3] = w_a; dd[4] = w_bf; dd[5] = w_t; dd[6] = clock64() - t0; }
  } else if (warp < EPI_WARP0) {
    // =============================================================== converters (TMEM lane quadrant = warp % 4, thread = position)
    const int quad = warp & 3;
    const uint32_t mine = (uint32_t)(warp - CONV_WARP0) >> 2;      // this warp converts the windows with cnt % 2 == mine
    const int m = quad * 32 + lane;
    uint32_t cnt = 0;                                              // A items: (window, tile of the pair)
    long long w_rf = 0, w_ae = 0, w_st = 0; const long long t0 = clock64();
    for (int u = blockIdx.x; u < nunits; u += gridDim.x) {
      const Unit un = unit_of(u, p.spt, p.tiles_per_slice);
      const int d = un.slice % p.D;
      const int nh = un.nh;
      for (int item = 0; item < 9 * nh; ++item) {
        const int widx = item / nh;
        const int dd = d + widx / 3 - 1;
        if ((unsigned)dd >= (unsigned)p.D) continue;
        if ((cnt & 1) != mine) { ++cnt; continue; }
        const uint32_t s = cnt % NR, aslot = cnt % NA;
        T3WAIT(w_rf, tc::mbar_wait(&rfull[s], (cnt / NR) & 1));
        const unsigned char* rowp = base + s * A_BYTES + m * 128;
        float4 v[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) v[c] = *reinterpret_cast<const float4*>(rowp + ((c ^ (m & 7)) << 4));
      const uint32_t dep = xor_all(v);                     // every lane's loads have RETURNED before the slot is released
      __syncwarp();
      if (lane == 0) mbar_arrive_after(&rempty[s], dep);   // (the scoreboard is per warp register: lane 0 waits for the whole warp-wide load)
        T3WAIT(w_ae, tc::mbar_wait(&aempty[aslot], ((cnt / NA) & 1) ^ 1));
        tc_fence_after();
        const uint32_t ta = tmem_base + ((uint32_t)(quad * 32) << 16) + TA_BASE + aslot * ACOLS;
        if (F16) {
          uint32_t hl[32];
          split_f16(v, hl);
          tmem_st32(ta, hl);
        } else {
          {
            uint32_t h[32];
#pragma unroll
            for (int c = 0; c < 8; ++c) {
              h[4 * c] = tc::tf32_hi_bits(v[c].x); h[4 * c + 1] = tc::tf32_hi_bits(v[c].y);
              h[4 * c + 2] = tc::tf32_hi_bits(v[c].z); h[4 * c + 3] = tc::tf32_hi_bits(v[c].w);
            }
            tmem_st32(ta, h);
          }
          if (p.passes == 3) {
            uint32_t l[32];
#pragma unroll
            for (int c = 0; c < 8; ++c) {
              const float4 x4 = v[c];
              l[4 * c] = __float_as_uint(x4.x - __uint_as_float(tc::tf32_hi_bits(x4.x)));
              l[4 * c + 1] = __float_as_uint(x4.y - __uint_as_float(tc::tf32_hi_bits(x4.y)));
              l[4 * c + 2] = __float_as_uint(x4.z - __uint_as_float(tc::tf32_hi_bits(x4.z)));
              l[4 * c + 3] = __float_as_uint(x4.w - __uint_as_float(tc::tf32_hi_bits(x4.w)));
            }
            tmem_st32(ta + 32, l);
          }
        }
        T3WAIT(w_st, tmem_wait_st());
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&afull[aslot]);
        ++cnt;
      }
    }
    if (p.dbg && warp == CONV_WARP0 && lane == 0) { long long* dd = p.dbg + blockIdx.x * 16; dd[7] = w_rf; dd[8] = w_ae; dd[9] = w_st; dd[10] = clock64() - t0; }
  } else {
    // =============================================================== epilogue (8 warps, two per TMEM lane quadrant)
    // Step 1: TMEM -> smem (Y0|Y1|Y2 tiles, rows XOR-swizzled).  Step 2: out[r] = Y0[r-1] + Y1[r] + Y2[r+1] (+bias, stats, BN
    // scale/shift, LeakyReLU, residual) with 8 lanes per 128-B position -> fully coalesced stores.
    const int ew = warp - EPI_WARP0;                 // 0..7
    const int quad = warp & 3;
    const int half = ew >> 2;                        // which 16 of the 32 output channels this warp moves out of TMEM
    const int m = quad * 32 + lane;
    const int et = tid - EPI_WARP0 * 32;
    const int chunk = et & 7, rg = et >> 3;          // rows rg + 32*j, 16-B chunk `chunk`
    float* sY = reinterpret_cast<float*>(sOutB);
    const snb_conv_epilogue& e = p.e;
    const bool has_res = e.residual != nullptr, has_stats = e.stats != nullptr;
    float4 bias4 = make_float4(0.f, 0.f, 0.f, 0.f), sc4 = make_float4(1.f, 1.f, 1.f, 1.f), sh4 = bias4;
    if (e.bias) bias4 = reinterpret_cast<const float4*>(e.bias)[chunk];
    if (e.scale) { sc4 = reinterpret_cast<const float4*>(e.scale)[chunk]; sh4 = reinterpret_cast<const float4*>(e.shift)[chunk]; }
    const float winv = F16 ? __ldg(p.wimg + WIMG_SCALE_SLOT) : 1.f;    // 2^-s of the weight image (power of two: exact)
    int it = 0;
    long long w_tf = 0; const long long t0 = clock64();
    uint32_t accpar = 0;                             // same bookkeeping as the MMA warp (a single-tile pair skips one accumulator)
    for (int u = blockIdx.x; u < nunits; u += gridDim.x, ++it) {
     const Unit un = unit_of(u, p.spt, p.tiles_per_slice);
     const int slice = un.slice, sst = un.sst, nh = un.nh;
     for (int hh2 = 0; hh2 < nh; ++hh2) {
      const int acc = (2 * it + hh2) % NACC3;
      const uint32_t accphase = (accpar >> acc) & 1;
      accpar ^= 1u << acc;
      const int tt = 2 * sst + hh2;
      const int tile = slice * p.tiles_per_slice + tt;
      const int q0 = tt * p.step - 1;
      int ww[4]; bool okr[4]; size_t goff[4];
      {
        const int q = q0 + rg;
        int h = (q >= 0) ? q / p.W : -1, w = (q >= 0) ? q - h * p.W : p.W - 1;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int r = rg + 32 * j, qq = q + 32 * j;
          ww[j] = w;
          okr[j] = r >= 1 && r < 127 && qq >= 0 && qq < HW;
          goff[j] = ((size_t)slice * HW + (size_t)(qq < 0 ? 0 : qq)) * 32 + chunk * 4;
          w += 32;
          while (w >= p.W) { w -= p.W; ++h; }
        }
      }
      float4 res[4];
      if (has_res) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          res[j] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (okr[j]) res[j] = __ldcg(reinterpret_cast<const float4*>(e.residual + goff[j]));
        }
      }
      T3WAIT(w_tf, tc::mbar_wait(&tfull[acc], accphase));
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + acc * ACC_STRIDE + half * 16;
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
      if (lane == 0) mbar_arrive(&tempty[acc]);
      epi_bar3();
      float s1[4] = {0.f, 0.f, 0.f, 0.f}, s2[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int r = rg + 32 * j;
        const int r0 = max(r - 1, 0), r2 = min(r + 1, 127);
        float4 a = *reinterpret_cast<const float4*>(sY + r0 * 32 + ((chunk ^ (r0 & 7)) << 2));
        const float4 bb = *reinterpret_cast<const float4*>(sY + 128 * 32 + r * 32 + ((chunk ^ (r & 7)) << 2));
        float4 c = *reinterpret_cast<const float4*>(sY + 2 * 128 * 32 + r2 * 32 + ((chunk ^ (r2 & 7)) << 2));
        // un-padded flat index: the kw = 0 / kw = 2 taps of the first / last column wrapped into the neighbouring image row
        if (ww[j] == 0) a = make_float4(0.f, 0.f, 0.f, 0.f);
        if (ww[j] == p.W - 1) c = make_float4(0.f, 0.f, 0.f, 0.f);
        float4 o;
        if (F16) {
          o.x = fmaf((a.x + bb.x) + c.x, winv, bias4.x); o.y = fmaf((a.y + bb.y) + c.y, winv, bias4.y);
          o.z = fmaf((a.z + bb.z) + c.z, winv, bias4.z); o.w = fmaf((a.w + bb.w) + c.w, winv, bias4.w);
        } else {
          o.x = (a.x + bb.x) + c.x + bias4.x; o.y = (a.y + bb.y) + c.y + bias4.y;
          o.z = (a.z + bb.z) + c.z + bias4.z; o.w = (a.w + bb.w) + c.w + bias4.w;
        }
        if (has_stats && okr[j]) {
          s1[0] += o.x; s1[1] += o.y; s1[2] += o.z; s1[3] += o.w;
          s2[0] = fmaf(o.x, o.x, s2[0]); s2[1] = fmaf(o.y, o.y, s2[1]); s2[2] = fmaf(o.z, o.z, s2[2]); s2[3] = fmaf(o.w, o.w, s2[3]);
        }
        if (e.scale) { o.x = fmaf(o.x, sc4.x, sh4.x); o.y = fmaf(o.y, sc4.y, sh4.y); o.z = fmaf(o.z, sc4.z, sh4.z); o.w = fmaf(o.w, sc4.w, sh4.w); }
        if (e.lrelu) { o.x = lrelu(o.x); o.y = lrelu(o.y); o.z = lrelu(o.z); o.w = lrelu(o.w); }
        if (has_res) { o.x += res[j].x; o.y += res[j].y; o.z += res[j].z; o.w += res[j].w; }
        if (okr[j]) __stcg(reinterpret_cast<float4*>(p.y + goff[j]), o);
      }
      if (has_stats) {
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          s1[c] += __shfl_xor_sync(0xffffffffu, s1[c], 8); s1[c] += __shfl_xor_sync(0xffffffffu, s1[c], 16);
          s2[c] += __shfl_xor_sync(0xffffffffu, s2[c], 8); s2[c] += __shfl_xor_sync(0xffffffffu, s2[c], 16);
        }
        if (lane < 8) {
#pragma unroll
          for (int c = 0; c < 4; ++c) { sRed[ew * 64 + lane * 4 + c] = s1[c]; sRed[ew * 64 + 32 + lane * 4 + c] = s2[c]; }
        }
        epi_bar3();
        if (et < 64) {
          float a = 0.f;
#pragma unroll
          for (int wq = 0; wq < NUM_EPI_WARPS; ++wq) a += sRed[wq * 64 + et];
          e.stats[(size_t)tile * 64 + et] = a;
        }
      }
      epi_bar3();                                    // sY / sRed are rewritten by the next tile
     }
    }
    if (p.dbg && et == 0) { long long* dd = p.dbg + blockIdx.x * 16; dd[11] = w_tf; dd[12] = clock64() - t0; }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;\n" :: "r"(tmem_base) : "memory");
  }
}

}  // namespace tc3

// ---- host side
int snb_conv3d_tma_setup(const snb_conv_geom* g, tc3::Params3& p, const char* who) {
  SNB_REQUIRE(g != nullptr, "%s: null geometry", who);
  SNB_REQUIRE(g->transposed == 0 && g->stride == 1 && g->KD == 3 && g->KH == 3 && g->KW == 3 && g->dil == 1, "%s: needs a stride-1 3x3x3 conv", who);
  SNB_REQUIRE(g->OD == g->D && g->OH == g->H && g->OW == g->W && g->ph == 1 && g->pw == 1 && g->pd == 1, "%s: needs 'same' padding", who);
  p.B = g->B; p.D = g->D; p.H = g->H; p.W = g->W;
  p.step = 126;
  p.tiles_per_slice = snb_ceil_div((long long)g->H * g->W, p.step);
  const long long nt = (long long)g->B * g->D * p.tiles_per_slice;
  SNB_REQUIRE(nt < (1ll << 30), "%s: too many tiles", who);
  p.ntiles = (int)nt;
  p.spt = (p.tiles_per_slice + 1) / 2;
  p.nsuper = g->B * g->D * p.spt;
  return 0;
}

int snb_conv3d_tma_num_tiles(const snb_conv_geom* g) {
  tc3::Params3 p;
  if (snb_conv3d_tma_setup(g, p, "snb_conv_c32_tc_num_tiles")) return -1;
  return p.ntiles;
}

int snb_conv3d_tma_launch(const float* x, const float* wimg, float* y, const snb_conv_geom* g, const snb_conv_epilogue* e,
                          int passes, long long* dbg, void* stream) {
  tc3::Params3 p;
  if (int rc = snb_conv3d_tma_setup(g, p, "snb_conv_c32_tc")) return rc;
  SNB_REQUIRE((reinterpret_cast<uintptr_t>(x) & 15) == 0, "snb_conv_c32_tc: x must be 16-byte aligned");
  const bool f16 = (passes & SNB_CONV_F16) != 0;       // weight image in the fp16-split format (snb_prep_conv_weights_tc mode | 0x10)
  passes &= 0xf;
  p.wimg = wimg; p.y = y; p.passes = passes; p.e = *e; p.dbg = dbg;
  snb_encode_tiled_fn enc = snb_get_encode_tiled();
  SNB_REQUIRE(enc != nullptr, "snb_conv_c32_tc: cuTensorMapEncodeTiled is not available from the driver");
  CUtensorMap tmap;
  const cuuint64_t HW = (cuuint64_t)g->H * g->W;
  const cuuint64_t dims[4] = {32, HW, (cuuint64_t)g->D, (cuuint64_t)g->B};
  const cuuint64_t strides[3] = {128, HW * 128, HW * 128 * (cuuint64_t)g->D};      // bytes, dims 1..3
  const cuuint32_t box[4] = {32, 128, 1, 1};
  const cuuint32_t estr[4] = {1, 1, 1, 1};
  const CUresult cr = enc(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(x), dims, strides, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  SNB_REQUIRE(cr == CUDA_SUCCESS, "snb_conv_c32_tc: cuTensorMapEncodeTiled failed (%d)", (int)cr);
  int dev = 0, sms = 148;
  SNB_CUDA(cudaGetDevice(&dev));
  SNB_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const int grid = p.nsuper < sms ? p.nsuper : sms;
  if (f16) {
    SNB_CUDA(cudaFuncSetAttribute(tc3::conv3d_c32_tma_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc3::SMEM_BYTES3));
    snb_launch(tc3::conv3d_c32_tma_kernel<true>, grid, tc3::NTHREADS3, tc3::SMEM_BYTES3, stream, tmap, p);
  } else {
    SNB_CUDA(cudaFuncSetAttribute(tc3::conv3d_c32_tma_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc3::SMEM_BYTES3));
    snb_launch(tc3::conv3d_c32_tma_kernel<false>, grid, tc3::NTHREADS3, tc3::SMEM_BYTES3, stream, tmap, p);
  }
  SNB_LAUNCH_CHECK("conv3d_c32_tma_kernel");
  return 0;
}
