// Microbenchmark: issue rate of tcgen05.mma kind::tf32 M128 x N x K8 for several N, A operand from smem or TMEM.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_rate mma_rate.cu && ./mma_rate
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t a) {
  return (uint64_t)((a & 0x3FFFF) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n.reg .pred p;\nelect.sync _|p, 0xffffffff;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(pred));
  return pred != 0;
}

template <int N, bool TA, int NACC, int BG, int H16 = 0>
__global__ void __launch_bounds__(288, 1) rate_kernel(long long* out, int iters, int rnd, const float* gsrc) {
  __shared__ volatile int stop_flag;
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ uint64_t bar;
  __shared__ uint64_t bar2;      // target of the periodic commits (BG == 5/6)
  __shared__ uint32_t tslot;
  const uint32_t base = (smem_u32(smem) + 1023u) & ~1023u;
  for (int i = threadIdx.x; i < (16384 + 32768 + 65536) / 4; i += 288) {
    uint32_t h = (uint32_t)i * 2654435761u + blockIdx.x * 40503u; h ^= h >> 15; h *= 2246822519u; h ^= h >> 13;
    reinterpret_cast<float*>(smem)[i] = rnd ? ((float)(h & 0xffffff) / 8388608.0f - 1.0f) : 1.0f;     // random operands toggle far more bits than constants
  }
  if (threadIdx.x == 0) stop_flag = 0;
  if (threadIdx.x < 32) {
    if (threadIdx.x == 0) { asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" :: "r"(smem_u32(&bar))); asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" :: "r"(smem_u32(&bar2))); asm volatile("fence.mbarrier_init.release.cluster;\n"); }
    __syncwarp();
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;\n" :: "r"(smem_u32(&tslot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  const uint32_t tmem = tslot;
  if (rnd && threadIdx.x >= 32 && threadIdx.x < 160) {       // random A operand in TMEM columns 448..479 (4 warps = 4 lane quadrants)
    const int w = threadIdx.x >> 5, quad = w & 3;
    uint32_t r[16];
    for (int half = 0; half < 2; ++half) {
      for (int i = 0; i < 16; ++i) { uint32_t h = (threadIdx.x * 64 + half * 16 + i) * 2654435761u; h ^= h >> 15; h *= 2246822519u; r[i] = __float_as_uint((float)(h & 0xffffff) / 8388608.0f - 1.0f); }
      asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};\n"
                   :: "r"(tmem + ((uint32_t)(quad * 32) << 16) + 448 + half * 16), "r"(r[0]),"r"(r[1]),"r"(r[2]),"r"(r[3]),"r"(r[4]),"r"(r[5]),"r"(r[6]),"r"(r[7]),"r"(r[8]),"r"(r[9]),"r"(r[10]),"r"(r[11]),"r"(r[12]),"r"(r[13]),"r"(r[14]),"r"(r[15]) : "memory");
    }
    asm volatile("tcgen05.wait::st.sync.aligned;\n" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  }
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  constexpr uint32_t IDESC = H16 ? ((1u << 4) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24))
                                 : ((1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24));
  long long t0 = 0, t1 = 0;
  if (threadIdx.x < 32) {
    const uint32_t sa = base, sb = base + 16384;
    const uint32_t ta = tmem + 448;                       // A operand columns (garbage values are fine)
    t0 = clock64();
    if (elect_one()) {
      for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
#pragma unroll
          for (int a = 0; a < NACC; ++a) {
            const uint32_t td = tmem + a * (N <= 96 ? 96 : (N <= 128 ? 128 : 256)) % 448;
            const uint64_t bd = make_desc(sb + ks * 32 + (BG == 4 ? ((it * 4 + ks) % 2) * 12288 : 0));
            if (H16) asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n}\n"
                                 :: "r"(td), "r"(ta + ks * 8), "l"(bd), "r"(IDESC), "r"(1u) : "memory");
            else if (TA) asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n}\n"
                                 :: "r"(td), "r"(ta + ks * 8), "l"(bd), "r"(IDESC), "r"(1u) : "memory");
            else    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}\n"
                                 :: "r"(td), "l"(make_desc(sa + ks * 32)), "l"(bd), "r"(IDESC), "r"(1u) : "memory");
          }
          // BG == 5: one commit per 12 MMAs (as the convolution kernels do per A slot); BG == 6: one per 3 MMAs
          if (BG == 6 || (BG == 5 && ks == 3)) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" :: "r"(smem_u32(&bar2)) : "memory");
        }
      }
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" :: "r"(smem_u32(&bar)) : "memory");
    }
    __syncwarp();
    uint32_t ok = 0;
    while (!ok) asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(ok) : "r"(smem_u32(&bar)) : "memory");
    t1 = clock64();
    stop_flag = 1;
  } else if (threadIdx.x >= 32) {
    // background warps 1..8 (two per TMEM lane quadrant)
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31, quad = w & 3;
    float4* sp = reinterpret_cast<float4*>(smem + 16384 + 32768 + 1024) + (w - 1) * 512 + lane;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    uint32_t r[16];
    for (int i = 0; i < 16; ++i) r[i] = i;
    while (!stop_flag) {
      if (BG == 1) {
#pragma unroll
        for (int k = 0; k < 8; ++k) { float4 v = sp[k * 32]; acc.x += v.x; sp[k * 32 + 256] = acc; }
      } else if (BG == 2) {
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n"
                     : "=r"(r[0]),"=r"(r[1]),"=r"(r[2]),"=r"(r[3]),"=r"(r[4]),"=r"(r[5]),"=r"(r[6]),"=r"(r[7]),"=r"(r[8]),"=r"(r[9]),"=r"(r[10]),"=r"(r[11]),"=r"(r[12]),"=r"(r[13]),"=r"(r[14]),"=r"(r[15])
                     : "r"(tmem + ((uint32_t)(quad * 32) << 16) + 256) : "memory");
        asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
      } else if (BG == 3) {
        asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};\n"
                     :: "r"(tmem + ((uint32_t)(quad * 32) << 16) + 480), "r"(r[0]),"r"(r[1]),"r"(r[2]),"r"(r[3]),"r"(r[4]),"r"(r[5]),"r"(r[6]),"r"(r[7]),"r"(r[8]),"r"(r[9]),"r"(r[10]),"r"(r[11]),"r"(r[12]),"r"(r[13]),"r"(r[14]),"r"(r[15]) : "memory");
        asm volatile("tcgen05.wait::st.sync.aligned;\n" ::: "memory");
      } else if (BG == 7) {
        // one thread streams 16 KB TMA bulk copies global -> smem (async proxy writes, as the A windows / weight stages do)
        if (threadIdx.x == 32) {
          __shared__ uint64_t tbar;
          asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" :: "r"(smem_u32(&tbar)));
          asm volatile("fence.mbarrier_init.release.cluster;\n");
          uint32_t ph = 0; int slot = 0;
          while (!stop_flag) {
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" :: "r"(smem_u32(&tbar)), "r"(2 * 16384) : "memory");
            for (int k = 0; k < 2; ++k)
              asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n"
                           :: "r"(smem_u32(smem + 49152 + 1024 + ((slot + k) & 3) * 16384)), "l"(gsrc + (size_t)((blockIdx.x * 8 + slot + k) & 1023) * 4096), "r"(16384), "r"(smem_u32(&tbar)) : "memory");
            uint32_t ok = 0;
            while (!ok) asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(ok) : "r"(smem_u32(&tbar)), "r"(ph) : "memory");
            ph ^= 1; slot += 2;
          }
        } else __nanosleep(200);
      } else {
        __nanosleep(200);
      }
    }
    if (acc.x == 123.f && r[3] == 77) out[0] = 1;
  }
  __syncthreads();
  if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;\n" :: "r"(tmem) : "memory");
}

template <int N, bool TA, int NACC, int BG, int H16 = 0>
void run(const char* tag, int rnd = 0, int iters = 200) {
  long long* d; cudaMalloc(&d, 148 * sizeof(long long));
  static float* gsrc = nullptr;
  if (!gsrc) { cudaMalloc(&gsrc, 1024 * 16384); cudaMemset(gsrc, 0, 1024 * 16384); }
  const int smem = 16384 + 32768 + 65536 + 2048;
  cudaFuncSetAttribute(rate_kernel<N, TA, NACC, BG, H16>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  for (int rep = 0; rep < 2; ++rep) rate_kernel<N, TA, NACC, BG, H16><<<148, 288, smem>>>(d, iters, rnd, gsrc);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[148]; cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  double s = 0; for (int i = 0; i < 148; ++i) s += h[i];
  const double cyc = s / 148 / (iters * 4 * NACC);
  printf("%-28s N=%3d acc=%d: %7.1f cycles/MMA  %7.1f MAC/clk/SM  (%s)\n", tag, N, NACC, cyc, 128.0 * N * (H16 ? 16 : 8) / cyc, cudaGetErrorString(e));
  cudaFree(d);
}

int main(int argc, char** argv) {
  if (argc > 1) {      // round 2: kind::f16 (K = 16), A operand in TMEM
    run<32, true, 1, 0, 1>("f16 K16, A tmem"); run<64, true, 1, 0, 1>("f16 K16, A tmem"); run<96, true, 1, 0, 1>("f16 K16, A tmem");
    run<128, true, 1, 0, 1>("f16 K16, A tmem"); run<256, true, 1, 0, 1>("f16 K16, A tmem");
    run<32, true, 3, 0, 1>("f16 K16, A tmem, 3 acc"); run<96, true, 3, 0, 1>("f16 K16, A tmem, 3 acc"); run<96, true, 3, 1, 1>("f16 K16, A tmem + LDS/STS");
    run<32, true, 1, 0, 0>("tf32 K8, A tmem"); run<64, true, 1, 0, 0>("tf32 K8, A tmem");
    return 0;
  }
  run<32, false, 1, 0>("A smem"); run<64, false, 1, 0>("A smem"); run<96, false, 1, 0>("A smem"); run<128, false, 1, 0>("A smem"); run<192, false, 1, 0>("A smem"); run<256, false, 1, 0>("A smem");
  run<96, true, 1, 0>("A tmem"); run<128, true, 1, 0>("A tmem"); run<192, true, 1, 0>("A tmem"); run<256, true, 1, 0>("A tmem");
  run<96, false, 3, 0>("A smem, 3 accumulators"); run<96, true, 3, 0>("A tmem, 3 accumulators");
  run<96, true, 3, 4>("A tmem, alternating B images");
  run<96, true, 3, 1>("A tmem + 8 warps LDS/STS"); run<96, false, 3, 1>("A smem + 8 warps LDS/STS");
  run<96, true, 3, 2>("A tmem + 8 warps tcgen05.ld"); run<96, true, 3, 3>("A tmem + 8 warps tcgen05.st");
  // operand data and duration: constants vs random values, 0.8k vs 20k MMAs per SM (power management reacts to both)
  run<96, true, 3, 0>("A tmem, random data", 1); run<96, true, 3, 0>("A tmem, random, 20k MMAs", 1, 1700); run<96, true, 3, 0>("A tmem, ones, 20k MMAs", 0, 1700);
  run<96, true, 3, 1>("A tmem, random + LDS/STS", 1, 1700);
  run<96, true, 3, 7>("A tmem + TMA bulk g2s stream", 1, 1700);
  run<96, true, 3, 5>("A tmem, commit per 12 MMAs"); run<96, true, 3, 6>("A tmem, commit per 3 MMAs");
  return 0;
}
