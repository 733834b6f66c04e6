// Polyphase (space-to-depth) helpers for the stride-2 5x5 convolutions of the feature extractor (stereo_net.py:64-70,81).
// A 5x5 / stride-2 / pad-2 convolution equals the sum of four stride-1 'same' 3x3 convolutions over the four phase
// images P_ab[i][j] = x[2i+a][2j+b] (a,b in {0,1}) with the sub-kernels w[.,.,a::2,b::2] (zero padded to 3x3), so the
// forward AND the data gradient run on the tensor-core 3x3 kernel (snb_conv2d_c32_tc) instead of the FFMA kernel:
//   forward : y  = sum_ab conv3x3_same(P_ab, W_ab)            (chained through the epilogue's `residual` input)
//   dgrad   : dP_ab = conv3x3_same(dy, flip/transpose(W_ab)),  dx[2i+a][2j+b] = dP_ab[i][j]
// These two kernels only move data: split x into the four phase images and merge four phase gradients back.
#include "common.cuh"

namespace {

// x [B,H,W,32] -> out [4][B,OH,OW,32] (phase index a*2+b), rows/cols past the image are zero.  One thread = one float4.
__global__ void __launch_bounds__(256)
phase_split_kernel(const float4* __restrict__ x, float4* __restrict__ out, int B, int H, int W, int OH, int OW, long long n4) {
  pdl_launch(); pdl_wait();
  for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n4; i += (long long)gridDim.x * 256) {
    const int chunk = (int)(i & 7);
    long long p = i >> 3;
    const int j = (int)(p % OW); p /= OW;
    const int ii = (int)(p % OH); p /= OH;
    const int b = (int)(p % B); const int ph = (int)(p / B);
    const int y = 2 * ii + (ph >> 1), xx = 2 * j + (ph & 1);
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (y < H && xx < W) v = x[(((size_t)b * H + y) * W + xx) * 8 + chunk];
    out[i] = v;
  }
}

// in [4][B,OH,OW,32] -> dx [B,H,W,32]
__global__ void __launch_bounds__(256)
phase_merge_kernel(const float4* __restrict__ in, float4* __restrict__ dx, int B, int H, int W, int OH, int OW, long long n4) {
  pdl_launch(); pdl_wait();
  for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n4; i += (long long)gridDim.x * 256) {
    const int chunk = (int)(i & 7);
    long long p = i >> 3;
    const int xx = (int)(p % W); p /= W;
    const int y = (int)(p % H); const int b = (int)(p / H);
    const int ph = (y & 1) * 2 + (xx & 1);
    dx[i] = in[((((size_t)ph * B + b) * OH + (y >> 1)) * OW + (xx >> 1)) * 8 + chunk];
  }
}

}  // namespace

static int phase_grid(long long n4) {
  long long blocks = (n4 + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  return (int)(blocks < 1 ? 1 : blocks);
}

extern "C" int snb_phase_split(const float* x, float* out, int B, int H, int W, void* stream) {
  SNB_REQUIRE(x && out && B > 0 && H > 0 && W > 0, "snb_phase_split: bad args");
  const int OH = (H + 1) / 2, OW = (W + 1) / 2;
  const long long n4 = 4ll * B * OH * OW * 8;
  snb_launch(phase_split_kernel, phase_grid(n4), 256, 0, stream, (const float4*)x, (float4*)out, B, H, W, OH, OW, n4);
  SNB_LAUNCH_CHECK("phase_split_kernel");
  return 0;
}

extern "C" int snb_phase_merge(const float* in, float* dx, int B, int H, int W, void* stream) {
  SNB_REQUIRE(in && dx && B > 0 && H > 0 && W > 0, "snb_phase_merge: bad args");
  const int OH = (H + 1) / 2, OW = (W + 1) / 2;
  const long long n4 = (long long)B * H * W * 8;
  snb_launch(phase_merge_kernel, phase_grid(n4), 256, 0, stream, (const float4*)in, (float4*)dx, B, H, W, OH, OW, n4);
  SNB_LAUNCH_CHECK("phase_merge_kernel");
  return 0;
}
