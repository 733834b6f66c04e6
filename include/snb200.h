/*
 * snb200.h — C-ABI of the B200-native StereoNet hot path (libsnb200.so).
 *
 * Drop-in boundary for miloknowles/adaptive-stereo-icra-2021's StereoNet forward / online-adaptation step.
 * The reference has no native interface (it is 100 % PyTorch); every entry point below replaces the ATen/cuDNN
 * calls issued by one span of adaptive_stereo/models/stereo_net.py (cited per function, paths relative to the
 * reference checkout).  The host side (stereonet_b200/models/stereo_net.py) mirrors the reference nn.Module API
 * and calls these through ctypes with raw device pointers.
 *
 * Conventions
 *  - plain C types only: device pointers, ints, floats, an opaque cudaStream_t passed as void*.
 *  - the caller owns every buffer (inputs, outputs, workspaces); the library never allocates, frees or
 *    synchronises, and every launch is ordered on the given stream (CUDA-graph capturable).
 *  - activations are fp32 channels-last: [B,H,W,32] (2-D) or [B,D,H,W,32] (3-D); images are NCHW fp32 as the
 *    reference's callers pass them; single-channel maps are [B,H,W] / [B,D,H,W].
 *  - return value 0 on success, non-zero on failure; snb_last_error() returns a thread-local message.
 *  - re-entrant: no global mutable state (the reference may be driven from a ROS callback thread,
 *    ros/stereo_depth_node.py:113).
 */
#ifndef SNB200_H_
#define SNB200_H_

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define SNB_API __attribute__((visibility("default")))
#else
#define SNB_API
#endif

#define SNB_C 32 /* channel width of every hidden activation (stereo_net.py:59,64,155) */

/* Geometry of one convolution over channels-last tensors. 2-D convs use D = OD = KD = 1, pd = 0. */
typedef struct {
  int B, D, H, W;     /* input  [B,D,H,W,C] */
  int OD, OH, OW;     /* output [B,OD,OH,OW,C] */
  int KD, KH, KW;     /* taps */
  int stride;         /* H/W stride (depth stride is always 1) */
  int dil;            /* H/W dilation (depth dilation is always 1) */
  int pd, ph, pw;     /* zero padding */
} snb_conv_geom;

/* Epilogue of the 32->32 convolution: z = conv + bias; stats += (sum z, sum z^2) per channel;
 * y = z*scale + shift (if scale); y = LeakyReLU_0.2(y) (if lrelu); y += residual (if residual). */
typedef struct {
  const float* bias;      /* [32] or NULL */
  const float* scale;     /* [32] or NULL : folded eval-mode BatchNorm (stereo_net.py:17,29) */
  const float* shift;     /* [32] (required when scale != NULL) */
  const float* residual;  /* channels-last tensor shaped like the output, or NULL (BasicBlock, stereo_net.py:50) */
  float* stats;           /* [ntiles][2][32] per-tile partial sums for train-mode BN, or NULL */
  int lrelu;              /* 1: LeakyReLU(0.2) (stereo_net.py:39,94,159) */
} snb_conv_epilogue;

SNB_API const char* snb_last_error(void);
SNB_API int snb_version(void);

/* Number of 128-position tiles (= rows of the `stats` partial buffer) snb_conv_c32 uses for this geometry. */
SNB_API int snb_conv_c32_num_tiles(const snb_conv_geom* g);

/* Repack a PyTorch conv weight [Cout][Cin][taps] into the kernels' [taps][Cin][Cout] layout.
 * mode 0: forward weights.  mode 1: data-gradient weights (taps flipped, Cin/Cout swapped). */
SNB_API int snb_prep_conv_weights(const float* w, float* out, int cout, int cin, int taps, int mode, void* stream);

/* Difference cost volume — replaces the host-side zero+H2D+24-slice loop of stereo_net.py:173-184.
 * left/right [B,H,W,32] -> cost [B,D,H,W,32]; cost[b,d,y,x,c] = left[b,y,x,c] - right[b,y,x-d,c] (x >= d) else 0. */
SNB_API int snb_cost_volume_fwd(const float* left, const float* right, float* cost, int B, int D, int H, int W, void* stream);
/* Adjoint of the above (autograd of the slice assignments, stereo_net.py:178-182). */
SNB_API int snb_cost_volume_bwd(const float* dcost, float* dleft, float* dright, int B, int D, int H, int W, void* stream);

/* 32->32 convolution, fp32 CUDA-core (FFMA) implicit GEMM — replaces nn.Conv2d / nn.Conv3d (+ folded BN, LeakyReLU,
 * residual add) at stereo_net.py:10-17,23-29,44-51,64-70,77,81-85,185-186.  wprep from snb_prep_conv_weights. */
SNB_API int snb_conv_c32(const float* x, const float* wprep, float* y, const snb_conv_geom* g, const snb_conv_epilogue* e, void* stream);

/* Same contract as snb_conv_c32 for stride-1 "same" 3x3 (dilated) / 3x3x3 convolutions, on the tcgen05 tensor cores
 * (TF32 operands from shared memory, fp32 accumulators in TMEM; kw folded into N = 96).  passes = 3: error-compensated
 * 3xTF32 split (fp32-grade results); passes = 1: plain single-pass TF32.  wimg from snb_prep_conv_weights_tc.
 * `stats` rows are indexed by this kernel's own tiles: snb_conv_c32_tc_num_tiles(). */
SNB_API int snb_conv_c32_tc(const float* x, const float* wimg, float* y, const snb_conv_geom* g, const snb_conv_epilogue* e,
                    int passes, void* stream);
SNB_API int snb_conv_c32_tc_num_tiles(const snb_conv_geom* g);
/* Diagnostics: same launch, plus per-CTA cycle counters [grid][16] = {loader wait-empty, loader total, MMA wait-full,
 * MMA wait-tmem, MMA total, epilogue wait-tmem-full, epilogue total, epilogue barrier} (clock64 ticks). */
SNB_API int snb_conv_c32_tc_profile(const float* x, const float* wimg, float* y, const snb_conv_geom* g,
                            const snb_conv_epilogue* e, int passes, long long* counters, void* stream);
/* Repack [32][32][kd*3*3] weights into the tensor-core B-operand smem image (hi/lo TF32 split, SWIZZLE_128B K-major,
 * one 24 KB block per (kd,kh) window).  kd = 1 (2-D) or 3 (3-D).  mode 0: forward, 1: data gradient. */
SNB_API int snb_prep_conv_weights_tc(const float* w, float* out, int kd, int mode, void* stream);
SNB_API int snb_conv_weights_tc_floats(int kd);

/* Small-Cin first layers.
 * snb_conv5x5s2_c3: FeatureExtractorNetwork.downsample[0] (stereo_net.py:64-70,81): NCHW image [B,3,H,W] -> [B,OH,OW,32]. */
SNB_API int snb_conv5x5s2_c3(const float* img, const float* w /*[32][3][5][5]*/, const float* bias, float* y,
                     int B, int H, int W, void* stream);
/* snb_refine_in_conv: EdgeAwareRefinement stereo_net.py:105-117 fused: bilinear upsample of coarse [B,h,w] to HxW
 * (align_corners=False), * disp_scale, concat with NCHW rgb, Conv2d(4->32,3x3,p1) + bias.  Writes the upsampled
 * disparity to up [B,H,W] and z [B,H,W,32]; epilogue fields as in snb_conv_c32 (residual unused). */
SNB_API int snb_refine_in_conv(const float* coarse, const float* rgb, const float* w /*[32][4][3][3]*/, float* up, float* z,
                       int B, int h, int w_, int H, int W, float disp_scale, const snb_conv_epilogue* e, void* stream);
SNB_API int snb_refine_in_conv_num_tiles(int B, int H, int W);

/* Channel contraction of a 32->1 convolution: taps[b,(d),t,y,x] = sum_c w[t][c] * x[b,(d),y,x,c].
 * First half of conv3d_alone (stereo_net.py:162,187) and conv2d_out (stereo_net.py:102,121). ntaps = 27 or 9.
 * npos = B*D*H*W positions, plane = H*W. */
SNB_API int snb_conv_c32_taps(const float* x, const float* w /*[1][32][ntaps]*/, float* taps, long long nslices, int plane, int ntaps, void* stream);

/* Second half of conv3d_alone fused with the soft-argmin (stereo_net.py:187-192, DisparityRegression :124-134):
 * cost[b,d,y,x] = bias + sum over the 27 shifted tap planes; p = softmax_d(+cost); pred = sum_d d * p_d.
 * cost_out may be NULL (stereo_net.py:197-198 returns it when requested). */
SNB_API int snb_tapsum_softargmin(const float* taps /*[B,D,27,H,W]*/, const float* bias /*[1]*/, float* cost_out /*[B,D,H,W]*/,
                          float* pred /*[B,H,W]*/, int B, int D, int H, int W, void* stream);

/* Second half of conv2d_out fused with the residual add + ReLU (stereo_net.py:121): out = relu(up + bias + sum taps). */
SNB_API int snb_tapsum_refine_out(const float* taps /*[B,9,H,W]*/, const float* bias, const float* up, float* out,
                          int B, int H, int W, void* stream);

/* out[b,y,x] = mul * bilinear(in[b], size (H,W), align_corners=False) — stereo_net.py:201-202 (mul = 2**k). */
SNB_API int snb_upsample_bilinear(const float* in, float* out, int B, int h, int w, int H, int W, float mul, void* stream);
/* adjoint: din[b,y,x] (+)= mul * sum of bilinear weights * dout ; din must be zeroed by the caller when accumulate=0 is not used */
SNB_API int snb_upsample_bilinear_bwd(const float* dout, float* din, int B, int h, int w, int H, int W, float mul, void* stream);

/* Train-mode BatchNorm (batch statistics, adapt.py:313-314).  Reduces the per-tile partials written by a conv epilogue,
 * updates running stats in place (momentum, unbiased variance) and emits scale = gamma*invstd, shift = beta - mean*scale,
 * plus mean / invstd for the backward pass. */
SNB_API int snb_bn_finalize(const float* stats, int ntiles, long long count, const float* gamma, const float* beta,
                    float* running_mean, float* running_var, float momentum, float eps,
                    float* scale, float* shift, float* mean, float* invstd, void* stream);
/* y = [residual +] LeakyReLU(z*scale + shift) over n positions x 32 channels. */
SNB_API int snb_bn_apply(const float* z, const float* scale, const float* shift, const float* residual, float* y,
                 long long npos, int lrelu, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SNB200_H_ */
