/*
 * crdr_b200 -- C ABI of the B200-native CRDR codec hot path (libcrdr_sm100.so).
 *
 * The reference (iwa-shi/CRDR) is pure PyTorch and has no FFI of its own; every
 * entry point below replaces a *library call site* of the reference, cited as
 * file:line relative to the reference tree.  Plain pointers and sizes only; all
 * device buffers are owned by the caller (the PyTorch caching allocator in the
 * shipped host code).  Every launch is asynchronous on the given stream.
 *
 * Return value: 0 on success, otherwise a crdr_status; a human readable message
 * for the calling thread is available from crdr_last_error().
 *
 * Activation storage inside the path ("planes"): NHWC, fp16.  A tensor is a pair
 * of planes (hi, lo) with  x ~= hi + lo * 2^-11  (22 significant bits); the lo
 * plane is NULL for single-term (fp16) tensors.  Pixel stride (`cs`, in
 * elements) and channel offset (`coff`) let kernels read/write channel ranges of
 * a wider tensor in place (the ChARM support tensor, SURVEY 2.1 "torch.cat").
 */
#ifndef CRDR_B200_H_
#define CRDR_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CRDR_ABI_VERSION 1
#define CRDR_MAX_TAPS 25

typedef enum {
  CRDR_OK = 0,
  CRDR_ERR_BAD_SHAPE = 1,
  CRDR_ERR_MISALIGNED = 2,
  CRDR_ERR_UNSUPPORTED_ARCH = 3,
  CRDR_ERR_CUDA = 4,
  CRDR_ERR_DEVICE_FLAG = 5 /* a kernel raised the device status word (fp16 overflow / pipeline timeout / int16 symbol range) */
} crdr_status;

/* Arithmetic of the contraction. */
typedef enum {
  CRDR_PREC_F16X3 = 0, /* error-compensated 3-term fp16 split, fp32 accumulate: fp32-class */
  CRDR_PREC_F16X1 = 1  /* plain fp16 operands, fp32 accumulate */
} crdr_precision;

/* Which kernel executes the contraction. */
typedef enum {
  CRDR_ENGINE_TCGEN05 = 0,       /* tcgen05.mma + TMEM accumulators, weights by TMA (the product path) */
  CRDR_ENGINE_TCGEN05_NOTMA = 1, /* same, weights staged by cp.async (bring-up / bisect) */
  CRDR_ENGINE_SIMT = 2           /* scalar fp32 CUDA-core kernel over the same operands (on-device cross-check) */
} crdr_engine;

typedef enum {
  CRDR_EPI_NONE = 0,
  CRDR_EPI_RESIDUAL = 1,  /* v += res                                   (BaseBlock / NLAMResBlock skip) */
  CRDR_EPI_GATE = 2,      /* v = res + trunk * sigmoid(v)               (ChengNLAM.forward, cheng_nlam.py:23-29) */
  CRDR_EPI_HALF_TANH = 3  /* v = res + 0.5 * tanh(v)                    (LRP, minnen20_charm_context_model.py:127-131) */
} crdr_epilogue_mode;

typedef struct {
  const void* hi; /* fp16 plane, NHWC */
  const void* lo; /* fp16 plane holding (x - hi) * 2^11, or NULL */
  int32_t cs;     /* elements per pixel */
  int32_t coff;   /* first channel */
} crdr_planes;

/*
 * One convolution-shaped contraction  out[m, co] = sum_{t, ci} in[pix(m) + tap_t, ci] * W[co, t*Cin + ci]
 * followed by the fused epilogue
 *     v = acc + bias[co];  if relu: v = max(v, 0);  v += add_vec[co];
 *     mode (residual / gate / half-tanh);  v = v * scale[co] + shift[co]
 * Replaces nn.Conv2d / nn.ConvTranspose2d (+ the element-wise ops that follow) at:
 *   elic_autoencoder.py:42-52 (g_a convs), elic_layers.py:14-36 (up_conv, BaseBlock),
 *   elic_interpca_beta_cond_autoencoder.py:42-66 (BetaCondBaseBlock), cheng_nlam.py:5-47,
 *   minnen20_hyperprior.py:17-19,50-52, minnen20_charm_context_model.py:26-38,
 *   interp_channel_attention.py:54-73 (gain/bias folded into scale/shift).
 *
 * Geometry: the GEMM rows m enumerate a base grid N x Hb x Wb.  Input pixel for tap t is
 * (b_h * in_stride + dh[t], b_w * in_stride + dw[t]) (zero outside the image); the output pixel is
 * (b_h * out_stride + out_ph, b_w * out_stride + out_pw).  A stride-2 Conv2d has in_stride 2; a stride-2
 * ConvTranspose2d is four launches (one per output phase) with out_stride 2.
 * Input channels are the concatenation of up to two channel ranges of the same NHWC tensor.
 */
typedef struct {
  /* input */
  crdr_planes in;
  int32_t n, hin, win;
  int32_t seg0_off, seg0_len, seg1_off, seg1_len; /* Cin = seg0_len + seg1_len, each a multiple of 8 */
  /* base grid and taps */
  int32_t hb, wb, in_stride;
  int32_t ntaps;
  int8_t dh[CRDR_MAX_TAPS], dw[CRDR_MAX_TAPS];
  /* packed weights [cout_pad][k_pad] fp16 (hi, lo*2^11), k = t*Cin + ci, k_pad % 64 == 0 */
  const void* w_hi;
  const void* w_lo;
  int32_t k_pad, cout_pad, cout;
  int32_t tile_n; /* N tile (multiple of 16, <= 256, divides cout_pad) */
  /* K order of the packed weights.  0: k = t*Cin + ci (operands gathered per tap; any stride).
   * 1: 64-channel blocks cb of each input range, k = (cb*ntaps + t)*64 + ci%64, zero padded -- selects the
   *    halo-patch engine (one TMA box per channel block serves every tap); needs in_stride == 1. */
  int32_t k_order;
  /* output */
  int32_t hout, wout, out_stride, out_ph, out_pw;
  crdr_planes out;  /* may have hi == NULL */
  float* out_f32;   /* optional NHWC fp32 output */
  int32_t out_f32_cs, out_f32_coff;
  /* epilogue */
  const float* bias;
  int32_t relu;
  const float* add_vec;
  int32_t mode;
  crdr_planes res;      /* residual as planes ... */
  const float* res_f32; /* ... or as NHWC fp32 (takes precedence) */
  int32_t res_f32_cs, res_f32_coff;
  crdr_planes trunk;
  const float* scale;
  const float* shift;
  /* arithmetic */
  int32_t precision; /* crdr_precision */
  int32_t engine;    /* crdr_engine */
} crdr_conv_desc;

int crdr_abi_version(void);
const char* crdr_last_error(void);

/* Device status word (one per device, lives in device memory owned by the library:
 * bit0 fp16 overflow in an epilogue, bit1 pipeline wait timeout, bit2 a symbol outside the compact int16 range).  `crdr_status_read`
 * synchronises the stream. */
int crdr_status_reset(void* stream);
int crdr_status_read(uint32_t* flags, void* stream);
/* Asynchronous copy of the status word into caller-owned (page-locked) host memory, no synchronisation; and an
 * in-stream clear of selected bits (a handled condition must not poison later calls). */
int crdr_status_peek_async(uint32_t* host_flags, void* stream);
int crdr_status_clear_bits(uint32_t bits, void* stream);

int crdr_conv2d(const crdr_conv_desc* d, void* stream);

/*
 * Fused tail of a residual bottleneck (F16X1 tensors), one launch instead of two crdr_conv2d calls:
 *     t2  = relu(conv3x3(t1; w2) + bias2) + add2              (never leaves the SM)
 *     out = (conv1x1(t2; w3) + bias3 + add3 + res) * scale + shift
 * Replaces the second and third convolution (and the element-wise ops between / after them) of
 *   BaseBlock.forward (elic_layers.py:23-36), NLAMResBlock.forward (cheng_nlam.py:32-47) and
 *   BetaCondBaseBlock.forward (elic_interpca_beta_cond_autoencoder.py:52-66; add2 / add3 = proj_2(c) / proj_3(c)).
 * The result is bit-identical to the two separate launches (same fp16 mid tensor, same K order, same epilogue).
 * in: t1 [n, h, w, mid] (the first 1x1's output); res / out: [n, h, w, cout] channel ranges of NHWC planes.
 * w2: packed [mid_pad][k2_pad] in k_order 1 (k = (cb*9 + tap)*64 + ci%64, taps row-major from (-1,-1)), w3: packed
 * [cout_pad][k3_pad] (k = cb*64 + ci%64) -- the matrices crdr_conv2d takes for the same layers.
 * Limits: mid % 32 == 0, mid <= 128, cout % 32 == 0, cout <= 256 (512 TMEM columns hold two 3x3 and one 1x1 accumulator).
 */
typedef struct {
  crdr_planes in;   /* t1 */
  int32_t n, h, w;
  int32_t mid, cout;
  const void* w2;
  int32_t k2_pad, mid_pad;
  const void* w3;
  int32_t k3_pad, cout_pad;
  const float* bias2;
  const float* add2; /* may be NULL */
  const float* bias3;
  const float* add3; /* may be NULL */
  crdr_planes res;   /* skip connection x */
  crdr_planes out;
  const float* scale; /* may be NULL */
  const float* shift; /* may be NULL */
  int32_t precision;  /* CRDR_PREC_F16X1 */
} crdr_bottleneck_desc;
int crdr_bottleneck_bc(const crdr_bottleneck_desc* d, void* stream);

/* fp32 NHWC [m][c] -> planes, v = x*scale[c] + shift[c] (InterpChAtt on a tensor that has no producing
 * conv: decoder input, elic_interpca_beta_cond_autoencoder.py:153-156).  scale/shift may be NULL. */
int crdr_affine_to_planes(const float* x, int32_t x_cs, int32_t x_coff, int64_t m, int32_t c,
                          const float* scale, const float* shift, crdr_planes out, void* stream);

/* Image pre-processing: NCHW fp32 in [-1,1] (n,3,h,w) -> reflect-pad bottom/right to (hp,wp) -> NHWC planes with
 * 8 channels (3 real + 5 zero).  Replaces BaseModel.data_preprocess/_pad_image (base_model.py:35-43,145-152). */
int crdr_image_to_planes(const float* img, int32_t n, int32_t h, int32_t w, int32_t hp, int32_t wp,
                         crdr_planes out, void* stream);
/* Same pre-processing fused with the im2col of the first analysis layer (ElicEncoder.conv1: 5x5, stride 2, padding 2,
 * elic_autoencoder.py:42): output NHWC planes (n, hp/2, wp/2, 80..128) whose channel (kh*5+kw)*3+c is the padded image at
 * (2i+kh-2, 2j+kw-2, c) (zero outside), channels 75.. zero.  conv1 then runs as a 1x1 convolution over these channels. */
int crdr_image_to_patches(const float* img, int32_t n, int32_t h, int32_t w, int32_t hp, int32_t wp,
                          crdr_planes out, void* stream);
/* Post-processing: NHWC fp32 (n,hp,wp,cs) first 3 channels -> crop (h,w) -> clamp(-1,1) -> NCHW fp32.
 * Replaces data_postprocess/_crop_image (base_model.py:45-57,165-167). */
int crdr_planes_to_image(const float* x, int32_t x_cs, int32_t n, int32_t hp, int32_t wp, int32_t h, int32_t w,
                         float* img, void* stream);

/* Same post-processing for the final stride-2 ConvTranspose2d (g_s conv4, 256 -> 3) evaluated as ONE stride-1
 * launch whose 12 output channels are (phase_y*2 + phase_x)*3 + c on the (hb, wb) input grid: pixel shuffle +
 * crop + clamp -> NCHW fp32 (h <= 2*hb, w <= 2*wb). */
int crdr_phases_to_image(const float* x, int32_t x_cs, int32_t n, int32_t hb, int32_t wb, int32_t h, int32_t w,
                         float* img, void* stream);
/* Same with the clamp optional: training-mode forward returns the unclamped reconstruction
 * (beta_cond_interpca_hyperprior_charm_model.py:54-56 clamps only when not is_train). */
int crdr_phases_to_image_ex(const float* x, int32_t x_cs, int32_t n, int32_t hb, int32_t wb, int32_t h, int32_t w,
                            float* img, int32_t clamp, void* stream);

/* uint8 image boundary.  Input: NCHW uint8 RGB as PIL / cv2 deliver it; the kernel applies the reference's
 * ToTensor + Normalize(0.5, 0.5) arithmetic (scripts/compress.py:54-57: (u/255 - 0.5)/0.5 in fp32) before the reflect
 * pad + im2col of crdr_image_to_patches.  Output: crop + clamp(-1,1) followed by the reference's PNG conversion
 * (img_utils.py:30-42,66-76: ((x+1)/2*255) in fp32, astype(uint8) = truncation), NCHW uint8. */
int crdr_image_u8_to_patches(const uint8_t* img, int32_t n, int32_t h, int32_t w, int32_t hp, int32_t wp,
                             crdr_planes out, void* stream);
int crdr_phases_to_image_u8(const float* x, int32_t x_cs, int32_t n, int32_t hb, int32_t wb, int32_t h, int32_t w,
                            uint8_t* img, void* stream);

/* NHWC fp32 [n][hw][cs](coff..coff+c) -> NCHW fp32 [n][c][hw]  (API-facing tensors). */
int crdr_nhwc_to_nchw(const float* x, int32_t x_cs, int32_t x_coff, int32_t n, int32_t hw, int32_t c,
                      float* out, void* stream);

/*
 * GaussianConditional, evaluation mode, one channel slice (CompressAI GaussianConditional.forward /
 * quantize / _likelihood / build_indexes as called at minnen20_charm_context_model.py:118,123,170,186).
 *   q = rint(y - mu) (half-to-even);  yq = q + mu;  s = max(sigma, bound)
 *   L = Phi((.5-|q|)/s) - Phi((-.5-|q|)/s), Phi(x) = .5*erfc(-x/sqrt2);  L = max(L, 1e-9)
 *   index = (ntable-1) - #{k < ntable-1 : s <= table[k]}
 * Inputs NHWC fp32: y [m][y_cs] at y_coff; mu, sigma [m][ms_cs] at mu_coff / sigma_coff.  c channels (<= 32 per
 * call is typical).  Outputs: yq as planes (+ optional NHWC fp32), symbols / indexes int32 and likelihood fp32
 * in NCHW order [n][c_total][hw] at channel offset nchw_coff.
 */
typedef struct {
  const float* y;
  int32_t y_cs, y_coff;
  const float* mu;
  const float* sigma;
  int32_t ms_cs, mu_coff, sigma_coff;
  int32_t n, hw, c;
  float scale_bound;
  const float* scale_table;
  int32_t ntable;
  crdr_planes yq_planes;
  float* yq_f32;
  int32_t yq_f32_cs, yq_f32_coff;
  int32_t* symbols;   /* NCHW, may be NULL */
  int32_t* indexes;   /* NCHW, may be NULL */
  float* likelihood;  /* NCHW, may be NULL */
  int32_t c_total, nchw_coff;
  /* Compact copies for the host range coder (same NCHW order and offsets; a quarter of the PCIe and host-memory
   * traffic of the int32 tensors the reference moves with .cpu() / .tolist()): symbols as int16 (values outside the
   * int16 range saturate and raise device status bit 2, the caller then falls back to the int32 tensor), table
   * indexes as uint8.  crdr_gauss_dequantize reads `symbols16` when `symbols` is NULL.  May be NULL. */
  int16_t* symbols16;
  uint8_t* indexes8;
  /* Training-mode likelihood (SteGaussianMeanScaleConditional.forward(is_train=True), ste_gaussian_conditional.py:20-27;
   * CompressAI GaussianConditional.forward(training=True)): the likelihood of y + u, u ~ U(-1/2, 1/2), under N(mu, sigma).
   * The noise is an INPUT (NCHW fp32 like `likelihood`) so that a CPU oracle can replay it; the quantised outputs above
   * are the straight-through forward values (round(y - mu) + mu) and the quantised likelihood (q_likelihoods).  crdr_gauss_quantize
   * only; both NULL in evaluation mode. */
  const float* noise;
  float* likelihood_noisy;
} crdr_gauss_desc;
int crdr_gauss_quantize(const crdr_gauss_desc* d, void* stream);

/* Decoder side of the same slice: indexes from sigma only (build_indexes, :221) ... */
int crdr_gauss_indexes(const crdr_gauss_desc* d, void* stream);
/* ... and yq = symbol + mu (GaussianConditional.dequantize, :226); `symbols` is the NCHW int32 input. */
int crdr_gauss_dequantize(const crdr_gauss_desc* d, void* stream);

/*
 * EntropyBottleneck, evaluation mode (CompressAI EntropyBottleneck.forward/_likelihood/_logits_cumulative as
 * called at beta_cond_interpca_hyperprior_charm_model.py:43,58,95).  z: NHWC fp32 [m][c].
 * params: per channel 58 floats = softplus(matrix0)[3] bias0[3] tanh(factor0)[3] | softplus(matrix1)[9] bias1[3]
 * tanh(factor1)[3] | (same for 2, 3) | softplus(matrix4)[3] bias4[1];  medians[c].
 * Outputs: z_hat planes (NHWC), symbols int32 / z_hat fp32 / likelihood fp32 in NCHW.
 */
typedef struct {
  const float* z;
  int32_t z_cs;
  int32_t n, hw, c;
  const float* params;
  const float* medians;
  crdr_planes zhat_planes;
  int32_t* symbols;
  float* zhat_nchw;
  float* likelihood;
  /* Training mode (SteEntropyBottleneck.forward(is_train=True), entropy_bottleneck.py:23-30): likelihood of z + u with the
   * uniform noise u given as an NCHW fp32 input; z_hat stays round(z - median) + median.  crdr_eb_quantize only; may be NULL. */
  const float* noise;
  float* likelihood_noisy;
} crdr_eb_desc;
int crdr_eb_quantize(const crdr_eb_desc* d, void* stream);
/* Decoder: symbols (NCHW int32) -> z_hat planes + NCHW fp32. */
int crdr_eb_dequantize(const crdr_eb_desc* d, void* stream);

/* bits[i] = -sum_j log2(L[i][j]) over `per` contiguous elements, fixed summation order
 * (hyperprior_model.py:80-85 likelihood_to_bit). */
int crdr_bits_from_likelihood(const float* lik, int32_t n, int64_t per, float* bits, void* stream);

/* max |x| over `count` floats -> out[0] (header byte, codec_utils.py:88). */
int crdr_max_abs(const float* x, int64_t count, float* out, void* stream);
/* out[i] = max |x[i*per .. (i+1)*per)| for n consecutive images in one launch. */
int crdr_max_abs_batch(const float* x, int32_t n, int64_t per, float* out, void* stream);

/* =====================================================================================================================
 * Training step (backward).  The reference's backward is torch.autograd behind `l_total.backward()`
 * (rate_distortion_trainer.py:84) followed by clip_grad_norm_ / Adam (:85-87); the entry points below are the
 * contractions and element-wise steps that autograd would dispatch for this model.  Activation gradients are single
 * fp16 planes multiplied by the caller's loss scale; parameter gradients are fp32 in the parameter's own layout.
 * ===================================================================================================================== */

/* dgrad: the gradient w.r.t. a convolution's input is itself a convolution of the output gradient with the transposed
 * (and, for stride 1, flipped) weights -- the same kernel and descriptor as crdr_conv2d, fed with matrices packed by
 * crdr_pack_weights through a dgrad index map (torch: conv backward-data at every nn.Conv2d / nn.ConvTranspose2d). */
int crdr_conv_dgrad(const crdr_conv_desc* d, void* stream);

/* wgrad:  G[t][a][b] = sum over pixels p = (n, h, w) of the small grid   S[p][a] * B[n, h*stride + dh[t], w*stride + dw[t]][b]
 *         out[t*st + a*sa + b*sb] (+)= scale * G[t][a][b]
 * nn.Conv2d:          S = dY, B = X (input), stride = conv stride      -> weight.grad [co][ci][kh][kw]
 * nn.ConvTranspose2d: S = X (input), B = dY, stride = deconv stride    -> weight.grad [ci][co][kh][kw]
 * S / B: channel ranges [coff, coff + ca|cb) of NHWC fp16 planes (hi plane only).  Deterministic (fixed-order split-K). */
typedef struct {
  crdr_planes s;
  int32_t ca;
  crdr_planes b;
  int32_t cb;
  int32_t n, hs, ws, hb, wb, stride;
  int32_t ntaps;
  int8_t dh[CRDR_MAX_TAPS], dw[CRDR_MAX_TAPS];
  float* out;
  int64_t sa, sb, st;
  float scale;
  int32_t accumulate;
  void* workspace;        /* device scratch for the split-K partial sums */
  size_t workspace_bytes; /* >= crdr_conv_wgrad_workspace(d) for full parallelism; smaller values reduce the split count */
} crdr_wgrad_desc;
size_t crdr_conv_wgrad_workspace(const crdr_wgrad_desc* d);
int crdr_conv_wgrad(const crdr_wgrad_desc* d, void* stream);

/* packed[i] = split_fp16(master[map[i]]) (map[i] < 0: zero): re-packs a parameter into a K-major tensor-core matrix
 * (forward, dgrad or phase-packed form) after every optimiser step.  lo may be NULL (single-plane matrices). */
int crdr_pack_weights(const float* master, const int32_t* map, int64_t count, void* hi, void* lo, void* stream);
/* The same for a whole table of matrices in one launch; `jobs` is DEVICE memory (njobs entries, <= 65535). */
typedef struct {
  const float* master;
  const int32_t* map;
  int64_t count;
  void* hi;
  void* lo;
} crdr_pack_job;
int crdr_pack_weights_multi(const crdr_pack_job* jobs, int32_t njobs, void* stream);

/* Backward of the fused convolution epilogue  out = ([relu](acc + bias) [+ res | res + 0.5 tanh(.)]) * scale + shift:
 *   g1 = g * scale;  dres += g1;  dv = relu ? g1 * (out > 0) : (half-tanh ? g1 * 0.5 * (1 - (2 (f32_out - f32_res))^2) : g1)
 * and per-channel sums over the pixels, written per row block to partial[blocks][3][c]:
 *   0: sum dv (bias.grad)   1: sum g (InterpChAtt bias path; always computed)   2: sum g * (out - shift) / scale (InterpChAtt weight path) */
typedef struct {
  crdr_planes g;
  crdr_planes out;
  int64_t m;
  int32_t c;
  int32_t relu;
  const float* scale;
  const float* shift;
  const float* f32_out; /* half-tanh (LRP): the layer's fp32 result and its residual operand, NHWC */
  const float* f32_res;
  int32_t f32_cs, f32_coff;
  crdr_planes dv;       /* may have hi == NULL (sums only) */
  crdr_planes dres;     /* may have hi == NULL */
  float* partial;       /* may be NULL */
  int32_t blocks;       /* row blocks = grid size (1..4096) */
  const float* add_vec; /* forward added this per-channel vector AFTER the ReLU (beta conditioning, single-plane tensors): the
                           mask is out != fp16(add_vec) instead of out > 0; its gradient is sum 1 (no gain) or sum 0 (no ReLU) */
  float leaky_slope;    /* with relu = 1: LeakyReLU of this negative slope (0: plain ReLU) */
} crdr_epi_bwd_desc;
int crdr_epilogue_backward(const crdr_epi_bwd_desc* d, void* stream);
/* out[c] (+)= scale * sum_b partial[b][which][c], blocks in ascending order */
int crdr_colsum_finish(const float* partial, int32_t blocks, int32_t nsums, int32_t which, int32_t c, float* out, float scale,
                       int32_t accumulate, void* stream);

/* ChengNLAM gate as its own step (training keeps the logits): out = (x + t * sigmoid(a)) * scale + shift (cheng_nlam.py:23-29).
 * Backward: dx += g * scale; dt = g * scale * sig; da = g * scale * t * sig * (1 - sig); partial[blocks][2][c]: sum g | sum g * pre */
typedef struct {
  crdr_planes x, t, a;
  int64_t m;
  int32_t c;
  const float* scale;
  const float* shift;
  crdr_planes out;    /* forward */
  float* out_f32;
  int32_t out_f32_cs, out_f32_coff;
  crdr_planes g, dx, dt, da; /* backward */
  float* partial;
  int32_t blocks;
} crdr_gate_desc;
int crdr_gate_forward(const crdr_gate_desc* d, void* stream);
int crdr_gate_backward(const crdr_gate_desc* d, void* stream);

/* Rate term of one slice group: loss += -coef * ln max(L(y + noise; mu, sigma), lik_bound)  (GaussianConditional._likelihood +
 * LowerBound; ste_gaussian_conditional.py:20-27, minnen20_charm_context_model.py:118).  dy = dL/dy + gpre (the
 * straight-through path of the quantised slice), dmu, dsigma. */
typedef struct {
  const float* y;
  int32_t y_cs, y_coff;
  const float* noise; /* NCHW [n][c_total][hw], this group's channels start at nchw_coff */
  const float* ms;    /* NHWC fp32: mu at mu_coff, sigma at sigma_coff */
  int32_t ms_cs, mu_coff, sigma_coff;
  int32_t n, hw, c, c_total, nchw_coff;
  float scale_bound, lik_bound, coef;
  crdr_planes gpre;   /* may have hi == NULL */
  crdr_planes dy, dmu, dsigma;
  const float* coef_scale; /* optional device scalar multiplied into coef (the HiFiC rate weight decided on the device) */
} crdr_gauss_bwd_desc;
int crdr_gauss_backward(const crdr_gauss_bwd_desc* d, void* stream);

/* Gradient of coef/2 * sum (fake - real)^2 on the phase-packed output of the last up-convolution (distortion_loss.py:40-46):
 * fake [n, hb, wb, >=16] fp32 (channel (ph*2+pw)*3 + c), real [n, 3, h, w] fp32 -> g [n, hb, wb, g_cs] fp16. */
int crdr_mse_backward(const float* fake, int32_t fake_cs, const float* real, int32_t n, int32_t hb, int32_t wb, int32_t h,
                      int32_t w, float coef, void* g, int32_t g_cs, void* stream);

/* nn.LeakyReLU(slope) in place on channels [coff, coff + c) of a single fp16 plane (clic21_gvae_discriminator.py:12-25;
 * its backward is crdr_epilogue_backward with relu = 1 and leaky_slope). */
int crdr_leaky_relu(crdr_planes x, int64_t m, int32_t c, float slope, void* stream);
/* g[n, a, b, (ph*2+pw)*3 + c] += scale * g8[n, 2a+ph, 2b+pw, c]: the gradient a discriminator sends back to the image
 * (8-channel NHWC fp16 planes, 3 used) added to the phase-packed gradient of the reconstruction (see crdr_mse_backward). */
int crdr_planes_grad_to_phases(const void* g8, int32_t g8_cs, int32_t n, int32_t hb, int32_t wb, float scale, void* g,
                               int32_t g_cs, void* stream);

/* torch.optim.Adam step on flat fp32 buffers; the gradient is multiplied by gscale (* gscale_ptr[0] when given: the
 * clip coefficient computed on the device).  step >= 1, or `hyper` = device float[4] {lr, 1 - beta1^step, sqrt(1 - beta2^step),
 * skip} overriding lr / step (a CUDA-graph replay reads the schedule from device memory); skip != 0 leaves everything
 * untouched (the reference skips the update when the loss is nan / inf / huge, base_trainer.py:228-238). */
int crdr_adam_step(float* p, const float* g, float* m, float* v, int64_t count, float lr, float beta1, float beta2, float eps,
                   int32_t step, const float* gscale_ptr, float gscale, const float* hyper, void* stream);
/* out[0] = sum x^2 (fixed order); partial: 1024 floats of scratch. */
int crdr_sum_squares(const float* x, int64_t count, float* partial, float* out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* CRDR_B200_H_ */
