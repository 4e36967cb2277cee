// extern "C" surface of libcrdr_sm100.so (declared in include/crdr_b200.h).
#include <cstdarg>
#include <cstdio>
#include <mutex>

#include "common.cuh"

namespace crdr {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("%s: launch failed: %s", what, cudaGetErrorString(e));
    return e == cudaErrorNoKernelImageForDevice ? CRDR_ERR_UNSUPPORTED_ARCH : CRDR_ERR_CUDA;
  }
  return CRDR_OK;
}

static std::mutex g_status_mutex;
static uint32_t* g_status[64] = {nullptr};

uint32_t* device_status_word() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) {
    set_error("cannot query the current CUDA device");
    return nullptr;
  }
  std::lock_guard<std::mutex> lk(g_status_mutex);
  if (!g_status[dev]) {
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, dev) != cudaSuccess || prop.major != 10) {
      set_error("libcrdr_sm100 needs an sm_100 class GPU (found compute %d.%d)", prop.major, prop.minor);
      return nullptr;
    }
    uint32_t* p = nullptr;
    // 256 bytes of status / cycle counters + a 64 KB event-trace area (bring-up: CRDR_CONV_TRACE=2)
    if (cudaMalloc(&p, 256 + 65536) != cudaSuccess || cudaMemset(p, 0, 256 + 65536) != cudaSuccess) {
      set_error("cannot allocate the device status word: %s", cudaGetErrorString(cudaGetLastError()));
      return nullptr;
    }
    g_status[dev] = p;
  }
  return g_status[dev];
}

int conv2d_launch(const crdr_conv_desc* d, cudaStream_t stream);
void conv_set_lean(int enabled, int swizzle);
int bottleneck_bc_launch(const crdr_bottleneck_desc* d, cudaStream_t stream);
int gauss_launch(const crdr_gauss_desc* d, int mode, cudaStream_t st);
int eb_launch(const crdr_eb_desc* d, int dequant, cudaStream_t st);
int nhwc_to_nchw_launch(const float* x, int x_cs, int x_coff, int n, int hw, int c, float* out, cudaStream_t st);
int affine_to_planes_launch(const float* x, int x_cs, int x_coff, int64_t m, int c, const float* scale,
                            const float* shift, crdr_planes out, cudaStream_t st);
int image_to_planes_launch(const float* img, int n, int h, int w, int hp, int wp, crdr_planes out, cudaStream_t st);
int image_to_patches_launch(const float* img, int n, int h, int w, int hp, int wp, crdr_planes out, cudaStream_t st);
int planes_to_image_launch(const float* x, int x_cs, int n, int hp, int wp, int h, int w, float* img, cudaStream_t st);
int phases_to_image_launch(const float* x, int x_cs, int n, int hb, int wb, int h, int w, float* img, int clamp, cudaStream_t st);
int image_u8_to_patches_launch(const uint8_t* img, int n, int h, int w, int hp, int wp, crdr_planes out, cudaStream_t st);
int phases_to_image_u8_launch(const float* x, int x_cs, int n, int hb, int wb, int h, int w, uint8_t* img, cudaStream_t st);
int max_abs_batch_launch(const float* x, int n, int64_t per, float* out, cudaStream_t st);
int bits_launch(const float* lik, int n, int64_t per, float* bits, cudaStream_t st);
int status_clear_bits_launch(uint32_t* status, uint32_t bits, cudaStream_t st);
int max_abs_launch(const float* x, int64_t count, float* out, cudaStream_t st);
size_t wgrad_workspace_bytes(const crdr_wgrad_desc* d);
int wgrad_launch(const crdr_wgrad_desc* d, cudaStream_t st);
int pack_weights_launch(const float* master, const int32_t* map, int64_t count, void* hi, void* lo, cudaStream_t st);
int pack_weights_multi_launch(const crdr_pack_job* jobs, int njobs, cudaStream_t st);
int epi_bwd_launch(const crdr_epi_bwd_desc* d, cudaStream_t st);
int colsum_finish_launch(const float* partial, int blocks, int nsums, int which, int c, float* out, float scale, int accumulate,
                         cudaStream_t st);
int gate_launch(const crdr_gate_desc* d, int backward, cudaStream_t st);
int gauss_bwd_launch(const crdr_gauss_bwd_desc* d, cudaStream_t st);
int mse_bwd_launch(const float* fake, int fake_cs, const float* real, int n, int hb, int wb, int h, int w, float coef, void* g,
                   int g_cs, cudaStream_t st);
int adam_launch(float* p, const float* g, float* m, float* v, int64_t count, float lr, float b1, float b2, float eps, int step,
                const float* gscale_ptr, float gscale, const float* hyper, cudaStream_t st);
int sumsq_launch(const float* x, int64_t count, float* partial1024, float* out, cudaStream_t st);
int leaky_relu_launch(crdr_planes x, int64_t m, int c, float slope, cudaStream_t st);
int planes_grad_to_phases_launch(const void* g8, int g8_cs, int n, int hb, int wb, float scale, void* g, int g_cs, cudaStream_t st);

}  // namespace crdr

using namespace crdr;

extern "C" {

int crdr_abi_version(void) { return CRDR_ABI_VERSION; }
const char* crdr_last_error(void) { return g_err; }

int crdr_status_reset(void* stream) {
  uint32_t* p = device_status_word();
  if (!p) return CRDR_ERR_CUDA;
  cudaError_t e = cudaMemsetAsync(p, 0, 4, (cudaStream_t)stream);
  if (e != cudaSuccess) { set_error("status_reset: %s", cudaGetErrorString(e)); return CRDR_ERR_CUDA; }
  return CRDR_OK;
}

int crdr_status_read(uint32_t* flags, void* stream) {
  uint32_t* p = device_status_word();
  if (!p || !flags) return CRDR_ERR_CUDA;
  cudaError_t e = cudaMemcpyAsync(flags, p, 4, cudaMemcpyDeviceToHost, (cudaStream_t)stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize((cudaStream_t)stream);
  if (e != cudaSuccess) { set_error("status_read: %s", cudaGetErrorString(e)); return CRDR_ERR_CUDA; }
  if (*flags) {
    set_error("device status flags 0x%x (%s%s%s)", *flags, (*flags & kFlagOverflow) ? "fp16 overflow " : "",
              (*flags & kFlagTimeout) ? "pipeline timeout " : "", (*flags & kFlagSymRange) ? "int16 symbol range" : "");
    return CRDR_ERR_DEVICE_FLAG;
  }
  return CRDR_OK;
}

int crdr_status_peek_async(uint32_t* host_flags, void* stream) {
  uint32_t* p = device_status_word();
  if (!p || !host_flags) return CRDR_ERR_CUDA;
  cudaError_t e = cudaMemcpyAsync(host_flags, p, 4, cudaMemcpyDeviceToHost, (cudaStream_t)stream);
  if (e != cudaSuccess) { set_error("status_peek: %s", cudaGetErrorString(e)); return CRDR_ERR_CUDA; }
  return CRDR_OK;
}

int crdr_status_clear_bits(uint32_t bits, void* stream) {
  uint32_t* p = device_status_word();
  if (!p) return CRDR_ERR_CUDA;
  return status_clear_bits_launch(p, bits, (cudaStream_t)stream);
}

/* bring-up aid (not part of the documented ABI): MMA-thread cycle counters of CTA 0 of the last traced launch */
int crdr_debug_counters(unsigned long long* out10) {
  uint32_t* p = device_status_word();
  if (!p) return CRDR_ERR_CUDA;
  cudaDeviceSynchronize();
  return cudaMemcpy(out10, p + 16, 80, cudaMemcpyDeviceToHost) == cudaSuccess ? CRDR_OK : CRDR_ERR_CUDA;
}

/* bring-up aid: event trace of CTA 0 of the last CRDR_CONV_TRACE=2 launch: three regions of 2600 (tag, clock) records;
 * unused records are zero */
int crdr_debug_events(uint32_t* out, int32_t max_words) {
  uint32_t* p = device_status_word();
  if (!p) return CRDR_ERR_CUDA;
  cudaDeviceSynchronize();
  const size_t bytes = (size_t)(max_words < 16384 ? max_words : 16384) * 4;
  if (cudaMemcpy(out, p + 64, bytes, cudaMemcpyDeviceToHost) != cudaSuccess) return CRDR_ERR_CUDA;
  cudaMemset(p + 64, 0, 65536);
  return CRDR_OK;
}

/* bring-up / test aid (not part of the documented ABI): select the convolution epilogue at run time.
 * enabled: 1 = TMA-in / TMA-out LEAN epilogue where eligible (default), 0 = staged epilogue everywhere, -1 = keep;
 * swizzle: 1 = SWIZZLE_64B staging units (default), 0 = linear, -1 = keep.  Both epilogues produce identical bits. */
void crdr_debug_conv_epilogue(int32_t enabled, int32_t swizzle) { conv_set_lean(enabled, swizzle); }

int crdr_conv2d(const crdr_conv_desc* d, void* stream) {
  if (!d) { set_error("conv2d: null descriptor"); return CRDR_ERR_BAD_SHAPE; }
  return conv2d_launch(d, (cudaStream_t)stream);
}

int crdr_bottleneck_bc(const crdr_bottleneck_desc* d, void* stream) {
  if (!d) { set_error("bottleneck: null descriptor"); return CRDR_ERR_BAD_SHAPE; }
  return bottleneck_bc_launch(d, (cudaStream_t)stream);
}

int crdr_affine_to_planes(const float* x, int32_t x_cs, int32_t x_coff, int64_t m, int32_t c, const float* scale,
                          const float* shift, crdr_planes out, void* stream) {
  return affine_to_planes_launch(x, x_cs, x_coff, m, c, scale, shift, out, (cudaStream_t)stream);
}

int crdr_image_to_planes(const float* img, int32_t n, int32_t h, int32_t w, int32_t hp, int32_t wp, crdr_planes out,
                         void* stream) {
  return image_to_planes_launch(img, n, h, w, hp, wp, out, (cudaStream_t)stream);
}

int crdr_image_to_patches(const float* img, int32_t n, int32_t h, int32_t w, int32_t hp, int32_t wp, crdr_planes out,
                          void* stream) {
  return image_to_patches_launch(img, n, h, w, hp, wp, out, (cudaStream_t)stream);
}

int crdr_planes_to_image(const float* x, int32_t x_cs, int32_t n, int32_t hp, int32_t wp, int32_t h, int32_t w,
                         float* img, void* stream) {
  return planes_to_image_launch(x, x_cs, n, hp, wp, h, w, img, (cudaStream_t)stream);
}

int crdr_phases_to_image(const float* x, int32_t x_cs, int32_t n, int32_t hb, int32_t wb, int32_t h, int32_t w,
                         float* img, void* stream) {
  return phases_to_image_launch(x, x_cs, n, hb, wb, h, w, img, 1, (cudaStream_t)stream);
}

int crdr_phases_to_image_ex(const float* x, int32_t x_cs, int32_t n, int32_t hb, int32_t wb, int32_t h, int32_t w,
                            float* img, int32_t clamp, void* stream) {
  return phases_to_image_launch(x, x_cs, n, hb, wb, h, w, img, clamp, (cudaStream_t)stream);
}

int crdr_image_u8_to_patches(const uint8_t* img, int32_t n, int32_t h, int32_t w, int32_t hp, int32_t wp, crdr_planes out,
                             void* stream) {
  return image_u8_to_patches_launch(img, n, h, w, hp, wp, out, (cudaStream_t)stream);
}

int crdr_phases_to_image_u8(const float* x, int32_t x_cs, int32_t n, int32_t hb, int32_t wb, int32_t h, int32_t w,
                            uint8_t* img, void* stream) {
  return phases_to_image_u8_launch(x, x_cs, n, hb, wb, h, w, img, (cudaStream_t)stream);
}

int crdr_max_abs_batch(const float* x, int32_t n, int64_t per, float* out, void* stream) {
  return max_abs_batch_launch(x, n, per, out, (cudaStream_t)stream);
}

int crdr_nhwc_to_nchw(const float* x, int32_t x_cs, int32_t x_coff, int32_t n, int32_t hw, int32_t c, float* out,
                      void* stream) {
  return nhwc_to_nchw_launch(x, x_cs, x_coff, n, hw, c, out, (cudaStream_t)stream);
}

int crdr_gauss_quantize(const crdr_gauss_desc* d, void* stream) { return gauss_launch(d, 0, (cudaStream_t)stream); }
int crdr_gauss_indexes(const crdr_gauss_desc* d, void* stream) { return gauss_launch(d, 1, (cudaStream_t)stream); }
int crdr_gauss_dequantize(const crdr_gauss_desc* d, void* stream) { return gauss_launch(d, 2, (cudaStream_t)stream); }
int crdr_eb_quantize(const crdr_eb_desc* d, void* stream) { return eb_launch(d, 0, (cudaStream_t)stream); }
int crdr_eb_dequantize(const crdr_eb_desc* d, void* stream) { return eb_launch(d, 1, (cudaStream_t)stream); }

int crdr_bits_from_likelihood(const float* lik, int32_t n, int64_t per, float* bits, void* stream) {
  return bits_launch(lik, n, per, bits, (cudaStream_t)stream);
}
int crdr_max_abs(const float* x, int64_t count, float* out, void* stream) {
  return max_abs_launch(x, count, out, (cudaStream_t)stream);
}

/* ---- training step ---- */
int crdr_conv_dgrad(const crdr_conv_desc* d, void* stream) {
  if (!d) { set_error("conv_dgrad: null descriptor"); return CRDR_ERR_BAD_SHAPE; }
  return conv2d_launch(d, (cudaStream_t)stream);
}
size_t crdr_conv_wgrad_workspace(const crdr_wgrad_desc* d) { return d ? wgrad_workspace_bytes(d) : 0; }
int crdr_conv_wgrad(const crdr_wgrad_desc* d, void* stream) {
  if (!d) { set_error("conv_wgrad: null descriptor"); return CRDR_ERR_BAD_SHAPE; }
  return wgrad_launch(d, (cudaStream_t)stream);
}
int crdr_pack_weights(const float* master, const int32_t* map, int64_t count, void* hi, void* lo, void* stream) {
  return pack_weights_launch(master, map, count, hi, lo, (cudaStream_t)stream);
}
int crdr_pack_weights_multi(const crdr_pack_job* jobs, int32_t njobs, void* stream) {
  return pack_weights_multi_launch(jobs, njobs, (cudaStream_t)stream);
}
int crdr_epilogue_backward(const crdr_epi_bwd_desc* d, void* stream) {
  if (!d) { set_error("epilogue_backward: null descriptor"); return CRDR_ERR_BAD_SHAPE; }
  return epi_bwd_launch(d, (cudaStream_t)stream);
}
int crdr_colsum_finish(const float* partial, int32_t blocks, int32_t nsums, int32_t which, int32_t c, float* out, float scale,
                       int32_t accumulate, void* stream) {
  return colsum_finish_launch(partial, blocks, nsums, which, c, out, scale, accumulate, (cudaStream_t)stream);
}
int crdr_gate_forward(const crdr_gate_desc* d, void* stream) {
  if (!d) { set_error("gate: null descriptor"); return CRDR_ERR_BAD_SHAPE; }
  return gate_launch(d, 0, (cudaStream_t)stream);
}
int crdr_gate_backward(const crdr_gate_desc* d, void* stream) {
  if (!d) { set_error("gate: null descriptor"); return CRDR_ERR_BAD_SHAPE; }
  return gate_launch(d, 1, (cudaStream_t)stream);
}
int crdr_gauss_backward(const crdr_gauss_bwd_desc* d, void* stream) {
  if (!d) { set_error("gauss_backward: null descriptor"); return CRDR_ERR_BAD_SHAPE; }
  return gauss_bwd_launch(d, (cudaStream_t)stream);
}
int crdr_mse_backward(const float* fake, int32_t fake_cs, const float* real, int32_t n, int32_t hb, int32_t wb, int32_t h,
                      int32_t w, float coef, void* g, int32_t g_cs, void* stream) {
  return mse_bwd_launch(fake, fake_cs, real, n, hb, wb, h, w, coef, g, g_cs, (cudaStream_t)stream);
}
int crdr_adam_step(float* p, const float* g, float* m, float* v, int64_t count, float lr, float beta1, float beta2, float eps,
                   int32_t step, const float* gscale_ptr, float gscale, const float* hyper, void* stream) {
  return adam_launch(p, g, m, v, count, lr, beta1, beta2, eps, step, gscale_ptr, gscale, hyper, (cudaStream_t)stream);
}
int crdr_leaky_relu(crdr_planes x, int64_t m, int32_t c, float slope, void* stream) {
  return leaky_relu_launch(x, m, c, slope, (cudaStream_t)stream);
}
int crdr_planes_grad_to_phases(const void* g8, int32_t g8_cs, int32_t n, int32_t hb, int32_t wb, float scale, void* g,
                               int32_t g_cs, void* stream) {
  return planes_grad_to_phases_launch(g8, g8_cs, n, hb, wb, scale, g, g_cs, (cudaStream_t)stream);
}
int crdr_sum_squares(const float* x, int64_t count, float* partial, float* out, void* stream) {
  return sumsq_launch(x, count, partial, out, (cudaStream_t)stream);
}

}  // extern "C"
