// Shared device/host helpers for libcrdr_sm100.so (sm_100a only).
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/crdr_b200.h"

namespace crdr {

constexpr float kLoScale = 2048.0f;          // lo plane stores (x - hi) * 2^11
constexpr float kLoInv = 1.0f / 2048.0f;
constexpr uint32_t kFlagOverflow = 1u;       // fp16 overflow in an epilogue
constexpr uint32_t kFlagTimeout = 2u;        // mbarrier wait timed out (pipeline bug guard)
constexpr uint32_t kFlagSymRange = 4u;       // a quantised symbol did not fit the compact int16 copy

void set_error(const char* fmt, ...);
uint32_t* device_status_word();              // per-device, lazily allocated
int check_launch(const char* what);

// Split an fp32 value into the (hi, lo) fp16 pair; raises the overflow flag instead of producing inf.
__device__ __forceinline__ void split_f16(float x, __half& hi, __half& lo, uint32_t* status) {
  if (fabsf(x) > 65504.0f) {
    atomicOr(status, kFlagOverflow);
    x = copysignf(65504.0f, x);
  }
  hi = __float2half_rn(x);
  lo = __float2half_rn((x - __half2float(hi)) * kLoScale);
}

__device__ __forceinline__ float join_f16(__half hi, __half lo) {
  return fmaf(__half2float(lo), kLoInv, __half2float(hi));
}

__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }

// The fused epilogue arithmetic (see crdr_conv_desc in include/crdr_b200.h).
__device__ __forceinline__ float epilogue_math(float acc, float bias, int relu, float addv, int mode, float res,
                                               float trunk, float scale, float shift) {
  float v = acc + bias;
  if (relu) v = fmaxf(v, 0.0f);
  v += addv;
  if (mode == CRDR_EPI_RESIDUAL) v += res;
  else if (mode == CRDR_EPI_GATE) v = fmaf(trunk, sigmoidf_(v), res);
  else if (mode == CRDR_EPI_HALF_TANH) v = fmaf(0.5f, tanhf(v), res);
  return fmaf(v, scale, shift);
}

}  // namespace crdr
