// Element-wise / layout kernels of the codec hot path (HBM-bound; coalesced on both sides through
// 32x32 shared-memory tile transposes between the NHWC working layout and the NCHW API / coder order).
#include "common.cuh"

namespace crdr {

constexpr int kTile = 32;

// tile coordinates shared by the NHWC<->NCHW kernels: blockIdx.x enumerates (image, 32-pixel tile),
// blockIdx.y enumerates 32-channel tiles; block = (32, 8).
struct TileCoord {
  int n, p0, c0;
};
__device__ __forceinline__ TileCoord tile_coord(int hw) {
  const int tiles_per_img = (hw + kTile - 1) / kTile;
  TileCoord t;
  t.n = blockIdx.x / tiles_per_img;
  t.p0 = (blockIdx.x % tiles_per_img) * kTile;
  t.c0 = blockIdx.y * kTile;
  return t;
}

// ---------------------------------------------------------------------------------------------
// GaussianConditional (eval): quantise + likelihood + CDF index, one channel slice per launch.
// MODE 0: full encoder-side op; MODE 1: indexes only; MODE 2: dequantise symbols.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float std_cumulative(float x) {
  // CompressAI GaussianConditional._standardized_cumulative: 0.5 * erfc(-(2^-0.5) * x)
  return 0.5f * erfcf(-0.70710678118654752440f * x);
}

// TRAIN (MODE 0 only): additionally the likelihood of y + noise (noise: NCHW input, replayable by the oracle).
template <int MODE, bool TRAIN = false>
__global__ void __launch_bounds__(256) gauss_kernel(const __grid_constant__ crdr_gauss_desc d, uint32_t* status) {
  __shared__ float s_table[64];
  __shared__ int s_sym[kTile][kTile + 1];
  __shared__ int s_idx[kTile][kTile + 1];
  __shared__ float s_lik[kTile][kTile + 1];
  __shared__ float s_noise[TRAIN ? kTile : 1][kTile + 1];   // in: noise, out: noisy likelihood
  const TileCoord tc = tile_coord(d.hw);
  const int tid = threadIdx.y * 32 + threadIdx.x;
  if (tid < d.ntable && tid < 64) s_table[tid] = d.scale_table[tid];
  if (TRAIN) {
    for (int cc = threadIdx.y; cc < kTile; cc += 8) {
      const int c = tc.c0 + cc, p = tc.p0 + threadIdx.x;
      if (c < d.c && p < d.hw) s_noise[cc][threadIdx.x] = d.noise[((int64_t)tc.n * d.c_total + d.nchw_coff + c) * d.hw + p];
    }
  }
  if (MODE == 2) {
    // load symbols NCHW (pixel-contiguous) into the tile: s_sym[channel][pixel]
    for (int cc = threadIdx.y; cc < kTile; cc += 8) {
      const int c = tc.c0 + cc, p = tc.p0 + threadIdx.x;
      if (c < d.c && p < d.hw) {
        const int64_t o = ((int64_t)tc.n * d.c_total + d.nchw_coff + c) * d.hw + p;
        s_sym[cc][threadIdx.x] = d.symbols ? d.symbols[o] : (int)d.symbols16[o];
      }
    }
  }
  __syncthreads();
  // phase 1: lane = channel (NHWC contiguous), rows = pixels
  for (int pp = threadIdx.y; pp < kTile; pp += 8) {
    const int p = tc.p0 + pp, c = tc.c0 + threadIdx.x;
    if (p >= d.hw || c >= d.c) continue;
    const int64_t m = (int64_t)tc.n * d.hw + p;
    const float mu = d.mu ? d.mu[m * d.ms_cs + d.mu_coff + c] : 0.f;
    float q = 0.f, yq = 0.f;
    if (MODE == 0) {
      const float y = d.y[m * d.y_cs + d.y_coff + c];
      q = rintf(y - mu);
      yq = q + mu;
    } else if (MODE == 2) {
      q = (float)s_sym[threadIdx.x][pp];
      yq = q + mu;
    }
    if (MODE != 1) {
      if (d.yq_f32) d.yq_f32[m * d.yq_f32_cs + d.yq_f32_coff + c] = yq;
      if (d.yq_planes.hi) {
        const int64_t o = m * d.yq_planes.cs + d.yq_planes.coff + c;
        __half h, l;
        split_f16(yq, h, l, status);
        ((__half*)d.yq_planes.hi)[o] = h;
        if (d.yq_planes.lo) ((__half*)d.yq_planes.lo)[o] = l;
      }
    }
    if (MODE != 2) {
      const float sg = fmaxf(d.sigma[m * d.ms_cs + d.sigma_coff + c], d.scale_bound);
      // build_indexes: (ntable-1) - #{k < ntable-1 : sg <= table[k]} = #{k < ntable-1 : table[k] < sg}; the table is
      // strictly ascending, so the count is a lower bound: six branch-free steps instead of 63 compares (same result)
      const int nt = d.ntable - 1;
      int idx = 0;
#pragma unroll
      for (int step = 32; step > 0; step >>= 1) {
        const int t = idx + step;
        if (t <= nt && s_table[t - 1] < sg) idx = t;
      }
      s_idx[threadIdx.x][pp] = idx;
      if (MODE == 0) {
        const float v = fabsf(yq - mu);
        const float upper = std_cumulative((0.5f - v) / sg);
        const float lower = std_cumulative((-0.5f - v) / sg);
        s_lik[threadIdx.x][pp] = fmaxf(upper - lower, 1e-9f);
        s_sym[threadIdx.x][pp] = (int)q;
        if (TRAIN) {
          // quantize(inputs, "noise") = inputs + noise;  _likelihood: values = |outputs - means|
          const float vn = fabsf((d.y[m * d.y_cs + d.y_coff + c] + s_noise[threadIdx.x][pp]) - mu);
          const float un = std_cumulative((0.5f - vn) / sg);
          const float ln = std_cumulative((-0.5f - vn) / sg);
          s_noise[threadIdx.x][pp] = fmaxf(un - ln, 1e-9f);
        }
      }
    }
  }
  if (MODE == 2) return;
  __syncthreads();
  // phase 2: lane = pixel (NCHW contiguous)
  for (int cc = threadIdx.y; cc < kTile; cc += 8) {
    const int c = tc.c0 + cc, p = tc.p0 + threadIdx.x;
    if (c >= d.c || p >= d.hw) continue;
    const int64_t o = ((int64_t)tc.n * d.c_total + d.nchw_coff + c) * d.hw + p;
    if (d.indexes) d.indexes[o] = s_idx[cc][threadIdx.x];
    if (d.indexes8) d.indexes8[o] = (uint8_t)s_idx[cc][threadIdx.x];   // ntable <= 64
    if (MODE == 0) {
      const int sym = s_sym[cc][threadIdx.x];
      if (d.symbols) d.symbols[o] = sym;
      if (d.symbols16) {
        const int sat = max(-32768, min(32767, sym));
        if (sat != sym) atomicOr(status, kFlagSymRange);
        d.symbols16[o] = (int16_t)sat;
      }
      if (d.likelihood) d.likelihood[o] = s_lik[cc][threadIdx.x];
      if (TRAIN) d.likelihood_noisy[o] = s_noise[cc][threadIdx.x];
    }
  }
}

// ---------------------------------------------------------------------------------------------
// EntropyBottleneck (eval): quantise about the medians + factorised-prior likelihood.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float eb_logits(const float* __restrict__ p, float x) {
  float h[3], g[3];
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    float t = p[j] * x + p[3 + j];
    h[j] = t + p[6 + j] * tanhf(t);
  }
  p += 9;
#pragma unroll
  for (int l = 0; l < 3; ++l) {
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      float t = p[3 * j] * h[0];
      t += p[3 * j + 1] * h[1];
      t += p[3 * j + 2] * h[2];
      t += p[9 + j];
      g[j] = t + p[12 + j] * tanhf(t);
    }
#pragma unroll
    for (int j = 0; j < 3; ++j) h[j] = g[j];
    p += 15;
  }
  float t = p[0] * h[0];
  t += p[1] * h[1];
  t += p[2] * h[2];
  return t + p[3];
}

__device__ __forceinline__ float eb_likelihood(const float* __restrict__ prm, float v) {
  const float lower = eb_logits(prm, v - 0.5f);
  const float upper = eb_logits(prm, v + 0.5f);
  const float sum = lower + upper;
  const float sgn = sum > 0.f ? -1.f : (sum < 0.f ? 1.f : 0.f);
  return fmaxf(fabsf(sigmoidf_(sgn * upper) - sigmoidf_(sgn * lower)), 1e-9f);
}

template <int DEQUANT>
__global__ void __launch_bounds__(256) eb_kernel(const __grid_constant__ crdr_eb_desc d, uint32_t* status) {
  __shared__ int s_sym[kTile][kTile + 1];
  __shared__ float s_zh[kTile][kTile + 1];
  __shared__ float s_lik[kTile][kTile + 1];
  __shared__ float s_noise[kTile][kTile + 1];   // training mode: noise in, noisy likelihood out
  const TileCoord tc = tile_coord(d.hw);
  const bool train = !DEQUANT && d.noise != nullptr;
  if (train) {
    for (int cc = threadIdx.y; cc < kTile; cc += 8) {
      const int c = tc.c0 + cc, p = tc.p0 + threadIdx.x;
      if (c < d.c && p < d.hw) s_noise[cc][threadIdx.x] = d.noise[((int64_t)tc.n * d.c + c) * d.hw + p];
    }
    __syncthreads();
  }
  if (DEQUANT) {
    for (int cc = threadIdx.y; cc < kTile; cc += 8) {
      const int c = tc.c0 + cc, p = tc.p0 + threadIdx.x;
      if (c < d.c && p < d.hw) s_sym[cc][threadIdx.x] = d.symbols[((int64_t)tc.n * d.c + c) * d.hw + p];
    }
    __syncthreads();
  }
  for (int pp = threadIdx.y; pp < kTile; pp += 8) {
    const int p = tc.p0 + pp, c = tc.c0 + threadIdx.x;
    if (p >= d.hw || c >= d.c) continue;
    const int64_t m = (int64_t)tc.n * d.hw + p;
    const float med = d.medians[c];
    float q;
    if (DEQUANT) q = (float)s_sym[threadIdx.x][pp];
    else q = rintf(d.z[m * d.z_cs + c] - med);
    const float zh = q + med;
    if (d.zhat_planes.hi) {
      const int64_t o = m * d.zhat_planes.cs + d.zhat_planes.coff + c;
      __half h, l;
      split_f16(zh, h, l, status);
      ((__half*)d.zhat_planes.hi)[o] = h;
      if (d.zhat_planes.lo) ((__half*)d.zhat_planes.lo)[o] = l;
    }
    s_zh[threadIdx.x][pp] = zh;
    if (!DEQUANT) {
      s_sym[threadIdx.x][pp] = (int)q;
      const float* prm = d.params + (int64_t)c * 58;
      s_lik[threadIdx.x][pp] = eb_likelihood(prm, zh);
      if (train) s_noise[threadIdx.x][pp] = eb_likelihood(prm, d.z[m * d.z_cs + c] + s_noise[threadIdx.x][pp]);
    }
  }
  __syncthreads();
  for (int cc = threadIdx.y; cc < kTile; cc += 8) {
    const int c = tc.c0 + cc, p = tc.p0 + threadIdx.x;
    if (c >= d.c || p >= d.hw) continue;
    const int64_t o = ((int64_t)tc.n * d.c + c) * d.hw + p;
    if (d.zhat_nchw) d.zhat_nchw[o] = s_zh[cc][threadIdx.x];
    if (!DEQUANT) {
      if (d.symbols) d.symbols[o] = s_sym[cc][threadIdx.x];
      if (d.likelihood) d.likelihood[o] = s_lik[cc][threadIdx.x];
      if (train && d.likelihood_noisy) d.likelihood_noisy[o] = s_noise[cc][threadIdx.x];
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Layout / pre / post kernels
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) nhwc_to_nchw_kernel(const float* __restrict__ x, int x_cs, int x_coff, int hw,
                                                           int c, float* __restrict__ out) {
  __shared__ float s[kTile][kTile + 1];
  const TileCoord tc = tile_coord(hw);
  for (int pp = threadIdx.y; pp < kTile; pp += 8) {
    const int p = tc.p0 + pp, ch = tc.c0 + threadIdx.x;
    if (p < hw && ch < c) s[threadIdx.x][pp] = x[((int64_t)tc.n * hw + p) * x_cs + x_coff + ch];
  }
  __syncthreads();
  for (int cc = threadIdx.y; cc < kTile; cc += 8) {
    const int ch = tc.c0 + cc, p = tc.p0 + threadIdx.x;
    if (ch < c && p < hw) out[((int64_t)tc.n * c + ch) * hw + p] = s[cc][threadIdx.x];
  }
}

__global__ void __launch_bounds__(256) affine_to_planes_kernel(const float* __restrict__ x, int x_cs, int x_coff,
                                                               int64_t m, int c, const float* __restrict__ scale,
                                                               const float* __restrict__ shift, crdr_planes out,
                                                               uint32_t* status) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= m * c) return;
  const int ch = (int)(idx % c);
  const int64_t pix = idx / c;
  float v = x[pix * x_cs + x_coff + ch];
  v = fmaf(v, scale ? scale[ch] : 1.f, shift ? shift[ch] : 0.f);
  const int64_t o = pix * out.cs + out.coff + ch;
  __half h, l;
  split_f16(v, h, l, status);
  ((__half*)out.hi)[o] = h;
  if (out.lo) ((__half*)out.lo)[o] = l;
}

// one thread per padded pixel: 3 strided reads (NCHW), two 16-byte writes (8-channel NHWC planes)
__global__ void __launch_bounds__(256) image_to_planes_kernel(const float* __restrict__ img, int n, int h, int w,
                                                              int hp, int wp, crdr_planes out, uint32_t* status) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t total = (int64_t)n * hp * wp;
  if (idx >= total) return;
  const int x = (int)(idx % wp);
  const int64_t t = idx / wp;
  const int y = (int)(t % hp);
  const int b = (int)(t / hp);
  const int sy = y < h ? y : 2 * (h - 1) - y;  // 'reflect' padding on the bottom / right edges
  const int sx = x < w ? x : 2 * (w - 1) - x;
  __align__(16) __half hh[8], ll[8];
#pragma unroll
  for (int c = 0; c < 8; ++c) { hh[c] = __float2half_rn(0.f); ll[c] = __float2half_rn(0.f); }
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const float v = img[(((int64_t)b * 3 + c) * h + sy) * w + sx];
    split_f16(v, hh[c], ll[c], status);
  }
  const int64_t o = idx * out.cs + out.coff;
  *(uint4*)((__half*)out.hi + o) = *(const uint4*)hh;
  if (out.lo) *(uint4*)((__half*)out.lo + o) = *(const uint4*)ll;
}

// im2col of the first analysis layer (5x5, stride 2, padding 2 on the reflect-padded image): one thread per
// (output pixel, group of 8 patch channels); patch channel k = (kh * 5 + kw) * 3 + c for k < 75, zero up to 128.
// The layer then runs as a 1x1 convolution with K = 128 on the patch engine (2 K blocks instead of a 25-tap gather).
__global__ void __launch_bounds__(256) image_to_patches_kernel(const float* __restrict__ img, int n, int h, int w,
                                                               int hp, int wp, crdr_planes out, int ng, uint32_t* status) {
  // ng groups of 8 channels per pixel (10 .. 16): 75 patch channels + zero padding up to the tensor's channel count
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int ho = hp / 2, wo = wp / 2;
  const int64_t total = (int64_t)n * ho * wo * ng;
  if (idx >= total) return;
  const int64_t pix = idx / ng;
  const int grp = (int)(idx - pix * ng);
  const int j = (int)(pix % wo);
  const int64_t t = pix / wo;
  const int i = (int)(t % ho);
  const int b = (int)(t / ho);
  __align__(16) __half hh[8], ll[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    const int k = grp * 8 + e;
    float v = 0.f;
    if (k < 75) {
      const int tap = k / 3, c = k - 3 * tap;
      const int kh = tap / 5, kw = tap - 5 * kh;
      const int y = 2 * i + kh - 2, x = 2 * j + kw - 2;
      if (y >= 0 && y < hp && x >= 0 && x < wp) {     // the convolution's own zero padding
        const int sy = y < h ? y : 2 * (h - 1) - y;   // 'reflect' padding on the bottom / right edges
        const int sx = x < w ? x : 2 * (w - 1) - x;
        v = img[(((int64_t)b * 3 + c) * h + sy) * w + sx];
      }
    }
    split_f16(v, hh[e], ll[e], status);
  }
  const int64_t o = pix * out.cs + out.coff + grp * 8;
  *(uint4*)((__half*)out.hi + o) = *(const uint4*)hh;
  if (out.lo) *(uint4*)((__half*)out.lo + o) = *(const uint4*)ll;
}

// uint8 variant of image_to_patches_kernel: ToTensor + Normalize(0.5, 0.5) of the reference data path, in its fp32 order
__global__ void __launch_bounds__(256) image_u8_to_patches_kernel(const uint8_t* __restrict__ img, int n, int h, int w,
                                                                  int hp, int wp, crdr_planes out, int ng, uint32_t* status) {
  // ng groups of 8 channels per pixel (10 .. 16): 75 patch channels + zero padding up to the tensor's channel count
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int ho = hp / 2, wo = wp / 2;
  const int64_t total = (int64_t)n * ho * wo * ng;
  if (idx >= total) return;
  const int64_t pix = idx / ng;
  const int grp = (int)(idx - pix * ng);
  const int j = (int)(pix % wo);
  const int64_t t = pix / wo;
  const int i = (int)(t % ho);
  const int b = (int)(t / ho);
  __align__(16) __half hh[8], ll[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    const int k = grp * 8 + e;
    float v = 0.f;
    if (k < 75) {
      const int tap = k / 3, c = k - 3 * tap;
      const int kh = tap / 5, kw = tap - 5 * kh;
      const int y = 2 * i + kh - 2, x = 2 * j + kw - 2;
      if (y >= 0 && y < hp && x >= 0 && x < wp) {
        const int sy = y < h ? y : 2 * (h - 1) - y;
        const int sx = x < w ? x : 2 * (w - 1) - x;
        const float u = (float)img[(((int64_t)b * 3 + c) * h + sy) * w + sx];
        v = __fdiv_rn(__fsub_rn(__fdiv_rn(u, 255.0f), 0.5f), 0.5f);   // ToTensor: u / 255; Normalize: (t - 0.5) / 0.5
      }
    }
    split_f16(v, hh[e], ll[e], status);
  }
  const int64_t o = pix * out.cs + out.coff + grp * 8;
  *(uint4*)((__half*)out.hi + o) = *(const uint4*)hh;
  if (out.lo) *(uint4*)((__half*)out.lo + o) = *(const uint4*)ll;
}

__global__ void __launch_bounds__(256) planes_to_image_kernel(const float* __restrict__ x, int x_cs, int n, int hp,
                                                              int wp, int h, int w, float* __restrict__ img) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t total = (int64_t)n * 3 * h * w;
  if (idx >= total) return;
  const int xx = (int)(idx % w);
  int64_t t = idx / w;
  const int yy = (int)(t % h);
  t /= h;
  const int c = (int)(t % 3);
  const int b = (int)(t / 3);
  const float v = x[(((int64_t)b * hp + yy) * wp + xx) * x_cs + c];
  img[idx] = fminf(fmaxf(v, -1.f), 1.f);
}

// Same, for a stride-2 transposed convolution evaluated as ONE stride-1 launch whose output channels are the four
// output phases: x[(b, y/2, x/2), ((y%2)*2 + x%2)*3 + c]  (pixel shuffle + crop + clamp in one pass).
__global__ void __launch_bounds__(256) phases_to_image_kernel(const float* __restrict__ x, int x_cs, int n, int hb,
                                                              int wb, int h, int w, float* __restrict__ img, int clamp) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t total = (int64_t)n * 3 * h * w;
  if (idx >= total) return;
  const int xx = (int)(idx % w);
  int64_t t = idx / w;
  const int yy = (int)(t % h);
  t /= h;
  const int c = (int)(t % 3);
  const int b = (int)(t / 3);
  const float v = x[(((int64_t)b * hb + (yy >> 1)) * wb + (xx >> 1)) * x_cs + ((yy & 1) * 2 + (xx & 1)) * 3 + c];
  img[idx] = clamp ? fminf(fmaxf(v, -1.f), 1.f) : v;
}

// uint8 variant: clamp, then the reference's PNG conversion ((x + 1) / 2 * 255 in fp32, truncated by astype(uint8))
__global__ void __launch_bounds__(256) phases_to_image_u8_kernel(const float* __restrict__ x, int x_cs, int n, int hb,
                                                                 int wb, int h, int w, uint8_t* __restrict__ img) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t total = (int64_t)n * 3 * h * w;
  if (idx >= total) return;
  const int xx = (int)(idx % w);
  int64_t t = idx / w;
  const int yy = (int)(t % h);
  t /= h;
  const int c = (int)(t % 3);
  const int b = (int)(t / 3);
  float v = x[(((int64_t)b * hb + (yy >> 1)) * wb + (xx >> 1)) * x_cs + ((yy & 1) * 2 + (xx & 1)) * 3 + c];
  v = fminf(fmaxf(v, -1.f), 1.f);
  v = __fmul_rn(__fdiv_rn(__fadd_rn(v, 1.0f), 2.0f), 255.0f);
  img[idx] = (uint8_t)(int)v;   // v in [0, 255]: truncation
}

// one block per image; fixed summation order -> run-to-run and batch-size invariant
__global__ void __launch_bounds__(1024) bits_kernel(const float* __restrict__ lik, int64_t per, float* __restrict__ bits) {
  __shared__ double s[1024];
  const float* p = lik + (int64_t)blockIdx.x * per;
  double acc = 0.0;
  for (int64_t i = threadIdx.x; i < per; i += 1024) acc += (double)log2f(p[i]);
  s[threadIdx.x] = acc;
  __syncthreads();
  for (int st = 512; st > 0; st >>= 1) {
    if ((int)threadIdx.x < st) s[threadIdx.x] += s[threadIdx.x + st];
    __syncthreads();
  }
  if (threadIdx.x == 0) bits[blockIdx.x] = (float)(-s[0]);
}

__global__ void __launch_bounds__(256) max_abs_kernel(const float* __restrict__ x, int64_t count, float* out) {
  float m = 0.f;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (int64_t)gridDim.x * blockDim.x)
    m = fmaxf(m, fabsf(x[i]));
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) atomicMax((unsigned int*)out, __float_as_uint(m));  // m >= 0: uint order == float order
}

// gridDim.y = image; per-image maximum through an atomicMax on the (non-negative) float bit pattern
__global__ void __launch_bounds__(256) max_abs_batch_kernel(const float* __restrict__ x, int64_t per, float* out) {
  const float* p = x + (int64_t)blockIdx.y * per;
  float m = 0.f;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < per; i += (int64_t)gridDim.x * blockDim.x)
    m = fmaxf(m, fabsf(p[i]));
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) atomicMax((unsigned int*)out + blockIdx.y, __float_as_uint(m));
}

// ---------------------------------------------------------------------------------------------
// launchers
// ---------------------------------------------------------------------------------------------
static dim3 tile_grid(int n, int hw, int c) { return dim3((unsigned)(n * ((hw + kTile - 1) / kTile)), (unsigned)((c + kTile - 1) / kTile)); }

int gauss_launch(const crdr_gauss_desc* d, int mode, cudaStream_t st) {
  if (d->n <= 0 || d->hw <= 0 || d->c <= 0 || d->ntable < 1 || d->ntable > 64) {
    set_error("gauss: bad shape (n=%d hw=%d c=%d ntable=%d)", d->n, d->hw, d->c, d->ntable);
    return CRDR_ERR_BAD_SHAPE;
  }
  uint32_t* status = device_status_word();
  if (!status) return CRDR_ERR_CUDA;
  const dim3 g = tile_grid(d->n, d->hw, d->c), b(32, 8);
  if (mode == 0 && (d->noise || d->likelihood_noisy)) {
    if (!d->noise || !d->likelihood_noisy || !d->y) { set_error("gauss: training mode needs y, noise and likelihood_noisy"); return CRDR_ERR_BAD_SHAPE; }
    gauss_kernel<0, true><<<g, b, 0, st>>>(*d, status);
  } else if (mode == 0) gauss_kernel<0><<<g, b, 0, st>>>(*d, status);
  else if (mode == 1) gauss_kernel<1><<<g, b, 0, st>>>(*d, status);
  else gauss_kernel<2><<<g, b, 0, st>>>(*d, status);
  return check_launch("gauss_kernel");
}

int eb_launch(const crdr_eb_desc* d, int dequant, cudaStream_t st) {
  if (d->n <= 0 || d->hw <= 0 || d->c <= 0) { set_error("eb: bad shape"); return CRDR_ERR_BAD_SHAPE; }
  uint32_t* status = device_status_word();
  if (!status) return CRDR_ERR_CUDA;
  const dim3 g = tile_grid(d->n, d->hw, d->c), b(32, 8);
  if (dequant) eb_kernel<1><<<g, b, 0, st>>>(*d, status);
  else eb_kernel<0><<<g, b, 0, st>>>(*d, status);
  return check_launch("eb_kernel");
}

int nhwc_to_nchw_launch(const float* x, int x_cs, int x_coff, int n, int hw, int c, float* out, cudaStream_t st) {
  if (n <= 0 || hw <= 0 || c <= 0) { set_error("nhwc_to_nchw: bad shape"); return CRDR_ERR_BAD_SHAPE; }
  nhwc_to_nchw_kernel<<<tile_grid(n, hw, c), dim3(32, 8), 0, st>>>(x, x_cs, x_coff, hw, c, out);
  return check_launch("nhwc_to_nchw_kernel");
}

int affine_to_planes_launch(const float* x, int x_cs, int x_coff, int64_t m, int c, const float* scale,
                            const float* shift, crdr_planes out, cudaStream_t st) {
  if (m <= 0 || c <= 0 || !out.hi) { set_error("affine_to_planes: bad arguments"); return CRDR_ERR_BAD_SHAPE; }
  uint32_t* status = device_status_word();
  if (!status) return CRDR_ERR_CUDA;
  const int64_t total = m * c;
  affine_to_planes_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(x, x_cs, x_coff, m, c, scale, shift, out, status);
  return check_launch("affine_to_planes_kernel");
}

int image_to_planes_launch(const float* img, int n, int h, int w, int hp, int wp, crdr_planes out, cudaStream_t st) {
  if (n <= 0 || h <= 0 || w <= 0 || hp < h || wp < w || hp - h >= h || wp - w >= w || out.cs % 8 || out.coff % 8 || !out.hi) {
    set_error("image_to_planes: bad shape (h=%d w=%d hp=%d wp=%d cs=%d)", h, w, hp, wp, out.cs);
    return CRDR_ERR_BAD_SHAPE;
  }
  uint32_t* status = device_status_word();
  if (!status) return CRDR_ERR_CUDA;
  const int64_t total = (int64_t)n * hp * wp;
  image_to_planes_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(img, n, h, w, hp, wp, out, status);
  return check_launch("image_to_planes_kernel");
}

int image_to_patches_launch(const float* img, int n, int h, int w, int hp, int wp, crdr_planes out, cudaStream_t st) {
  if (n <= 0 || h <= 0 || w <= 0 || hp < h || wp < w || hp - h >= h || wp - w >= w || (hp & 1) || (wp & 1) ||
      out.cs % 8 || out.coff % 8 || out.cs < out.coff + 80 || !out.hi) {
    set_error("image_to_patches: bad shape (h=%d w=%d hp=%d wp=%d cs=%d)", h, w, hp, wp, out.cs);
    return CRDR_ERR_BAD_SHAPE;
  }
  uint32_t* status = device_status_word();
  if (!status) return CRDR_ERR_CUDA;
  const int ng = (out.cs - out.coff) / 8 < 16 ? (out.cs - out.coff) / 8 : 16;
  const int64_t total = (int64_t)n * (hp / 2) * (wp / 2) * ng;
  image_to_patches_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(img, n, h, w, hp, wp, out, ng, status);
  return check_launch("image_to_patches_kernel");
}

int planes_to_image_launch(const float* x, int x_cs, int n, int hp, int wp, int h, int w, float* img, cudaStream_t st) {
  if (n <= 0 || h <= 0 || w <= 0 || hp < h || wp < w) { set_error("planes_to_image: bad shape"); return CRDR_ERR_BAD_SHAPE; }
  const int64_t total = (int64_t)n * 3 * h * w;
  planes_to_image_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(x, x_cs, n, hp, wp, h, w, img);
  return check_launch("planes_to_image_kernel");
}

int phases_to_image_launch(const float* x, int x_cs, int n, int hb, int wb, int h, int w, float* img, int clamp, cudaStream_t st) {
  if (n <= 0 || h <= 0 || w <= 0 || 2 * hb < h || 2 * wb < w || x_cs < 12) {
    set_error("phases_to_image: bad shape");
    return CRDR_ERR_BAD_SHAPE;
  }
  const int64_t total = (int64_t)n * 3 * h * w;
  phases_to_image_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(x, x_cs, n, hb, wb, h, w, img, clamp);
  return check_launch("phases_to_image_kernel");
}

int image_u8_to_patches_launch(const uint8_t* img, int n, int h, int w, int hp, int wp, crdr_planes out, cudaStream_t st) {
  if (n <= 0 || h <= 0 || w <= 0 || hp < h || wp < w || hp - h >= h || wp - w >= w || (hp & 1) || (wp & 1) ||
      out.cs % 8 || out.coff % 8 || out.cs < out.coff + 80 || !out.hi || !img) {
    set_error("image_u8_to_patches: bad shape (h=%d w=%d hp=%d wp=%d cs=%d)", h, w, hp, wp, out.cs);
    return CRDR_ERR_BAD_SHAPE;
  }
  uint32_t* status = device_status_word();
  if (!status) return CRDR_ERR_CUDA;
  const int ng = (out.cs - out.coff) / 8 < 16 ? (out.cs - out.coff) / 8 : 16;
  const int64_t total = (int64_t)n * (hp / 2) * (wp / 2) * ng;
  image_u8_to_patches_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(img, n, h, w, hp, wp, out, ng, status);
  return check_launch("image_u8_to_patches_kernel");
}

int phases_to_image_u8_launch(const float* x, int x_cs, int n, int hb, int wb, int h, int w, uint8_t* img, cudaStream_t st) {
  if (n <= 0 || h <= 0 || w <= 0 || 2 * hb < h || 2 * wb < w || x_cs < 12 || !img) {
    set_error("phases_to_image_u8: bad shape");
    return CRDR_ERR_BAD_SHAPE;
  }
  const int64_t total = (int64_t)n * 3 * h * w;
  phases_to_image_u8_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(x, x_cs, n, hb, wb, h, w, img);
  return check_launch("phases_to_image_u8_kernel");
}

int max_abs_batch_launch(const float* x, int n, int64_t per, float* out, cudaStream_t st) {
  if (n <= 0 || per <= 0) { set_error("max_abs_batch: bad shape"); return CRDR_ERR_BAD_SHAPE; }
  cudaError_t e = cudaMemsetAsync(out, 0, sizeof(float) * (size_t)n, st);
  if (e != cudaSuccess) { set_error("max_abs_batch: memset failed: %s", cudaGetErrorString(e)); return CRDR_ERR_CUDA; }
  int bx = (int)((per + 255) / 256);
  const int cap = (148 * 8 + n - 1) / n;
  if (bx > cap) bx = cap;
  if (bx < 1) bx = 1;
  max_abs_batch_kernel<<<dim3((unsigned)bx, (unsigned)n), 256, 0, st>>>(x, per, out);
  return check_launch("max_abs_batch_kernel");
}

__global__ void status_clear_bits_kernel(uint32_t* status, uint32_t bits) { atomicAnd(status, ~bits); }

int status_clear_bits_launch(uint32_t* status, uint32_t bits, cudaStream_t st) {
  status_clear_bits_kernel<<<1, 1, 0, st>>>(status, bits);
  return check_launch("status_clear_bits_kernel");
}

int bits_launch(const float* lik, int n, int64_t per, float* bits, cudaStream_t st) {
  if (n <= 0 || per <= 0) { set_error("bits: bad shape"); return CRDR_ERR_BAD_SHAPE; }
  bits_kernel<<<n, 1024, 0, st>>>(lik, per, bits);
  return check_launch("bits_kernel");
}

int max_abs_launch(const float* x, int64_t count, float* out, cudaStream_t st) {
  if (count <= 0) { set_error("max_abs: bad shape"); return CRDR_ERR_BAD_SHAPE; }
  cudaError_t e = cudaMemsetAsync(out, 0, sizeof(float), st);
  if (e != cudaSuccess) { set_error("max_abs: memset failed: %s", cudaGetErrorString(e)); return CRDR_ERR_CUDA; }
  int blocks = (int)((count + 255) / 256);
  if (blocks > 148 * 8) blocks = 148 * 8;
  max_abs_kernel<<<blocks, 256, 0, st>>>(x, count, out);
  return check_launch("max_abs_kernel");
}

}  // namespace crdr
