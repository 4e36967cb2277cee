// Training-side element-wise kernels of libcrdr_sm100.so: device weight packing, the backward of the fused convolution
// epilogues, the NLAM gate, the GaussianConditional rate term, the MSE term, per-channel sum finalisation, fused Adam
// and the gradient-norm reduction.  Roofline: HBM (every kernel streams its operands once, 128-bit accesses on the
// NHWC fp16 planes).  Gradients of activations are single fp16 planes scaled by the caller's loss scale.
#include "common.cuh"

namespace crdr {

// ----------------------------------------------------------------------------------------------
// packed[i] = split(master[map[i]])  (map[i] < 0: zero padding).  One gather per packed element: the same launch
// serves forward matrices, dgrad matrices (flipped / transposed) and phase-packed forms; the maps are built once.
// ----------------------------------------------------------------------------------------------
__global__ void pack_weights_kernel(const float* __restrict__ master, const int32_t* __restrict__ map, int64_t count,
                                    __half* __restrict__ hi, __half* __restrict__ lo) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < count; i += (int64_t)gridDim.x * blockDim.x) {
    const int32_t m = map[i];
    float v = m >= 0 ? master[m] : 0.0f;
    v = fminf(fmaxf(v, -65504.0f), 65504.0f);
    const __half h = __float2half_rn(v);
    hi[i] = h;
    if (lo) lo[i] = __float2half_rn((v - __half2float(h)) * kLoScale);
  }
}

int pack_weights_launch(const float* master, const int32_t* map, int64_t count, void* hi, void* lo, cudaStream_t st) {
  if (!master || !map || !hi || count <= 0) { set_error("pack_weights: missing operand"); return CRDR_ERR_BAD_SHAPE; }
  int64_t blocks = (count + 255) / 256;
  if (blocks > 148 * 32) blocks = 148 * 32;
  pack_weights_kernel<<<(unsigned)blocks, 256, 0, st>>>(master, map, count, (__half*)hi, (__half*)lo);
  return check_launch("pack_weights_kernel");
}

// All matrices of a model in ONE launch: blockIdx.y selects the job (a device-resident table built once, the pointers are
// stable), blockIdx.x strides over its elements.
__global__ void pack_weights_multi_kernel(const crdr_pack_job* __restrict__ jobs) {
  const crdr_pack_job j = jobs[blockIdx.y];
  const float* __restrict__ master = j.master;
  const int32_t* __restrict__ map = j.map;
  __half* __restrict__ hi = (__half*)j.hi;
  __half* __restrict__ lo = (__half*)j.lo;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < j.count; i += (int64_t)gridDim.x * blockDim.x) {
    const int32_t m = map[i];
    float v = m >= 0 ? master[m] : 0.0f;
    v = fminf(fmaxf(v, -65504.0f), 65504.0f);
    const __half h = __float2half_rn(v);
    hi[i] = h;
    if (lo) lo[i] = __float2half_rn((v - __half2float(h)) * kLoScale);
  }
}

int pack_weights_multi_launch(const crdr_pack_job* jobs, int njobs, cudaStream_t st) {
  if (!jobs || njobs <= 0 || njobs > 65535) { set_error("pack_weights_multi: bad job table"); return CRDR_ERR_BAD_SHAPE; }
  pack_weights_multi_kernel<<<dim3(96, (unsigned)njobs), 256, 0, st>>>(jobs);
  return check_launch("pack_weights_multi_kernel");
}

// ----------------------------------------------------------------------------------------------
// 8-channel vector helpers on fp16 planes
// ----------------------------------------------------------------------------------------------
struct F8 { float v[8]; };

__device__ __forceinline__ F8 load_h8(const __half* p) {
  const uint4 u = *reinterpret_cast<const uint4*>(p);
  const uint32_t w[4] = {u.x, u.y, u.z, u.w};
  F8 r;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&w[i]));
    r.v[2 * i] = f.x;
    r.v[2 * i + 1] = f.y;
  }
  return r;
}
__device__ __forceinline__ void store_h8(__half* p, const F8& x, uint32_t* status) {
  uint32_t w[4];
  bool over = false;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float a = x.v[2 * i], b = x.v[2 * i + 1];
    if (!(fabsf(a) <= 65504.0f) || !(fabsf(b) <= 65504.0f)) {
      over = true;
      a = fminf(fmaxf(a, -65504.0f), 65504.0f);
      b = fminf(fmaxf(b, -65504.0f), 65504.0f);
    }
    const __half2 h = __floats2half2_rn(a, b);
    w[i] = *reinterpret_cast<const uint32_t*>(&h);
  }
  if (over) atomicOr(status, kFlagOverflow);
  *reinterpret_cast<uint4*>(p) = make_uint4(w[0], w[1], w[2], w[3]);
}
__device__ __forceinline__ F8 load_planes8(const crdr_planes& t, int64_t row, int c0) {
  const int64_t o = row * t.cs + t.coff + c0;
  F8 r = load_h8(reinterpret_cast<const __half*>(t.hi) + o);
  if (t.lo) {
    const F8 l = load_h8(reinterpret_cast<const __half*>(t.lo) + o);
#pragma unroll
    for (int e = 0; e < 8; ++e) r.v[e] = fmaf(l.v[e], kLoInv, r.v[e]);
  }
  return r;
}
__device__ __forceinline__ F8 load_vec8(const float* p, int c0, float dflt) {
  F8 r;
  if (p) {
    const float4 a = *reinterpret_cast<const float4*>(p + c0), b = *reinterpret_cast<const float4*>(p + c0 + 4);
    r.v[0] = a.x; r.v[1] = a.y; r.v[2] = a.z; r.v[3] = a.w; r.v[4] = b.x; r.v[5] = b.y; r.v[6] = b.z; r.v[7] = b.w;
  } else {
#pragma unroll
    for (int e = 0; e < 8; ++e) r.v[e] = dflt;
  }
  return r;
}

// Block (32, 8): threadIdx.x walks 8-channel groups, threadIdx.y rows; per-block per-channel sums of up to 3 quantities
// leave through shared memory in a fixed order (deterministic).
constexpr int kBwRowsPerBlockY = 8;
template <int NS>
__device__ __forceinline__ void block_colsum_store(float (&acc)[NS][8], float* s_red, float* partial, int c, int c0) {
  // s_red: [8][32][NS*8]
  float* mine = s_red + ((threadIdx.y * 32 + threadIdx.x) * NS) * 8;
#pragma unroll
  for (int k = 0; k < NS; ++k)
#pragma unroll
    for (int e = 0; e < 8; ++e) mine[k * 8 + e] = acc[k][e];
  __syncthreads();
  if (threadIdx.y == 0 && c0 < c) {
#pragma unroll
    for (int k = 0; k < NS; ++k)
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        float s = 0.f;
        for (int y = 0; y < kBwRowsPerBlockY; ++y) s += s_red[((y * 32 + threadIdx.x) * NS) * 8 + k * 8 + e];
        partial[((int64_t)blockIdx.x * NS + k) * c + c0 + e] = s;
      }
  }
  __syncthreads();
}

// ----------------------------------------------------------------------------------------------
// Backward of the fused convolution epilogue  out = affine( [relu](acc + bias) [+ res | res + 0.5 tanh(.)] )
// ----------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) epi_bwd_kernel(const crdr_epi_bwd_desc d, uint32_t* status) {
  __shared__ float s_red[8 * 32 * 3 * 8];
  const int64_t rows_per = (d.m + gridDim.x - 1) / gridDim.x;
  const int64_t r0 = (int64_t)blockIdx.x * rows_per;
  const int64_t r1 = r0 + rows_per < d.m ? r0 + rows_per : d.m;
  const int groups = d.c >> 3;
  for (int cg0 = 0; cg0 < groups; cg0 += 32) {
    const int cg = cg0 + threadIdx.x;
    const int c0 = cg * 8;
    const bool live = cg < groups;
    float acc[3][8];
#pragma unroll
    for (int k = 0; k < 3; ++k)
#pragma unroll
      for (int e = 0; e < 8; ++e) acc[k][e] = 0.f;
    if (live) {
      const F8 sc = load_vec8(d.scale, c0, 1.0f), sh = load_vec8(d.shift, c0, 0.0f);
      const bool affine = d.scale != nullptr || d.shift != nullptr;
      // ReLU followed by a per-channel bias (beta conditioning): the unit was clipped iff the stored value IS the bias
      F8 zero_level = load_vec8(d.add_vec, c0, 0.0f);
#pragma unroll
      for (int e = 0; e < 8; ++e) zero_level.v[e] = __half2float(__float2half_rn(zero_level.v[e]));
      for (int64_t r = r0 + threadIdx.y; r < r1; r += kBwRowsPerBlockY) {
        const F8 g = load_h8(reinterpret_cast<const __half*>(d.g.hi) + r * d.g.cs + d.g.coff + c0);
        F8 g1 = g, dv;
#pragma unroll
        for (int e = 0; e < 8; ++e) acc[1][e] += g.v[e];
        if (affine) {
          const F8 o = load_planes8(d.out, r, c0);
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            const float pre = (o.v[e] - sh.v[e]) / sc.v[e];
            acc[2][e] = fmaf(g.v[e], pre, acc[2][e]);
            g1.v[e] = g.v[e] * sc.v[e];
          }
        }
        dv = g1;
        if (d.relu) {
          const F8 o = load_h8(reinterpret_cast<const __half*>(d.out.hi) + r * d.out.cs + d.out.coff + c0);
          if (d.add_vec) {
#pragma unroll
            for (int e = 0; e < 8; ++e) dv.v[e] = o.v[e] != zero_level.v[e] ? g1.v[e] : 0.f;
          } else if (d.leaky_slope != 0.f) {   // LeakyReLU (discriminator): the stored output has the sign of the input
#pragma unroll
            for (int e = 0; e < 8; ++e) dv.v[e] = o.v[e] > 0.f ? g1.v[e] : d.leaky_slope * g1.v[e];
          } else {
#pragma unroll
            for (int e = 0; e < 8; ++e) dv.v[e] = o.v[e] > 0.f ? g1.v[e] : 0.f;
          }
        } else if (d.f32_out) {
          // half-tanh (LRP): out = res + 0.5 tanh(v)  =>  dv = g * 0.5 * (1 - tanh^2),  tanh = 2 (out - res)
          const float* po = d.f32_out + r * d.f32_cs + d.f32_coff + c0;
          const float* pr = d.f32_res + r * d.f32_cs + d.f32_coff + c0;
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            const float t = 2.0f * (po[e] - pr[e]);
            dv.v[e] = g1.v[e] * 0.5f * (1.0f - t * t);
          }
        }
#pragma unroll
        for (int e = 0; e < 8; ++e) acc[0][e] += dv.v[e];
        if (d.dv.hi) store_h8(const_cast<__half*>(reinterpret_cast<const __half*>(d.dv.hi)) + r * d.dv.cs + d.dv.coff + c0, dv, status);
        if (d.dres.hi) {
          __half* p = const_cast<__half*>(reinterpret_cast<const __half*>(d.dres.hi)) + r * d.dres.cs + d.dres.coff + c0;
          F8 t = load_h8(p);
#pragma unroll
          for (int e = 0; e < 8; ++e) t.v[e] += g1.v[e];
          store_h8(p, t, status);
        }
      }
    }
    if (d.partial) block_colsum_store<3>(acc, s_red, d.partial, d.c, live ? c0 : d.c);
  }
}

int epi_bwd_launch(const crdr_epi_bwd_desc* dp, cudaStream_t st) {
  const crdr_epi_bwd_desc& d = *dp;
  if (!d.g.hi || d.m <= 0 || d.c <= 0 || d.c % 8 || d.g.cs % 8 || d.g.coff % 8 || d.blocks < 1 || d.blocks > 4096 ||
      (d.dv.hi && (d.dv.cs % 8 || d.dv.coff % 8)) || (d.dres.hi && (d.dres.cs % 8 || d.dres.coff % 8)) ||
      ((d.relu || d.scale || d.shift) && (!d.out.hi || d.out.cs % 8 || d.out.coff % 8)) ||
      (d.f32_out && (!d.f32_res || d.f32_cs % 4 || d.f32_coff % 4))) {
    set_error("epi_bwd: missing operand or channel counts / strides / offsets not multiples of 8");
    return CRDR_ERR_BAD_SHAPE;
  }
  uint32_t* status = device_status_word();
  if (!status) return CRDR_ERR_CUDA;
  epi_bwd_kernel<<<d.blocks, dim3(32, 8), 0, st>>>(d, status);
  return check_launch("epi_bwd_kernel");
}

// LeakyReLU in place on a single fp16 plane (CLIC21GVAEDiscriminator, clic21_gvae_discriminator.py:12-25)
__global__ void leaky_relu_kernel(__half* __restrict__ x, int64_t m, int c, int cs, int coff, float slope) {
  const int groups = c >> 3;
  const int64_t total = m * groups;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / groups;
    const int c0 = (int)(i % groups) * 8;
    __half* p = x + r * cs + coff + c0;
    F8 v = load_h8(p);
    uint32_t w[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float a = v.v[2 * k] > 0.f ? v.v[2 * k] : slope * v.v[2 * k];
      const float b = v.v[2 * k + 1] > 0.f ? v.v[2 * k + 1] : slope * v.v[2 * k + 1];
      const __half2 h = __floats2half2_rn(a, b);
      w[k] = *reinterpret_cast<const uint32_t*>(&h);
    }
    *reinterpret_cast<uint4*>(p) = make_uint4(w[0], w[1], w[2], w[3]);
  }
}

int leaky_relu_launch(crdr_planes x, int64_t m, int c, float slope, cudaStream_t st) {
  if (!x.hi || x.lo || m <= 0 || c <= 0 || c % 8 || x.cs % 8 || x.coff % 8) { set_error("leaky_relu: single fp16 plane, channels / stride / offset multiples of 8"); return CRDR_ERR_BAD_SHAPE; }
  int64_t blocks = (m * (c >> 3) + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  leaky_relu_kernel<<<(unsigned)blocks, 256, 0, st>>>(const_cast<__half*>(reinterpret_cast<const __half*>(x.hi)), m, c, x.cs, x.coff, slope);
  return check_launch("leaky_relu_kernel");
}

// Gradient w.r.t. an image held as 8-channel NHWC planes (the discriminator's input) added into the phase-packed gradient
// of the last up-convolution's output:  g[n, a, b, (ph*2+pw)*3 + c] += scale * g8[n, 2a+ph, 2b+pw, c]
__global__ void planes_grad_to_phases_kernel(const __half* __restrict__ g8, int g8_cs, int n, int hb, int wb, float scale,
                                             __half* __restrict__ g, int g_cs, uint32_t* status) {
  const int64_t total = (int64_t)n * hb * wb * 12;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int ch = (int)(i % 12);
    const int64_t pix = i / 12;
    const int b = (int)(pix % wb);
    const int a = (int)((pix / wb) % hb);
    const int64_t img = pix / ((int64_t)wb * hb);
    const int phase = ch / 3, c = ch % 3;
    const int64_t src = ((img * (2 * hb) + 2 * a + (phase >> 1)) * (2 * wb) + 2 * b + (phase & 1)) * g8_cs + c;
    float v = __half2float(g[pix * g_cs + ch]) + scale * __half2float(g8[src]);
    if (!(fabsf(v) <= 65504.f)) { atomicOr(status, kFlagOverflow); v = fminf(fmaxf(v, -65504.f), 65504.f); }
    g[pix * g_cs + ch] = __float2half_rn(v);
  }
}

int planes_grad_to_phases_launch(const void* g8, int g8_cs, int n, int hb, int wb, float scale, void* g, int g_cs, cudaStream_t st) {
  if (!g8 || !g || n <= 0 || hb <= 0 || wb <= 0 || g8_cs < 3 || g_cs < 12) { set_error("planes_grad_to_phases: bad arguments"); return CRDR_ERR_BAD_SHAPE; }
  uint32_t* status = device_status_word();
  if (!status) return CRDR_ERR_CUDA;
  const int64_t total = (int64_t)n * hb * wb * 12;
  int64_t blocks = (total + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  planes_grad_to_phases_kernel<<<(unsigned)blocks, 256, 0, st>>>((const __half*)g8, g8_cs, n, hb, wb, scale, (__half*)g, g_cs, status);
  return check_launch("planes_grad_to_phases_kernel");
}

// out[c] (+)= scale * sum over blocks of partial[b][which][c].  Block (32, 8): lane x owns a channel, the 8 rows take
// interleaved subsets of the row blocks and are combined through shared memory in a fixed order (deterministic).
__global__ void __launch_bounds__(256) colsum_finish_kernel(const float* __restrict__ partial, int blocks, int nsums, int which, int c,
                                                            float* out, float scale, int accumulate) {
  __shared__ float s_red[8][32];
  const int ch = blockIdx.x * 32 + threadIdx.x;
  float s = 0.f;
  if (ch < c)
    for (int b = threadIdx.y; b < blocks; b += 8) s += partial[((int64_t)b * nsums + which) * c + ch];
  s_red[threadIdx.y][threadIdx.x] = s;
  __syncthreads();
  if (threadIdx.y == 0 && ch < c) {
    float t = 0.f;
#pragma unroll
    for (int y = 0; y < 8; ++y) t += s_red[y][threadIdx.x];
    out[ch] = accumulate ? fmaf(t, scale, out[ch]) : t * scale;
  }
}

int colsum_finish_launch(const float* partial, int blocks, int nsums, int which, int c, float* out, float scale, int accumulate,
                         cudaStream_t st) {
  if (!partial || !out || blocks < 1 || c < 1 || which < 0 || which >= nsums) { set_error("colsum_finish: bad arguments"); return CRDR_ERR_BAD_SHAPE; }
  colsum_finish_kernel<<<(c + 31) / 32, dim3(32, 8), 0, st>>>(partial, blocks, nsums, which, c, out, scale, accumulate);
  return check_launch("colsum_finish_kernel");
}

// ----------------------------------------------------------------------------------------------
// NLAM gate (cheng_nlam.py:23-29) as its own step in training mode (the logits `a` are kept for the backward):
//   forward   out = (x + t * sigmoid(a)) * scale + shift
//   backward  g1 = g * scale;  dx += g1;  dt = g1 * sig;  da = g1 * t * sig * (1 - sig);  sums: g | g * pre
// ----------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) gate_kernel(const crdr_gate_desc d, int backward, uint32_t* status) {
  __shared__ float s_red[8 * 32 * 2 * 8];
  const int64_t rows_per = (d.m + gridDim.x - 1) / gridDim.x;
  const int64_t r0 = (int64_t)blockIdx.x * rows_per;
  const int64_t r1 = r0 + rows_per < d.m ? r0 + rows_per : d.m;
  const int groups = d.c >> 3;
  for (int cg0 = 0; cg0 < groups; cg0 += 32) {
    const int cg = cg0 + threadIdx.x;
    const int c0 = cg * 8;
    const bool live = cg < groups;
    float acc[2][8];
#pragma unroll
    for (int k = 0; k < 2; ++k)
#pragma unroll
      for (int e = 0; e < 8; ++e) acc[k][e] = 0.f;
    if (live) {
      const F8 sc = load_vec8(d.scale, c0, 1.0f), sh = load_vec8(d.shift, c0, 0.0f);
      for (int64_t r = r0 + threadIdx.y; r < r1; r += kBwRowsPerBlockY) {
        const F8 x = load_planes8(d.x, r, c0), t = load_planes8(d.t, r, c0), a = load_planes8(d.a, r, c0);
        F8 sig, pre;
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          sig.v[e] = sigmoidf_(a.v[e]);
          pre.v[e] = fmaf(t.v[e], sig.v[e], x.v[e]);
        }
        if (!backward) {
          F8 o;
#pragma unroll
          for (int e = 0; e < 8; ++e) o.v[e] = fmaf(pre.v[e], sc.v[e], sh.v[e]);
          const int64_t oo = r * d.out.cs + d.out.coff + c0;
          if (d.out.lo) {
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              __half h, l;
              split_f16(o.v[e], h, l, status);
              const_cast<__half*>(reinterpret_cast<const __half*>(d.out.hi))[oo + e] = h;
              const_cast<__half*>(reinterpret_cast<const __half*>(d.out.lo))[oo + e] = l;
            }
          } else {
            store_h8(const_cast<__half*>(reinterpret_cast<const __half*>(d.out.hi)) + oo, o, status);
          }
          if (d.out_f32) {
            float* p = d.out_f32 + r * d.out_f32_cs + d.out_f32_coff + c0;
#pragma unroll
            for (int e = 0; e < 8; ++e) p[e] = o.v[e];
          }
        } else {
          const F8 g = load_h8(reinterpret_cast<const __half*>(d.g.hi) + r * d.g.cs + d.g.coff + c0);
          F8 dt, da;
          __half* px = const_cast<__half*>(reinterpret_cast<const __half*>(d.dx.hi)) + r * d.dx.cs + d.dx.coff + c0;
          F8 dx = load_h8(px);
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            const float g1 = g.v[e] * sc.v[e];
            acc[0][e] += g.v[e];
            acc[1][e] = fmaf(g.v[e], pre.v[e], acc[1][e]);
            dx.v[e] += g1;
            dt.v[e] = g1 * sig.v[e];
            da.v[e] = g1 * t.v[e] * sig.v[e] * (1.0f - sig.v[e]);
          }
          store_h8(px, dx, status);
          store_h8(const_cast<__half*>(reinterpret_cast<const __half*>(d.dt.hi)) + r * d.dt.cs + d.dt.coff + c0, dt, status);
          store_h8(const_cast<__half*>(reinterpret_cast<const __half*>(d.da.hi)) + r * d.da.cs + d.da.coff + c0, da, status);
        }
      }
    }
    if (backward && d.partial) block_colsum_store<2>(acc, s_red, d.partial, d.c, live ? c0 : d.c);
  }
}

int gate_launch(const crdr_gate_desc* dp, int backward, cudaStream_t st) {
  const crdr_gate_desc& d = *dp;
  auto ok = [](const crdr_planes& p) { return p.hi && p.cs % 8 == 0 && p.coff % 8 == 0; };
  if (d.m <= 0 || d.c <= 0 || d.c % 8 || !ok(d.x) || !ok(d.t) || !ok(d.a) || d.blocks < 1 || d.blocks > 4096 ||
      (!backward && !ok(d.out)) || (backward && (!ok(d.g) || !ok(d.dx) || !ok(d.dt) || !ok(d.da)))) {
    set_error("gate: missing operand or channel counts / strides / offsets not multiples of 8");
    return CRDR_ERR_BAD_SHAPE;
  }
  uint32_t* status = device_status_word();
  if (!status) return CRDR_ERR_CUDA;
  gate_kernel<<<d.blocks, dim3(32, 8), 0, st>>>(d, backward, status);
  return check_launch("gate_kernel");
}

// ----------------------------------------------------------------------------------------------
// Rate term of the GaussianConditional (CompressAI GaussianConditional._likelihood + LowerBound, restated in
// oracle/shims/compressai; call site minnen20_charm_context_model.py:118):  loss += -coef * ln L(y + u; mu, sigma)
// ----------------------------------------------------------------------------------------------
__global__ void gauss_bwd_kernel(const crdr_gauss_bwd_desc d, uint32_t* status) {
  const int64_t total = (int64_t)d.n * d.hw * d.c;
  const float coef = d.coef_scale ? d.coef * d.coef_scale[0] : d.coef;   // device-resident rate weight (graph replay)
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int ch = (int)(i % d.c);
    const int64_t pix = i / d.c;
    const int64_t img = pix / d.hw, p = pix % d.hw;
    const float y = d.y[pix * d.y_cs + d.y_coff + ch];
    const float u = d.noise[(img * d.c_total + d.nchw_coff + ch) * d.hw + p];
    const float mu = d.ms[pix * d.ms_cs + d.mu_coff + ch];
    const float sg = d.ms[pix * d.ms_cs + d.sigma_coff + ch];
    const float s = fmaxf(sg, d.scale_bound);
    const float x = y + u - mu;
    const float v = fabsf(x);
    const float up = (0.5f - v) / s, lo = (-0.5f - v) / s;
    const float L = 0.5f * erfcf(-0.70710678118654752f * up) - 0.5f * erfcf(-0.70710678118654752f * lo);
    const float gL = -coef / fmaxf(L, d.lik_bound);        // LowerBound passes every negative gradient
    const float pu = 0.3989422804014327f * expf(-0.5f * up * up), pl = 0.3989422804014327f * expf(-0.5f * lo * lo);
    const float dLdv = -(pu - pl) / s;
    const float dLds = -(up * pu - lo * pl) / s;
    const float sgn = x > 0.f ? 1.f : (x < 0.f ? -1.f : 0.f);
    const float dx = gL * dLdv * sgn;
    float ds = gL * dLds;
    if (!(sg >= d.scale_bound) && !(ds < 0.f)) ds = 0.f;   // LowerBound on sigma
    float gp = 0.f;
    if (d.gpre.hi) gp = __half2float(reinterpret_cast<const __half*>(d.gpre.hi)[pix * d.gpre.cs + d.gpre.coff + ch]);
    float dy = dx + gp, dm = -dx;
    if (!(fabsf(dy) <= 65504.f) || !(fabsf(dm) <= 65504.f) || !(fabsf(ds) <= 65504.f)) {
      atomicOr(status, kFlagOverflow);
      dy = fminf(fmaxf(dy, -65504.f), 65504.f); dm = fminf(fmaxf(dm, -65504.f), 65504.f); ds = fminf(fmaxf(ds, -65504.f), 65504.f);
    }
    const_cast<__half*>(reinterpret_cast<const __half*>(d.dy.hi))[pix * d.dy.cs + d.dy.coff + ch] = __float2half_rn(dy);
    const_cast<__half*>(reinterpret_cast<const __half*>(d.dmu.hi))[pix * d.dmu.cs + d.dmu.coff + ch] = __float2half_rn(dm);
    const_cast<__half*>(reinterpret_cast<const __half*>(d.dsigma.hi))[pix * d.dsigma.cs + d.dsigma.coff + ch] = __float2half_rn(ds);
  }
}

int gauss_bwd_launch(const crdr_gauss_bwd_desc* dp, cudaStream_t st) {
  const crdr_gauss_bwd_desc& d = *dp;
  if (!d.y || !d.noise || !d.ms || !d.dy.hi || !d.dmu.hi || !d.dsigma.hi || d.n <= 0 || d.hw <= 0 || d.c <= 0) {
    set_error("gauss_bwd: missing operand or bad shape");
    return CRDR_ERR_BAD_SHAPE;
  }
  uint32_t* status = device_status_word();
  if (!status) return CRDR_ERR_CUDA;
  const int64_t total = (int64_t)d.n * d.hw * d.c;
  int64_t blocks = (total + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  gauss_bwd_kernel<<<(unsigned)blocks, 256, 0, st>>>(d, status);
  return check_launch("gauss_bwd_kernel");
}

// ----------------------------------------------------------------------------------------------
// MSE term on the phase-packed reconstruction (distortion_loss.py:40-46 on the output of the last up-convolution):
//   g[n, a, b, (ph*2+pw)*3 + c] = coef * (fake - real[n, c, 2a+ph, 2b+pw])   (zero outside the h x w crop, channels 12-15)
// ----------------------------------------------------------------------------------------------
__global__ void mse_bwd_kernel(const float* __restrict__ fake, int fake_cs, const float* __restrict__ real, int n, int hb, int wb,
                               int h, int w, float coef, __half* __restrict__ g, int g_cs, uint32_t* status) {
  const int64_t total = (int64_t)n * hb * wb * 16;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int ch = (int)(i & 15);
    const int64_t pix = i >> 4;
    const int b = (int)(pix % wb);
    const int a = (int)((pix / wb) % hb);
    const int img = (int)(pix / ((int64_t)wb * hb));
    float v = 0.f;
    if (ch < 12) {
      const int phase = ch / 3, c = ch % 3;
      const int yy = 2 * a + (phase >> 1), xx = 2 * b + (phase & 1);
      if (yy < h && xx < w) v = coef * (fake[pix * fake_cs + ch] - real[(((int64_t)img * 3 + c) * h + yy) * w + xx]);
    }
    if (!(fabsf(v) <= 65504.f)) { atomicOr(status, kFlagOverflow); v = fminf(fmaxf(v, -65504.f), 65504.f); }
    g[pix * g_cs + ch] = __float2half_rn(v);
  }
}

int mse_bwd_launch(const float* fake, int fake_cs, const float* real, int n, int hb, int wb, int h, int w, float coef, void* g,
                   int g_cs, cudaStream_t st) {
  if (!fake || !real || !g || n <= 0 || hb <= 0 || wb <= 0 || fake_cs < 16 || g_cs < 16) { set_error("mse_bwd: bad arguments"); return CRDR_ERR_BAD_SHAPE; }
  uint32_t* status = device_status_word();
  if (!status) return CRDR_ERR_CUDA;
  const int64_t total = (int64_t)n * hb * wb * 16;
  int64_t blocks = (total + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  mse_bwd_kernel<<<(unsigned)blocks, 256, 0, st>>>(fake, fake_cs, real, n, hb, wb, h, w, coef, (__half*)g, g_cs, status);
  return check_launch("mse_bwd_kernel");
}

// ----------------------------------------------------------------------------------------------
// Optimiser step on flat fp32 buffers (torch.optim.Adam semantics, rate_distortion_trainer.py:87) and the squared
// gradient norm (nn.utils.clip_grad_norm_, :85-86) as a fixed-order two-stage reduction.
// ----------------------------------------------------------------------------------------------
__global__ void adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                            int64_t count, float lr, float b1, float b2, float eps, float bc1, float bc2_sqrt,
                            const float* __restrict__ gscale_ptr, float gscale, const float* __restrict__ hyper) {
  const float gs = gscale_ptr ? gscale * gscale_ptr[0] : gscale;
  if (hyper) {   // device-resident schedule (graph replay); hyper[3] != 0: this step is skipped (non-finite / huge loss)
    if (hyper[3] != 0.f) return;
    lr = hyper[0]; bc1 = hyper[1]; bc2_sqrt = hyper[2];
  }
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < count; i += (int64_t)gridDim.x * blockDim.x) {
    const float gi = g[i] * gs;
    const float mi = fmaf(b1, m[i], (1.f - b1) * gi);
    const float vi = fmaf(b2, v[i], (1.f - b2) * gi * gi);
    m[i] = mi;
    v[i] = vi;
    const float denom = sqrtf(vi) / bc2_sqrt + eps;
    p[i] -= (lr / bc1) * (mi / denom);
  }
}

int adam_launch(float* p, const float* g, float* m, float* v, int64_t count, float lr, float b1, float b2, float eps, int step,
                const float* gscale_ptr, float gscale, const float* hyper, cudaStream_t st) {
  if (!p || !g || !m || !v || count <= 0 || (step < 1 && !hyper)) { set_error("adam: bad arguments"); return CRDR_ERR_BAD_SHAPE; }
  if (step < 1) step = 1;
  const float bc1 = 1.f - powf(b1, (float)step), bc2 = 1.f - powf(b2, (float)step);
  int64_t blocks = (count + 255) / 256;
  if (blocks > 148 * 32) blocks = 148 * 32;
  adam_kernel<<<(unsigned)blocks, 256, 0, st>>>(p, g, m, v, count, lr, b1, b2, eps, bc1, sqrtf(bc2), gscale_ptr, gscale, hyper);
  return check_launch("adam_kernel");
}

constexpr int kSumsqBlocks = 1024;
__global__ void sumsq_kernel(const float* __restrict__ x, int64_t count, float* __restrict__ partial) {
  __shared__ float s[256];
  const int64_t per = (count + gridDim.x - 1) / gridDim.x;
  const int64_t i0 = (int64_t)blockIdx.x * per, i1 = i0 + per < count ? i0 + per : count;
  float a = 0.f;
  for (int64_t i = i0 + threadIdx.x; i < i1; i += blockDim.x) a = fmaf(x[i], x[i], a);
  s[threadIdx.x] = a;
  __syncthreads();
  for (int k = 128; k > 0; k >>= 1) {
    if ((int)threadIdx.x < k) s[threadIdx.x] += s[threadIdx.x + k];
    __syncthreads();
  }
  if (threadIdx.x == 0) partial[blockIdx.x] = s[0];
}
__global__ void sumsq_finish_kernel(const float* __restrict__ partial, int blocks, float* out) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    float s = 0.f;
    for (int b = 0; b < blocks; ++b) s += partial[b];
    out[0] = s;
  }
}

int sumsq_launch(const float* x, int64_t count, float* partial1024, float* out, cudaStream_t st) {
  if (!x || !partial1024 || !out || count <= 0) { set_error("sumsq: bad arguments"); return CRDR_ERR_BAD_SHAPE; }
  sumsq_kernel<<<kSumsqBlocks, 256, 0, st>>>(x, count, partial1024);
  int rc = check_launch("sumsq_kernel");
  if (rc) return rc;
  sumsq_finish_kernel<<<1, 32, 0, st>>>(partial1024, kSumsqBlocks, out);
  return check_launch("sumsq_finish_kernel");
}

}  // namespace crdr
