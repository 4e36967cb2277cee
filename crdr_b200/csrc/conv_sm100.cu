// Implicit-GEMM convolution for sm_100a: tcgen05.mma (kind::f16) with TMEM accumulators.
//
//   out[m, co] = sum_k A[m, k] * W[co, k],   k = tap * Cin + ci,   m = pixel of the base grid
//
// * A (im2col of the NHWC fp16 activation planes) is gathered straight into the canonical
//   K-major / 128B-swizzled UMMA shared-memory layout by 128 producer threads with 16-byte
//   cp.async (zero-fill outside the image), so any kernel size / stride / transposed phase /
//   two-range channel concatenation is the same code path.
// * W tiles ([tile_n][64] fp16, K-major) arrive by TMA (cp.async.bulk.tensor.2d, SWIZZLE_128B)
//   issued by one thread and complete on the same mbarrier as the gather.
// * One thread issues tcgen05.mma.cta_group::1 (M=128, N=tile_n, K=16); accumulators live in TMEM.
//   F16X3 precision keeps fp32-class accuracy on fp16 tensor cores: operands are stored as
//   (hi, lo*2^11) fp16 pairs, D0 += Ahi*Bhi and D1 += Ahi*Blo + Alo*Bhi run as three MMAs per
//   K step, and the epilogue forms D0 + 2^-11 * D1.
// * Tensor-core accumulation truncates, and the bias grows with the length of the accumulate chain
//   (measured: 7e-6 of max|out| at K=8800 vs 2e-6 for fp32 FMA, enough to flip CDF indexes).  F16X3
//   therefore accumulates D0 in short chunks (kChunkKB K-blocks) that ping-pong between two TMEM
//   buffers; warps 4-7 drain each finished chunk with tcgen05.ld, add it to a running fp32 total
//   (round-to-nearest) kept in a fourth TMEM region via tcgen05.st, and hand the buffer back.
// * The same warps then run the epilogue: tcgen05.ld 32x32b.x16 -> fused bias / ReLU /
//   beta-bias / residual / sigmoid-gate / half-tanh / gain -> fp16 planes (+ optional fp32) NHWC.
//
// Deterministic: fixed K order, no atomics on data, no split-K; a given output pixel sees the same
// arithmetic whatever the batch size or tile position.
#include <atomic>
#include <cstdlib>
#include <mutex>
#include <unordered_map>

#include "common.cuh"
#include "sm100_device.cuh"

namespace crdr {

constexpr uint32_t kAPlaneBytes = kTileM * 128;
constexpr int kMaxStages = 10;
constexpr int kSlotKB = 2;                 // PATCH mode: K blocks of weights per ring slot (one full / empty barrier round per slot)
constexpr int kMaxPatchStages = 8;         // PATCH mode: halo-patch ring (deep for small-K launches: a patch is a whole tile's A operand)
constexpr int kThreads = 512;              // gather variant: warps 0-3 gather, 4-11 drain/epilogue, 12 TMA + TMEM alloc, 13 MMA, 14-15 idle
constexpr int kThreadsPatch = 512;         // patch variant, F16X1: warps 0-11 drain/epilogue, 12 TMA + TMEM alloc, 13 MMA, 14 halo
                                           // patches, 15 idle; setmaxnreg 384 * 152 + 128 * 56 == 65536.  F16X3 (register totals)
                                           // runs 384 threads: warps 0-7 epilogue at 216 registers, 8-10 control, 11 idle.
// epilogue warps = 4 TMEM lane quarters x column groups (the epilogue is latency bound: 12 warps issue more than 8 do, but
// the F16X3 drain totals do not fit the 152-register budget of a 12-warp layout)
__host__ __device__ constexpr int epi_groups(bool patch, bool three) { return (patch && !three) ? 3 : 2; }
constexpr int kChunkKB = 6;                // F16X3: K blocks per D0 accumulate chain (24 MMAs of K=16).  Measured conv error vs
                                           // float64, relative to max|out|: 2 -> 2.3e-7, 4 -> 3.0e-7, 6 -> 3.3e-7, 8 -> 4.5e-7, whole K
                                           // (8800) -> 6.9e-6; all codec parity tests stay byte-identical up to 8; step time on one box
                                           // 88.9 / 85.9 / 84.8 / 84.4 ms for 2 / 4 / 6 / 8
constexpr uint32_t kSmemLimit = 227 * 1024;
constexpr uint32_t kDynSmemMax = kSmemLimit - 8 * 1024;  // static smem: parameter cache (5 KB) + barriers

constexpr int kMaxLeanRes = 6;             // LEAN epilogue: residual slots (per warp group a private ring of 1-2 slots)
constexpr int kMaxCBlocks = 12;            // PATCH mode: 64-channel blocks of the (two-range) input

struct alignas(64) ConvKParams {
  CUtensorMap tm_hi;
  CUtensorMap tm_lo;
  CUtensorMap tm_in_hi;  // PATCH mode: 4-D (C, W, H, N) view of the input planes
  CUtensorMap tm_in_lo;
  CUtensorMap tm_out_hi; // LEAN epilogue: 4-D (C, W, H, N) views of the output / residual planes, box (32, 8, 16, 1)
  CUtensorMap tm_out_lo;
  CUtensorMap tm_res_hi;
  CUtensorMap tm_res_lo;
  crdr_conv_desc d;
  // PATCH mode geometry
  int32_t ph, pw, dh_min, dw_min, ncb, tiles_h, tiles_w, patch_stages;
  // PATCH mode with an input stride s (1 or 2): the taps are grouped into s*s parity classes; class c owns one halo patch
  // per channel block, loaded through a tensor map that steps s pixels (elementStrides), from pixel
  // (s*h0 + cls_h0[c], s*w0 + cls_w0[c]); its taps are tapoff[cls_end[c-1] .. cls_end[c]).
  int32_t ncls, in_mul;
  int32_t cls_end[4], cls_h0[4], cls_w0[4];
  int32_t cb_c0[kMaxCBlocks];
  int32_t cb_ksteps[kMaxCBlocks];   // PATCH: K steps (of 16 channels) of a channel block that hold real channels (1..4)
  uint32_t tapoff[CRDR_MAX_TAPS];  // PATCH: byte offset of a tap's start row inside the halo patch
  int32_t m_total, nkb, cin, k_real, nplanes, stages, tmem_cols, use_tma;
  int32_t vec_planes_out, vec_f32_out, vec_res_planes, vec_res_f32, vec_trunk;
  int32_t has_bias, has_add, has_affine, chunk_kb, trace, split, fast_epi, res_stage_pitch, out_stage_pitch, dbg;
  int32_t lean_swz, lean_res_slots;   // LEAN epilogue: staging units are SWIZZLE_64B (1) or linear (0); residual ring depth
  uint32_t* status;
};

// ----------------------------------------------------------------------------------------------
// Epilogue for 16 consecutive output channels of one output pixel.
// Split in two so the global loads of a chunk can be issued before its accumulator is read:
//   epi_load   : residual / trunk operands -> registers (raw fp16 / fp32 vectors)
//   epi_finish : fused arithmetic + stores
// Per-channel vectors (bias, beta bias, gain, shift) come from shared memory (loaded once per CTA).
// ----------------------------------------------------------------------------------------------
// Raw epilogue operands of one 16-channel chunk, NR 16-byte registers:
//   NR == 4 (F16X3): planes -> r[0..1] = 16 hi halfs, r[2..3] = 16 lo halfs;  fp32 residual -> 16 floats
//   NR == 2 (F16X1): hi halfs only (single-term tensors; an fp32 residual is read in epi_finish instead)

// streaming 16-byte load that does not allocate in the (tiny, smem-carved) L1
__device__ __forceinline__ uint4 ld_stream16(const void* p) {
  uint4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
               : "l"(p));
  return v;
}
template <int NR>
__device__ __forceinline__ void epi_load_planes(const crdr_planes& pl, int64_t opix, int co0, uint4 (&r)[NR]) {
  const int64_t o = opix * pl.cs + pl.coff + co0;
  const __half* ph = (const __half*)pl.hi + o;
  r[0] = ld_stream16(ph);
  r[1] = ld_stream16(ph + 8);
  if (NR == 4 && pl.lo) {
    const __half* plo = (const __half*)pl.lo + o;
    r[NR - 2] = ld_stream16(plo);
    r[NR - 1] = ld_stream16(plo + 8);
  }
}

template <int NR>
__device__ __forceinline__ void epi_load_res(const ConvKParams& P, int64_t opix, int co0, uint4 (&r)[NR]) {
  const crdr_conv_desc& d = P.d;
  if (d.mode == CRDR_EPI_NONE || !P.fast_epi) return;
  if (d.res_f32) {
    if (NR == 4) {
      const float* p = d.res_f32 + opix * d.res_f32_cs + d.res_f32_coff + co0;
#pragma unroll
      for (int q = 0; q < NR; ++q) r[q] = ld_stream16(p + 4 * q);
    }
  } else {
    epi_load_planes<NR>(d.res, opix, co0, r);
  }
}

template <int NR>
__device__ __forceinline__ void epi_load_trunk(const ConvKParams& P, int64_t opix, int co0, uint4 (&t)[NR]) {
  const crdr_conv_desc& d = P.d;
  if (d.mode != CRDR_EPI_GATE || !P.fast_epi) return;
  epi_load_planes<NR>(d.trunk, opix, co0, t);
}

__device__ __forceinline__ float plane_at(const crdr_planes& pl, int64_t o) {
  return pl.lo ? join_f16(((const __half*)pl.hi)[o], ((const __half*)pl.lo)[o]) : __half2float(((const __half*)pl.hi)[o]);
}

// 16 halfs (two uint4) [+ 16 lo halfs] -> 16 floats; everything stays in registers (static indexing only)
template <int NR>
__device__ __forceinline__ void unpack16(const uint4 (&p)[NR], bool has_lo, float (&o)[16]) {
  const uint32_t hw[8] = {p[0].x, p[0].y, p[0].z, p[0].w, p[1].x, p[1].y, p[1].z, p[1].w};
  const uint32_t lw[8] = {p[NR - 2].x, p[NR - 2].y, p[NR - 2].z, p[NR - 2].w, p[NR - 1].x, p[NR - 1].y, p[NR - 1].z, p[NR - 1].w};
  has_lo = has_lo && NR == 4;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const float2 a = unpack_h2(hw[i]);
    o[2 * i] = a.x;
    o[2 * i + 1] = a.y;
    if (has_lo) {
      const float2 b = unpack_h2(lw[i]);
      o[2 * i] = fmaf(b.x, kLoInv, a.x);
      o[2 * i + 1] = fmaf(b.y, kLoInv, a.y);
    }
  }
}

// s_par: [4][kMaxCout] = bias | add_vec | scale | shift, indexed by absolute output channel
// Generic (rare) epilogue: partial channel chunks, unaligned strides, fp32 residual in F16X1 launches.  Kept out of
// line so the unrolled hot path stays small (instruction-fetch stalls were ~1/3 of the epilogue warps' stall samples).
__device__ __noinline__ void epi_finish_generic(const ConvKParams* Pp, int64_t opix, int co0, const float* acc,
                                                const float* s_par) {
  const ConvKParams& P = *Pp;
  const crdr_conv_desc& d = P.d;
  const int nvalid = min(16, d.cout - co0);
  for (int e = 0; e < nvalid; ++e) {
    const int co = co0 + e;
    float res = 0.f, trunk = 0.f;
    if (d.mode != CRDR_EPI_NONE) {
      if (d.res_f32) res = d.res_f32[opix * d.res_f32_cs + d.res_f32_coff + co];
      else res = plane_at(d.res, opix * d.res.cs + d.res.coff + co);
      if (d.mode == CRDR_EPI_GATE) trunk = plane_at(d.trunk, opix * d.trunk.cs + d.trunk.coff + co);
    }
    const float v = epilogue_math(acc[e], s_par[co], d.relu, s_par[kMaxCout + co], d.mode, res, trunk,
                                  s_par[2 * kMaxCout + co], s_par[3 * kMaxCout + co]);
    if (d.out_f32) d.out_f32[opix * d.out_f32_cs + d.out_f32_coff + co] = v;
    if (d.out.hi) {
      const int64_t o = opix * d.out.cs + d.out.coff + co;
      if (d.out.lo) {
        __half h, l;
        split_f16(v, h, l, P.status);
        ((__half*)d.out.hi)[o] = h;
        ((__half*)d.out.lo)[o] = l;
      } else {
        float x = v;
        if (fabsf(x) > 65504.0f) { atomicOr(P.status, kFlagOverflow); x = copysignf(65504.0f, x); }
        ((__half*)d.out.hi)[o] = __float2half_rn(x);
      }
    }
  }
}

// Hot path: a full 16-channel chunk with every operand vector-aligned (P.fast_epi); residual / trunk operands were
// preloaded into registers (rr / rt).
// ost != 0: the fp16 planes of this chunk go to the warp's staging tile in shared memory (32 bytes at `ost`, the lo
// plane `ost_plane` bytes further) and are written to global memory by the warp's coalesced copy-out.
// Per-launch epilogue switches, gathered once per thread into a register bit mask: read from the parameter bank inside the
// chunk loop, every test was an LDC -> ISETP -> BRA chain whose load latency three warps per scheduler cannot hide.
__device__ __forceinline__ uint32_t epi_flags(const ConvKParams& P) {
  const crdr_conv_desc& d = P.d;
  return (P.has_bias ? kEfBias : 0u) | (d.relu ? kEfRelu : 0u) | (P.has_add ? kEfAdd : 0u) | (P.has_affine ? kEfAffine : 0u) |
         (d.out_f32 ? kEfOutF32 : 0u) | (d.out.hi ? kEfOutHi : 0u) | (d.out.lo ? kEfOutLo : 0u) | (d.res_f32 ? kEfResF32 : 0u) |
         (d.res.lo ? kEfResLo : 0u) | (d.trunk.lo ? kEfTrunkLo : 0u) | ((uint32_t)d.mode << kEfModeShift);
}

template <int NR>
__device__ __forceinline__ void epi_finish(const ConvKParams& P, int64_t opix, int co0, const float (&acc)[16],
                                           const uint4 (&rr)[NR], const uint4 (&rt)[NR], const float* s_par,
                                           uint32_t ost, uint32_t ost_plane, uint32_t ef) {
  const int mode = (int)(ef >> kEfModeShift);
  const crdr_conv_desc& d = P.d;
  float res[16], trunk[16];
#pragma unroll
  for (int e = 0; e < 16; ++e) { res[e] = 0.f; trunk[e] = 0.f; }
  if (mode != CRDR_EPI_NONE) {
    if (NR == 4 && (ef & kEfResF32)) {
#pragma unroll
      for (int q = 0; q < NR; ++q) {
        res[4 * q] = __uint_as_float(rr[q].x); res[4 * q + 1] = __uint_as_float(rr[q].y);
        res[4 * q + 2] = __uint_as_float(rr[q].z); res[4 * q + 3] = __uint_as_float(rr[q].w);
      }
    } else {
      unpack16<NR>(rr, (ef & kEfResLo) != 0, res);
    }
    if (mode == CRDR_EPI_GATE) unpack16<NR>(rt, (ef & kEfTrunkLo) != 0, trunk);
  }
  // uniform (per-launch) switches outside the element loops; per-channel vectors as 128-bit shared loads
  float v[16];
#pragma unroll
  for (int e = 0; e < 16; ++e) v[e] = acc[e];
  const float4* par = reinterpret_cast<const float4*>(s_par + co0);
  constexpr int kVecStride = kMaxCout / 4;
  if (ef & kEfBias) {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const float4 t = par[q];
      v[4 * q] += t.x; v[4 * q + 1] += t.y; v[4 * q + 2] += t.z; v[4 * q + 3] += t.w;
    }
  }
  if (ef & kEfRelu) {
#pragma unroll
    for (int e = 0; e < 16; ++e) v[e] = fmaxf(v[e], 0.0f);
  }
  if (ef & kEfAdd) {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const float4 t = par[kVecStride + q];
      v[4 * q] += t.x; v[4 * q + 1] += t.y; v[4 * q + 2] += t.z; v[4 * q + 3] += t.w;
    }
  }
  if (mode == CRDR_EPI_RESIDUAL) {
#pragma unroll
    for (int e = 0; e < 16; ++e) v[e] += res[e];
  } else if (mode == CRDR_EPI_GATE) {
#pragma unroll
    for (int e = 0; e < 16; ++e) v[e] = fmaf(trunk[e], sigmoidf_(v[e]), res[e]);
  } else if (mode == CRDR_EPI_HALF_TANH) {
#pragma unroll
    for (int e = 0; e < 16; ++e) v[e] = fmaf(0.5f, tanhf(v[e]), res[e]);
  }
  if (ef & kEfAffine) {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const float4 sc = par[2 * kVecStride + q], sh = par[3 * kVecStride + q];
      v[4 * q] = fmaf(v[4 * q], sc.x, sh.x); v[4 * q + 1] = fmaf(v[4 * q + 1], sc.y, sh.y);
      v[4 * q + 2] = fmaf(v[4 * q + 2], sc.z, sh.z); v[4 * q + 3] = fmaf(v[4 * q + 3], sc.w, sh.w);
    }
  }
  if (ef & kEfOutF32) {
    float* o = d.out_f32 + opix * d.out_f32_cs + d.out_f32_coff + co0;
#pragma unroll
    for (int q = 0; q < 4; ++q) ((float4*)o)[q] = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
  }
  if (ef & kEfOutHi) {
    const bool want_lo = (ef & kEfOutLo) != 0;
    // one range check per chunk instead of one branch per element
    float amax = 0.f;
#pragma unroll
    for (int e = 0; e < 16; ++e) amax = fmaxf(amax, fabsf(v[e]));
    if (!(amax <= 65504.0f)) {
      atomicOr(P.status, kFlagOverflow);
#pragma unroll
      for (int e = 0; e < 16; ++e) v[e] = fminf(fmaxf(v[e], -65504.0f), 65504.0f);
    }
    uint32_t hw[8], lw[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const __half2 h = __floats2half2_rn(v[2 * i], v[2 * i + 1]);   // one cvt.rn.f16x2.f32
      hw[i] = *reinterpret_cast<const uint32_t*>(&h);
      if (want_lo) {
        const float2 hf = __half22float2(h);
        // (x - hi) * 2^11 is exact in fp32: fma(x, 2048, -2048*hi) has a single, exact rounding
        const __half2 l = __floats2half2_rn(fmaf(v[2 * i], kLoScale, -kLoScale * hf.x),
                                            fmaf(v[2 * i + 1], kLoScale, -kLoScale * hf.y));
        lw[i] = *reinterpret_cast<const uint32_t*>(&l);
      }
    }
    if (ost) {
      st_shared16(ost, hw[0], hw[1], hw[2], hw[3]);
      st_shared16(ost + 16u, hw[4], hw[5], hw[6], hw[7]);
      if (want_lo) {
        st_shared16(ost + ost_plane, lw[0], lw[1], lw[2], lw[3]);
        st_shared16(ost + ost_plane + 16u, lw[4], lw[5], lw[6], lw[7]);
      }
    } else {
      const int64_t o = opix * d.out.cs + d.out.coff + co0;
      __half* ph = (__half*)d.out.hi + o;
      ((uint4*)ph)[0] = make_uint4(hw[0], hw[1], hw[2], hw[3]);
      ((uint4*)ph)[1] = make_uint4(hw[4], hw[5], hw[6], hw[7]);
      if (want_lo) {
        __half* pl = (__half*)d.out.lo + o;
        ((uint4*)pl)[0] = make_uint4(lw[0], lw[1], lw[2], lw[3]);
        ((uint4*)pl)[1] = make_uint4(lw[4], lw[5], lw[6], lw[7]);
      }
    }
  }
}

// Output pixel (linear NHW index in the output tensor) of GEMM row m; -1 if m is past the end.
__device__ __forceinline__ int64_t out_pixel_of_row(const crdr_conv_desc& d, int m, int m_total) {
  if (m >= m_total) return -1;
  const int bw = m % d.wb;
  const int t = m / d.wb;
  const int bh = t % d.hb;
  const int n = t / d.hb;
  return ((int64_t)n * d.hout + (bh * d.out_stride + d.out_ph)) * d.wout + (bw * d.out_stride + d.out_pw);
}

// PATCH mode tiling: m-tile index -> (image, top-left base pixel); row r of the tile is pixel (h0 + r/8, w0 + r%8)
__device__ __forceinline__ void patch_tile_origin(const ConvKParams& P, int mt, int& n, int& h0, int& w0) {
  const int tw = mt % P.tiles_w;
  const int t = mt / P.tiles_w;
  const int th = t % P.tiles_h;
  n = t / P.tiles_h;
  h0 = th * kPatchTH;
  w0 = tw * kPatchTW;
}
__device__ __forceinline__ int64_t out_pixel_of_patch_row(const ConvKParams& P, int mt, int row) {
  int n, h0, w0;
  patch_tile_origin(P, mt, n, h0, w0);
  const int bh = h0 + (row >> 3), bw = w0 + (row & 7);
  const crdr_conv_desc& d = P.d;
  if (bh >= d.hb || bw >= d.wb || n >= d.n) return -1;
  return ((int64_t)n * d.hout + (bh * d.out_stride + d.out_ph)) * d.wout + (bw * d.out_stride + d.out_pw);
}

__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

// Pull the residual / trunk operands of one output pixel (channels [c_begin, c_end)) towards L2.
__device__ __forceinline__ void prefetch_epilogue_operands(const crdr_conv_desc& d, int64_t opix, int c_begin, int c_end) {
  if (d.mode == CRDR_EPI_NONE || opix < 0) return;
  if (d.res_f32) {
    const float* p = d.res_f32 + opix * d.res_f32_cs + d.res_f32_coff;
    for (int c = c_begin; c < c_end; c += 32) prefetch_l2(p + c);
  } else {
    const int64_t o = opix * d.res.cs + d.res.coff;
    for (int c = c_begin; c < c_end; c += 64) {
      prefetch_l2((const __half*)d.res.hi + o + c);
      if (d.res.lo) prefetch_l2((const __half*)d.res.lo + o + c);
    }
  }
  if (d.mode == CRDR_EPI_GATE) {
    const int64_t o = opix * d.trunk.cs + d.trunk.coff;
    for (int c = c_begin; c < c_end; c += 64) {
      prefetch_l2((const __half*)d.trunk.hi + o + c);
      if (d.trunk.lo) prefetch_l2((const __half*)d.trunk.lo + o + c);
    }
  }
}

#ifdef CRDR_TRACE_EVENTS
// bring-up event trace (CRDR_CONV_TRACE=2, CTA 0 only): (tag << 24 | payload, clock) records behind the status words.
// Three writer threads (epilogue warp 0 lane 0, MMA lane, patch producer) own one region each and count their records
// in a register, so a record is two fire-and-forget stores (an atomic slot counter cost ~600 cycles per event).
constexpr uint32_t kTraceRegion = 2600;  // records per region; 3 regions * 2600 * 8 B < 64 KB
__device__ __forceinline__ void trace_event(uint32_t* status, uint32_t region, uint32_t& count, uint32_t tag, uint32_t payload) {
  if (count < kTraceRegion) {
    uint32_t* rec = status + 64 + 2 * (region * kTraceRegion + count);
    rec[0] = (tag << 24) | (payload & 0xFFFFFFu);
    rec[1] = (uint32_t)clock64();
    ++count;
  }
}
#define CRDR_EV(...) __VA_ARGS__
#else
// event-trace hooks and the MMA-thread cycle counters compile to nothing in the product build: present, they cost 4 %
// of the step (registers, predicated instructions and code in the hot loops); build with
// CRDR_BUILD_TRACE=1 python -m crdr_b200.build to use tools/conv_events.py and the CRDR_CONV_TRACE=1 counters
#define CRDR_EV(...)
#endif

// ----------------------------------------------------------------------------------------------
// The tcgen05 kernel: persistent CTAs (one per SM), static round-robin tile schedule.
//   warps 0-3  : im2col gather producers (cp.async), also L2-prefetch the epilogue operands of the tile
//   warps 4-11 : drain D0 chunks into fp32 register totals (F16X3) and run the epilogue; two warps per
//                TMEM lane quarter, each owning half of the tile's column chunks
//   warp 12    : weight tiles by TMA, TMEM alloc / dealloc
//   warp 13    : MMA issue (one thread)
// Accumulators are double buffered per tile (F16X1: D0[2]; F16X3: D1[2] + the D0 chunk ping-pong), so the
// epilogue of tile j overlaps the main loop of tile j+1.
// MAXCH = column chunks (of 16) per drain warp in F16X3 mode (register totals); 0 selects F16X1.
// ----------------------------------------------------------------------------------------------
// PATCH (stride-1 inputs): instead of gathering one [128 x 64] im2col tile per (tap, channel block) from L2, one
// TMA box load brings the (16 + span_h) x (8 + span_w) halo patch of a 64-channel block into shared memory
// (zero fill outside the image) and every tap's A operand is just a different start row of that patch, addressed
// through the UMMA descriptor (group stride = patch row pitch).  L2 -> SM traffic for A drops by ~ntaps.
// CG2 (PATCH only): the kernel runs as clusters of two CTAs (a CTA pair on one TPC).  The pair works on two
// consecutive M tiles and the same N tile; the leader (cluster rank 0) issues
// tcgen05.mma.cta_group::2 with M = 256 over both CTAs' shared memory / TMEM, each CTA loads its own halo patch and
// HALF of the weight tile.  One 128 x N x 16 cta_group::1 MMA occupies the tensor pipe for ~N cycles (measured,
// twice the 4096 MAC/clk rate); the pair form is how sm_100 reaches the full rate.
template <int MAXCH, bool PATCH, bool CG2 = false, bool DIRECT = false, bool LEAN = false>
// (setmaxnreg re-distributes the CTA's LAUNCH allocation: F16X3 patch kernels are bounded at 384 threads so that ptxas
// gives them 168 registers per thread at launch, enough for 256 x 216 + 128 x 56 afterwards)
__global__ void __launch_bounds__(PATCH ? (MAXCH > 0 ? 384 : kThreadsPatch) : kThreads, 1)
conv_tcgen05_kernel(const __grid_constant__ ConvKParams P) {
  static_assert(PATCH || !CG2, "the CTA-pair form exists for the patch variant only");
  // DIRECT: F16X3 for short K (<= kDirectMaxKB K blocks): D0 is one accumulate chain per tile like an F16X1 accumulator,
  // so there is no chunk drain, no register totals, and the launch gets the 12-warp epilogue.
  static_assert(!DIRECT || (MAXCH == 0 && PATCH), "DIRECT is a patch-variant mode without register totals");
  // LEAN: TMA-in / TMA-out epilogue (see lean_unit_math) for the CTA-pair patch kernels without a chunk drain
  static_assert(!LEAN || (MAXCH == 0 && PATCH && CG2), "LEAN needs the CTA-pair patch variant without register totals");
  constexpr bool three = MAXCH > 0 || DIRECT;
  constexpr bool drain = MAXCH > 0;
  constexpr int kEpiWarp0 = PATCH ? 0 : 4;      // first of the 8 drain / epilogue warps (a multiple of 4: TMEM lane quarters)
  constexpr int kEpiGroups = epi_groups(PATCH, drain);
  constexpr int kEpiWarps = 4 * kEpiGroups;
  constexpr int kTmaWarp = kEpiWarp0 + kEpiWarps;
  constexpr int kMmaWarp = kTmaWarp + 1;
  constexpr int kPatchWarp = kTmaWarp + 2;
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[kMaxStages];
  __shared__ __align__(8) uint64_t empty_bar[kMaxStages];
  __shared__ __align__(8) uint64_t d0_full_bar[2];    // F16X3: MMA -> drain warps, a D0 chunk is complete
  __shared__ __align__(8) uint64_t d0_empty_bar[2];   // F16X3: drain warps -> MMA, the D0 buffer may be overwritten
  __shared__ __align__(8) uint64_t acc_full_bar[2];   // F16X1: MMA -> epilogue, the tile accumulator is complete
  __shared__ __align__(8) uint64_t acc_empty_bar[2];  // epilogue -> MMA, the tile accumulator (D0 / D1) was consumed
  __shared__ __align__(8) uint64_t patch_full_bar[kMaxPatchStages];
  __shared__ __align__(8) uint64_t patch_empty_bar[kMaxPatchStages];
  __shared__ __align__(8) uint64_t lean_res_full[kMaxLeanRes];   // LEAN: residual unit landed (TMA)
  __shared__ __align__(8) uint64_t lean_res_empty[kMaxLeanRes];  // LEAN: the unit's four consumer warps have read it
  __shared__ uint32_t tmem_slot;
  __shared__ int s_dh[CRDR_MAX_TAPS + 1], s_dw[CRDR_MAX_TAPS + 1];
  __shared__ __align__(16) float s_par[4 * kMaxCout];

  asm volatile("griddepcontrol.launch_dependents;");  // the next kernel's CTAs may take over SMs as ours exit
  const crdr_conv_desc& d = P.d;
  // warp-uniform values are laundered through a lane-0 broadcast so the compiler keeps everything derived from them
  // (descriptors, barrier addresses, TMEM addresses) in uniform registers: the MMA issue loop is a serial instruction
  // stream and every R2UR / waterfall loop in it costs tensor-pipe time.
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  const int S = P.stages;
  constexpr int nplanes = three ? 2 : 1;
  const int BN = d.tile_n;
  const uint32_t cta_rank = CG2 ? __shfl_sync(0xffffffffu, cluster_ctarank(), 0) : 0u;
  const int tile0 = CG2 ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;  // tile schedule of this CTA (pair)
  const int tstep = CG2 ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  const int BNL = CG2 ? BN / 2 : BN;                                 // weight-tile rows held by this CTA
  const uint32_t b_bytes = (uint32_t)BNL * 128u;
  // shared memory: [patch stages (PATCH only)] [ring of S stages: (A tile, gather mode only) + B tile]
  const uint32_t patch_plane_bytes = PATCH ? (uint32_t)(P.ph * P.pw) * 128u : 0u;
  const uint32_t patch_stage_bytes = (patch_plane_bytes * nplanes + 1023u) & ~1023u;
  const uint32_t a_bytes = PATCH ? 0u : kAPlaneBytes;
  constexpr int KPS = PATCH ? kSlotKB : 1;   // K blocks per ring slot
  const uint32_t kb_bytes = (uint32_t)nplanes * (a_bytes + b_bytes);
  const uint32_t stage_bytes = (uint32_t)KPS * kb_bytes;
  const uint32_t smem_patch = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t smem_base = smem_patch + (PATCH ? (uint32_t)P.patch_stages * patch_stage_bytes : 0u);
  const uint32_t smem_res = smem_base + (uint32_t)P.stages * stage_bytes;  // residual staging slots (128 rows)
  const uint32_t smem_ost = smem_res + (uint32_t)kTileM * (uint32_t)P.res_stage_pitch;  // output staging (per warp)
  const int nkb = P.nkb;
  const int n_tiles = d.cout_pad / BN;
  const int m_tiles = PATCH ? d.n * P.tiles_h * P.tiles_w : (P.m_total + kTileM - 1) / kTileM;
  const int num_tiles = (CG2 ? (m_tiles + 1) / 2 : m_tiles) * n_tiles;  // CG2: tiles of the pair (M = 256)
#define CRDR_MTILE(TILE) (CG2 ? 2 * ((TILE) / n_tiles) + (int)cta_rank : (TILE) / n_tiles)
  const int chunk_kb = drain ? P.chunk_kb : nkb;
  const int nchunks = (nkb + chunk_kb - 1) / chunk_kb;  // D0 chunks per tile

  for (int i = threadIdx.x; i < 4 * kMaxCout; i += (int)blockDim.x) {
    const int which = i / kMaxCout, co = i % kMaxCout;
    const float* src = which == 0 ? d.bias : which == 1 ? d.add_vec : which == 2 ? d.scale : d.shift;
    s_par[i] = (src && co < d.cout) ? src[co] : (which == 2 ? 1.f : 0.f);
  }
  if (threadIdx.x < d.ntaps) {
    s_dh[threadIdx.x] = d.dh[threadIdx.x];
    s_dw[threadIdx.x] = d.dw[threadIdx.x];
  }
  if (threadIdx.x == 0) {
    s_dh[d.ntaps] = 0;  // tap index of the zero-padded K tail
    s_dw[d.ntaps] = 0;
    const uint32_t full_count = PATCH ? 1u : (P.use_tma ? 129u : 128u);
    for (int b = 0; b < kMaxPatchStages; ++b) {
      mbar_init(smem_u32(&patch_full_bar[b]), 1u);
      mbar_init(smem_u32(&patch_empty_bar[b]), 1u);
    }
    for (int s = 0; s < S; ++s) {
      mbar_init(smem_u32(&full_bar[s]), full_count);
      mbar_init(smem_u32(&empty_bar[s]), 1u);
    }
    for (int b = 0; b < kMaxLeanRes; ++b) {
      mbar_init(smem_u32(&lean_res_full[b]), 1u);
      mbar_init(smem_u32(&lean_res_empty[b]), 4u);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(smem_u32(&d0_full_bar[b]), 1u);
      mbar_init(smem_u32(&d0_empty_bar[b]), (CG2 ? 2u : 1u) * kEpiWarps);   // one arrive per drain warp (of both CTAs)
      mbar_init(smem_u32(&acc_full_bar[b]), 1u);
      mbar_init(smem_u32(&acc_empty_bar[b]), (CG2 ? 2u : 1u) * kEpiWarps);
    }
    fence_barrier_init();
  }
  if (CG2) cluster_sync_all();  // both CTAs resident, barriers initialised, before any remote arrive / pair alloc
  if (warp == kTmaWarp) {
    if (lane == 0 && (P.use_tma || PATCH)) {
      prefetch_tmap(&P.tm_hi);
      if (three) prefetch_tmap(&P.tm_lo);
      if (PATCH) {
        prefetch_tmap(&P.tm_in_hi);
        if (three) prefetch_tmap(&P.tm_in_lo);
      }
      if (LEAN) {
        prefetch_tmap(&P.tm_out_hi);
        if (three) prefetch_tmap(&P.tm_out_lo);
        if (d.mode == CRDR_EPI_RESIDUAL) {
          prefetch_tmap(&P.tm_res_hi);
          if (three) prefetch_tmap(&P.tm_res_lo);
        }
      }
    }
    __syncwarp();
    if (CG2) tmem_alloc_cg2(smem_u32(&tmem_slot), 512u);
    else tmem_alloc(smem_u32(&tmem_slot), 512u);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, tmem_slot, 0);
  // Programmatic dependent launch: everything above (parameter cache, barriers, TMEM, cluster sync) overlapped the tail
  // of the previous kernel in the stream; its outputs (our inputs / residuals) are visible only after this wait.
  asm volatile("griddepcontrol.wait;" ::: "memory");
  // TMEM columns.  F16X3: D0 chunk ping-pong at 0 / 128, D1 per-tile double buffer at 256 / 384.
  //                F16X1: per-tile accumulator double buffer at 0 / 256.
  // Dependent tcgen05.mma on one accumulator issue only every ~207 cycles (measured), longer than a 128 x N x 16
  // MMA takes, so long-K launches (P.split) trade the per-tile double buffer for independent accumulate chains:
  //   F16X3 split: D0 ping-pong at 0 / 128, D1a (Ahi*Blo) at 256, D1b (Alo*Bhi) at 384  -> three chains
  //   F16X1 split: even K steps at 0, odd K steps at 256                                -> two chains
  constexpr uint32_t kAccStride = three ? 128u : 256u;
  constexpr uint32_t kD1Base = 256u;
#ifdef CRDR_TRACE_EVENTS
  const bool split = P.split != 0;   // tuning knobs (CRDR_CONV_SPLIT, CRDR_EPI_DEBUG) exist in the trace build only:
  const int dbg = P.dbg;             // their per-MMA / per-chunk tests cost issue slots in the product kernel
#else
  constexpr bool split = false;
  constexpr int dbg = 0;
#endif
  const int TB = split ? 1 : 2;  // per-tile accumulator buffers

  // Register re-balancing between the warp groups (512 threads x 128 = the whole register file):
  // 128 * (72 + 192 + 192 + 56) = 65536.
  // (gather variant only; each setmaxnreg sits at the top of its role branch)
  if (!PATCH && warp < 4) {
   asm volatile("setmaxnreg.dec.sync.aligned.u32 72;");
   {
    // ------------------------------------------------------------------ gather (im2col) producers
    const int t = threadIdx.x;
    const int j = t & 7;        // 16-byte chunk of the 128-byte K row
    const int rbase = t >> 3;   // rows rbase + 16*i
    const uint32_t row_off = (uint32_t)(rbase >> 3) * 1024u + (uint32_t)(rbase & 7) * 128u +
                             (uint32_t)((j ^ (rbase & 7)) << 4);
    const __half* in_hi = (const __half*)d.in.hi;
    const __half* in_lo = (const __half*)d.in.lo;
    const __half* w_hi = (const __half*)d.w_hi;
    const __half* w_lo = (const __half*)d.w_lo;
    const int cin = P.cin;
    const int lookahead = S - 1;
    int g = 0;     // K blocks issued so far (all tiles)
    int gpub = 0;  // K blocks published to the MMA thread
    for (int tile = tile0; tile < num_tiles; tile += tstep) {
      const int m0 = (tile / n_tiles) * kTileM;
      const int n0 = (tile % n_tiles) * BN;
      int64_t rowoff[8];   // element offset of the row's tap-(0,0) input pixel (channel coff included)
      uint32_t tapmask[8]; // bit t set <=> tap t of this row lies inside the image
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int m = m0 + rbase + 16 * i;
        rowoff[i] = 0;
        tapmask[i] = 0u;
        if (m < P.m_total) {
          const int bw = m % d.wb;
          const int tt = m / d.wb;
          const int bh = tt % d.hb;
          const int n = tt / d.hb;
          const int ih0 = bh * d.in_stride, iw0 = bw * d.in_stride;
          rowoff[i] = ((int64_t)(n * d.hin + ih0) * d.win + iw0) * d.in.cs + d.in.coff;
          for (int tp = 0; tp < d.ntaps; ++tp) {
            const int ih = ih0 + s_dh[tp], iw = iw0 + s_dw[tp];
            if ((unsigned)ih < (unsigned)d.hin && (unsigned)iw < (unsigned)d.win) tapmask[i] |= 1u << tp;
          }
          if (d.mode != CRDR_EPI_NONE) {
            // pull this row's epilogue operands towards L2 while the main loop runs (line j of the row segment)
            const int64_t opix = ((int64_t)n * d.hout + (bh * d.out_stride + d.out_ph)) * d.wout + (bw * d.out_stride + d.out_pw);
            if (d.res_f32) {
              if (j * 32 < BN) prefetch_l2(d.res_f32 + opix * d.res_f32_cs + d.res_f32_coff + n0 + j * 32);
            } else if (j * 64 < BN) {
              const int64_t o = opix * d.res.cs + d.res.coff + n0 + j * 64;
              prefetch_l2((const __half*)d.res.hi + o);
              if (d.res.lo) prefetch_l2((const __half*)d.res.lo + o);
            }
            if (d.mode == CRDR_EPI_GATE && j * 64 < BN) {
              const int64_t o = opix * d.trunk.cs + d.trunk.coff + n0 + j * 64;
              prefetch_l2((const __half*)d.trunk.hi + o);
              if (d.trunk.lo) prefetch_l2((const __half*)d.trunk.lo + o);
            }
          }
        }
      }
      int tap = 0, c = j * 8;  // position of this thread's chunk inside K for the current block
      while (c >= cin) { c -= cin; ++tap; }
      for (int kb = 0; kb < nkb; ++kb, ++g) {
        if (g - gpub >= lookahead) {
          // the block issued `lookahead` iterations ago has landed: publish it to the MMA thread
          switch (lookahead) {
            case 1: cp_async_wait<0>(); break;
            case 2: cp_async_wait<1>(); break;
            case 3: cp_async_wait<2>(); break;
            case 4: cp_async_wait<3>(); break;
            default: cp_async_wait<4>(); break;
          }
          fence_proxy_async();
          mbar_arrive(smem_u32(&full_bar[gpub % S]));
          ++gpub;
        }
        const int s = g % S;
        mbar_wait(smem_u32(&empty_bar[s]), ((uint32_t)(g / S) & 1u) ^ 1u, P.status);
        const uint32_t stage = smem_base + (uint32_t)s * stage_bytes;
        const int tapc = min(tap, d.ntaps);  // == ntaps in the zero-padded K tail: no mask bit is set there
        const int chan = (c < d.seg0_len) ? d.seg0_off + c : d.seg1_off + (c - d.seg0_len);
        const int64_t delta = (int64_t)((s_dh[tapc] * d.win + s_dw[tapc]) * d.in.cs + chan);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const bool ok = (tapmask[i] >> tapc) & 1u;
          const int64_t off = ok ? rowoff[i] + delta : 0;
          const uint32_t dst = stage + row_off + (uint32_t)i * 2048u;
          cp_async16(dst, in_hi + off, ok ? 16u : 0u);
          if (three) cp_async16(dst + kAPlaneBytes, in_lo + off, ok ? 16u : 0u);
        }
        if (!P.use_tma) {
          const uint32_t bst = stage + (uint32_t)nplanes * kAPlaneBytes;
          for (int i = 0; i < BN / 16; ++i) {
            const int64_t off = (int64_t)(n0 + rbase + 16 * i) * d.k_pad + (int64_t)kb * kKBlk + j * 8;
            const uint32_t dst = bst + row_off + (uint32_t)i * 2048u;
            cp_async16(dst, w_hi + off, 16u);
            if (three) cp_async16(dst + b_bytes, w_lo + off, 16u);
          }
        }
        cp_async_commit();
        c += kKBlk;
        while (c >= cin) { c -= cin; ++tap; }
      }
    }
    // publish what is still in flight
    cp_async_wait<0>();
    fence_proxy_async();
    for (; gpub < g; ++gpub) mbar_arrive(smem_u32(&full_bar[gpub % S]));
   }
  } else if (warp >= kEpiWarp0 && warp < kEpiWarp0 + kEpiWarps) {
    if (PATCH && !drain) asm volatile("setmaxnreg.inc.sync.aligned.u32 152;");
    else if (PATCH) asm volatile("setmaxnreg.inc.sync.aligned.u32 216;");
    else asm volatile("setmaxnreg.inc.sync.aligned.u32 192;");
    if constexpr (LEAN) {
    // ------------------------------------------------------------------ LEAN epilogue: 32-channel units, TMA in / out
    const int q = warp & 3;                    // TMEM lane quarter this warp may access
    const int cgrp = (warp - kEpiWarp0) >> 2;  // warp group: units cgrp, cgrp + kEpiGroups, ... of every tile
    const int row = q * 32 + lane;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16);
    const int units = BN >> 5;
    const uint32_t unit_bytes = (uint32_t)nplanes * kLeanUnitPlane;
    // Residual units: every warp group owns a private ring of D slots, filled in the group's own unit order.  (A ring
    // shared by the groups races: a fast group can run two phases ahead of a slot's mbarrier, and a parity wait cannot
    // tell phase k from phase k-2.)
    const int D = P.lean_res_slots;
    const uint32_t res_ring = smem_res + (uint32_t)(cgrp * D) * unit_bytes;
    const uint32_t stage_unit = smem_res + (uint32_t)(kEpiGroups * D) * unit_bytes + (uint32_t)cgrp * unit_bytes;
    int rslot = 0;            // this group's next residual slot and its phase
    uint32_t rpar = 0u;
    // the row's four 16-byte pieces inside a unit (SWIZZLE_64B: piece index ^ address bits 7-8)
    const uint32_t sx = P.lean_swz ? (uint32_t)((row >> 1) & 3) : 0u;
    uint32_t choff[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) choff[j] = (uint32_t)row * 64u + (((uint32_t)j ^ sx) << 4);
    const uint32_t ef = epi_flags(P);
    const bool has_res = (int)(ef >> kEfModeShift) == CRDR_EPI_RESIDUAL;
    const bool issuer = q == 0 && lane == 0;   // issues the group's TMA stores
    auto arrive_leader = [&](uint64_t* bar) {
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(mapa_shared(smem_u32(bar), 0u));
    };
    int jt = 0;
    for (int tile = tile0; tile < num_tiles; tile += tstep, ++jt) {
      const int n0 = (tile % n_tiles) * BN;
      const int tb = jt & 1;
      mbar_wait(smem_u32(&acc_full_bar[tb]), (uint32_t)(jt >> 1) & 1u, P.status);
      tc_fence_after();
      if (cgrp >= units) {   // narrow tile: this group has no unit, but every epilogue warp hands the buffer back
        arrive_leader(&acc_empty_bar[tb]);
        continue;
      }
#pragma unroll 1
      for (int u = cgrp; u < units; u += kEpiGroups) {
        uint32_t r0[32], r1[32];
        const uint32_t col = (uint32_t)tb * kAccStride + (uint32_t)u * 32u;
        tmem_ld32_issue(lane_addr + col, r0);
        if (three) tmem_ld32_issue(lane_addr + kD1Base + col, r1);
        uint4 rh[4], rl[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) { rh[j] = make_uint4(0u, 0u, 0u, 0u); rl[j] = rh[j]; }
        if (has_res) {
          const int slot = cgrp * D + rslot;
          mbar_wait(smem_u32(&lean_res_full[slot]), rpar, P.status);
          const uint32_t rbase = res_ring + (uint32_t)rslot * unit_bytes;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            rh[j] = lds128(rbase + choff[j]);
            if (three) rl[j] = lds128(rbase + kLeanUnitPlane + choff[j]);
          }
          // The slot goes back to the TMA (async proxy) while these generic-proxy loads may still be queued behind bank
          // conflicts: without the proxy fence the refill of a slot the producer is already waiting for overtook them
          // (measured: wrong residual rows in the first unit of a tile, linear staging layout only).
          fence_proxy_async();
          __syncwarp();
          if (lane == 0) mbar_arrive(smem_u32(&lean_res_empty[slot]));
          if (++rslot == D) { rslot = 0; rpar ^= 1u; }
        }
        tmem_wait_ld();
        if (u + kEpiGroups >= units) arrive_leader(&acc_empty_bar[tb]);   // this warp's last read of the tile's accumulators
        float v[32];
#pragma unroll
        for (int e = 0; e < 32; ++e)
          v[e] = three ? fmaf(__uint_as_float(r1[e]), kLoInv, __uint_as_float(r0[e])) : __uint_as_float(r0[e]);
        uint32_t hw[16], lw[16];
        lean_unit_math<three>(v, rh, rl, s_par, n0 + u * 32, ef, P.status, hw, lw);
        if (issuer) bulk_wait_read0();              // the previous store of this group has left the staging unit
        bar_sync_named(1u + (uint32_t)cgrp, 128u);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          st_shared16(stage_unit + choff[j], hw[4 * j], hw[4 * j + 1], hw[4 * j + 2], hw[4 * j + 3]);
          if (three) st_shared16(stage_unit + kLeanUnitPlane + choff[j], lw[4 * j], lw[4 * j + 1], lw[4 * j + 2], lw[4 * j + 3]);
        }
        fence_proxy_async();                        // generic-proxy writes -> visible to the TMA (async proxy)
        bar_sync_named(1u + (uint32_t)cgrp, 128u);
        if (issuer) {
          int n, h0, w0;
          patch_tile_origin(P, CRDR_MTILE(tile), n, h0, w0);
          if (n < d.n) {   // a CTA pair's odd tail tile lies past the end: nothing to store
            const int c = d.out.coff + n0 + u * 32;
            tma_store_4d(&P.tm_out_hi, stage_unit, c, w0, h0, n);
            if (three) tma_store_4d(&P.tm_out_lo, stage_unit + kLeanUnitPlane, c, w0, h0, n);
          }
          bulk_commit();
        }
      }
    }
    if (issuer) bulk_wait0();
    } else {
    // ------------------------------------------------------------------ drain D0 chunks + epilogue
    const int q = warp & 3;                    // TMEM lane quarter this warp may access
    const int cgrp = (warp - kEpiWarp0) >> 2;  // which group of the tile's column chunks
    const int row = q * 32 + lane;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16);
    const int nch_tile = BN / 16;
    const int ch_per = (nch_tile + kEpiGroups - 1) / kEpiGroups;
    const int ch_begin = min(cgrp * ch_per, nch_tile);
    const int ch_end = min(ch_begin + ch_per, nch_tile);
    constexpr int NR = three ? 4 : 2;         // 16-byte registers per residual chunk
    // Residual operands are software-pipelined ACROSS tiles without spending registers on it: every thread copies
    // the residual bytes of its own output row for tile j+1 into a private shared-memory slot with cp.async while
    // tile j is being finished (a whole tile per SM in flight; the one-chunk-ahead register ring kept only 16 KB
    // in flight = 0.9 TB/s).  At the start of tile j+1 the slot is read into registers and immediately refilled
    // for tile j+2.  Each thread only ever touches its own slot, so no barrier is involved.
    const bool staged = P.res_stage_pitch != 0;
    const uint32_t ef = epi_flags(P);
    const uint32_t my_slot = smem_res + (uint32_t)row * (uint32_t)P.res_stage_pitch + (uint32_t)(ch_begin * NR * 16);
    float total[drain ? MAXCH * 16 : 1];
    uint32_t r0[16], r1[16];
    int gc = 0;  // D0 chunks seen so far (all tiles)

    // Tile cursor: (N tile, image, tile row, tile column) of the current tile, advanced by the persistent schedule's
    // stride with adds and carries only (the per-tile integer divisions were ~25 % of the epilogue's instructions).
    struct Cursor { int tile, nt, n, th, tw; };
    const int m_per = CG2 ? 2 : 1;                                  // M tiles per schedule step in the M direction
    const int step_nt = tstep % n_tiles;
    const int step_m = (tstep / n_tiles) * m_per;
    const int tw_div = PATCH ? P.tiles_w : 1, th_div = PATCH ? P.tiles_h : 1;
    const int a_tw = step_m % tw_div, a_th = (step_m / tw_div) % th_div, a_n = (step_m / tw_div) / th_div;
    auto cursor_at = [&](int tile) {
      Cursor c;
      c.tile = tile;
      c.nt = tile % n_tiles;
      const int mt = CRDR_MTILE(tile);
      c.tw = mt % tw_div;
      c.th = (mt / tw_div) % th_div;
      c.n = (mt / tw_div) / th_div;
      return c;
    };
    auto cursor_advance = [&](Cursor& c) {
      c.tile += tstep;
      if (!PATCH) return;
      c.nt += step_nt;
      int extra = 0;
      if (c.nt >= n_tiles) { c.nt -= n_tiles; extra = m_per; }
      c.tw += a_tw;
      if (c.tw >= tw_div) { c.tw -= tw_div; ++c.th; }
      c.th += a_th;
      if (c.th >= th_div) { c.th -= th_div; ++c.n; }
      c.n += a_n;
      for (; extra > 0; --extra)
        if (++c.tw == tw_div) { c.tw = 0; if (++c.th == th_div) { c.th = 0; ++c.n; } }
    };
    // output pixel (linear NHW index) of this thread's row in the cursor's tile; -1 outside the tensor
    auto cursor_pixel = [&](const Cursor& c) -> int64_t {
      if (c.tile >= num_tiles) return -1;
      if (!PATCH) return out_pixel_of_row(d, (c.tile / n_tiles) * kTileM + row, P.m_total);
      const int bh = c.th * kPatchTH + (row >> 3), bw = c.tw * kPatchTW + (row & 7);
      if (bh >= d.hb || bw >= d.wb || c.n >= d.n) return -1;
      return ((int64_t)c.n * d.hout + (bh * d.out_stride + d.out_ph)) * d.wout + (bw * d.out_stride + d.out_pw);
    };
    // Global <-> shared traffic of the epilogue is warp-cooperative: a lane that moved only its own row's 16-byte pieces
    // touched 32 different 128-byte lines per instruction and the LSU (one line per cycle) became the limit of every
    // memory-bound layer (measured ~1.3 TB/s).  Instead the 16-byte pieces of the warp's 32 rows x column range are
    // dealt out lane by lane in memory order (consecutive lanes = consecutive pieces of a row), staged in shared memory
    // in the per-row layout the owning lanes use, and handed over with __syncwarp().
    const int nmine = ch_end - ch_begin;
    const uint32_t warp_slot0 = smem_res + (uint32_t)(q * 32) * (uint32_t)P.res_stage_pitch + (uint32_t)(ch_begin * NR * 16);
    // walker over pieces i = j * 32 + lane of a [32 rows][ppr pieces] block: (row, p) advance without division
    struct Walk { int row0, p0, drow, dp, ppr; };
    auto make_walk = [&](int ppr) {
      Walk w;
      w.ppr = ppr > 0 ? ppr : 1;
      w.row0 = lane / w.ppr; w.p0 = lane % w.ppr; w.drow = 32 / w.ppr; w.dp = 32 % w.ppr;
      return w;
    };
    const Walk wk_planes = make_walk(2 * nmine);   // fp16 plane: 32 bytes per chunk
    // cp.async the residual operands of chunk c (16 channels) of the warp's 32 rows into the rows' slots; consecutive
    // lanes fetch consecutive 16-byte pieces of a row (2 or 4 pieces per row and plane).  No commit.
    // loop invariants of the refill / copy-out helpers, held in registers (same reason as the epilogue flag mask)
    const uint32_t rpitch = (uint32_t)P.res_stage_pitch;
    const __half* const res_hi = (const __half*)d.res.hi;
    const __half* const res_lo = (const __half*)d.res.lo;
    const int res_cs = d.res.cs, res_coff = d.res.coff;
    __half* const out_hi = (__half*)d.out.hi;
    __half* const out_lo = (__half*)d.out.lo;
    const int out_cs = d.out.cs, out_coff = d.out.coff;
    auto refill_chunk = [&](int c, int opi, int c0) {
      if (ef & kEfResF32) {
#pragma unroll
        for (int it = 0; it < 4; ++it) {
          const int i = it * 32 + lane, row = i >> 2, pp = i & 3;
          const int orow = __shfl_sync(0xffffffffu, opi, row);
          if (orow >= 0)
            cp_async16(warp_slot0 + (uint32_t)row * rpitch + (uint32_t)((c * NR + pp) * 16),
                       d.res_f32 + (int64_t)orow * d.res_f32_cs + d.res_f32_coff + c0 + c * 16 + 4 * pp, 16u);
        }
      } else {
        const bool has_lo = NR == 4 && (ef & kEfResLo);
#pragma unroll
        for (int it = 0; it < 2; ++it) {
          const int i = it * 32 + lane, row = i >> 1, pp = i & 1;
          const int orow = __shfl_sync(0xffffffffu, opi, row);
          if (orow >= 0) {
            const int64_t o = (int64_t)orow * res_cs + res_coff + c0 + c * 16 + 8 * pp;
            const uint32_t dst = warp_slot0 + (uint32_t)row * rpitch + (uint32_t)((c * NR + pp) * 16);
            cp_async16(dst, res_hi + o, 16u);
            if (has_lo) cp_async16(dst + 32u, res_lo + o, 16u);
          }
        }
      }
    };
    // coalesced copy-out of the warp's staged fp16 output planes
    const bool ostaged = P.out_stage_pitch != 0;
    const uint32_t opitch = (uint32_t)P.out_stage_pitch;
    const uint32_t ost_plane = 32u * opitch;
    const uint32_t ost_warp = smem_ost + (uint32_t)(warp - kEpiWarp0) * (d.out.lo ? 2u : 1u) * ost_plane;
    auto copy_out = [&](int64_t op, int n0) {
      __syncwarp();  // every lane's chunks are staged
      const int opi = (int)op;
      const int c0 = n0 + ch_begin * 16;
      const bool want_lo = (ef & kEfOutLo) != 0;
      int row = wk_planes.row0, pp = wk_planes.p0;
      for (int j = 0; j < wk_planes.ppr; ++j) {
        const int orow = __shfl_sync(0xffffffffu, opi, row);
        if (orow >= 0 && nmine > 0 && !(dbg & 1)) {
          const uint32_t src = ost_warp + (uint32_t)row * opitch + (uint32_t)pp * 16u;
          const int64_t o = (int64_t)orow * out_cs + out_coff + c0 + 8 * pp;
          uint4 v;
          asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(src));
          *reinterpret_cast<uint4*>(out_hi + o) = v;
          if (want_lo) {
            asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(src + ost_plane));
            *reinterpret_cast<uint4*>(out_lo + o) = v;
          }
        }
        pp += wk_planes.dp; row += wk_planes.drow;
        if (pp >= wk_planes.ppr) { pp -= wk_planes.ppr; ++row; }
      }
      __syncwarp();  // staging tile free for the next tile
    };
    Cursor cur = cursor_at(tile0);
    int64_t opix = cursor_pixel(cur);
    if (staged) {
      const int c0 = (PATCH ? cur.nt : cur.tile % n_tiles) * BN + ch_begin * 16;
      for (int c = 0; c < nmine; ++c) refill_chunk(c, (int)opix, c0);
      cp_async_commit();
    }
    // hand a TMEM buffer back to the MMA thread (of the pair's leader): one arrive per warp
    auto arrive_leader = [&](uint64_t* bar) {
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (CG2) mbar_arrive_cluster(mapa_shared(smem_u32(bar), 0u));
        else mbar_arrive(smem_u32(bar));
      }
    };

    int jt = 0;  // tiles processed by this CTA
    CRDR_EV(uint32_t ev_count = 0;)
    for (; cur.tile < num_tiles; ++jt) {
      const int n0 = (PATCH ? cur.nt : cur.tile % n_tiles) * BN;
      const int tb = split ? 0 : (jt & 1);
      Cursor nxt = cur;
      cursor_advance(nxt);
      const int64_t opix_next = cursor_pixel(nxt);
      CRDR_EV(const bool ev = P.trace == 2 && blockIdx.x == 0 && warp == kEpiWarp0 && lane == 0;)
      CRDR_EV(if (ev) trace_event(P.status, 0u, ev_count, 1u, (uint32_t)jt);)  // epilogue: tile start
      const int c0_next = (PATCH ? nxt.nt : nxt.tile % n_tiles) * BN + ch_begin * 16;
      if (staged) {
        cp_async_wait<0>();  // this lane's copies for `tile` have landed ...
        __syncwarp();        // ... and so have the other lanes' (they fill this lane's row)
      }
      if (drain) {
        for (int ch = 0; ch < nchunks; ++ch, ++gc) {
          const int b = gc & 1;
          mbar_wait(smem_u32(&d0_full_bar[b]), (uint32_t)(gc >> 1) & 1u, P.status);
          tc_fence_after();
          const uint32_t src = lane_addr + (uint32_t)b * kAccStride;
          // two tcgen05.ld in flight per wait; the buffer goes back to the MMA thread as soon as the last load has
          // landed in registers, before the adds (the drain round trip bounds short-N F16X3 launches)
#pragma unroll
          for (int c = 0; c < MAXCH; c += 2) {
            const bool has0 = ch_begin + c < ch_end, has1 = c + 1 < MAXCH && ch_begin + c + 1 < ch_end;
            if (has0) tmem_ld16_issue(src + (uint32_t)(ch_begin + c) * 16u, r0);
            if (has1) tmem_ld16_issue(src + (uint32_t)(ch_begin + c + 1) * 16u, r1);
            tmem_wait_ld();
            if (c + 2 >= MAXCH) arrive_leader(&d0_empty_bar[b]);
            if (has0) {
#pragma unroll
              for (int e = 0; e < 16; ++e)
                total[c * 16 + e] = (ch == 0) ? __uint_as_float(r0[e]) : total[c * 16 + e] + __uint_as_float(r0[e]);
            }
            if (has1) {
#pragma unroll
              for (int e = 0; e < 16; ++e)
                total[(c + 1 < MAXCH ? c + 1 : c) * 16 + e] =
                    (ch == 0) ? __uint_as_float(r1[e]) : total[(c + 1 < MAXCH ? c + 1 : c) * 16 + e] + __uint_as_float(r1[e]);
            }
          }
        }
      } else {
        mbar_wait(smem_u32(&acc_full_bar[tb]), (uint32_t)(jt / TB) & 1u, P.status);
        tc_fence_after();
        CRDR_EV(if (ev) trace_event(P.status, 0u, ev_count, 2u, (uint32_t)jt);)  // epilogue: accumulator full observed
      }
      // The chunk loop is deliberately NOT unrolled: one copy of the (large) fused epilogue body instead of up to
      // eight keeps the tile loop inside the instruction cache (the unrolled form ran at 7-19 cycles per instruction
      // with every line of the epilogue collecting stall samples).
      uint4 rtrunk[NR], rr[NR];
#pragma unroll 1
      for (int c = 0; c < nmine; ++c) {
        const int chn = ch_begin + c;
        // second chain of a split launch lives one accumulator stride further (D1b / odd K steps)
        if (dbg & 4) {
        } else if (three) tmem_ld16_issue(lane_addr + kD1Base + (uint32_t)tb * kAccStride + (uint32_t)chn * 16u, r1);
        else tmem_ld16_issue(lane_addr + (uint32_t)tb * kAccStride + (uint32_t)chn * 16u, r0);
        if (DIRECT) tmem_ld16_issue(lane_addr + (uint32_t)tb * kAccStride + (uint32_t)chn * 16u, r0);  // D0 of the tile
        else if (split) tmem_ld16_issue(lane_addr + (three ? kD1Base : 0u) + kAccStride + (uint32_t)chn * 16u, three ? r0 : r1);
        CRDR_EV(if (ev && c == 0) trace_event(P.status, 0u, ev_count, 16u, (uint32_t)jt);)  // chunk 0: TMEM load issued
        if (staged) {
#pragma unroll
          for (int q2 = 0; q2 < NR; ++q2) {
            asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];"
                         : "=r"(rr[q2].x), "=r"(rr[q2].y), "=r"(rr[q2].z), "=r"(rr[q2].w)
                         : "r"(my_slot + (uint32_t)((c * NR + q2) * 16)));
          }
          __syncwarp();
          // chunk c of every row has been read: refill it for the next tile (in flight during the rest of this tile)
          if (nxt.tile < num_tiles) refill_chunk(c, (int)opix_next, c0_next);
        } else if (opix >= 0) {
          epi_load_res<NR>(P, opix, n0 + chn * 16, rr);
        }
        if (opix >= 0) epi_load_trunk<NR>(P, opix, n0 + chn * 16, rtrunk);
        CRDR_EV(if (ev && c == 0) trace_event(P.status, 0u, ev_count, 17u, (uint32_t)jt);)  // chunk 0: residual read + refill issued
        tmem_wait_ld();
        CRDR_EV(if (ev && c == 0) trace_event(P.status, 0u, ev_count, 18u, (uint32_t)jt);)  // chunk 0: accumulator in registers
        float acc[16];
        if (three) {
          float tsel[16];
#pragma unroll
          for (int k = 0; k < (drain ? MAXCH : 1); ++k) {
            if (k == 0 || c == k) {
#pragma unroll
              for (int e = 0; e < 16; ++e) tsel[e] = DIRECT ? __uint_as_float(r0[e]) : total[(drain ? k : 0) * 16 + e];
            }
          }
#pragma unroll
          for (int e = 0; e < 16; ++e) {
            const float cross = (split && !DIRECT) ? __uint_as_float(r1[e]) + __uint_as_float(r0[e]) : __uint_as_float(r1[e]);
            acc[e] = fmaf(cross, kLoInv, tsel[e]);
          }
        } else {
#pragma unroll
          for (int e = 0; e < 16; ++e)
            acc[e] = split ? __uint_as_float(r0[e]) + __uint_as_float(r1[e]) : __uint_as_float(r0[e]);
        }
        if (opix >= 0 && !(dbg & 2)) {
          if (P.fast_epi)
            epi_finish<NR>(P, opix, n0 + chn * 16, acc, rr, rtrunk, s_par,
                           ostaged ? ost_warp + (uint32_t)lane * opitch + (uint32_t)c * 32u : 0u, ost_plane, ef);
          else epi_finish_generic(&P, opix, n0 + chn * 16, acc, s_par);
        }
        CRDR_EV(if (ev && c == 0) trace_event(P.status, 0u, ev_count, 19u, (uint32_t)jt);)  // chunk 0: fused epilogue done
      }
      if (staged) cp_async_commit();
      CRDR_EV(if (ev) trace_event(P.status, 0u, ev_count, 3u, (uint32_t)jt);)  // epilogue: chunks done
      arrive_leader(&acc_empty_bar[tb]);
      if (ostaged) copy_out(opix, n0);
      CRDR_EV(if (ev) trace_event(P.status, 0u, ev_count, 4u, (uint32_t)jt);)  // epilogue: copy-out done
      cur = nxt;
      opix = opix_next;
    }
    }  // !LEAN
  } else {
   asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
   if (warp == kTmaWarp) {
    // ------------------------------------------------------------------ weight tiles by TMA
    if (lane == 0 && (P.use_tma || PATCH)) {
      int g = 0;
      CRDR_EV(const bool trp = P.trace && blockIdx.x == 0;)
      CRDR_EV(long long tp_wait = 0, tp0 = 0, tp_all0 = trp ? clock64() : 0;)
      int s = 0;
      uint32_t empty_par = 1u;  // a fresh barrier passes a wait on the previous phase
      uint32_t bst = smem_base + (uint32_t)nplanes * a_bytes;
      for (int tile = tile0; tile < num_tiles; tile += tstep) {
        const int n0 = (tile % n_tiles) * BN;
        int kk = 0;  // K block inside the ring slot
        for (int kb = 0; kb < nkb; ++kb, ++g) {
          const uint32_t bar = smem_u32(&full_bar[s]);
          if (kk == 0) {
            CRDR_EV(if (trp) tp0 = clock64();)
            mbar_wait(smem_u32(&empty_bar[s]), empty_par, P.status);
            CRDR_EV(if (trp) tp_wait += clock64() - tp0;)
            // the slot's K blocks (both halves of the weight tile for a CTA pair) complete on one (the leader's) barrier
            const uint32_t kcount = (uint32_t)min(KPS, nkb - kb);
            if (!CG2 || cta_rank == 0) mbar_arrive_expect_tx(bar, (CG2 ? 2u : 1u) * kcount * (uint32_t)nplanes * b_bytes);
          }
          const uint32_t dst = bst + (uint32_t)kk * kb_bytes;
          if (CG2) {
            const uint32_t lbar = mapa_shared(bar, 0u);
            const int nrow = n0 + (int)cta_rank * BNL;
            tma_load_2d_cg2(dst, &P.tm_hi, kb * kKBlk, nrow, lbar);
            if (three) tma_load_2d_cg2(dst + b_bytes, &P.tm_lo, kb * kKBlk, nrow, lbar);
          } else {
            tma_load_2d(dst, &P.tm_hi, kb * kKBlk, n0, bar);
            if (three) tma_load_2d(dst + b_bytes, &P.tm_lo, kb * kKBlk, n0, bar);
          }
          if (++kk == KPS || kb == nkb - 1) {
            kk = 0;
            bst += stage_bytes;
            if (++s == S) { s = 0; empty_par ^= 1u; bst = smem_base + (uint32_t)nplanes * a_bytes; }
          }
        }
      }
#ifdef CRDR_TRACE_EVENTS
      if (trp) {
        unsigned long long* c = reinterpret_cast<unsigned long long*>(P.status + 16);
        c[8] = (unsigned long long)(clock64() - tp_all0);
        c[9] = (unsigned long long)tp_wait;
      }
#endif
    }
  } else if (warp == kPatchWarp) {
    // ------------------------------------------------------------------ PATCH: halo patches by TMA (one thread)
    if (PATCH && lane == 0) {
      int pb = 0;
      uint32_t empty_par = 1u;
      CRDR_EV(uint32_t ev_count = 0;)
      for (int tile = tile0; tile < num_tiles; tile += tstep) {
        int n, h0, w0;
        patch_tile_origin(P, CRDR_MTILE(tile), n, h0, w0);  // CG2: a tile past the end has n == d.n -> zero fill
        for (int cb = 0; cb < P.ncb; ++cb) {
         for (int cls = 0; cls < P.ncls; ++cls, ++pb) {
          if (pb == P.patch_stages) { pb = 0; empty_par ^= 1u; }
          mbar_wait(smem_u32(&patch_empty_bar[pb]), empty_par, P.status);
          CRDR_EV(if (P.trace == 2 && blockIdx.x == 0) trace_event(P.status, 2u, ev_count, 12u, (uint32_t)cb);)  // patch: slot free, TMA issued
          const uint32_t bar = smem_u32(&patch_full_bar[pb]);
          const uint32_t dst = smem_patch + (uint32_t)pb * patch_stage_bytes;
          const int pw0 = w0 * P.in_mul + P.cls_w0[cls], ph0 = h0 * P.in_mul + P.cls_h0[cls];
          if (CG2) {
            if (cta_rank == 0) mbar_arrive_expect_tx(bar, 2u * (uint32_t)nplanes * patch_plane_bytes);
            const uint32_t lbar = mapa_shared(bar, 0u);
            tma_load_4d_cg2(dst, &P.tm_in_hi, P.cb_c0[cb], pw0, ph0, n, lbar);
            if (three) tma_load_4d_cg2(dst + patch_plane_bytes, &P.tm_in_lo, P.cb_c0[cb], pw0, ph0, n, lbar);
          } else {
            mbar_arrive_expect_tx(bar, (uint32_t)nplanes * patch_plane_bytes);
            tma_load_4d(dst, &P.tm_in_hi, P.cb_c0[cb], pw0, ph0, n, bar);
            if (three) tma_load_4d(dst + patch_plane_bytes, &P.tm_in_lo, P.cb_c0[cb], pw0, ph0, n, bar);
          }
         }
        }
      }
    }
  } else if (LEAN && warp == kPatchWarp + 1) {
    // ------------------------------------------------------------------ LEAN: residual units by TMA (one thread)
    if (lane == 0 && d.mode == CRDR_EPI_RESIDUAL) {
      const int units = BN >> 5;
      const uint32_t unit_bytes = (uint32_t)nplanes * kLeanUnitPlane;
      const int D = P.lean_res_slots;
      int gslot[kEpiGroups];        // per warp group: next slot of its private ring and the ring's phase
      uint32_t gpar[kEpiGroups];
#pragma unroll
      for (int g = 0; g < kEpiGroups; ++g) { gslot[g] = 0; gpar[g] = 1u; }
      for (int tile = tile0; tile < num_tiles; tile += tstep) {
        int n, h0, w0;
        patch_tile_origin(P, CRDR_MTILE(tile), n, h0, w0);   // a tile past the end has n == d.n -> zero fill
        const int c0 = d.res.coff + (tile % n_tiles) * BN;
        int g = 0;
        for (int u = 0; u < units; ++u) {
          int slot = 0;
          uint32_t par = 0u;
#pragma unroll
          for (int k = 0; k < kEpiGroups; ++k)   // static indexing keeps the cursors in registers
            if (k == g) { slot = g * D + gslot[k]; par = gpar[k]; if (++gslot[k] == D) { gslot[k] = 0; gpar[k] ^= 1u; } }
          mbar_wait(smem_u32(&lean_res_empty[slot]), par, P.status);
          const uint32_t bar = smem_u32(&lean_res_full[slot]);
          const uint32_t dst = smem_res + (uint32_t)slot * unit_bytes;
          mbar_arrive_expect_tx(bar, unit_bytes);
          tma_load_4d(dst, &P.tm_res_hi, c0 + u * 32, w0, h0, n, bar);
          if (three) tma_load_4d(dst + kLeanUnitPlane, &P.tm_res_lo, c0 + u * 32, w0, h0, n, bar);
          if (++g == kEpiGroups) g = 0;
        }
      }
    }
  } else if (warp == kMmaWarp) {
    // ------------------------------------------------------------------ MMA issue (whole warp, one elected lane issues)
    // The loop is a serial instruction stream that has to sustain one K block per few hundred cycles, so
    // everything that varies per K block is carried incrementally (ring slot / phase, tap, chunk position), the
    // descriptors are a constant plus a shifted shared-memory address, and all of it is warp-uniform (uniform
    // datapath); no division, no shared-memory loads, no warp-level sync.
    if (!CG2 || cta_rank == 0) {
      const bool elected = elect_one();
      const uint32_t idesc = umma_idesc_f16((uint32_t)BN, CG2 ? 256u : (uint32_t)kTileM);
      const int ntaps = d.ntaps;
      CRDR_EV(const bool tr = P.trace && blockIdx.x == 0 && elected;)
      CRDR_EV(long long t_full = 0, t_d0 = 0, t_acc = 0, t_patch = 0, t_mma = 0, t_commit = 0, t_all0 = tr ? clock64() : 0, t0 = 0;)
      const uint64_t desc_b = umma_desc_sw128(0u);                                            // + (addr >> 4)
      const uint64_t desc_a = PATCH ? umma_desc_sw128_rows(0u, (uint32_t)P.pw * 128u) : desc_b;
      const uint32_t a_planes = (uint32_t)nplanes * a_bytes;
      int g = 0, gc = 0, jt = 0;
      CRDR_EV(uint32_t ev_count = 0;)
      int s = 0;                 // weight / gather ring slot and its phase
      uint32_t ring_par = 0;
      uint32_t stage = smem_base;
      int pb = 0;                // patch buffer and its phase
      uint32_t patch_par = 0;
      uint32_t patch_addr = smem_patch;
      for (int tile = tile0; tile < num_tiles; tile += tstep, ++jt) {
        const int tb = split ? 0 : (jt & 1);
        if (jt >= TB) {
          // the epilogue of tile jt-TB must have consumed this per-tile accumulator buffer
          CRDR_EV(if (tr) t0 = clock64();)
          mbar_wait(smem_u32(&acc_empty_bar[tb]), (uint32_t)((jt - TB) / TB) & 1u, P.status);
          CRDR_EV(if (tr) t_acc += clock64() - t0;)
          tc_fence_after();
        }
        CRDR_EV(const bool ev = P.trace == 2 && blockIdx.x == 0 && elected;)
        CRDR_EV(if (ev) trace_event(P.status, 1u, ev_count, 8u, (uint32_t)jt);)  // MMA: accumulator buffer free
        const uint32_t d1 = tmem_base + kD1Base + (uint32_t)tb * kAccStride;
        int tap = 0, ck = 0, kk = 0;  // tap of this K block (PATCH), position inside the D0 chunk / the ring slot
        int cb = 0;                   // channel block of this K block (PATCH)
        int cls = 0;                  // parity class of the tap (PATCH with an input stride; one class otherwise)
        int tap_first = 0, tap_end = PATCH ? P.cls_end[0] : 0;   // taps of the class: [tap_first, tap_end)
        for (int kb = 0; kb < nkb; ++kb, ++g) {
          const bool chunk_first = ck == 0;
          const bool chunk_last = ck == chunk_kb - 1 || kb == nkb - 1;
          const int b = gc & 1;
          if (drain && chunk_first && gc >= 2) {
            // the drain warps must have emptied this D0 buffer (chunk gc-2) before it is overwritten
            CRDR_EV(if (tr) t0 = clock64();)
            mbar_wait(smem_u32(&d0_empty_bar[b]), (uint32_t)((gc - 2) >> 1) & 1u, P.status);
            CRDR_EV(if (tr) t_d0 += clock64() - t0;)
            tc_fence_after();
          }
          uint64_t a_hi, a_lo;
          if (PATCH) {
            // K block kb = (channel block, tap): the A operand is the patch shifted by the tap offset
            if (tap == tap_first) {
              CRDR_EV(if (tr) t0 = clock64();)
              mbar_wait(smem_u32(&patch_full_bar[pb]), patch_par, P.status);
              CRDR_EV(if (tr) t_patch += clock64() - t0;)
              CRDR_EV(if (ev) trace_event(P.status, 1u, ev_count, 10u, (uint32_t)jt);)  // MMA: patch landed
            }
            const uint32_t pa = patch_addr + P.tapoff[tap];
            a_hi = desc_a + (uint64_t)(pa >> 4);
            a_lo = desc_a + (uint64_t)((pa + patch_plane_bytes) >> 4);
          } else {
            a_hi = desc_a + (uint64_t)(stage >> 4);
            a_lo = desc_a + (uint64_t)((stage + kAPlaneBytes) >> 4);
          }
          if (kk == 0) {
            CRDR_EV(if (tr) t0 = clock64();)
            mbar_wait(smem_u32(&full_bar[s]), ring_par, P.status);
            CRDR_EV(if (tr) t_full += clock64() - t0;)
          }
          tc_fence_after();
          const uint32_t d0 = tmem_base + (uint32_t)(drain ? b : tb) * kAccStride;
          const uint32_t bsrc = stage + a_planes + (uint32_t)kk * kb_bytes;
          const uint64_t b_hi = desc_b + (uint64_t)(bsrc >> 4);
          const uint64_t b_lo = desc_b + (uint64_t)((bsrc + b_bytes) >> 4);
          CRDR_EV(if (tr) t0 = clock64();)
          const int nks = PATCH ? P.cb_ksteps[cb] : kKBlk / 16;  // K steps of this block that hold real channels
          if (elected) {
#pragma unroll
          for (int k = 0; k < kKBlk / 16; ++k) {
            if (k >= nks) break;
            const uint64_t adv = (uint64_t)(k * 2);  // 16 fp16 = 32 bytes = 2 descriptor units
            if (three) {
              umma_issue<CG2>(d0, a_hi + adv, b_hi + adv, idesc, (chunk_first && k == 0) ? 0u : 1u);
              umma_issue<CG2>(d1, a_hi + adv, b_lo + adv, idesc, (kb > 0 || k > 0) ? 1u : 0u);
              // split: the second cross term accumulates in its own chain (D1b)
              umma_issue<CG2>(split ? d1 + kAccStride : d1, a_lo + adv, b_hi + adv, idesc, (split && kb == 0 && k == 0) ? 0u : 1u);
            } else if (split) {
              // even / odd K steps alternate between two accumulators
              umma_issue<CG2>(d0 + (uint32_t)(k & 1) * kAccStride, a_hi + adv, b_hi + adv, idesc, (kb == 0 && k < 2) ? 0u : 1u);
            } else {
              umma_issue<CG2>(d0, a_hi + adv, b_hi + adv, idesc, (chunk_first && k == 0) ? 0u : 1u);
            }
          }
          }
          CRDR_EV(if (tr) { const long long t1 = clock64(); t_mma += t1 - t0; t0 = t1; })
          const bool slot_last = ++kk == KPS || kb == nkb - 1;
          if (slot_last && elected) umma_done<CG2>(smem_u32(&empty_bar[s]));
          if (PATCH) {
            if (++tap == tap_end) {  // all taps of this patch (channel block, parity class) issued: the buffer may be refilled
              if (elected) umma_done<CG2>(smem_u32(&patch_empty_bar[pb]));
              patch_addr += patch_stage_bytes;
              if (++pb == P.patch_stages) { pb = 0; patch_par ^= 1u; patch_addr = smem_patch; }
              if (tap == ntaps) {    // last class: next channel block
                tap = 0;
                cls = 0;
                ++cb;
              } else {
                ++cls;
              }
              tap_first = tap;
              tap_end = P.cls_end[cls];
            }
          }
          if (chunk_last) {
            CRDR_EV(if (ev && kb == nkb - 1) trace_event(P.status, 1u, ev_count, 9u, (uint32_t)jt);)  // MMA: last K block of the tile issued
            if (elected) umma_done<CG2>(smem_u32(drain ? &d0_full_bar[b] : &acc_full_bar[tb]));
            if (drain) ++gc;
            ck = 0;
          } else {
            ++ck;
          }
          CRDR_EV(if (tr) t_commit += clock64() - t0;)
          if (slot_last) {
            kk = 0;
            stage += stage_bytes;
            if (++s == S) { s = 0; ring_par ^= 1u; stage = smem_base; }
          }
        }
      }
#ifdef CRDR_TRACE_EVENTS
      if (tr) {  // bring-up counters (cycles): total, wait full, wait d0_empty, wait acc_empty, wait patch, k blocks
        unsigned long long* c = reinterpret_cast<unsigned long long*>(P.status + 16);
        c[0] = (unsigned long long)(clock64() - t_all0);
        c[1] = (unsigned long long)t_full; c[2] = (unsigned long long)t_d0; c[3] = (unsigned long long)t_acc;
        c[4] = (unsigned long long)t_patch; c[5] = (unsigned long long)g;
        c[6] = (unsigned long long)t_mma; c[7] = (unsigned long long)t_commit;
      }
#endif
    }
   }
  }
  tc_fence_before();
  __syncthreads();
  if (CG2) cluster_sync_all();  // the peer's TMEM and barriers are in use until both CTAs are done
  if (warp == kTmaWarp) {
    tc_fence_after();
    if (CG2) tmem_dealloc_cg2(tmem_base, 512u);
    else tmem_dealloc(tmem_base, 512u);
  }
#undef CRDR_MTILE
}

// ----------------------------------------------------------------------------------------------
// Scalar fp32 cross-check kernel over the same operands and epilogue (one thread per output element).
// ----------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) conv_simt_kernel(const __grid_constant__ ConvKParams P) {
  const crdr_conv_desc& d = P.d;
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int co = (int)(idx % d.cout);
  const int64_t m64 = idx / d.cout;
  if (m64 >= P.m_total) return;
  const int m = (int)m64;
  const int bw = m % d.wb;
  const int tt = m / d.wb;
  const int bh = tt % d.hb;
  const int n = tt / d.hb;
  const __half* in_hi = (const __half*)d.in.hi;
  const __half* in_lo = (const __half*)d.in.lo;
  const __half* w_hi = (const __half*)d.w_hi + (int64_t)co * d.k_pad;
  const __half* w_lo = d.w_lo ? (const __half*)d.w_lo + (int64_t)co * d.k_pad : nullptr;
  const bool three = P.nplanes == 2;
  float acc0 = 0.f, acc1 = 0.f;
  for (int tp = 0; tp < d.ntaps; ++tp) {
    const int ih = bh * d.in_stride + d.dh[tp], iw = bw * d.in_stride + d.dw[tp];
    if ((unsigned)ih >= (unsigned)d.hin || (unsigned)iw >= (unsigned)d.win) continue;
    const int64_t pbase = ((int64_t)(n * d.hin + ih) * d.win + iw) * d.in.cs + d.in.coff;
    for (int c = 0; c < P.cin; ++c) {
      const int chan = (c < d.seg0_len) ? d.seg0_off + c : d.seg1_off + (c - d.seg0_len);
      int k;
      if (d.k_order == 0) {
        k = tp * P.cin + c;
      } else {
        const int in0 = c < d.seg0_len;
        const int cs_ = in0 ? c : c - d.seg0_len;
        const int cb = (in0 ? 0 : (d.seg0_len + 63) / 64) + cs_ / 64;
        k = (cb * d.ntaps + tp) * 64 + (cs_ & 63);
      }
      const float ah = __half2float(in_hi[pbase + chan]);
      const float bhv = __half2float(w_hi[k]);
      acc0 = fmaf(ah, bhv, acc0);
      if (three) {
        const float al = __half2float(in_lo[pbase + chan]);
        const float bl = __half2float(w_lo[k]);
        acc1 = fmaf(ah, bl, acc1);
        acc1 = fmaf(al, bhv, acc1);
      }
    }
  }
  const float acc = fmaf(acc1, kLoInv, acc0);
  const int64_t opix = out_pixel_of_row(d, m, P.m_total);
  float res = 0.f, trunk = 0.f;
  if (d.mode != CRDR_EPI_NONE) {
    if (d.res_f32) {
      res = d.res_f32[opix * d.res_f32_cs + d.res_f32_coff + co];
    } else {
      const int64_t o = opix * d.res.cs + d.res.coff + co;
      res = d.res.lo ? join_f16(((const __half*)d.res.hi)[o], ((const __half*)d.res.lo)[o])
                     : __half2float(((const __half*)d.res.hi)[o]);
    }
    if (d.mode == CRDR_EPI_GATE) {
      const int64_t o = opix * d.trunk.cs + d.trunk.coff + co;
      trunk = d.trunk.lo ? join_f16(((const __half*)d.trunk.hi)[o], ((const __half*)d.trunk.lo)[o])
                         : __half2float(((const __half*)d.trunk.hi)[o]);
    }
  }
  const float v = epilogue_math(acc, d.bias ? d.bias[co] : 0.f, d.relu, d.add_vec ? d.add_vec[co] : 0.f, d.mode, res,
                                trunk, d.scale ? d.scale[co] : 1.f, d.shift ? d.shift[co] : 0.f);
  if (d.out_f32) d.out_f32[opix * d.out_f32_cs + d.out_f32_coff + co] = v;
  if (d.out.hi) {
    const int64_t o = opix * d.out.cs + d.out.coff + co;
    if (d.out.lo) {
      __half h, l;
      split_f16(v, h, l, P.status);
      ((__half*)d.out.hi)[o] = h;
      ((__half*)d.out.lo)[o] = l;
    } else {
      float x = v;
      if (fabsf(x) > 65504.0f) { atomicOr(P.status, kFlagOverflow); x = copysignf(65504.0f, x); }
      ((__half*)d.out.hi)[o] = __float2half_rn(x);
    }
  }
}

// ----------------------------------------------------------------------------------------------
// Host side: descriptor validation, tensor-map cache, launch
// ----------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
  });
  return fn;
}

// (shared with bottleneck_sm100.cu)
int weight_tensor_map(const void* ptr, int k_pad, int rows, int box_rows, CUtensorMap* out);
int input_tensor_map(const void* ptr, int cs, int w, int h, int n, int pw, int ph, CUtensorMap* out, int box_c = 64,
                     int swz = 128, int estride = 1);

struct TmKey {
  const void* ptr;
  int32_t k_pad, rows, box_rows, dev;
  bool operator==(const TmKey& o) const {
    return ptr == o.ptr && k_pad == o.k_pad && rows == o.rows && box_rows == o.box_rows && dev == o.dev;
  }
};
struct TmKeyHash {
  size_t operator()(const TmKey& k) const {
    size_t h = std::hash<const void*>()(k.ptr);
    h ^= std::hash<int64_t>()(((int64_t)k.k_pad << 32) ^ ((int64_t)k.rows << 12) ^ ((int64_t)k.box_rows << 4) ^ k.dev) +
         0x9e3779b97f4a7c15ULL + (h << 6) + (h >> 2);
    return h;
  }
};
static std::mutex g_tm_mutex;
static std::unordered_map<TmKey, CUtensorMap, TmKeyHash> g_tm_cache;

int weight_tensor_map(const void* ptr, int k_pad, int rows, int box_rows, CUtensorMap* out) {
  int dev = 0;
  cudaGetDevice(&dev);
  TmKey key{ptr, k_pad, rows, box_rows, dev};
  {
    std::lock_guard<std::mutex> lk(g_tm_mutex);
    auto it = g_tm_cache.find(key);
    if (it != g_tm_cache.end()) { *out = it->second; return CRDR_OK; }
  }
  EncodeTiledFn fn = encode_fn();
  if (!fn) { set_error("cuTensorMapEncodeTiled is unavailable (driver too old?)"); return CRDR_ERR_CUDA; }
  cuuint64_t gdim[2] = {(cuuint64_t)k_pad, (cuuint64_t)rows};
  cuuint64_t gstride[1] = {(cuuint64_t)k_pad * 2};
  cuuint32_t box[2] = {(cuuint32_t)kKBlk, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUtensorMap tm;
  CUresult r = fn(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(ptr), gdim, gstride, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed (CUresult %d)", (int)r); return CRDR_ERR_CUDA; }
  {
    std::lock_guard<std::mutex> lk(g_tm_mutex);
    g_tm_cache[key] = tm;
  }
  *out = tm;
  return CRDR_OK;
}

struct InKey {
  const void* ptr;
  int32_t cs, w, h, n, pw, ph, dev, box_c, swz;
  bool operator==(const InKey& o) const {
    return ptr == o.ptr && cs == o.cs && w == o.w && h == o.h && n == o.n && pw == o.pw && ph == o.ph && dev == o.dev &&
           box_c == o.box_c && swz == o.swz;
  }
};
struct InKeyHash {
  size_t operator()(const InKey& k) const {
    size_t h = std::hash<const void*>()(k.ptr);
    const int32_t v[9] = {k.cs, k.w, k.h, k.n, k.pw, k.ph, k.dev, k.box_c, k.swz};
    for (int i = 0; i < 9; ++i) h = h * 1000003u ^ (size_t)v[i];
    return h;
  }
};
static std::unordered_map<InKey, CUtensorMap, InKeyHash> g_in_cache;

// 4-D (C, W, H, N) tensor map over NHWC fp16 planes with a (box_c, pw, ph, 1) box, zero fill out of bounds (loads) /
// clipping (stores).  swz: 128 (halo patches, 64-channel boxes), 64 (LEAN units, 32-channel boxes) or 0 (linear).
int input_tensor_map(const void* ptr, int cs, int w, int h, int n, int pw, int ph, CUtensorMap* out,
                     int box_c, int swz, int estride) {
  int dev = 0;
  cudaGetDevice(&dev);
  InKey key{ptr, cs, w, h, n, pw, ph, dev, box_c, swz + 1000 * estride};
  {
    std::lock_guard<std::mutex> lk(g_tm_mutex);
    auto it = g_in_cache.find(key);
    if (it != g_in_cache.end()) { *out = it->second; return CRDR_OK; }
  }
  EncodeTiledFn fn = encode_fn();
  if (!fn) { set_error("cuTensorMapEncodeTiled is unavailable (driver too old?)"); return CRDR_ERR_CUDA; }
  cuuint64_t gdim[4] = {(cuuint64_t)cs, (cuuint64_t)w, (cuuint64_t)h, (cuuint64_t)n};
  cuuint64_t gstride[3] = {(cuuint64_t)cs * 2, (cuuint64_t)w * cs * 2, (cuuint64_t)h * w * cs * 2};
  // pixel step `estride` (a strided convolution's parity-class patch): the box spans pw * estride pixels, of which
  // every estride-th is loaded (cuTensorMapEncodeTiled: ceil(boxDim / elementStrides) elements per dimension)
  cuuint32_t box[4] = {(cuuint32_t)box_c, (cuuint32_t)(pw * estride), (cuuint32_t)(ph * estride), 1u};
  cuuint32_t estr[4] = {1, (cuuint32_t)estride, (cuuint32_t)estride, 1};
  CUtensorMap tm;
  const CUtensorMapSwizzle sw = swz == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : swz == 64 ? CU_TENSOR_MAP_SWIZZLE_64B
                                                                                    : CU_TENSOR_MAP_SWIZZLE_NONE;
  CUresult r = fn(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<void*>(ptr), gdim, gstride, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled(input) failed (CUresult %d)", (int)r); return CRDR_ERR_CUDA; }
  {
    std::lock_guard<std::mutex> lk(g_tm_mutex);
    if (g_in_cache.size() > 4096) g_in_cache.clear();
    g_in_cache[key] = tm;
  }
  *out = tm;
  return CRDR_OK;
}

static bool aligned16(const void* p) { return ((uintptr_t)p & 15) == 0; }
static bool planes_vec_ok(const crdr_planes& p) {
  return p.hi && aligned16(p.hi) && (!p.lo || aligned16(p.lo)) && p.cs % 8 == 0 && p.coff % 8 == 0;
}

// runtime switches of the LEAN epilogue (defaults from CRDR_CONV_LEAN / CRDR_LEAN_SWZ; tests flip them to compare the
// two epilogues bit for bit inside one process)
static int env_int(const char* name, int dflt) { const char* e = getenv(name); return e ? atoi(e) : dflt; }
static std::atomic<int> g_lean_enabled{env_int("CRDR_CONV_LEAN", 1)};
static std::atomic<int> g_lean_swizzle{env_int("CRDR_LEAN_SWZ", 1)};
void conv_set_lean(int enabled, int swizzle) {
  if (enabled >= 0) g_lean_enabled.store(enabled);
  if (swizzle >= 0) g_lean_swizzle.store(swizzle);
}

int conv2d_launch(const crdr_conv_desc* dp, cudaStream_t stream) {
  const crdr_conv_desc& d = *dp;
  const int cin = d.seg0_len + d.seg1_len;
  if (d.ntaps < 1 || d.ntaps > CRDR_MAX_TAPS || cin <= 0 || d.n <= 0 || d.hb <= 0 || d.wb <= 0 || d.cout <= 0) {
    set_error("conv2d: bad shape (ntaps=%d cin=%d n=%d hb=%d wb=%d cout=%d)", d.ntaps, cin, d.n, d.hb, d.wb, d.cout);
    return CRDR_ERR_BAD_SHAPE;
  }
  if (d.seg0_len % 8 || d.seg1_len % 8 || d.seg0_off % 8 || d.seg1_off % 8 || d.in.cs % 8 || d.in.coff % 8) {
    set_error("conv2d: input channel ranges / strides must be multiples of 8");
    return CRDR_ERR_MISALIGNED;
  }
  const bool patch = d.k_order == 1;
  int ncb = 0, cb_c0[kMaxCBlocks], cb_ksteps[kMaxCBlocks];
  if (patch) {
    for (int sgi = 0; sgi < 2; ++sgi) {
      const int off = sgi ? d.seg1_off : d.seg0_off, len = sgi ? d.seg1_len : d.seg0_len;
      for (int b = 0; b * 64 < len; ++b) {
        if (ncb >= kMaxCBlocks) { set_error("conv2d: more than %d channel blocks", kMaxCBlocks); return CRDR_ERR_BAD_SHAPE; }
        const int valid = len - 64 * b < 64 ? len - 64 * b : 64;
        cb_ksteps[ncb] = (valid + 15) / 16;   // the zero-padded tail of a partial block is never multiplied
        cb_c0[ncb++] = d.in.coff + off + 64 * b;
      }
    }
    if ((d.in_stride != 1 && d.in_stride != 2) || d.k_pad < ncb * d.ntaps * 64) {
      set_error("conv2d: k_order=1 needs in_stride 1 or 2 and k_pad >= blocks*taps*64 (in_stride=%d k_pad=%d need %d)",
                d.in_stride, d.k_pad, ncb * d.ntaps * 64);
      return CRDR_ERR_BAD_SHAPE;
    }
  }
  if (d.k_pad % kKBlk || (!patch && d.k_pad < d.ntaps * cin) || d.tile_n % 16 || d.tile_n < 16 || d.tile_n > 256 ||
      d.cout_pad % d.tile_n || d.cout > d.cout_pad) {
    set_error("conv2d: bad packed-weight geometry (k_pad=%d need>=%d, tile_n=%d, cout=%d, cout_pad=%d)", d.k_pad,
              d.ntaps * cin, d.tile_n, d.cout, d.cout_pad);
    return CRDR_ERR_BAD_SHAPE;
  }
  const bool three = d.precision == CRDR_PREC_F16X3;
  if (!d.in.hi || !d.w_hi || (three && (!d.in.lo || !d.w_lo))) {
    set_error("conv2d: missing operand plane for precision %d", d.precision);
    return CRDR_ERR_BAD_SHAPE;
  }
  if (!aligned16(d.in.hi) || !aligned16(d.in.lo) || !aligned16(d.w_hi) || !aligned16(d.w_lo)) {
    set_error("conv2d: operand pointers must be 16-byte aligned");
    return CRDR_ERR_MISALIGNED;
  }
  if (d.mode != CRDR_EPI_NONE && !d.res_f32 && !d.res.hi) {
    set_error("conv2d: epilogue mode %d needs a residual tensor", d.mode);
    return CRDR_ERR_BAD_SHAPE;
  }
  if (d.mode == CRDR_EPI_GATE && !d.trunk.hi) {
    set_error("conv2d: gate epilogue needs a trunk tensor");
    return CRDR_ERR_BAD_SHAPE;
  }
  const int64_t m_total = (int64_t)d.n * d.hb * d.wb;
  if (m_total > (1LL << 30) || (int64_t)d.n * d.hin * d.win > (1LL << 30)) {
    set_error("conv2d: pixel count exceeds 2^30");
    return CRDR_ERR_BAD_SHAPE;
  }

  ConvKParams P;
  memset(&P, 0, sizeof(P));
  P.d = d;
  P.m_total = (int32_t)m_total;
  P.cin = cin;
  P.k_real = d.ntaps * cin;
  P.nkb = patch ? ncb * d.ntaps : (P.k_real + kKBlk - 1) / kKBlk;
  P.nplanes = three ? 2 : 1;
  P.status = device_status_word();
  if (!P.status) return CRDR_ERR_CUDA;
  P.vec_planes_out = d.out.hi ? planes_vec_ok(d.out) : 0;
  P.vec_f32_out = d.out_f32 && aligned16(d.out_f32) && d.out_f32_cs % 4 == 0 && d.out_f32_coff % 4 == 0;
  P.vec_res_planes = d.res.hi ? planes_vec_ok(d.res) : 0;
  P.vec_res_f32 = d.res_f32 && aligned16(d.res_f32) && d.res_f32_cs % 4 == 0 && d.res_f32_coff % 4 == 0;
  P.vec_trunk = d.trunk.hi ? planes_vec_ok(d.trunk) : 0;
  if (!three) {
    P.d.res.lo = nullptr;    // F16X1 tensors are single-term: the epilogue reads the hi plane only
    P.d.trunk.lo = nullptr;
    P.vec_res_f32 = 0;
  }
  P.fast_epi = d.cout % 16 == 0 && (!d.out.hi || P.vec_planes_out) && (!d.out_f32 || P.vec_f32_out) &&
               (d.mode == CRDR_EPI_NONE || (d.res_f32 ? (three && P.vec_res_f32) : P.vec_res_planes)) &&
               (d.mode != CRDR_EPI_GATE || P.vec_trunk);
  P.has_bias = d.bias != nullptr;
  P.has_add = d.add_vec != nullptr;
  P.has_affine = d.scale != nullptr || d.shift != nullptr;
  {
    static int chunk_env = -1;  // tuning knob: K blocks per D0 accumulate chain
    if (chunk_env < 0) { const char* e = getenv("CRDR_CHUNK_KB"); chunk_env = e ? atoi(e) : kChunkKB; if (chunk_env < 1) chunk_env = 1; }
    P.chunk_kb = chunk_env;
    static int dbg_env = -1;  // bring-up: bit0 skip copy-out stores, bit1 skip the epilogue arithmetic, bit2 skip TMEM loads
    if (dbg_env < 0) { const char* e = getenv("CRDR_EPI_DEBUG"); dbg_env = e ? atoi(e) : 0; }
    P.dbg = dbg_env;
    static int hint_env = -1;
    if (hint_env < 0) {
      const char* e = getenv("CRDR_WAIT_HINT_NS");
      hint_env = e ? atoi(e) : 1000000;
      const uint32_t h = (uint32_t)hint_env;
      cudaMemcpyToSymbol(g_wait_hint_ns, &h, sizeof(h));
    }
    static int trace_env = -1;
    if (trace_env < 0) { const char* e = getenv("CRDR_CONV_TRACE"); trace_env = e ? atoi(e) : 0; }
    P.trace = trace_env;
    static int split_env = -1;  // -1 auto, 0 never, 1 always (tuning knob)
    if (split_env == -1) { const char* e = getenv("CRDR_CONV_SPLIT"); split_env = e ? atoi(e) : 0; }
    P.split = split_env == 2 ? -1 : split_env;
  }

  if (d.engine == CRDR_ENGINE_SIMT) {
    const int64_t total = m_total * d.cout;
    const int64_t blocks = (total + 255) / 256;
    if (blocks > 0x7fffffffLL) { set_error("conv2d(simt): problem too large"); return CRDR_ERR_BAD_SHAPE; }
    conv_simt_kernel<<<(unsigned)blocks, 256, 0, stream>>>(P);
    return check_launch("conv_simt_kernel");
  }

  if (d.cout_pad > kMaxCout) {
    set_error("conv2d: cout_pad=%d exceeds the per-CTA parameter cache (%d)", d.cout_pad, kMaxCout);
    return CRDR_ERR_BAD_SHAPE;
  }
  if (three && d.tile_n > 128) {
    set_error("conv2d: F16X3 needs tile_n <= 128 (D0 ping-pong and double-buffered D1 share 512 TMEM columns)");
    return CRDR_ERR_BAD_SHAPE;
  }
  // template instance: register-total chunks per drain warp (F16X3) or 0 (F16X1); PATCH or gather producers
  const bool use_patch = patch && d.engine != CRDR_ENGINE_TCGEN05_NOTMA;
  if (patch && !use_patch) {
    set_error("conv2d: k_order=1 weights need the TMA engine");
    return CRDR_ERR_BAD_SHAPE;
  }
  // short-K F16X3 patch launches run without the chunked D0 drain (one accumulate chain of <= kDirectMaxKB K blocks)
  static int direct_env = -1;
  if (direct_env < 0) { const char* e = getenv("CRDR_CONV_DIRECT_KB"); direct_env = e ? atoi(e) : 4; }
  static int cg2_env = -1;
  if (cg2_env < 0) { const char* e = getenv("CRDR_CONV_CG2"); cg2_env = e ? atoi(e) : 1; }
  const bool cg2 = use_patch && cg2_env != 0;
  const bool direct = three && cg2 && P.nkb <= direct_env && P.split == 0;
  const int groups = epi_groups(use_patch, three && !direct);
  const int ch_per_warp = (d.tile_n / 16 + groups - 1) / groups;  // column chunks per epilogue warp
  const int maxch = (three && !direct) ? ch_per_warp : 0;
  // LEAN epilogue (TMA-in / TMA-out, see lean_unit_math): CTA-pair patch launches without a chunk drain whose output is
  // a stride-1 set of fp16 planes in 32-channel units; residual (planes) and per-channel vectors supported.
  bool lean = g_lean_enabled.load() != 0 && cg2 && (!three || direct) && P.fast_epi && d.out.hi && !d.out_f32 &&
              d.out_stride == 1 && d.out_ph == 0 && d.out_pw == 0 && d.hb == d.hout && d.wb == d.wout &&
              (d.mode == CRDR_EPI_NONE || (d.mode == CRDR_EPI_RESIDUAL && !d.res_f32)) && d.tile_n % 32 == 0 &&
              d.cout % 32 == 0 && (three ? d.out.lo != nullptr : d.out.lo == nullptr) &&
              (d.mode == CRDR_EPI_NONE || !three || d.res.lo != nullptr);
  // residual slots per warp group: two for single-plane tensors, one for the (twice as large) two-plane units
  const int lean_res_slots = (lean && d.mode == CRDR_EPI_RESIDUAL) ? (three ? 1 : 2) : 0;
  const uint32_t lean_bytes = lean ? (uint32_t)(lean_res_slots * groups + groups) * (uint32_t)(three ? 2 : 1) * kLeanUnitPlane : 0u;
  typedef void (*KernelFn)(const ConvKParams);
  KernelFn fn = nullptr;
  // CTA-pair form (cta_group::2, M = 256 per MMA) for the patch variant; CRDR_CONV_CG2=0 falls back to single CTAs
  if (lean) {
    fn = direct ? conv_tcgen05_kernel<0, true, true, true, true> : conv_tcgen05_kernel<0, true, true, false, true>;
  } else if (direct) {
    fn = conv_tcgen05_kernel<0, true, true, true>;
  } else if (cg2) {
    switch (maxch) {
      case 0: fn = conv_tcgen05_kernel<0, true, true>; break;
      case 1: fn = conv_tcgen05_kernel<1, true, true>; break;
      case 2: fn = conv_tcgen05_kernel<2, true, true>; break;
      case 3: fn = conv_tcgen05_kernel<3, true, true>; break;
      default: fn = conv_tcgen05_kernel<4, true, true>; break;
    }
  } else
  switch (maxch * 2 + (use_patch ? 1 : 0)) {
    case 0: fn = conv_tcgen05_kernel<0, false>; break;
    case 1: fn = conv_tcgen05_kernel<0, true>; break;
    case 2: fn = conv_tcgen05_kernel<1, false>; break;
    case 3: fn = conv_tcgen05_kernel<1, true>; break;
    case 4: fn = conv_tcgen05_kernel<2, false>; break;
    case 5: fn = conv_tcgen05_kernel<2, true>; break;
    case 6: fn = conv_tcgen05_kernel<3, false>; break;
    case 7: fn = conv_tcgen05_kernel<3, true>; break;
    case 8: fn = conv_tcgen05_kernel<4, false>; break;
    default: fn = conv_tcgen05_kernel<4, true>; break;
  }
  const int variant = lean ? (direct ? 17 : 16) : direct ? 15 : cg2 ? 10 + maxch : maxch * 2 + (use_patch ? 1 : 0);
  static std::mutex attr_mutex;
  static bool attr_done[18] = {false};
  static int num_sms = 0;
  {
    std::lock_guard<std::mutex> lk(attr_mutex);
    if (!attr_done[variant]) {
      cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kDynSmemMax);
      if (e != cudaSuccess) {
        set_error("conv2d: cannot opt in to large shared memory: %s", cudaGetErrorString(e));
        return CRDR_ERR_UNSUPPORTED_ARCH;
      }
      attr_done[variant] = true;
    }
    if (num_sms == 0) {
      int dev = 0;
      cudaGetDevice(&dev);
      cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
      if (num_sms <= 0) num_sms = 148;
    }
  }
  // residual staging: one private slot per output row (pitch padded by 16 B against bank conflicts)
  uint32_t res_need = 0, res_pitch = 0;
  if (lean) {
    res_need = lean_bytes;   // residual ring + staging units (mandatory for this variant)
  } else if (d.mode != CRDR_EPI_NONE && P.fast_epi) {
    res_pitch = (uint32_t)d.tile_n * (three ? 4u : 2u) + 16u;
    res_need = kTileM * res_pitch;
  }
  // output staging: per epilogue warp a [32 rows][columns of the warp] tile per fp16 plane (pitch padded by 16 B)
  uint32_t out_need = 0, out_pitch = 0;
  {
    static int ost_env = -1;
    if (ost_env < 0) { const char* e = getenv("CRDR_CONV_OSTAGE"); ost_env = e ? atoi(e) : 1; }
    if (!lean && ost_env && P.fast_epi && d.out.hi && !d.out_f32) {
      out_pitch = (uint32_t)((d.tile_n / 16 + groups - 1) / groups) * 32u + 16u;
      out_need = (uint32_t)(4 * groups) * 32u * out_pitch * (d.out.lo ? 2u : 1u);
    }
  }
  uint32_t patch_total = 0;
  if (use_patch) {
    // Tap (dh, dw) of a stride-s layer reads pixel (s*b + dh): with dh = s*i + ph (0 <= ph < s) that is pixel (b + i)
    // of the image sub-sampled at phase ph -- a stride-1 tap on the parity class (ph, pw).  Classes must be contiguous
    // in the tap list (the packed K order follows the tap list); one halo patch per (channel block, class).
    const int sdiv = d.in_stride;
    auto fdiv = [sdiv](int v) { return v >= 0 ? v / sdiv : -((-v + sdiv - 1) / sdiv); };
    int dh_min = 127, dh_max = -127, dw_min = 127, dw_max = -127;
    for (int t = 0; t < d.ntaps; ++t) {
      const int i = fdiv(d.dh[t]), j = fdiv(d.dw[t]);
      dh_min = i < dh_min ? i : dh_min; dh_max = i > dh_max ? i : dh_max;
      dw_min = j < dw_min ? j : dw_min; dw_max = j > dw_max ? j : dw_max;
    }
    P.dh_min = dh_min; P.dw_min = dw_min;
    P.ph = kPatchTH + dh_max - dh_min;
    P.pw = kPatchTW + dw_max - dw_min;
    P.ncb = ncb;
    P.in_mul = sdiv;
    for (int i = 0; i < ncb; ++i) { P.cb_c0[i] = cb_c0[i]; P.cb_ksteps[i] = cb_ksteps[i]; }
    int ncls = 0, last_id = -1;
    for (int t = 0; t < d.ntaps; ++t) {
      const int i = fdiv(d.dh[t]), j = fdiv(d.dw[t]);
      const int phs = d.dh[t] - sdiv * i, pws = d.dw[t] - sdiv * j;
      const int id = phs * sdiv + pws;
      if (id != last_id) {
        if (id < last_id || ncls >= 4) { set_error("conv2d: strided patch taps must be grouped by parity class in ascending order"); return CRDR_ERR_BAD_SHAPE; }
        P.cls_h0[ncls] = sdiv * dh_min + phs;
        P.cls_w0[ncls] = sdiv * dw_min + pws;
        ++ncls;
        last_id = id;
      }
      P.cls_end[ncls - 1] = t + 1;
      P.tapoff[t] = (uint32_t)((i - dh_min) * P.pw + (j - dw_min)) * 128u;
    }
    P.ncls = ncls;
    P.tiles_h = (d.hb + kPatchTH - 1) / kPatchTH;
    P.tiles_w = (d.wb + kPatchTW - 1) / kPatchTW;
    if (P.ph > 64 || P.pw > 64 || d.in.cs < 64) {
      set_error("conv2d: patch engine needs tap spans <= 48 and >= 64 stored channels (ph=%d pw=%d cs=%d)", P.ph, P.pw, d.in.cs);
      return CRDR_ERR_BAD_SHAPE;
    }
    const uint32_t plane = (uint32_t)(P.ph * P.pw) * 128u;
    const uint32_t pstage = (plane * (uint32_t)P.nplanes + 1023u) & ~1023u;
    const uint32_t bstage = (uint32_t)kSlotKB * (uint32_t)P.nplanes * (uint32_t)(cg2 ? d.tile_n / 2 : d.tile_n) * 128u;
    // Patch ring depth: the patches of about two tiles in flight (small-K launches are bound by the load latency
    // of the next tile's patch, not by the MMAs), as long as four weight stages still fit; at least one.
    {
      const uint32_t budget = kDynSmemMax - 1024 - res_need - out_need;
      static int pmul_env = -1;
      if (pmul_env < 0) { const char* e = getenv("CRDR_PATCH_TILES"); pmul_env = e ? atoi(e) : 2; }
      int want = pmul_env * ncb * P.ncls;
      if (want < 2) want = 2;
      if (want > kMaxPatchStages) want = kMaxPatchStages;
      int ps = want;
      // weight slots to keep free: three, but a short-K launch (1x1 layers: one or two slots hold its whole weight tile,
      // re-streamed per tile from L2) is bound by its patch loads instead -- with three slots reserved the two-plane
      // 1x1 + skip layers were left with ONE patch stage, i.e. one 64-channel block in flight per SM
      const int nslots_all = (P.nkb + kSlotKB - 1) / kSlotKB;
      const uint32_t wres = (uint32_t)(nslots_all + 1 < 3 ? nslots_all + 1 : 3) * bstage;
      while (ps > 1 && (uint32_t)ps * pstage + wres > budget) --ps;
      if (ps == 1 && pstage + 2 * bstage > budget) ps = 1;
      P.patch_stages = ps;
    }
    patch_total = (uint32_t)P.patch_stages * pstage;
    int rc = input_tensor_map(d.in.hi, d.in.cs, d.win, d.hin, d.n, P.pw, P.ph, &P.tm_in_hi, 64, 128, sdiv);
    if (rc) return rc;
    if (three) {
      rc = input_tensor_map(d.in.lo, d.in.cs, d.win, d.hin, d.n, P.pw, P.ph, &P.tm_in_lo, 64, 128, sdiv);
      if (rc) return rc;
    }
  }
  const int b_rows = cg2 ? d.tile_n / 2 : d.tile_n;  // weight-tile rows per CTA
  const uint32_t stage_bytes = (uint32_t)P.nplanes * (use_patch ? (uint32_t)kSlotKB * (uint32_t)b_rows * 128u
                                                                : kAPlaneBytes + (uint32_t)b_rows * 128u);
  if (patch_total + 2 * stage_bytes + 1024 > kDynSmemMax) {
    set_error("conv2d: tile_n=%d does not leave two pipeline stages (patch %u B)", d.tile_n, patch_total);
    return CRDR_ERR_BAD_SHAPE;
  }
  uint32_t res_total = 0;
  P.res_stage_pitch = 0;
  P.out_stage_pitch = 0;
  if (lean) {
    if (patch_total + 2 * stage_bytes + 1024 + res_need > kDynSmemMax) {
      set_error("conv2d: the LEAN epilogue does not fit (patch %u B, stage %u B, units %u B)", patch_total, stage_bytes, res_need);
      return CRDR_ERR_BAD_SHAPE;
    }
    res_total = res_need;
    P.lean_res_slots = lean_res_slots;
    P.lean_swz = g_lean_swizzle.load();
    const int swz = P.lean_swz ? 64 : 0;
    int rc = input_tensor_map(d.out.hi, d.out.cs, d.wout, d.hout, d.n, kPatchTW, kPatchTH, &P.tm_out_hi, 32, swz);
    if (!rc && three) rc = input_tensor_map(d.out.lo, d.out.cs, d.wout, d.hout, d.n, kPatchTW, kPatchTH, &P.tm_out_lo, 32, swz);
    if (!rc && d.mode == CRDR_EPI_RESIDUAL) {
      rc = input_tensor_map(d.res.hi, d.res.cs, d.wout, d.hout, d.n, kPatchTW, kPatchTH, &P.tm_res_hi, 32, swz);
      if (!rc && three) rc = input_tensor_map(d.res.lo, d.res.cs, d.wout, d.hout, d.n, kPatchTW, kPatchTH, &P.tm_res_lo, 32, swz);
    }
    if (rc) return rc;
  } else if (res_need && patch_total + 2 * stage_bytes + 1024 + res_need <= kDynSmemMax) {
    P.res_stage_pitch = (int32_t)res_pitch;
    res_total = res_need;
  }
  if (out_need && patch_total + 3 * stage_bytes + 1024 + res_total + out_need <= kDynSmemMax) {
    P.out_stage_pitch = (int32_t)out_pitch;
    res_total += out_need;  // laid out right behind the residual slots
  }
  int stages = (int)((kDynSmemMax - 1024 - patch_total - res_total) / stage_bytes);
  if (stages > kMaxStages) stages = kMaxStages;
  P.stages = stages;
  P.tmem_cols = 512;
  if (P.split < 0) P.split = P.nkb >= 16;  // long accumulate chains: independent accumulators beat epilogue overlap
  P.use_tma = d.engine == CRDR_ENGINE_TCGEN05;
  if (P.use_tma) {
    int rc = weight_tensor_map(d.w_hi, d.k_pad, d.cout_pad, b_rows, &P.tm_hi);
    if (rc) return rc;
    if (three) {
      rc = weight_tensor_map(d.w_lo, d.k_pad, d.cout_pad, b_rows, &P.tm_lo);
      if (rc) return rc;
    }
  }
  const uint32_t smem = patch_total + (uint32_t)stages * stage_bytes + res_total + 1024;
  const int64_t m_tiles = use_patch ? (int64_t)d.n * P.tiles_h * P.tiles_w : (m_total + kTileM - 1) / kTileM;
  const int64_t num_tiles = m_tiles * (d.cout_pad / d.tile_n);
  // programmatic dependent launch (CRDR_CONV_PDL=1 enables): the prologue of launch i+1 overlaps the tail of launch i.
  // Measured neutral (87.6 vs 87.9 ms per step; a CUDA graph of the whole step gains 2 %), so it stays off by default.
  static int pdl_env = -1;
  if (pdl_env < 0) { const char* e = getenv("CRDR_CONV_PDL"); pdl_env = e ? atoi(e) : 0; }
  if (cg2) {
    // persistent CTA pairs: one cluster of two per TPC
    const int64_t pair_tiles = ((m_tiles + 1) / 2) * (d.cout_pad / d.tile_n);
    const int64_t pairs = pair_tiles < num_sms / 2 ? pair_tiles : num_sms / 2;
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3((unsigned)(2 * pairs));
    cfg.blockDim = dim3(32 * (4 * groups + 4));
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl_env ? 2 : 1;
    cudaError_t e = cudaLaunchKernelEx(&cfg, fn, P);
    if (e != cudaSuccess) {
      set_error("conv2d: cluster launch failed: %s", cudaGetErrorString(e));
      return CRDR_ERR_CUDA;
    }
    return check_launch("conv_tcgen05_kernel(cg2)");
  }
  const unsigned grid = (unsigned)(num_tiles < num_sms ? num_tiles : num_sms);  // persistent: one CTA per SM
  {
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(use_patch ? 32 * (4 * groups + 4) : kThreads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr;
    attr.id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr.val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = &attr;
    cfg.numAttrs = pdl_env ? 1 : 0;
    cudaError_t e = cudaLaunchKernelEx(&cfg, fn, P);
    if (e != cudaSuccess) {
      set_error("conv2d: launch failed: %s", cudaGetErrorString(e));
      return CRDR_ERR_CUDA;
    }
  }
  return check_launch("conv_tcgen05_kernel");
}

}  // namespace crdr
