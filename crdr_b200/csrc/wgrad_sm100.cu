// Weight-gradient contraction of a (possibly strided / transposed) convolution for sm_100a:
//
//     G[t][a][b] = sum over pixels p = (n, h, w) of the SMALL grid   S[p][a] * B[n, h*s + dh_t, w*s + dw_t][b]
//
// S and B are NHWC fp16 planes (the tensors the forward pass already holds: no transposed copies are made).
//   nn.Conv2d (stride s):           S = dY (gradient of the output), B = X (saved input)      -> dW[co][ci][t]
//   nn.ConvTranspose2d (stride s):  S = X (saved input, low resolution), B = dY                -> dW[ci][co][t]
// (torch.autograd's conv backward for the weight, called from loss.backward() in rate_distortion_trainer.py:84.)
//
// GEMM view: M = channels of S (128 per tile), N = channels of B (<= 256 per tile), K = pixels.  Both operands are
// "MN-major" for the tensor core: a TMA box [64 pixels][64 channels] of an NHWC plane lands in shared memory as 64 rows
// of 128 bytes (SWIZZLE_128B), which IS the canonical MN-major UMMA layout ((8,n),(8,k)):((1,LBO),(8,SBO)) in 16-byte
// units: 64 contiguous channels per K row, 8-row groups SBO = 1024 B apart, 64-channel atoms LBO = one box apart.  The
// tap shift and the stride are TMA coordinates / element strides of the B box; out-of-image pixels are zero filled, which
// is exactly the convolution's zero padding.  One CTA = one (tap, M tile, N tile) and one K range (split-K over pixel
// blocks); partial sums go to an fp32 workspace and wgrad_reduce_kernel adds the splits in a fixed order (deterministic,
// no atomics) while scattering into the parameter's own layout.
//
// HALO form (stride 1, several taps, enough pixels): one CTA = (group of <= 8 taps, M tile, 64-channel block of B); a K block
// is an 8 x 8 pixel tile whose halo patch (one TMA box, zero filled outside the image) serves every tap of the group: the B
// operand of a tap is the same patch seen from a shifted start row with the 8-row K groups one PATCH row apart (the 128B
// swizzle is a function of the absolute shared-memory address, so any 128-byte start row works), and every tap accumulates
// into its own 64 TMEM columns.  L2 -> shared-memory traffic per MAC drops 5.6x against one tap per CTA.
//
// Warp roles: 0 TMA producer, 1 TMEM allocation + MMA issue, 2-5 epilogue (TMEM -> workspace).
#include <cstdlib>
#include <cstring>
#include <mutex>

#include "common.cuh"
#include "sm100_device.cuh"

namespace crdr {

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_fn();

constexpr int kWgThreads = 192;
constexpr int kWgMaxStages = 8;
constexpr uint32_t kWgBox = 64u * 128u;          // one TMA box: 64 pixels x 64 channels fp16
constexpr uint32_t kWgSmemMax = 200u * 1024u;

struct alignas(64) WgParams {
  CUtensorMap tm_s;     // 4-D (C, W, H, N) over S, box (64, bw, bh, bn)
  CUtensorMap tm_b;     // 4-D over B, box (64, bw*s, bh*s, bn), element strides (1, s, s, 1)
  float* ws;            // [splits][ntaps][a_pad][b_pad]
  int32_t ntaps, a_tiles, b_tiles, nb;
  int32_t a_pad, b_pad;
  int32_t kw_blocks, kh_blocks;
  int32_t bw, bh, bn, stride;
  int32_t kb_total, splits, stages;
  uint32_t tmem_cols;
  int8_t dh[CRDR_MAX_TAPS], dw[CRDR_MAX_TAPS];
  uint32_t* status;
  // splits == 1: the epilogue writes the parameter gradient itself (no workspace round trip, no reduction launch)
  float* out;
  int64_t sa, sb, st;
  int32_t ca, cb, accumulate;
  float scale;
  // HALO variant (stride 1, several taps): one CTA = (tap group, M tile, 64-channel block of B); a K block is an 8 x 8 pixel
  // tile whose halo patch (box (64, pw, ph, 1), zero filled outside the image) serves every tap of the group from shared
  // memory, each tap accumulating into its own 64 TMEM columns
  CUtensorMap tm_patch;
  int32_t ph, pw, dh_min, dw_min, group_taps, ngroups, b_blocks;
  uint32_t patch_bytes;
  uint32_t tapoff[CRDR_MAX_TAPS];
};

// MN-major SWIZZLE_128B operand: `lbo` bytes between 64-element atoms along M / N, `sbo` bytes between 8-row K groups
__device__ __forceinline__ uint64_t umma_desc_mn_sw128(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = (uint64_t)((saddr & 0x3FFFFu) >> 4);
  d |= (uint64_t)(lbo >> 4) << 16;
  d |= (uint64_t)(sbo >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

template <bool HALO>
__global__ void __launch_bounds__(kWgThreads, 1) wgrad_kernel(const __grid_constant__ WgParams P) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full[kWgMaxStages], empty[kWgMaxStages];
  __shared__ __align__(8) uint64_t acc_full;
  __shared__ uint32_t tmem_slot;

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  const int tile = (int)blockIdx.x;
  const int nbt = HALO ? P.b_blocks : P.b_tiles;
  const int bt = tile % nbt;
  const int at = (tile / nbt) % P.a_tiles;
  const int tap = tile / (nbt * P.a_tiles);                       // HALO: the tap GROUP
  const int tap0 = HALO ? tap * P.group_taps : tap;
  const int gtaps = HALO ? min(P.group_taps, P.ntaps - tap0) : 1;
  const int split = (int)blockIdx.y;
  const int kb0 = (int)(((int64_t)split * P.kb_total) / P.splits);
  const int kb1 = (int)(((int64_t)(split + 1) * P.kb_total) / P.splits);
  const int nkb = kb1 - kb0;
  const int nb = HALO ? 1 : P.nb;
  const uint32_t stage_bytes = HALO ? 2u * kWgBox + P.patch_bytes : (uint32_t)(2 + nb) * kWgBox;
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kWgMaxStages; ++s) { mbar_init(smem_u32(&full[s]), 1u); mbar_init(smem_u32(&empty[s]), 1u); }
    mbar_init(smem_u32(&acc_full), 1u);
    fence_barrier_init();
  }
  __syncthreads();
  if (warp == 1) tmem_alloc(smem_u32(&tmem_slot), P.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, tmem_slot, 0);

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      prefetch_tmap(&P.tm_s);
      prefetch_tmap(HALO ? &P.tm_patch : &P.tm_b);
      const int dh = HALO ? P.dh_min : P.dh[tap], dw = HALO ? P.dw_min : P.dw[tap];
      int s = 0;
      uint32_t par = 1u;
      for (int i = 0; i < nkb; ++i) {
        const int kb = kb0 + i;
        const int ww = kb % P.kw_blocks;
        const int t = kb / P.kw_blocks;
        const int hh = t % P.kh_blocks;
        const int nn = t / P.kh_blocks;
        mbar_wait(smem_u32(&empty[s]), par, P.status);
        const uint32_t bar = smem_u32(&full[s]);
        // bytes the TMA actually delivers (the patch slot is rounded up to the swizzle atom, the box is not)
        mbar_arrive_expect_tx(bar, HALO ? 2u * kWgBox + (uint32_t)(P.ph * P.pw) * 128u : stage_bytes);
        const uint32_t dst = smem_base + (uint32_t)s * stage_bytes;
        const int w0 = ww * P.bw, h0 = hh * P.bh, n0 = nn * P.bn;
        tma_load_4d(dst, &P.tm_s, at * 128, w0, h0, n0, bar);
        tma_load_4d(dst + kWgBox, &P.tm_s, at * 128 + 64, w0, h0, n0, bar);
        if (HALO) {
          tma_load_4d(dst + 2u * kWgBox, &P.tm_patch, bt * 64, w0 + dw, h0 + dh, n0, bar);
        } else {
          for (int j = 0; j < nb; ++j)
            tma_load_4d(dst + (uint32_t)(2 + j) * kWgBox, &P.tm_b, (bt * nb + j) * 64, w0 * P.stride + dw, h0 * P.stride + dh, n0, bar);
        }
        if (++s == P.stages) { s = 0; par ^= 1u; }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issue (one elected lane)
    const bool elected = elect_one();
    const uint32_t idesc = (1u << 4) | (1u << 15) | (1u << 16) | ((uint32_t)((64 * nb) >> 3) << 17) | ((128u >> 4) << 24);
    const uint64_t desc0 = umma_desc_mn_sw128(0u, kWgBox, 1024u);
    // HALO: the B operand of a tap is the patch seen from a shifted start row; its 8-row K groups (one tile row of 8 pixels)
    // are one PATCH row apart.  As for the K-major operands of the convolution kernel, the 128B swizzle is a function of the
    // absolute shared-memory address, so any 128-byte start row of the 1024-byte aligned patch works.
    const uint64_t desc_p = umma_desc_mn_sw128(0u, kWgBox, (uint32_t)P.pw * 128u);
    const uint32_t krow = (uint32_t)P.pw * 256u;                  // two tile rows of the patch = one K16 step
    int s = 0;
    uint32_t par = 0u;
    for (int i = 0; i < nkb; ++i) {
      mbar_wait(smem_u32(&full[s]), par, P.status);
      tc_fence_after();
      const uint32_t st = smem_base + (uint32_t)s * stage_bytes;
      const uint64_t a = desc0 + (uint64_t)(st >> 4);
      const uint64_t b = desc0 + (uint64_t)((st + 2u * kWgBox) >> 4);
      if (elected) {
        if (HALO) {
          for (int j = 0; j < gtaps; ++j) {
            const uint32_t pb = st + 2u * kWgBox + P.tapoff[tap0 + j];
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma_f16(tmem_base + (uint32_t)j * 64u, a + (uint64_t)(k * 128), desc_p + (uint64_t)((pb + (uint32_t)k * krow) >> 4), idesc,
                       (i > 0 || k > 0) ? 1u : 0u);
          }
        } else {
#pragma unroll
          for (int k = 0; k < 4; ++k)   // 64 pixels = 4 x K16; 16 K rows = 2048 bytes
            umma_f16(tmem_base, a + (uint64_t)(k * 128), b + (uint64_t)(k * 128), idesc, (i > 0 || k > 0) ? 1u : 0u);
        }
        umma_commit(smem_u32(&empty[s]));
      }
      __syncwarp();
      if (++s == P.stages) { s = 0; par ^= 1u; }
    }
    if (elected) umma_commit(smem_u32(&acc_full));
    __syncwarp();
  } else {
    // ------------------------------------------------------------------ epilogue: TMEM -> workspace
    const int q = warp & 3;
    const int row = q * 32 + lane;
    mbar_wait(smem_u32(&acc_full), 0u, P.status);
    tc_fence_after();
    const uint32_t lane_base = tmem_base + ((uint32_t)(q * 32) << 16);
    for (int tj = 0; tj < gtaps; ++tj) {
    const int tap = tap0 + tj;                                      // (shadows the group index: from here on the real tap)
    const uint32_t lane_addr = lane_base + (HALO ? (uint32_t)tj * 64u : 0u);
    if (P.splits == 1) {
      // direct: out[t*st + a*sa + b*sb] (+)= scale * G   (one writer per element: deterministic)
      const int a = at * 128 + row;
      const int b0 = bt * 64 * nb;
      float* dst = P.out + (int64_t)a * P.sa + (int64_t)tap * P.st;
      for (int c0 = 0; c0 < 64 * nb; c0 += 32) {
        if (b0 + c0 >= P.cb) break;           // warp-uniform: the rest of this N tile is padding
        uint32_t r[32];
        tmem_ld32_issue(lane_addr + (uint32_t)c0, r);
        tmem_wait_ld();
        if (a < P.ca) {
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const int b = b0 + c0 + j;
            if (b < P.cb) {
              float* o = dst + (int64_t)b * P.sb;
              const float v = __uint_as_float(r[j]) * P.scale;
              *o = P.accumulate ? *o + v : v;
            }
          }
        }
      }
    } else {
      float* dst = P.ws + (((int64_t)split * P.ntaps + tap) * P.a_pad + (at * 128 + row)) * (int64_t)P.b_pad + (int64_t)bt * 64 * nb;
      for (int c0 = 0; c0 < 64 * nb; c0 += 32) {
        uint32_t r[32];
        tmem_ld32_issue(lane_addr + (uint32_t)c0, r);
        tmem_wait_ld();
        float4* o = reinterpret_cast<float4*>(dst + c0);
#pragma unroll
        for (int j = 0; j < 8; ++j)
          o[j] = make_float4(__uint_as_float(r[4 * j]), __uint_as_float(r[4 * j + 1]), __uint_as_float(r[4 * j + 2]),
                             __uint_as_float(r[4 * j + 3]));
      }
    }
    }   // taps of the group
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, P.tmem_cols);
  }
}

// out[t*st + a*sa + b*sb] (+)= scale * sum over splits (ascending) of ws[split][t][a][b]
__global__ void wgrad_reduce_kernel(const float* __restrict__ ws, int splits, int ntaps, int a_pad, int b_pad, int ca, int cb,
                                    float* __restrict__ out, int64_t sa, int64_t sb, int64_t st, float scale, int accumulate) {
  const int64_t total = (int64_t)ca * cb * ntaps;
  const int64_t plane = (int64_t)a_pad * b_pad;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int t = (int)(i % ntaps);
    const int64_t r = i / ntaps;
    const int b = (int)(r % cb);
    const int a = (int)(r / cb);
    const float* p = ws + (int64_t)t * plane + (int64_t)a * b_pad + b;
    float v = 0.f;
    for (int s = 0; s < splits; ++s) v += p[(int64_t)s * ntaps * plane];
    const int64_t o = (int64_t)a * sa + (int64_t)b * sb + (int64_t)t * st;
    out[o] = accumulate ? fmaf(v, scale, out[o]) : v * scale;
  }
}

// ----------------------------------------------------------------------------------------------
// Host side
// ----------------------------------------------------------------------------------------------
static int nhwc_box_map(const void* ptr, int clen, int cs, int w, int h, int n, int bw, int bh, int bn, int estride,
                        CUtensorMap* out) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) { set_error("cuTensorMapEncodeTiled is unavailable (driver too old?)"); return CRDR_ERR_CUDA; }
  cuuint64_t gdim[4] = {(cuuint64_t)clen, (cuuint64_t)w, (cuuint64_t)h, (cuuint64_t)n};
  cuuint64_t gstride[3] = {(cuuint64_t)cs * 2, (cuuint64_t)w * cs * 2, (cuuint64_t)h * w * cs * 2};
  cuuint32_t box[4] = {64u, (cuuint32_t)(bw * estride), (cuuint32_t)(bh * estride), (cuuint32_t)bn};
  cuuint32_t estr[4] = {1u, (cuuint32_t)estride, (cuuint32_t)estride, 1u};
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<void*>(ptr), gdim, gstride, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled(wgrad) failed (CUresult %d)", (int)r); return CRDR_ERR_CUDA; }
  return CRDR_OK;
}

static int pow2_le(int x) { int p = 1; while (p * 2 <= x) p *= 2; return p; }

struct WgPlan {
  int bw, bh, bn, kw_blocks, kh_blocks, kn_blocks, kb_total;
  int a_tiles, b_tiles, nb, a_pad, b_pad, splits;
  int halo, ph, pw, dh_min, dw_min, group_taps, ngroups, b_blocks;
};

static int g_wgrad_halo = -1;   // CRDR_WGRAD_HALO (default 1): 0 forces the one-tap-per-CTA form (bisecting / comparison)

static int wgrad_plan(const crdr_wgrad_desc& d, size_t ws_bytes, WgPlan* pl) {
  if (d.n <= 0 || d.hs <= 0 || d.ws <= 0 || d.hb <= 0 || d.wb <= 0 || d.ca <= 0 || d.cb <= 0 || d.ntaps < 1 ||
      d.ntaps > CRDR_MAX_TAPS || d.stride < 1 || d.stride > 4) {
    set_error("wgrad: bad shape (n=%d hs=%d ws=%d ca=%d cb=%d ntaps=%d stride=%d)", d.n, d.hs, d.ws, d.ca, d.cb, d.ntaps, d.stride);
    return CRDR_ERR_BAD_SHAPE;
  }
  if (g_wgrad_halo < 0) { const char* e = getenv("CRDR_WGRAD_HALO"); g_wgrad_halo = e ? atoi(e) : 1; }
  int dh0 = 127, dh1 = -127, dw0 = 127, dw1 = -127;
  for (int t = 0; t < d.ntaps; ++t) {
    dh0 = d.dh[t] < dh0 ? d.dh[t] : dh0; dh1 = d.dh[t] > dh1 ? d.dh[t] : dh1;
    dw0 = d.dw[t] < dw0 ? d.dw[t] : dw0; dw1 = d.dw[t] > dw1 ? d.dw[t] : dw1;
  }
  // Measured (tools/wgrad_bench.py, profiles/wgrad_bench_r02.txt): with >= 128 pixel blocks the halo form is 1.9x faster
  // (3x3 128->128 on 131 072 pixels: 119 -> 63 us, 615 TFLOP/s; 914 TFLOP/s at eight times the pixels); on the 16 x 16 latent
  // grids of a training crop (32 pixel blocks) a launch is fixed cost + partial-sum traffic and the wide one-tap tiles win.
  const int64_t halo_kb = (int64_t)((d.ws + 7) / 8) * ((d.hs + 7) / 8) * d.n;
  pl->halo = g_wgrad_halo && d.stride == 1 && d.ntaps > 1 && d.hs >= 8 && d.ws >= 8 && dh1 - dh0 <= 8 && dw1 - dw0 <= 8 &&
             (halo_kb >= 128 || g_wgrad_halo > 1);
  if (pl->halo) {
    pl->ph = 8 + dh1 - dh0; pl->pw = 8 + dw1 - dw0; pl->dh_min = dh0; pl->dw_min = dw0;
    pl->ngroups = (d.ntaps + 7) / 8;
    pl->group_taps = (d.ntaps + pl->ngroups - 1) / pl->ngroups;
    pl->ngroups = (d.ntaps + pl->group_taps - 1) / pl->group_taps;
    pl->bw = 8; pl->bh = 8; pl->bn = 1;
    pl->kw_blocks = (d.ws + 7) / 8; pl->kh_blocks = (d.hs + 7) / 8; pl->kn_blocks = d.n;
    const int64_t kbh = (int64_t)pl->kw_blocks * pl->kh_blocks * pl->kn_blocks;
    if (kbh > (1 << 30)) { set_error("wgrad: too many pixel blocks"); return CRDR_ERR_BAD_SHAPE; }
    pl->kb_total = (int)kbh;
    pl->a_tiles = (d.ca + 127) / 128;
    pl->a_pad = pl->a_tiles * 128;
    pl->b_blocks = (d.cb + 63) / 64;
    pl->b_tiles = pl->b_blocks; pl->nb = 1;
    pl->b_pad = pl->b_blocks * 64;
    const int64_t tiles = (int64_t)pl->ngroups * pl->a_tiles * pl->b_blocks;
    static int waves_x2 = -1;   // CRDR_WGRAD_HALO_WAVES2: CTAs per launch in half waves of 148 (tuning knob)
    if (waves_x2 < 0) { const char* e = getenv("CRDR_WGRAD_HALO_WAVES2"); waves_x2 = e ? atoi(e) : 2; }
    int64_t splits = (waves_x2 * 74 + tiles - 1) / tiles;
    const int64_t by_k = pl->kb_total / 4 > 0 ? pl->kb_total / 4 : 1;
    if (splits > by_k) splits = by_k;
    if (splits > 128) splits = 128;
    const int64_t per = (int64_t)d.ntaps * pl->a_pad * pl->b_pad * 4;
    if (ws_bytes) {
      if ((int64_t)ws_bytes < per) { set_error("wgrad: workspace too small (%zu < %lld bytes)", ws_bytes, (long long)per); return CRDR_ERR_BAD_SHAPE; }
      if (splits * per > (int64_t)ws_bytes) splits = (int64_t)ws_bytes / per;
    }
    pl->splits = (int)splits;
    return CRDR_OK;
  }
  pl->bw = pow2_le(d.ws < 8 ? d.ws : 8);
  pl->bh = pow2_le(d.hs < 64 / pl->bw ? d.hs : 64 / pl->bw);
  pl->bn = 64 / (pl->bw * pl->bh);
  pl->kw_blocks = (d.ws + pl->bw - 1) / pl->bw;
  pl->kh_blocks = (d.hs + pl->bh - 1) / pl->bh;
  pl->kn_blocks = (d.n + pl->bn - 1) / pl->bn;
  const int64_t kb = (int64_t)pl->kw_blocks * pl->kh_blocks * pl->kn_blocks;
  if (kb > (1 << 30)) { set_error("wgrad: too many pixel blocks"); return CRDR_ERR_BAD_SHAPE; }
  pl->kb_total = (int)kb;
  pl->a_tiles = (d.ca + 127) / 128;
  pl->a_pad = pl->a_tiles * 128;
  const int nb_total = (d.cb + 63) / 64;
  pl->nb = nb_total < 4 ? nb_total : 4;
  // balance the N tiles (e.g. 5 boxes -> 3 + 2 instead of 4 + 1 keeps one tile shape: use ceil)
  pl->b_tiles = (nb_total + pl->nb - 1) / pl->nb;
  pl->nb = (nb_total + pl->b_tiles - 1) / pl->b_tiles;
  pl->b_pad = pl->b_tiles * pl->nb * 64;
  // The K loop is bound by the L2 -> shared-memory traffic of its operand boxes (no reuse across taps yet), which does not
  // depend on the split count, so what matters is keeping every SM loading: two waves of CTAs.  (One wave of longer CTAs was
  // tried: 26 -> 70 us per launch on the training step's shapes.)
  const int64_t tiles = (int64_t)d.ntaps * pl->a_tiles * pl->b_tiles;
  int64_t splits = (2 * 148 + tiles - 1) / tiles;
  const int64_t by_k = pl->kb_total / 8 > 0 ? pl->kb_total / 8 : 1;
  if (splits > by_k) splits = by_k;
  if (splits > 128) splits = 128;
  const int64_t per = (int64_t)d.ntaps * pl->a_pad * pl->b_pad * 4;
  if (ws_bytes) {
    if ((int64_t)ws_bytes < per) { set_error("wgrad: workspace too small (%zu < %lld bytes)", ws_bytes, (long long)per); return CRDR_ERR_BAD_SHAPE; }
    if (splits * per > (int64_t)ws_bytes) splits = (int64_t)ws_bytes / per;
  }
  pl->splits = (int)splits;
  return CRDR_OK;
}

size_t wgrad_workspace_bytes(const crdr_wgrad_desc* d) {
  WgPlan pl;
  if (wgrad_plan(*d, 0, &pl)) return 0;
  return (size_t)pl.splits * d->ntaps * pl.a_pad * pl.b_pad * 4;
}

int wgrad_launch(const crdr_wgrad_desc* dp, cudaStream_t stream) {
  const crdr_wgrad_desc& d = *dp;
  WgPlan pl;
  if (!d.workspace || !d.workspace_bytes) { set_error("wgrad: no workspace (see crdr_conv_wgrad_workspace)"); return CRDR_ERR_BAD_SHAPE; }
  int rc = wgrad_plan(d, d.workspace_bytes, &pl);
  if (rc) return rc;
  if (!d.s.hi || !d.b.hi || !d.out || d.s.cs % 8 || d.s.coff % 8 || d.b.cs % 8 || d.b.coff % 8 ||
      ((uintptr_t)d.s.hi & 15) || ((uintptr_t)d.b.hi & 15) || ((uintptr_t)d.workspace & 15)) {
    set_error("wgrad: missing operand, or channel strides / offsets not multiples of 8, or misaligned pointers");
    return CRDR_ERR_MISALIGNED;
  }
  WgParams P;
  memset(&P, 0, sizeof(P));
  P.status = device_status_word();
  if (!P.status) return CRDR_ERR_CUDA;
  const __half* sp = reinterpret_cast<const __half*>(d.s.hi) + d.s.coff;
  const __half* bp = reinterpret_cast<const __half*>(d.b.hi) + d.b.coff;
  rc = nhwc_box_map(sp, d.ca, d.s.cs, d.ws, d.hs, d.n, pl.bw, pl.bh, pl.bn, 1, &P.tm_s);
  if (!rc && !pl.halo) rc = nhwc_box_map(bp, d.cb, d.b.cs, d.wb, d.hb, d.n, pl.bw, pl.bh, pl.bn, d.stride, &P.tm_b);
  if (!rc && pl.halo) rc = nhwc_box_map(bp, d.cb, d.b.cs, d.wb, d.hb, d.n, pl.pw, pl.ph, 1, 1, &P.tm_patch);
  if (rc) return rc;
  P.ws = reinterpret_cast<float*>(d.workspace);
  P.ntaps = d.ntaps; P.a_tiles = pl.a_tiles; P.b_tiles = pl.b_tiles; P.nb = pl.nb;
  P.a_pad = pl.a_pad; P.b_pad = pl.b_pad;
  P.kw_blocks = pl.kw_blocks; P.kh_blocks = pl.kh_blocks;
  P.bw = pl.bw; P.bh = pl.bh; P.bn = pl.bn; P.stride = d.stride;
  P.kb_total = pl.kb_total; P.splits = pl.splits;
  P.out = d.out; P.sa = d.sa; P.sb = d.sb; P.st = d.st; P.ca = d.ca; P.cb = d.cb; P.accumulate = d.accumulate; P.scale = d.scale;
  for (int t = 0; t < d.ntaps; ++t) { P.dh[t] = d.dh[t]; P.dw[t] = d.dw[t]; }
  if (pl.halo) {
    P.ph = pl.ph; P.pw = pl.pw; P.dh_min = pl.dh_min; P.dw_min = pl.dw_min;
    P.group_taps = pl.group_taps; P.ngroups = pl.ngroups; P.b_blocks = pl.b_blocks;
    P.patch_bytes = ((uint32_t)(pl.ph * pl.pw) * 128u + 1023u) & ~1023u;
    for (int t = 0; t < d.ntaps; ++t)
      P.tapoff[t] = (uint32_t)((d.dh[t] - pl.dh_min) * pl.pw + (d.dw[t] - pl.dw_min)) * 128u;
  }
  const uint32_t stage_bytes = pl.halo ? 2u * kWgBox + P.patch_bytes : (uint32_t)(2 + pl.nb) * kWgBox;
  int stages = (int)((kWgSmemMax - 2048u) / stage_bytes);
  if (stages > kWgMaxStages) stages = kWgMaxStages;
  P.stages = stages;
  const int ncols = pl.halo ? 64 * pl.group_taps : 64 * pl.nb;
  P.tmem_cols = ncols <= 64 ? 64u : ncols <= 128 ? 128u : ncols <= 256 ? 256u : 512u;
  const uint32_t smem = 1024u + (uint32_t)stages * stage_bytes;

  static std::mutex mu;
  static bool attr_done = false;
  {
    std::lock_guard<std::mutex> lk(mu);
    if (!attr_done) {
      cudaError_t e = cudaFuncSetAttribute(wgrad_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kWgSmemMax);
      if (e == cudaSuccess) e = cudaFuncSetAttribute(wgrad_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kWgSmemMax);
      if (e != cudaSuccess) { set_error("wgrad: cannot opt in to large shared memory: %s", cudaGetErrorString(e)); return CRDR_ERR_UNSUPPORTED_ARCH; }
      attr_done = true;
    }
  }
  if (pl.halo) {
    dim3 grid((unsigned)(pl.ngroups * pl.a_tiles * pl.b_blocks), (unsigned)pl.splits);
    wgrad_kernel<true><<<grid, kWgThreads, smem, stream>>>(P);
  } else {
    dim3 grid((unsigned)(d.ntaps * pl.a_tiles * pl.b_tiles), (unsigned)pl.splits);
    wgrad_kernel<false><<<grid, kWgThreads, smem, stream>>>(P);
  }
  rc = check_launch("wgrad_kernel");
  if (rc || pl.splits == 1) return rc;
  const int64_t total = (int64_t)d.ca * d.cb * d.ntaps;
  int blocks = (int)((total + 255) / 256);
  if (blocks > 148 * 16) blocks = 148 * 16;
  wgrad_reduce_kernel<<<blocks, 256, 0, stream>>>(P.ws, pl.splits, d.ntaps, pl.a_pad, pl.b_pad, d.ca, d.cb, d.out, d.sa, d.sb,
                                                  d.st, d.scale, d.accumulate);
  return check_launch("wgrad_reduce_kernel");
}

}  // namespace crdr
