// Fused tail of a residual bottleneck for sm_100a:  out = affine( x + W3 * relu(W2 (*) t1 + b2 [+ a2]) + b3 [+ a3] )
//
//   t1 --3x3 conv (mid -> mid)--> ReLU (+ beta bias) --1x1 conv (mid -> C)--> (+ beta bias) + skip x  (+ InterpChAtt gain)
//
// i.e. the second and third convolution of BaseBlock (elic_layers.py:23-36), NLAMResBlock (cheng_nlam.py:32-47) and
// BetaCondBaseBlock (elic_interpca_beta_cond_autoencoder.py:42-66) in ONE launch (F16X1 tensors: the synthesis
// transform).  The 3x3 result never goes to HBM: its epilogue writes the fp16 tile straight into shared memory in the
// canonical K-major / 128B-swizzled UMMA layout, where it is the A operand of the 1x1 convolution whose weights stay
// resident in shared memory for the whole (persistent) kernel.  Per bottleneck this removes one launch and the write +
// read of the mid tensor (HBM traffic 10 -> 8 units of mid-channel planes), and hides the 1x1's epilogue (the
// HBM-bound part: residual in, result out) under the 3x3's MMAs.
//
// Structure = the CTA-pair patch kernel of conv_sm100.cu (halo-patch TMA, cta_group::2 MMAs with M = 256, weights
// streamed by TMA, LEAN TMA-in / TMA-out epilogue) plus a second, dependent GEMM per tile:
//   MMA thread:      b(0) b(1) c(0) b(2) c(1) ...      b(j) = 3x3 of tile j into acc_b[j & 1], c(j) = 1x1 of tile j into acc_c
//   epilogue warps:  eb(0) eb(1) ec(0) eb(2) ec(1) ... eb: acc_b -> bias/ReLU/add -> fp16 -> shared memory (t2)
//                                                      ec: acc_c -> bias/add/+x/gain -> fp16 -> TMA store
// so the tensor pipe works on b(j+1) while the epilogue warps finish tile j.  Same fp16 values, same K order and the same
// fp32 epilogue arithmetic as the two separate launches: the result is bit-identical to the unfused path.
//
// TMEM (512 columns): acc_b[2] at 0 / 128 (mid <= 128), acc_c at 256 (C <= 256).
#include <cstdlib>
#include <cstring>
#include <mutex>

#include "common.cuh"
#include "sm100_device.cuh"

namespace crdr {

int weight_tensor_map(const void* ptr, int k_pad, int rows, int box_rows, CUtensorMap* out);
int input_tensor_map(const void* ptr, int cs, int w, int h, int n, int pw, int ph, CUtensorMap* out, int box_c, int swz,
                     int estride);

constexpr int kBnThreads = 512;      // warps 0-11 epilogue, 12 weight TMA + TMEM alloc, 13 MMA, 14 halo patches, 15 residual units
constexpr int kBnGroups = 3;         // epilogue warp groups (4 TMEM lane quarters each)
constexpr int kBnEpiWarps = 4 * kBnGroups;
constexpr int kBnMaxStages = 8;      // weight ring slots (two K blocks each)
constexpr int kBnMaxPatch = 6;       // halo-patch ring
constexpr int kBnSlotKB = 2;
constexpr int kBnMaxRes = 6;         // residual unit slots (kBnGroups private rings of 1-2)
constexpr int kBnMaxCB = 2;          // 64-channel blocks of the mid tensor (mid <= 128)
constexpr uint32_t kBnAccB = 128u, kBnAccC = 256u;   // TMEM columns: acc_b stride, acc_c base
constexpr uint32_t kBnDynSmemMax = 227u * 1024u - 10u * 1024u;   // static: two parameter caches (7.5 KB) + barriers

struct alignas(64) BnParams {
  CUtensorMap tm_in;    // t1, 4-D (C, W, H, N), box (64, pw, ph, 1), SWIZZLE_128B
  CUtensorMap tm_w2;    // [mid_pad][k2_pad] fp16, box (64, mid / 2)
  CUtensorMap tm_w3;    // [cout_pad][k3_pad] fp16, box (64, cout / 2)
  CUtensorMap tm_res;   // x,   4-D, box (32, 8, 16, 1), SWIZZLE_64B
  CUtensorMap tm_out;   // out, 4-D, box (32, 8, 16, 1), SWIZZLE_64B
  const float* bias2;
  const float* add2;
  const float* bias3;
  const float* add3;
  const float* scale;
  const float* shift;
  int32_t n, h, w, tiles_h, tiles_w;
  int32_t mid, cout, ncb;
  int32_t cb_c0[kBnMaxCB], cb_ksteps[kBnMaxCB];
  int32_t ph, pw;
  uint32_t tapoff[9];
  int32_t stages, patch_stages, res_slots;
  int32_t res_coff, out_coff;
  uint32_t* status;
};

__device__ __forceinline__ void bn_tile_origin(const BnParams& P, int mt, int& n, int& h0, int& w0) {
  const int tw = mt % P.tiles_w;
  const int t = mt / P.tiles_w;
  const int th = t % P.tiles_h;
  n = t / P.tiles_h;
  h0 = th * kPatchTH;
  w0 = tw * kPatchTW;
}

__global__ void __launch_bounds__(kBnThreads, 1) bottleneck_bc_kernel(const __grid_constant__ BnParams P) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t wfull[kBnMaxStages], wempty[kBnMaxStages];
  __shared__ __align__(8) uint64_t patch_full[kBnMaxPatch], patch_empty[kBnMaxPatch];
  __shared__ __align__(8) uint64_t res_full[kBnMaxRes], res_empty[kBnMaxRes];
  __shared__ __align__(8) uint64_t accb_full[2], accb_empty[2];
  __shared__ __align__(8) uint64_t accc_full, accc_empty, t2_full, t2_empty, w3_full;
  __shared__ uint32_t tmem_slot;
  __shared__ __align__(16) float s_par3[4 * kMaxCout];   // 1x1: bias | add | scale | shift, by output channel
  __shared__ __align__(16) float s_par2[2 * kMaxCout];   // 3x3: bias | add, by mid channel

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  const uint32_t cta_rank = __shfl_sync(0xffffffffu, cluster_ctarank(), 0);
  const int tile0 = (int)(blockIdx.x >> 1), tstep = (int)(gridDim.x >> 1);
  const int m_tiles = P.n * P.tiles_h * P.tiles_w;
  const int num_tiles = (m_tiles + 1) / 2;                       // tiles of the CTA pair (M = 256)
  const int T = tile0 < num_tiles ? (num_tiles - tile0 + tstep - 1) / tstep : 0;   // tiles of this pair
  const int ncb = P.ncb;
  const int nkb = ncb * 9;                                       // K blocks of the 3x3: (channel block, tap)
  // shared memory map
  const uint32_t patch_stage_bytes = ((uint32_t)(P.ph * P.pw) * 128u + 1023u) & ~1023u;
  const uint32_t w2_kb_bytes = (uint32_t)(P.mid / 2) * 128u;     // one K block of this CTA's half of the 3x3 weight tile
  const uint32_t w2_slot_bytes = (uint32_t)kBnSlotKB * w2_kb_bytes;
  const uint32_t w3_kb_bytes = (uint32_t)(P.cout / 2) * 128u;    // one K block of this CTA's half of the 1x1 weights
  constexpr uint32_t t2_kb_bytes = 128u * 128u;
  const uint32_t smem_patch = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t smem_w2 = smem_patch + (uint32_t)P.patch_stages * patch_stage_bytes;
  const uint32_t smem_w3 = (smem_w2 + (uint32_t)P.stages * w2_slot_bytes + 1023u) & ~1023u;
  const uint32_t smem_t2 = (smem_w3 + (uint32_t)ncb * w3_kb_bytes + 1023u) & ~1023u;
  const uint32_t smem_res = smem_t2 + (uint32_t)ncb * t2_kb_bytes;
  const int D = P.res_slots;
  const uint32_t smem_stage = smem_res + (uint32_t)(kBnGroups * D) * kLeanUnitPlane;

  for (int i = threadIdx.x; i < 4 * kMaxCout; i += kBnThreads) {
    const int which = i / kMaxCout, co = i % kMaxCout;
    const float* src = which == 0 ? P.bias3 : which == 1 ? P.add3 : which == 2 ? P.scale : P.shift;
    s_par3[i] = (src && co < P.cout) ? src[co] : (which == 2 ? 1.f : 0.f);
  }
  for (int i = threadIdx.x; i < 2 * kMaxCout; i += kBnThreads) {
    const int which = i / kMaxCout, co = i % kMaxCout;
    const float* src = which == 0 ? P.bias2 : P.add2;
    s_par2[i] = (src && co < P.mid) ? src[co] : 0.f;
  }
  if (threadIdx.x == 0) {
    for (int b = 0; b < kBnMaxStages; ++b) { mbar_init(smem_u32(&wfull[b]), 1u); mbar_init(smem_u32(&wempty[b]), 1u); }
    for (int b = 0; b < kBnMaxPatch; ++b) { mbar_init(smem_u32(&patch_full[b]), 1u); mbar_init(smem_u32(&patch_empty[b]), 1u); }
    for (int b = 0; b < kBnMaxRes; ++b) { mbar_init(smem_u32(&res_full[b]), 1u); mbar_init(smem_u32(&res_empty[b]), 4u); }
    for (int b = 0; b < 2; ++b) {
      mbar_init(smem_u32(&accb_full[b]), 1u);
      mbar_init(smem_u32(&accb_empty[b]), 2u * kBnEpiWarps);   // one arrive per epilogue warp of both CTAs (on the leader)
    }
    mbar_init(smem_u32(&accc_full), 1u);
    mbar_init(smem_u32(&accc_empty), 2u * kBnEpiWarps);
    mbar_init(smem_u32(&t2_full), 2u * kBnEpiWarps);
    mbar_init(smem_u32(&t2_empty), 1u);
    mbar_init(smem_u32(&w3_full), 1u);
    fence_barrier_init();
  }
  cluster_sync_all();
  if (warp == kBnEpiWarps) {
    if (lane == 0) {
      prefetch_tmap(&P.tm_in); prefetch_tmap(&P.tm_w2); prefetch_tmap(&P.tm_w3); prefetch_tmap(&P.tm_res); prefetch_tmap(&P.tm_out);
    }
    __syncwarp();
    tmem_alloc_cg2(smem_u32(&tmem_slot), 512u);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, tmem_slot, 0);

  if (warp < kBnEpiWarps) {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 152;");
    // ------------------------------------------------------------------ epilogue warps
    const int q = warp & 3;                  // TMEM lane quarter
    const int cgrp = warp >> 2;              // warp group: units cgrp, cgrp + 3, ...
    const int row = q * 32 + lane;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16);
    const int units_b = P.mid >> 5, units_c = P.cout >> 5;
    const uint32_t ef_b = (P.bias2 ? kEfBias : 0u) | kEfRelu | (P.add2 ? kEfAdd : 0u) | kEfOutHi;
    const uint32_t ef_c = (P.bias3 ? kEfBias : 0u) | (P.add3 ? kEfAdd : 0u) | ((P.scale || P.shift) ? kEfAffine : 0u) | kEfOutHi |
                          ((uint32_t)CRDR_EPI_RESIDUAL << kEfModeShift);
    // LEAN unit addressing (SWIZZLE_64B units of [128 rows][64 B])
    const uint32_t sx = (uint32_t)((row >> 1) & 3);
    uint32_t choff[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) choff[j] = (uint32_t)row * 64u + (((uint32_t)j ^ sx) << 4);
    // t2 addressing (K-major SWIZZLE_128B rows of 128 B): 16-byte piece p of this thread's row
    const uint32_t t2_row = smem_t2 + (uint32_t)row * 128u;
    const uint32_t t2_x = (uint32_t)(row & 7);
    const uint32_t res_ring = smem_res + (uint32_t)(cgrp * D) * kLeanUnitPlane;
    const uint32_t stage_unit = smem_stage + (uint32_t)cgrp * kLeanUnitPlane;
    int rslot = 0;
    uint32_t rpar = 0u;
    const bool issuer = q == 0 && lane == 0;
    auto arrive_leader = [&](uint64_t* bar) {
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(mapa_shared(smem_u32(bar), 0u));
    };
    const uint4 z4 = make_uint4(0u, 0u, 0u, 0u);
    for (int j = 0; j <= T; ++j) {
      if (j < T) {
        // ---- eb(j): 3x3 accumulator -> bias / ReLU / beta bias -> fp16 -> t2 (the 1x1's A operand)
        const int tb = j & 1;
        mbar_wait(smem_u32(&accb_full[tb]), (uint32_t)(j >> 1) & 1u, P.status);
        tc_fence_after();
        bool first = true;
        // Unit ownership balances eb + ec work per warp group: the 1x1's units go round robin (group g: g, g + 3, ...), so
        // the 3x3's units beyond the first three go to the LAST groups (mid = 128, C = 256: 4 units per group and tile;
        // round robin for both left group 0 with 5 of the 12 and made it the critical path of every tile)
        const int nb_mine = (cgrp < units_b ? 1 : 0) + ((units_b > kBnGroups && cgrp == kBnGroups - 1) ? 1 : 0);
        if (nb_mine == 0) arrive_leader(&accb_empty[tb]);
#pragma unroll 1
        for (int k = 0; k < nb_mine; ++k) {
          const int u = k == 0 ? cgrp : units_b - 1;   // units_b <= 4: the only extra unit is the last one
          uint32_t r0[32];
          tmem_ld32_issue(lane_addr + (uint32_t)tb * kBnAccB + (uint32_t)u * 32u, r0);
          tmem_wait_ld();
          if (k + 1 == nb_mine) arrive_leader(&accb_empty[tb]);   // this warp's last read of acc_b[tb]
          float v[32];
#pragma unroll
          for (int e = 0; e < 32; ++e) v[e] = __uint_as_float(r0[e]);
          uint32_t hw[16], lw[16];
          const uint4 rz[4] = {z4, z4, z4, z4};
          lean_unit_math<false>(v, rz, rz, s_par2, u * 32, ef_b, P.status, hw, lw);
          if (first) {
            // t2 still feeds the 1x1 MMAs of the previous tile until their commit (a fresh barrier passes the first wait)
            mbar_wait(smem_u32(&t2_empty), ((uint32_t)j & 1u) ^ 1u, P.status);
            first = false;
          }
          const uint32_t base = t2_row + (uint32_t)(u >> 1) * t2_kb_bytes;
          const uint32_t p0 = (uint32_t)(u & 1) * 4u;
#pragma unroll
          for (int jj = 0; jj < 4; ++jj)
            st_shared16(base + (((p0 + (uint32_t)jj) ^ t2_x) << 4), hw[4 * jj], hw[4 * jj + 1], hw[4 * jj + 2], hw[4 * jj + 3]);
        }
        fence_proxy_async();            // generic-proxy writes -> visible to the tensor core (async proxy)
        arrive_leader(&t2_full);
      }
      if (j >= 1) {
        // ---- ec(j - 1): 1x1 accumulator -> bias / beta bias / + x / gain -> fp16 -> TMA store
        const int jc = j - 1;
        const int tile = tile0 + jc * tstep;
        mbar_wait(smem_u32(&accc_full), (uint32_t)jc & 1u, P.status);
        tc_fence_after();
        if (cgrp >= units_c) arrive_leader(&accc_empty);
#pragma unroll 1
        for (int u = cgrp; u < units_c; u += kBnGroups) {
          uint32_t r0[32];
          tmem_ld32_issue(lane_addr + kBnAccC + (uint32_t)u * 32u, r0);
          uint4 rh[4];
          {
            const int slot = cgrp * D + rslot;
            mbar_wait(smem_u32(&res_full[slot]), rpar, P.status);
            const uint32_t rbase = res_ring + (uint32_t)rslot * kLeanUnitPlane;
#pragma unroll
            for (int jj = 0; jj < 4; ++jj) rh[jj] = lds128(rbase + choff[jj]);
            fence_proxy_async();        // the slot goes back to the TMA while these loads may still be queued (see conv_sm100.cu)
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(&res_empty[slot]));
            if (++rslot == D) { rslot = 0; rpar ^= 1u; }
          }
          tmem_wait_ld();
          if (u + kBnGroups >= units_c) arrive_leader(&accc_empty);       // this warp's last read of acc_c
          float v[32];
#pragma unroll
          for (int e = 0; e < 32; ++e) v[e] = __uint_as_float(r0[e]);
          uint32_t hw[16], lw[16];
          lean_unit_math<false>(v, rh, rh, s_par3, u * 32, ef_c, P.status, hw, lw);
          if (issuer) bulk_wait_read0();
          bar_sync_named(1u + (uint32_t)cgrp, 128u);
#pragma unroll
          for (int jj = 0; jj < 4; ++jj)
            st_shared16(stage_unit + choff[jj], hw[4 * jj], hw[4 * jj + 1], hw[4 * jj + 2], hw[4 * jj + 3]);
          fence_proxy_async();
          bar_sync_named(1u + (uint32_t)cgrp, 128u);
          if (issuer) {
            int n, h0, w0;
            bn_tile_origin(P, 2 * tile + (int)cta_rank, n, h0, w0);
            if (n < P.n) tma_store_4d(&P.tm_out, stage_unit, P.out_coff + u * 32, w0, h0, n);
            bulk_commit();
          }
        }
      }
    }
    if (issuer) bulk_wait0();
  } else {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
    if (warp == kBnEpiWarps) {
      // ------------------------------------------------------------------ weights by TMA: the resident 1x1 tile, then the 3x3 ring
      if (lane == 0) {
        {
          const uint32_t bar = smem_u32(&w3_full);
          if (cta_rank == 0) mbar_arrive_expect_tx(bar, 2u * (uint32_t)ncb * w3_kb_bytes);
          const uint32_t lbar = mapa_shared(bar, 0u);
          for (int kb = 0; kb < ncb; ++kb)
            tma_load_2d_cg2(smem_w3 + (uint32_t)kb * w3_kb_bytes, &P.tm_w3, kb * kKBlk, (int)cta_rank * (P.cout / 2), lbar);
        }
        int s = 0;
        uint32_t empty_par = 1u;
        uint32_t dst0 = smem_w2;
        for (int j = 0; j < T; ++j) {
          int kk = 0;
          for (int kb = 0; kb < nkb; ++kb) {
            const uint32_t bar = smem_u32(&wfull[s]);
            if (kk == 0) {
              mbar_wait(smem_u32(&wempty[s]), empty_par, P.status);
              const uint32_t kcount = (uint32_t)min(kBnSlotKB, nkb - kb);
              if (cta_rank == 0) mbar_arrive_expect_tx(bar, 2u * kcount * w2_kb_bytes);
            }
            tma_load_2d_cg2(dst0 + (uint32_t)kk * w2_kb_bytes, &P.tm_w2, kb * kKBlk, (int)cta_rank * (P.mid / 2), mapa_shared(bar, 0u));
            if (++kk == kBnSlotKB || kb == nkb - 1) {
              kk = 0;
              dst0 += w2_slot_bytes;
              if (++s == P.stages) { s = 0; empty_par ^= 1u; dst0 = smem_w2; }
            }
          }
        }
      }
    } else if (warp == kBnEpiWarps + 2) {
      // ------------------------------------------------------------------ halo patches of t1 by TMA
      if (lane == 0) {
        int pb = 0;
        uint32_t empty_par = 1u;
        const uint32_t patch_bytes = (uint32_t)(P.ph * P.pw) * 128u;
        for (int j = 0; j < T; ++j) {
          int n, h0, w0;
          bn_tile_origin(P, 2 * (tile0 + j * tstep) + (int)cta_rank, n, h0, w0);   // a tile past the end has n == P.n -> zero fill
          for (int cb = 0; cb < ncb; ++cb) {
            mbar_wait(smem_u32(&patch_empty[pb]), empty_par, P.status);
            const uint32_t bar = smem_u32(&patch_full[pb]);
            if (cta_rank == 0) mbar_arrive_expect_tx(bar, 2u * patch_bytes);
            tma_load_4d_cg2(smem_patch + (uint32_t)pb * patch_stage_bytes, &P.tm_in, P.cb_c0[cb], w0 - 1, h0 - 1, n, mapa_shared(bar, 0u));
            if (++pb == P.patch_stages) { pb = 0; empty_par ^= 1u; }
          }
        }
      }
    } else if (warp == kBnEpiWarps + 3) {
      // ------------------------------------------------------------------ residual units (x) by TMA, per warp-group rings
      if (lane == 0) {
        const int units = P.cout >> 5;
        int gslot[kBnGroups];
        uint32_t gpar[kBnGroups];
#pragma unroll
        for (int g = 0; g < kBnGroups; ++g) { gslot[g] = 0; gpar[g] = 1u; }
        for (int j = 0; j < T; ++j) {
          int n, h0, w0;
          bn_tile_origin(P, 2 * (tile0 + j * tstep) + (int)cta_rank, n, h0, w0);
          int g = 0;
          for (int u = 0; u < units; ++u) {
            int slot = 0;
            uint32_t par = 0u;
#pragma unroll
            for (int k = 0; k < kBnGroups; ++k)
              if (k == g) { slot = g * D + gslot[k]; par = gpar[k]; if (++gslot[k] == D) { gslot[k] = 0; gpar[k] ^= 1u; } }
            mbar_wait(smem_u32(&res_empty[slot]), par, P.status);
            const uint32_t bar = smem_u32(&res_full[slot]);
            mbar_arrive_expect_tx(bar, kLeanUnitPlane);
            tma_load_4d(smem_res + (uint32_t)slot * kLeanUnitPlane, &P.tm_res, P.res_coff + u * 32, w0, h0, n, bar);
            if (++g == kBnGroups) g = 0;
          }
        }
      }
    } else if (warp == kBnEpiWarps + 1) {
      // ------------------------------------------------------------------ MMA issue (leader CTA, one elected lane)
      if (cta_rank == 0) {
        const bool elected = elect_one();
        const uint32_t idesc_b = umma_idesc_f16((uint32_t)P.mid, 256u);
        const uint32_t idesc_c = umma_idesc_f16((uint32_t)P.cout, 256u);
        const uint64_t desc_k = umma_desc_sw128(0u);                                   // + (addr >> 4): plain K-major tiles
        const uint64_t desc_patch = umma_desc_sw128_rows(0u, (uint32_t)P.pw * 128u);   // tap-shifted views of a halo patch
        int s = 0, pb = 0;
        uint32_t ring_par = 0u, patch_par = 0u;
        uint32_t stage = smem_w2, patch_addr = smem_patch;
        for (int j = 0; j <= T; ++j) {
          if (j < T) {
            // ---- b(j): the 3x3 convolution of tile j
            const int tb = j & 1;
            if (j >= 2) {
              mbar_wait(smem_u32(&accb_empty[tb]), (uint32_t)((j - 2) >> 1) & 1u, P.status);
              tc_fence_after();
            }
            const uint32_t d0 = tmem_base + (uint32_t)tb * kBnAccB;
            int kk = 0, kb = 0;
            for (int cb = 0; cb < ncb; ++cb) {
              mbar_wait(smem_u32(&patch_full[pb]), patch_par, P.status);
              const int nks = P.cb_ksteps[cb];
              for (int tap = 0; tap < 9; ++tap, ++kb) {
                if (kk == 0) mbar_wait(smem_u32(&wfull[s]), ring_par, P.status);
                tc_fence_after();
                const uint64_t a = desc_patch + (uint64_t)((patch_addr + P.tapoff[tap]) >> 4);
                const uint64_t b = desc_k + (uint64_t)((stage + (uint32_t)kk * w2_kb_bytes) >> 4);
                if (elected) {
#pragma unroll
                  for (int k = 0; k < kKBlk / 16; ++k) {
                    if (k >= nks) break;
                    umma_f16_cg2(d0, a + (uint64_t)(k * 2), b + (uint64_t)(k * 2), idesc_b, (kb > 0 || k > 0) ? 1u : 0u);
                  }
                }
                if (++kk == kBnSlotKB || kb == nkb - 1) {
                  if (elected) umma_commit_cg2(smem_u32(&wempty[s]));
                  kk = 0;
                  stage += w2_slot_bytes;
                  if (++s == P.stages) { s = 0; ring_par ^= 1u; stage = smem_w2; }
                }
              }
              if (elected) umma_commit_cg2(smem_u32(&patch_empty[pb]));
              patch_addr += patch_stage_bytes;
              if (++pb == P.patch_stages) { pb = 0; patch_par ^= 1u; patch_addr = smem_patch; }
            }
            if (elected) umma_commit_cg2(smem_u32(&accb_full[tb]));
          }
          if (j >= 1) {
            // ---- c(j - 1): the 1x1 convolution over the fp16 tile the epilogue warps left in shared memory
            const int jc = j - 1;
            if (jc == 0) mbar_wait(smem_u32(&w3_full), 0u, P.status);
            mbar_wait(smem_u32(&t2_full), (uint32_t)jc & 1u, P.status);
            if (jc >= 1) mbar_wait(smem_u32(&accc_empty), (uint32_t)(jc - 1) & 1u, P.status);
            tc_fence_after();
            if (elected) {
              for (int kb = 0; kb < ncb; ++kb) {
                const uint64_t a = desc_k + (uint64_t)((smem_t2 + (uint32_t)kb * t2_kb_bytes) >> 4);
                const uint64_t b = desc_k + (uint64_t)((smem_w3 + (uint32_t)kb * w3_kb_bytes) >> 4);
                const int nks = P.cb_ksteps[kb];
#pragma unroll
                for (int k = 0; k < kKBlk / 16; ++k) {
                  if (k >= nks) break;
                  umma_f16_cg2(tmem_base + kBnAccC, a + (uint64_t)(k * 2), b + (uint64_t)(k * 2), idesc_c, (kb > 0 || k > 0) ? 1u : 0u);
                }
              }
              umma_commit_cg2(smem_u32(&accc_full));
              umma_commit_cg2(smem_u32(&t2_empty));
            }
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();   // the peer's TMEM, shared memory and barriers are in use until both CTAs are done
  if (warp == kBnEpiWarps) {
    tc_fence_after();
    tmem_dealloc_cg2(tmem_base, 512u);
  }
}

// ----------------------------------------------------------------------------------------------
// Host side
// ----------------------------------------------------------------------------------------------
static bool bn_aligned16(const void* p) { return ((uintptr_t)p & 15) == 0; }

int bottleneck_bc_launch(const crdr_bottleneck_desc* dp, cudaStream_t stream) {
  const crdr_bottleneck_desc& d = *dp;
  if (d.n <= 0 || d.h <= 0 || d.w <= 0 || d.mid <= 0 || d.cout <= 0) {
    set_error("bottleneck: bad shape (n=%d h=%d w=%d mid=%d cout=%d)", d.n, d.h, d.w, d.mid, d.cout);
    return CRDR_ERR_BAD_SHAPE;
  }
  // TMEM: acc_b[2] (stride 128) + acc_c (256); units of 32 channels; two 64-channel blocks of mid channels at most
  if (d.mid % 32 || d.mid > 128 || d.cout % 32 || d.cout > 256 || d.mid < 32) {
    set_error("bottleneck: needs mid %% 32 == 0, mid <= 128, cout %% 32 == 0, cout <= 256 (mid=%d cout=%d)", d.mid, d.cout);
    return CRDR_ERR_BAD_SHAPE;
  }
  if (d.precision != CRDR_PREC_F16X1) {
    set_error("bottleneck: only F16X1 (single-plane fp16) tensors are fused; F16X3 blocks run as separate launches");
    return CRDR_ERR_BAD_SHAPE;
  }
  const int ncb = (d.mid + 63) / 64;
  if (!d.in.hi || !d.res.hi || !d.out.hi || !d.w2 || !d.w3 || d.in.cs < 64 || d.in.cs % 8 || d.in.coff % 8 || d.res.cs % 8 ||
      d.res.coff % 8 || d.out.cs % 8 || d.out.coff % 8 || d.k2_pad < ncb * 9 * 64 || d.k3_pad < ncb * 64 || d.k2_pad % 64 ||
      d.k3_pad % 64 || d.mid_pad < d.mid || d.cout_pad < d.cout) {
    set_error("bottleneck: missing operand or bad packed-weight geometry (k2_pad=%d k3_pad=%d in.cs=%d)", d.k2_pad, d.k3_pad, d.in.cs);
    return CRDR_ERR_BAD_SHAPE;
  }
  if (!bn_aligned16(d.in.hi) || !bn_aligned16(d.res.hi) || !bn_aligned16(d.out.hi) || !bn_aligned16(d.w2) || !bn_aligned16(d.w3)) {
    set_error("bottleneck: operand pointers must be 16-byte aligned");
    return CRDR_ERR_MISALIGNED;
  }
  if ((int64_t)d.n * d.h * d.w > (1LL << 30)) { set_error("bottleneck: pixel count exceeds 2^30"); return CRDR_ERR_BAD_SHAPE; }

  BnParams P;
  memset(&P, 0, sizeof(P));
  P.status = device_status_word();
  if (!P.status) return CRDR_ERR_CUDA;
  P.bias2 = d.bias2; P.add2 = d.add2; P.bias3 = d.bias3; P.add3 = d.add3; P.scale = d.scale; P.shift = d.shift;
  P.n = d.n; P.h = d.h; P.w = d.w;
  P.tiles_h = (d.h + kPatchTH - 1) / kPatchTH;
  P.tiles_w = (d.w + kPatchTW - 1) / kPatchTW;
  P.mid = d.mid; P.cout = d.cout; P.ncb = ncb;
  for (int b = 0; b < ncb; ++b) {
    const int valid = d.mid - 64 * b < 64 ? d.mid - 64 * b : 64;
    P.cb_c0[b] = d.in.coff + 64 * b;
    P.cb_ksteps[b] = (valid + 15) / 16;
  }
  P.ph = kPatchTH + 2; P.pw = kPatchTW + 2;
  for (int t = 0; t < 9; ++t) P.tapoff[t] = (uint32_t)((t / 3) * P.pw + (t % 3)) * 128u;   // taps (dh, dw) = (t/3 - 1, t%3 - 1)
  P.res_coff = d.res.coff; P.out_coff = d.out.coff;

  // shared memory: [patch ring][3x3 weight ring][resident 1x1 weights][t2][residual units][staging units]
  const uint32_t pstage = ((uint32_t)(P.ph * P.pw) * 128u + 1023u) & ~1023u;
  const uint32_t wslot = (uint32_t)kBnSlotKB * (uint32_t)(d.mid / 2) * 128u;
  const uint32_t w3 = ((uint32_t)ncb * (uint32_t)(d.cout / 2) * 128u + 1023u) & ~1023u;
  const uint32_t t2 = (uint32_t)ncb * 128u * 128u;
  int res_slots = 2;
  uint32_t fixed = 0;
  int stages = 0, pst = 0;
  for (; res_slots >= 1; --res_slots) {
    fixed = 2048 + w3 + 1024 + t2 + (uint32_t)(kBnGroups * res_slots + kBnGroups) * kLeanUnitPlane;
    // patches of one tile (ncb) at least, two tiles if four weight slots still fit
    pst = 2 * ncb;
    if (pst > kBnMaxPatch) pst = kBnMaxPatch;
    while (pst > ncb && fixed + (uint32_t)pst * pstage + 4 * wslot > kBnDynSmemMax) --pst;
    if (fixed + (uint32_t)pst * pstage + 3 * wslot <= kBnDynSmemMax) break;
  }
  if (res_slots < 1) { set_error("bottleneck: shared-memory budget exceeded (mid=%d cout=%d)", d.mid, d.cout); return CRDR_ERR_BAD_SHAPE; }
  stages = (int)((kBnDynSmemMax - fixed - (uint32_t)pst * pstage) / wslot);
  if (stages > kBnMaxStages) stages = kBnMaxStages;
  P.stages = stages; P.patch_stages = pst; P.res_slots = res_slots;
  const uint32_t smem = fixed + (uint32_t)pst * pstage + (uint32_t)stages * wslot;

  int rc = input_tensor_map(d.in.hi, d.in.cs, d.w, d.h, d.n, P.pw, P.ph, &P.tm_in, 64, 128, 1);
  if (!rc) rc = weight_tensor_map(d.w2, d.k2_pad, d.mid_pad, d.mid / 2, &P.tm_w2);
  if (!rc) rc = weight_tensor_map(d.w3, d.k3_pad, d.cout_pad, d.cout / 2, &P.tm_w3);
  if (!rc) rc = input_tensor_map(d.res.hi, d.res.cs, d.w, d.h, d.n, kPatchTW, kPatchTH, &P.tm_res, 32, 64, 1);
  if (!rc) rc = input_tensor_map(d.out.hi, d.out.cs, d.w, d.h, d.n, kPatchTW, kPatchTH, &P.tm_out, 32, 64, 1);
  if (rc) return rc;

  static std::mutex mu;
  static bool attr_done = false;
  static int num_sms = 0;
  {
    std::lock_guard<std::mutex> lk(mu);
    if (!attr_done) {
      cudaError_t e = cudaFuncSetAttribute(bottleneck_bc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kBnDynSmemMax);
      if (e != cudaSuccess) { set_error("bottleneck: cannot opt in to large shared memory: %s", cudaGetErrorString(e)); return CRDR_ERR_UNSUPPORTED_ARCH; }
      attr_done = true;
    }
    if (num_sms == 0) {
      int dev = 0;
      cudaGetDevice(&dev);
      cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
      if (num_sms <= 0) num_sms = 148;
    }
  }
  const int64_t m_tiles = (int64_t)d.n * P.tiles_h * P.tiles_w;
  const int64_t pair_tiles = (m_tiles + 1) / 2;
  const int64_t pairs = pair_tiles < num_sms / 2 ? pair_tiles : num_sms / 2;
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3((unsigned)(2 * pairs));
  cfg.blockDim = dim3(kBnThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, bottleneck_bc_kernel, P);
  if (e != cudaSuccess) { set_error("bottleneck: cluster launch failed: %s", cudaGetErrorString(e)); return CRDR_ERR_CUDA; }
  return check_launch("bottleneck_bc_kernel");
}

}  // namespace crdr
