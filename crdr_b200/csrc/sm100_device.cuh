// Device-side building blocks shared by the sm_100a tensor-core kernels (conv_sm100.cu, bottleneck_sm100.cu):
// PTX wrappers (mbarrier, TMA, tcgen05 / TMEM, clusters), UMMA descriptors and the LEAN epilogue arithmetic.
#pragma once
#include "common.cuh"

namespace crdr {

constexpr int kTileM = 128;
constexpr int kKBlk = 64;                  // fp16 elements per K block = one 128-byte swizzle row
constexpr int kMaxCout = 320;              // per-CTA cache of the per-channel epilogue vectors
constexpr uint32_t kLeanUnitPlane = 128u * 64u;  // LEAN: one unit = [128 rows][32 channels] fp16 per plane
constexpr int kPatchTH = 16, kPatchTW = 8; // PATCH mode: the 128 GEMM rows are a 16 x 8 pixel tile of the base grid

// ----------------------------------------------------------------------------------------------
// PTX wrappers
// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
static __constant__ uint32_t g_wait_hint_ns = 1000000u;  // CRDR_WAIT_HINT_NS (tuning knob)
// try_wait with a suspend-time hint: the thread sleeps in hardware until the phase completes (or the hint
// expires) instead of burning issue slots that the epilogue warps of the same SM sub-partition need.
__device__ __forceinline__ uint32_t mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity), "r"(g_wait_hint_ns)
      : "memory");
  return ok;
}
// Bounded wait: a pipeline bug must end in a trap with the status flag set, never in a hung GPU.
__device__ __forceinline__ uint32_t mbar_test_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity, uint32_t* status) {
  if (mbar_test_wait(bar, parity)) return;  // the common case in a filled pipeline: no suspend machinery
#pragma unroll 1
  for (uint32_t i = 0; i < 4000u; ++i)  // <= 4000 x 1 ms
    if (mbar_try_wait(bar, parity)) return;
  atomicOr(status, kFlagTimeout);
  __threadfence_system();
  __trap();
}
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, uint32_t src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }

__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* tm, int c0, int c1, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
      ::"r"(dst), "l"(tm), "r"(c0), "r"(c1), "r"(bar)
      : "memory");
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* tm, int c0, int c1, int c2, int c3,
                                            uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
      ::"r"(dst), "l"(tm), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(bar)
      : "memory");
}
// ---- CTA-pair (cta_group::2) variants: the mbarrier operand lives in the LEADER CTA of the pair ----
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of the same shared-memory location in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t mapa_shared(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  // default (.release.cta) semantics as in CUTLASS' ClusterBarrier::arrive(cta_id): the only thing handed over is
  // TMEM, ordered by tcgen05.fence::before_thread_sync; ".release.cluster" compiles to MEMBAR.ALL.GPU + ERRBAR and
  // was 27 % of all stall samples in the drain warps.
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void tma_load_2d_cg2(uint32_t dst, const CUtensorMap* tm, int c0, int c1, uint32_t leader_bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
      ::"r"(dst), "l"(tm), "r"(c0), "r"(c1), "r"(leader_bar)
      : "memory");
}
__device__ __forceinline__ void tma_load_4d_cg2(uint32_t dst, const CUtensorMap* tm, int c0, int c1, int c2, int c3,
                                                uint32_t leader_bar) {
  asm volatile(
      "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
      ::"r"(dst), "l"(tm), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(leader_bar)
      : "memory");
}
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* tm) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(tm) : "memory");
}

__device__ __forceinline__ void tmem_alloc(uint32_t slot_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(slot_smem), "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// CTA pair: one warp of EACH CTA of the pair executes these
__device__ __forceinline__ void tmem_alloc_cg2(uint32_t slot_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(slot_smem), "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_cg2(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// one lane of a fully converged warp (CUTLASS elect_one_sync)
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void umma_f16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                         uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Arrives on the mbarrier once every previously issued tcgen05.mma of this thread has completed.
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// CTA pair: M = 256 (rows 0-127 from the leader's shared memory / TMEM, 128-255 from the peer's), each CTA supplies
// half of the N rows of B; issued by the leader only.  The commit arrives on the barrier at this shared-memory
// offset in BOTH CTAs.
__device__ __forceinline__ void umma_f16_cg2(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit_cg2(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(bar), "h"((uint16_t)3)
               : "memory");
}
template <bool CG2>
__device__ __forceinline__ void umma_issue(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  if (CG2) umma_f16_cg2(d_tmem, adesc, bdesc, idesc, acc);
  else umma_f16(d_tmem, adesc, bdesc, idesc, acc);
}
template <bool CG2>
__device__ __forceinline__ void umma_done(uint32_t bar) {
  if (CG2) umma_commit_cg2(bar);
  else umma_commit(bar);
}
// 16 consecutive fp32 columns of this thread's TMEM lane; the values are valid after tmem_wait_ld().
__device__ __forceinline__ void tmem_ld16_issue(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
// 32 consecutive fp32 columns of this thread's TMEM lane
__device__ __forceinline__ void tmem_ld32_issue(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,"
      "%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]),
        "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]),
        "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// ---- LEAN epilogue plumbing: bulk tensor stores from shared memory, named barriers, 128-bit shared accesses ----
__device__ __forceinline__ void tma_store_4d(const CUtensorMap* tm, uint32_t src, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%1, %2, %3, %4}], [%5];"
               ::"l"(tm), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(src)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void bar_sync_named(uint32_t id, uint32_t nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ uint4 lds128(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
        "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// K-major, SWIZZLE_128B shared-memory matrix descriptor (sm_100 "version 1"): rows are 128 bytes,
// groups of 8 rows are 1024 bytes apart (SBO); LBO is unused for swizzled K-major operands.
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t saddr) {
  uint64_t d = (uint64_t)((saddr & 0x3FFFFu) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// Same layout for an operand that starts at an arbitrary 128-byte row of a swizzled buffer and whose 8-row
// groups are `sbo_bytes` apart (PATCH mode: one group per tile row of the halo patch).  Measured on B200: the
// swizzle XOR is taken from the absolute shared-memory address bits, so as long as the buffer itself (the TMA
// destination) is 1024-byte aligned the operand may start at any row and use any 128-byte-multiple group stride
// with base_offset = 0 (setting base_offset to the start row's phase gives wrong results).
__device__ __forceinline__ uint64_t umma_desc_sw128_rows(uint32_t saddr, uint32_t sbo_bytes) {
  uint64_t d = (uint64_t)((saddr & 0x3FFFFu) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(sbo_bytes >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// kind::f16 instruction descriptor: fp16 A/B (K-major), fp32 accumulate, M=128, N=n.
__device__ __forceinline__ uint32_t umma_idesc_f16(uint32_t n, uint32_t m = kTileM) {
  return (1u << 4) | ((n >> 3) << 17) | ((m >> 4) << 24);
}


__device__ __forceinline__ float2 unpack_h2(uint32_t u) {
  __half2 h = *reinterpret_cast<__half2*>(&u);
  return __half22float2(h);
}
__device__ __forceinline__ uint32_t pack_h2(__half a, __half b) {
  __half2 h = __halves2half2(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}

__device__ __forceinline__ void st_shared16(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}

// Per-launch epilogue switches as a register bit mask (see epi_flags in conv_sm100.cu)
constexpr uint32_t kEfBias = 1, kEfRelu = 2, kEfAdd = 4, kEfAffine = 8, kEfOutF32 = 16, kEfOutHi = 32, kEfOutLo = 64,
                   kEfResF32 = 128, kEfResLo = 256, kEfTrunkLo = 512, kEfModeShift = 12;

// ----------------------------------------------------------------------------------------------
// LEAN epilogue (patch + CTA-pair kernels, F16X1 or DIRECT F16X3).  A unit is 32 output channels of the tile's 128
// rows ([128][64 B] per fp16 plane in shared memory, the box of a 4-D tensor map over the NHWC planes).  Residual
// units arrive by TMA in a small ring, results leave through one staging unit per warp group and a TMA store (which
// also clips partial tiles): no per-thread global addressing, no cp.async bookkeeping, no shuffles, a third of the
// instructions of the staged epilogue above.  The arithmetic is the same sequence of fp32 operations as epi_finish,
// so both epilogues produce identical bits.
// ----------------------------------------------------------------------------------------------
template <bool THREE>
__device__ __forceinline__ void lean_unit_math(float (&v)[32], const uint4 (&rh)[4], const uint4 (&rl)[4],
                                               const float* s_par, int co0, uint32_t ef, uint32_t* status,
                                               uint32_t (&hw)[16], uint32_t (&lw)[16]) {
  const int mode = (int)(ef >> kEfModeShift);
  const float4* par = reinterpret_cast<const float4*>(s_par + co0);
  constexpr int kVecStride = kMaxCout / 4;
  if (ef & kEfBias) {
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const float4 t = par[q];
      v[4 * q] += t.x; v[4 * q + 1] += t.y; v[4 * q + 2] += t.z; v[4 * q + 3] += t.w;
    }
  }
  if (ef & kEfRelu) {
#pragma unroll
    for (int e = 0; e < 32; ++e) v[e] = fmaxf(v[e], 0.0f);
  }
  if (ef & kEfAdd) {
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const float4 t = par[kVecStride + q];
      v[4 * q] += t.x; v[4 * q + 1] += t.y; v[4 * q + 2] += t.z; v[4 * q + 3] += t.w;
    }
  }
  if (mode == CRDR_EPI_RESIDUAL) {
    const bool has_lo = THREE && (ef & kEfResLo);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const uint32_t h4[4] = {rh[j].x, rh[j].y, rh[j].z, rh[j].w};
      const uint32_t l4[4] = {rl[j].x, rl[j].y, rl[j].z, rl[j].w};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        float2 a = unpack_h2(h4[i]);
        if (has_lo) {
          const float2 b = unpack_h2(l4[i]);
          a.x = fmaf(b.x, kLoInv, a.x);
          a.y = fmaf(b.y, kLoInv, a.y);
        }
        v[8 * j + 2 * i] += a.x;
        v[8 * j + 2 * i + 1] += a.y;
      }
    }
  }
  if (ef & kEfAffine) {
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const float4 sc = par[2 * kVecStride + q], sh = par[3 * kVecStride + q];
      v[4 * q] = fmaf(v[4 * q], sc.x, sh.x); v[4 * q + 1] = fmaf(v[4 * q + 1], sc.y, sh.y);
      v[4 * q + 2] = fmaf(v[4 * q + 2], sc.z, sh.z); v[4 * q + 3] = fmaf(v[4 * q + 3], sc.w, sh.w);
    }
  }
  // one range check per unit (four independent max chains)
  float m0 = 0.f, m1 = 0.f, m2 = 0.f, m3 = 0.f;
#pragma unroll
  for (int e = 0; e < 32; e += 4) {
    m0 = fmaxf(m0, fabsf(v[e])); m1 = fmaxf(m1, fabsf(v[e + 1]));
    m2 = fmaxf(m2, fabsf(v[e + 2])); m3 = fmaxf(m3, fabsf(v[e + 3]));
  }
  const float amax = fmaxf(fmaxf(m0, m1), fmaxf(m2, m3));
  if (!(amax <= 65504.0f)) {   // inputs are finite fp16 planes, so a NaN can only follow an overflow flagged upstream
    atomicOr(status, kFlagOverflow);
#pragma unroll
    for (int e = 0; e < 32; ++e) v[e] = fminf(fmaxf(v[e], -65504.0f), 65504.0f);
  }
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const __half2 h = __floats2half2_rn(v[2 * i], v[2 * i + 1]);
    hw[i] = *reinterpret_cast<const uint32_t*>(&h);
    if (THREE) {
      const float2 hf = __half22float2(h);
      const __half2 l = __floats2half2_rn(fmaf(v[2 * i], kLoScale, -kLoScale * hf.x),
                                          fmaf(v[2 * i + 1], kLoScale, -kLoScale * hf.y));
      lw[i] = *reinterpret_cast<const uint32_t*>(&l);
    } else {
      lw[i] = 0u;
    }
  }
}

}  // namespace crdr
