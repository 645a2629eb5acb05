// tc_common.cuh -- PTX wrappers (mbarrier / TMA / tcgen05 / TMEM) and the staged epilogue shared by the tcgen05 GEMM
// kernels (gemm_tc.cu: 128-row tiles; gemm_tc2.cu: 256-row tiles with split A/B rings and x-halo reuse).
#pragma once
#include <cuda.h>

#include "common.cuh"

namespace dcb {

constexpr int TC_BM = 128;
constexpr int TC_BK = 64;                      // 64 bf16 = 128 B = one swizzle-128B row
constexpr int TC_A_BYTES = TC_BM * TC_BK * 2;  // 16 KB
constexpr int TC_SMEM_LIMIT = 227 * 1024;
constexpr int TC_BAR_BYTES = 256;
// staged epilogue: bf16 [128][128] tile + bias[256] + rowvec[128] + gate[128] + row ids [2][128] + residual rows [2][128]
constexpr int TC_EPI_BYTES = 128 * 128 * 2 + 512 * 4 + 4 * 128 * 4;

struct TcSeg {
  int map;  // which A tensor map
  int c0;   // channel coordinate (dim 0) of the first K block (includes the x-parity offset for stride 2)
  int dx, p, dy;
  int nkb;  // K blocks (of 64) in this segment
  int div;  // source sample = output sample / div (shared class-independent prefix), >= 1
};


// ---- PTX wrappers ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// arrive on the barrier at the same shared-memory offset in CTA `cta` of this cluster
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t bar, uint32_t cta) {
  asm volatile(
      "{\n\t.reg .b32 ra;\n\t"
      "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
      "mbarrier.arrive.shared::cluster.b64 _, [ra];\n\t}"
      ::"r"(bar), "r"(cta)
      : "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t cluster_cta_rank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint64_t globaltimer_ns() {
  uint64_t t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
// Bounded wait: a protocol bug becomes a trap (an error the host sees) instead of a hung GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  uint64_t t0 = 0;
  for (uint32_t spins = 0;; ++spins) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    if (done) return;
    if ((spins & 0x3ff) == 0x3ff) {
      uint64_t now = globaltimer_ns();
      if (t0 == 0) t0 = now;
      else if (now - t0 > 4000000000ull) {
        printf("dcb gemm_tc: mbarrier wait timeout (block %d thread %d bar %u parity %u)\n", blockIdx.x, threadIdx.x,
               bar, parity);
        __trap();
      }
    }
  }
}
// Waits that normally last thousands of cycles (an epilogue warp waiting for a whole main loop, a producer waiting for a ring
// slot, the transform warps waiting for a TMA box): the same try_wait with a suspend-time hint, so the warp sleeps in
// hardware instead of re-issuing the ~10-instruction poll loop every ~100 cycles.  ncu of the fused conv showed 47 % of all
// executed warp instructions in such loops -- issue slots taken from the transform / epilogue warps of the same scheduler,
// and power in a step that runs against the 1 kW cap.  Only the MMA warps' operand waits keep the tight form.
#ifndef DCB_WAIT_HINT_NS
#define DCB_WAIT_HINT_NS 1000
#endif
__device__ __forceinline__ void mbar_wait_long(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  uint64_t t0 = 0;
  for (uint32_t spins = 0;; ++spins) {
    if (DCB_WAIT_HINT_NS > 0)
      asm volatile(
          "{\n\t.reg .pred p;\n\t"
          "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
          "selp.u32 %0, 1, 0, p;\n\t}"
          : "=r"(done)
          : "r"(bar), "r"(parity), "r"((uint32_t)DCB_WAIT_HINT_NS)
          : "memory");
    else
      asm volatile(
          "{\n\t.reg .pred p;\n\t"
          "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
          "selp.u32 %0, 1, 0, p;\n\t}"
          : "=r"(done)
          : "r"(bar), "r"(parity)
          : "memory");
    if (done) return;
    if ((spins & 0x3ff) == 0x3ff) {
      uint64_t now = globaltimer_ns();
      if (t0 == 0) t0 = now;
      else if (now - t0 > 4000000000ull) {
        printf("dcb gemm_tc: mbarrier wait timeout (block %d thread %d bar %u parity %u)\n", blockIdx.x, threadIdx.x, bar, parity);
        __trap();
      }
    }
  }
}
// spin on the non-suspending form (producer-side waits on barriers signalled by tcgen05.commit)
__device__ __forceinline__ void mbar_spin(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  uint64_t t0 = 0;
  for (uint32_t spins = 0;; ++spins) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    if (done) return;
    if ((spins & 0xfff) == 0xfff) {
      uint64_t now = globaltimer_ns();
      if (t0 == 0) t0 = now;
      else if (now - t0 > 4000000000ull) {
        printf("dcb gemm_tc: mbarrier spin timeout (block %d thread %d bar %u parity %u)\n", blockIdx.x, threadIdx.x, bar, parity);
        __trap();
      }
    }
  }
}
__device__ __forceinline__ void tma_load_5d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2,
                                            int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
// ---- CTA pairs (tcgen05 cta_group::2) ----------------------------------------------------------------------
__device__ __forceinline__ void tma_load_2d_2sm(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  // executed by both CTAs of the pair: the peer bit of the barrier address is cleared so the bytes are counted on CTA 0's barrier
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar & 0xFEFFFFFFu), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void umma_f16_2sm(uint32_t tmem_d, uint32_t alo, uint32_t blo, uint32_t hi, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
      "mov.b64 da, {%1, %3};\n\tmov.b64 db, {%2, %3};\n\t"
      "setp.ne.b32 p, %5, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %4, p;\n\t}"
      ::"r"(tmem_d), "r"(alo), "r"(blo), "r"(hi), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit_2sm(uint32_t bar) {     // arrives on the barrier at this offset in BOTH CTAs
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(bar), "h"((uint16_t)3)
               : "memory");
}

__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                         uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// same MMA with the two 64-bit smem descriptors given as (low word, shared high word): keeps the issue loop free of
// 64-bit arithmetic and lets ptxas hold everything in uniform registers
__device__ __forceinline__ void umma_f16_lohi(uint32_t tmem_d, uint32_t alo, uint32_t blo, uint32_t hi, uint32_t idesc,
                                              uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
      "mov.b64 da, {%1, %3};\n\tmov.b64 db, {%2, %3};\n\t"
      "setp.ne.b32 p, %5, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %4, p;\n\t}"
      ::"r"(tmem_d), "r"(alo), "r"(blo), "r"(hi), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_f16_lohi2(uint32_t tmem_d, uint32_t alo, uint32_t ahi, uint32_t blo, uint32_t bhi,
                                               uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
      "mov.b64 da, {%1, %2};\n\tmov.b64 db, {%3, %4};\n\t"
      "setp.ne.b32 p, %6, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t}"
      ::"r"(tmem_d), "r"(alo), "r"(ahi), "r"(blo), "r"(bhi), "r"(idesc), "r"(accumulate)
      : "memory");
}
// 32 lanes x 16 consecutive fp32 columns: thread t of the warp receives lane (base_lane + t)
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* v) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld16_nowait(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// same, carrying a (fake) dependency on the loaded registers so that no use of them can be scheduled above the wait
__device__ __forceinline__ void tmem_ld_wait_dep(uint32_t* r) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "+r"(r[8]),
                 "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15])
               :
               : "memory");
}
// named barrier of one epilogue group (4 warps); group g uses barrier id 1 + g
__device__ __forceinline__ void epi_bar(int id) { asm volatile("bar.sync %0, 128;" ::"r"(id) : "memory"); }
__device__ __forceinline__ void bar_sync_n(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }
__device__ __forceinline__ void bar_arrive_n(int id, int n) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(n) : "memory"); }
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(pred));
  return pred != 0;
}

// K-major, 128B-swizzled operand tile ([rows][64 bf16], 8-row atoms of 1024 B): SBO = 1024 B, LBO unused (=1),
// descriptor version 1 (Blackwell), layout type 2 (SWIZZLE_128B).  cf. cute::UMMA::SmemDescriptor.
__device__ __forceinline__ uint64_t make_kmajor_sw128_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

// descriptor for an operand that starts at an arbitrary 128-byte row of a 1024-byte-aligned swizzled buffer (x-halo reuse):
// base_offset (bits 49-51) = (start address >> 7) & 7 when the start is not aligned to the swizzle repeat
__device__ __forceinline__ uint64_t make_kmajor_sw128_desc_off(uint32_t smem_addr, int base_off_mode) {
  uint64_t d = make_kmajor_sw128_desc(smem_addr);
  if (base_off_mode) d |= (uint64_t)((smem_addr >> 7) & 7) << 49;
  return d;
}

// 16 consecutive floats from shared memory (16-byte aligned address)
// explicit shared-space accesses for the staged epilogues: through C++ pointers derived from the dynamic shared memory
// base the compiler emits GENERIC loads / stores (LD.E / ST.E plus a 64-bit address per access: ncu source view of
// gemm_tc2x, one 534-instruction copy-out block = 39 % of the epilogue warps' instructions)
__device__ __forceinline__ uint4 lds128(uint32_t a) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a) : "memory");
  return v;
}
__device__ __forceinline__ void sts128(uint32_t a, const uint4& v) {
  asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(a), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ void sts128f(uint32_t a, float x, float y, float z, float w) {
  asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(a), "f"(x), "f"(y), "f"(z), "f"(w) : "memory");
}
__device__ __forceinline__ int lds32(uint32_t a) {
  int v;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(v) : "r"(a) : "memory");
  return v;
}
__device__ __forceinline__ void sts32(uint32_t a, int v) { asm volatile("st.shared.b32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ void sts32f(uint32_t a, float v) { asm volatile("st.shared.f32 [%0], %1;" ::"r"(a), "f"(v) : "memory"); }
__device__ __forceinline__ float2 lds64f(uint32_t a) {
  float2 v;
  asm volatile("ld.shared.v2.f32 {%0,%1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(a) : "memory");
  return v;
}
__device__ __forceinline__ void lds16f(uint32_t addr, float* f) {
#pragma unroll
  for (int i = 0; i < 4; ++i)
    asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(f[4 * i]), "=f"(f[4 * i + 1]), "=f"(f[4 * i + 2]), "=f"(f[4 * i + 3])
                 : "r"(addr + 16 * i));
}

// v[i] += bias[n0 + i] + rowvec[group(m)][n0 + i]: every tcgen05 epilogue associates the two vectors first (the staged
// epilogue keeps their sum in shared memory), so a layer's result does not depend on which epilogue its tile shape selects
__device__ __forceinline__ void epi_add_bias_rowvec16(const EpiDev& e, int m, int n0, float* v) {
  const float* rv = nullptr;
  if (e.rowvec) {
    const int grp = e.rows_per_group > 0 ? m / e.rows_per_group : 0;
    rv = e.rowvec + (int64_t)(e.rowvec_idx ? e.rowvec_idx[grp] : grp) * e.rowvec_ld + n0;
  }
#pragma unroll
  for (int i = 0; i < 16; ++i)
    if (n0 + i < e.n_out) v[i] += (e.bias ? e.bias[n0 + i] : 0.f) + (rv ? rv[i] : 0.f);
}

// ---- epilogue for 16 consecutive output columns of one row ------------------------------------------------
// v[] holds acc (+bias already added by caller for GEGLU); n0 is the first OUTPUT column.
__device__ __forceinline__ void epi_store16(const EpiDev& e, int m, int n0, float* v, float& mse_acc, int sample,
                                            int pix, bool add_rowvec = true) {
  const int grp = e.rows_per_group > 0 ? m / e.rows_per_group : 0;
  const int nvalid = min(16, e.n_out - n0);
  if (e.rowvec && add_rowvec) {
    const float* rv = e.rowvec + (int64_t)(e.rowvec_idx ? e.rowvec_idx[grp] : grp) * e.rowvec_ld + n0;
#pragma unroll
    for (int i = 0; i < 16; ++i)
      if (i < nvalid) v[i] += rv[i];
  }
  // bf16 outputs use the one-MUFU activations of the staged epilogue (a layer must not depend on which epilogue its tile
  // shape selects); fp32 outputs keep the exact ones
  const bool fast_act = e.out != nullptr && e.out_dtype == DCB_BF16;
  if (e.act == DCB_ACT_SILU || e.act == DCB_ACT_GELU_TANH) {
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = fast_act ? apply_act_fast(e.act, v[i]) : apply_act(e.act, v[i]);
  }
  if (e.gate) {
    const float* gt = e.gate + (int64_t)grp * e.gate_ld + n0;
#pragma unroll
    for (int i = 0; i < 16; ++i)
      if (i < nvalid) v[i] *= gt[i];
  }
  if (e.residual) {
    const int64_t r = e.res_idx ? e.res_idx[m] : (e.res_mod > 0 ? m % e.res_mod : m);
    const int64_t off = r * e.res_ld + n0;
    if (e.res_dtype == DCB_BF16 && nvalid == 16 && (off & 7) == 0) {
      const uint4* rp = reinterpret_cast<const uint4*>((const __nv_bfloat16*)e.residual + off);
      float f[8];
      unpack_bf16x8(rp[0], f);
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] += f[i];
      unpack_bf16x8(rp[1], f);
#pragma unroll
      for (int i = 0; i < 8; ++i) v[8 + i] += f[i];
    } else {
#pragma unroll
      for (int i = 0; i < 16; ++i)
        if (i < nvalid) v[i] += load_as_f(e.residual, e.res_dtype, off + i);
    }
  }
  if (e.act_post != DCB_ACT_NONE) {
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = fast_act ? apply_act_fast(e.act_post, v[i]) : apply_act(e.act_post, v[i]);
  }
  if (e.mse_part) {
    const float sc = e.mse_scale ? e.mse_scale[sample] : 1.f;
    const float* tg = e.mse_target + ((int64_t)(sample / e.mse_div) * e.rows_per_sample + pix) * e.mse_ld + n0;
#pragma unroll
    for (int i = 0; i < 16; ++i)
      if (i < nvalid) {
        float d = sc * v[i] - tg[i];
        mse_acc = fmaf(d, d, mse_acc);
      }
  }
  if (e.out) {
    const int64_t off = out_row_of(e, m) * e.out_ld + n0;
    if (e.out_dtype == DCB_BF16 && nvalid == 16 && (off & 7) == 0) {
      uint4* op = reinterpret_cast<uint4*>((__nv_bfloat16*)e.out + off);
      op[0] = pack_bf16x8(v);
      op[1] = pack_bf16x8(v + 8);
    } else if (e.out_dtype == DCB_F32 && nvalid == 16 && (off & 3) == 0) {
      float4* op = reinterpret_cast<float4*>((float*)e.out + off);
#pragma unroll
      for (int i = 0; i < 4; ++i) op[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
    } else {
#pragma unroll
      for (int i = 0; i < 16; ++i)
        if (i < nvalid) store_from_f(e.out, e.out_dtype, off + i, v[i]);
    }
  }
}

// ---- staged epilogue of one 128-row accumulator sub-tile (bf16 outputs, <= 128 output columns per tile) --------
//  0. per-row ids + tile-constant bias/rowvec/gate vectors -> smem
//  1. cp.async prefetch of the whole residual tile (32 KB in flight per SM) into the swizzled staging tile
//  2. thread-per-row: TMEM -> regs -> bias, rowvec, act, gate, +residual (smem), act_post -> bf16 in place
//  3. coalesced copy-out: 16 threads cover one 256-byte output row
// Called by the 4 warps of one epilogue group (128 threads, named barrier bar_id).  stg8: TC_EPI_BYTES of smem per group.
struct EpiGeom {
  int tiles_x, tiles_y, bw, bh, bn, OW, OH, NB, uniform;
};

__device__ __forceinline__ void staged_epilogue(const EpiGeom& gq, const EpiDev& e, uint8_t* stg8, int parity, int tm_lin,
                                                int tn, int BN, uint32_t taddr, uint32_t full_bar, uint32_t full_parity,
                                                bool do_wait, uint32_t empty_bar, bool do_release, int bar_id = 1) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q = warp & 3;
  const int r = q * 32 + lane;           // accumulator row (TMEM lane) owned in the thread-per-row pass
  const int et = (threadIdx.x & 127);    // any bijection onto 0..127 works for the coalesced passes
  const bool geglu = e.act == DCB_ACT_GEGLU;
  int tm = tm_lin;
  const int tx = tm % gq.tiles_x;
  tm /= gq.tiles_x;
  const int ty = tm % gq.tiles_y;
  const int tb = tm / gq.tiles_y;
  const int xl = r % gq.bw, yl = (r / gq.bw) % gq.bh, nl = r / (gq.bw * gq.bh);
  const int x = tx * gq.bw + xl, y = ty * gq.bh + yl, nb = tb * gq.bn + nl;
  const bool row_ok = x < gq.OW && y < gq.OH && nb < gq.NB;
  const int m = nb * e.rows_per_sample + y * gq.OW + x;
    // shared-space byte addresses (see lds128): staging tile [128 rows x 256 B], bias [256] fp32, rowvec [128], gate [128],
    // row ids [2][128], residual row ids [2][128]
    const uint32_t stg = smem_u32(stg8);
    const uint32_t s_bias = stg + 128 * 256;
    const uint32_t s_rowvec = s_bias + 256 * 4;
    const uint32_t s_gate = s_rowvec + 128 * 4;
    const uint32_t s_m = s_gate + 128 * 4 + parity * 128 * 4;
    const uint32_t s_res = s_gate + 128 * 4 + 256 * 4 + parity * 128 * 4;
            const int wrow0 = tn * BN;
    const int ncols_out = geglu ? 128 : BN;
    const int ocol0 = geglu ? tn * 128 : tn * BN;
    // output row: identity, or the (2y + a, 2x + b) position of a folded-upsample phase
    const int up_a = (e.up_phase - 1) >> 1, up_b = (e.up_phase - 1) & 1;
    const int m_out = e.up_phase ? ((nb * 2 * gq.OH + 2 * y + up_a) * (2 * gq.OW) + 2 * x + up_b) : m;
    sts32(s_m + r * 4, row_ok ? m_out : -1);
    if (e.residual) sts32(s_res + r * 4, row_ok ? (e.res_idx ? e.res_idx[m] : (e.res_mod > 0 ? m % e.res_mod : m)) : -1);
    const bool rv_folded = gq.uniform && e.rowvec != nullptr && !geglu;
    float my_bias = 0.f;     // s_bias[et] as this thread wrote it (first trip of the loop)
    for (int c = et; c < BN; c += 128) {
      const float bv = (e.bias && wrow0 + c < e.N) ? e.bias[wrow0 + c] : 0.f;
      if (c == et) my_bias = bv;
      sts32f(s_bias + c * 4, bv);
    }
    if (gq.uniform && (e.rowvec || e.gate) && et < ncols_out) {
      const int m0 = (tb * gq.bn) * e.rows_per_sample + (ty * gq.bh) * gq.OW + tx * gq.bw;
      const int grp0 = m0 / e.rows_per_group;
      const bool ok = ocol0 + et < e.n_out;
      if (e.rowvec) {
        const float rvv = ok ? e.rowvec[(int64_t)(e.rowvec_idx ? e.rowvec_idx[grp0] : grp0) * e.rowvec_ld + ocol0 + et] : 0.f;
        if (rv_folded) sts32f(s_bias + et * 4, my_bias + rvv);   // thread et wrote s_bias[et] just above: acc + (bias + rowvec)
        else sts32f(s_rowvec + et * 4, rvv);
      }
      if (e.gate) sts32f(s_gate + et * 4, ok ? e.gate[(int64_t)grp0 * e.gate_ld + ocol0 + et] : 0.f);
    }
    epi_bar(bar_id);  // ids/constants visible; everybody has finished copying the previous tile out of the staging tile
    const int cc = et & 15, rr0 = et >> 4;            // coalesced role: 16-byte chunk cc of rows rr0, rr0+8, ...
    const bool cc_ok = cc * 8 < ncols_out && ocol0 + cc * 8 + 8 <= e.n_out;
    if (e.residual && cc_ok) {
      const __nv_bfloat16* rbase = (const __nv_bfloat16*)e.residual + ocol0 + cc * 8;
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const int row = rr0 + 8 * i;
        const int rrow = lds32(s_res + row * 4);
        if (rrow >= 0)
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(stg + row * 256 + ((cc ^ (row & 7)) << 4)),
                       "l"(rbase + (int64_t)rrow * e.res_ld)
                       : "memory");
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
    }
    if (do_wait) {
      mbar_wait_long(full_bar, full_parity);
      tc_fence_after();
    }
    if (e.residual) {
      asm volatile("cp.async.wait_group 0;" ::: "memory");
      epi_bar(bar_id);
    }
    const int grp = (!gq.uniform && e.rows_per_group > 0 && row_ok) ? m / e.rows_per_group : 0;
    // ---- thread-per-row pass: 16 columns per step; the TMEM load of step i+1 is in flight while step i is processed ----
    auto process = [&](int c, const uint32_t (&ra)[16], const uint32_t (&rg)[16]) {
      float v[16], bz[16], bg[16];
      // tile-constant vectors through explicit 16-byte shared loads (the generic-pointer form compiles to LD.E);
      // s_bias already holds bias + rowvec when the tile has one rowvec group (rv_folded)
      lds16f(s_bias + c * 4, bz);
      if (geglu) lds16f(s_bias + (128 + c) * 4, bg);
      if (e.rowvec && !rv_folded && !geglu && row_ok) {   // row-dependent group: same (bias + rowvec) association
        const float* rv = e.rowvec + (int64_t)(e.rowvec_idx ? e.rowvec_idx[grp] : grp) * e.rowvec_ld + ocol0 + c;
#pragma unroll
        for (int i = 0; i < 16; ++i)
          if (ocol0 + c + i < e.n_out) bz[i] += rv[i];
      }
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        float a = __uint_as_float(ra[i]) + bz[i];
        if (geglu) a *= gelu_erf_fast_f(__uint_as_float(rg[i]) + bg[i]);
        v[i] = a;
      }
      if (e.rowvec && geglu) {
        if (gq.uniform) {
          float rvv[16];
          lds16f(s_rowvec + c * 4, rvv);
#pragma unroll
          for (int i = 0; i < 16; ++i) v[i] += rvv[i];
        } else if (row_ok) {
          const float* rv = e.rowvec + (int64_t)(e.rowvec_idx ? e.rowvec_idx[grp] : grp) * e.rowvec_ld + ocol0 + c;
#pragma unroll
          for (int i = 0; i < 16; ++i)
            if (ocol0 + c + i < e.n_out) v[i] += rv[i];
        }
      }
      if (e.act == DCB_ACT_SILU || e.act == DCB_ACT_GELU_TANH) {
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = apply_act_fast(e.act, v[i]);
      }
      if (e.gate) {
        if (gq.uniform) {
          float gvv[16];
          lds16f(s_gate + c * 4, gvv);
#pragma unroll
          for (int i = 0; i < 16; ++i) v[i] *= gvv[i];
        } else if (row_ok) {
          const float* gt = e.gate + (int64_t)grp * e.gate_ld + ocol0 + c;
#pragma unroll
          for (int i = 0; i < 16; ++i)
            if (ocol0 + c + i < e.n_out) v[i] *= gt[i];
        }
      }
      const uint32_t s0 = stg + r * 256 + ((((c >> 3)) ^ (r & 7)) << 4);
      const uint32_t s1 = stg + r * 256 + ((((c >> 3) + 1) ^ (r & 7)) << 4);
      if (e.residual && row_ok) {
        float f[8];
        unpack_bf16x8(lds128(s0), f);
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] += f[i];
        unpack_bf16x8(lds128(s1), f);
#pragma unroll
        for (int i = 0; i < 8; ++i) v[8 + i] += f[i];
      }
      if (e.act_post != DCB_ACT_NONE) {
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = apply_act_fast(e.act_post, v[i]);
      }
      sts128(s0, pack_bf16x8(v));
      sts128(s1, pack_bf16x8(v + 8));
    };
    if (geglu) {
      for (int c = 0; c < ncols_out; c += 16) {
        uint32_t ra[16], rg[16];
        tmem_ld16_nowait(taddr + (uint32_t)c, ra);
        tmem_ld16_nowait(taddr + (uint32_t)(128 + c), rg);
        tmem_ld_wait_dep(ra);
        process(c, ra, rg);
      }
    } else {
      uint32_t ra[16], rb[16];
      tmem_ld16_nowait(taddr, ra);
      for (int c = 0; c < ncols_out; c += 32) {
        tmem_ld_wait_dep(ra);
        const bool more1 = c + 16 < ncols_out;
        if (more1) tmem_ld16_nowait(taddr + (uint32_t)(c + 16), rb);
        process(c, ra, ra);
        if (more1) {
          tmem_ld_wait_dep(rb);
          if (c + 32 < ncols_out) tmem_ld16_nowait(taddr + (uint32_t)(c + 32), ra);
          process(c + 16, rb, rb);
        }
      }
    }
    // accumulator fully drained: hand the TMEM stage back to the MMA warp
    tc_fence_before();
    __syncwarp();
    if (do_release && lane == 0) mbar_arrive(empty_bar);
    epi_bar(bar_id);
    // ---- coalesced copy-out (+ per-column sum / sum of squares of the stored values for the next GroupNorm) ----
    // Canonical summation tree of a column over the tile's 128 rows (every staged epilogue follows it, so a layer's
    // statistics do not depend on which kernel wrote it): P[k] = rows k, k + 16, ..., k + 112 in order (k = 0..15);
    // Q[k] = P[k] + P[k + 8]; S[w] = Q[2w] + Q[2w + 1]; total = (S[0] + S[1]) + (S[2] + S[3]).
    float cs[8], cq[8], ds[8], dq[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) cs[i] = cq[i] = ds[i] = dq[i] = 0.f;
    if (cc_ok) {
      __nv_bfloat16* obase = (__nv_bfloat16*)e.out + ocol0 + cc * 8;
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const int row = rr0 + 8 * i;
        const int mm = lds32(s_m + row * 4);
        if (mm >= 0) {
          const uint4 val = lds128(stg + row * 256 + ((cc ^ (row & 7)) << 4));
          *reinterpret_cast<uint4*>(obase + (int64_t)mm * e.out_ld) = val;
          if (e.gn_part) {
            float f[8];
            unpack_bf16x8(val, f);
            if (i & 1) {        // rows rr0 + 8 + 16 j: P[rr0 + 8]
#pragma unroll
              for (int j = 0; j < 8; ++j) { ds[j] += f[j]; dq[j] = fmaf(f[j], f[j], dq[j]); }
            } else {            // rows rr0 + 16 j: P[rr0]
#pragma unroll
              for (int j = 0; j < 8; ++j) { cs[j] += f[j]; cq[j] = fmaf(f[j], f[j], cq[j]); }
            }
          }
        }
      }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) { cs[j] += ds[j]; cq[j] += dq[j]; }     // Q[rr0]
    if (e.gn_part) {
      // fixed-order reduction over the 8 row-slices: lanes cc / cc+16 by shuffle, the 4 warps through the (now free)
      // staging tile; thread c then owns output column c of this tile
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        cs[j] += __shfl_xor_sync(0xffffffffu, cs[j], 16);
        cq[j] += __shfl_xor_sync(0xffffffffu, cq[j], 16);
      }
      epi_bar(bar_id);  // every thread has read its rows out of the staging tile
      const uint32_t red = stg;  // [4 warps][128 columns][2] fp32
      const int wq = et >> 5;
      if ((et & 16) == 0) {
#pragma unroll
        for (int j = 0; j < 8; j += 2)
          sts128f(red + ((wq * 128 + cc * 8 + j) * 2) * 4, cs[j], cq[j], cs[j + 1], cq[j + 1]);
      }
      epi_bar(bar_id);
      if (et < ncols_out && ocol0 + et < e.n_out && tb * gq.bn < gq.NB) {  // (tc2's odd tail sub-tile has no slot)
        const float2 w0 = lds64f(red + (et * 2) * 4), w1 = lds64f(red + ((128 + et) * 2) * 4);
        const float2 w2 = lds64f(red + ((256 + et) * 2) * 4), w3 = lds64f(red + ((384 + et) * 2) * 4);
        const float s4 = (w0.x + w1.x) + (w2.x + w3.x);
        const float q4 = (w0.y + w1.y) + (w2.y + w3.y);
        // phases of a folded upsample interleave their tiles per sample: [n][phase][tile of the low-resolution grid]
        const int tps = gq.tiles_x * gq.tiles_y;
        const int64_t gtile = e.up_phase ? ((int64_t)(tb * 4 + e.up_phase - 1) * tps + (tm_lin - tb * tps)) : (int64_t)tm_lin;
        float* gp = e.gn_part + (gtile * e.n_out + ocol0 + et) * 2;
        gp[0] = s4;
        gp[1] = q4;
      }
    }
}


// ---- staged epilogue through a HALF-width staging tile (64 output columns at a time, 16 KB + constants per group) --------
// Same arithmetic, same stores and the same statistics tree as staged_epilogue (no GEGLU): gemm_tc2x_kernel uses it so that
// both epilogue groups own a staging tile next to six row boxes.  stg8: TC_EPI_HALF_BYTES of smem per group.
constexpr int TC_EPI_HALF_BYTES = 128 * 64 * 2 + 512 * 4 + 4 * 128 * 4;

__device__ __forceinline__ void staged_epilogue_half(const EpiGeom& gq, const EpiDev& e, uint8_t* stg8, int parity, int tm_lin,
                                                     int tn, int BN, uint32_t taddr, uint32_t full_bar, uint32_t full_parity,
                                                     uint32_t empty_bar, int bar_id, int release_cta = -1) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q = warp & 3;
  const int r = q * 32 + lane;           // accumulator row (TMEM lane) owned in the thread-per-row pass
  const int et = (threadIdx.x & 127);
  int tm = tm_lin;
  const int tx = tm % gq.tiles_x;
  tm /= gq.tiles_x;
  const int ty = tm % gq.tiles_y;
  const int tb = tm / gq.tiles_y;
  const int xl = r % gq.bw, yl = (r / gq.bw) % gq.bh, nl = r / (gq.bw * gq.bh);
  const int x = tx * gq.bw + xl, y = ty * gq.bh + yl, nb = tb * gq.bn + nl;
  const bool row_ok = x < gq.OW && y < gq.OH && nb < gq.NB;
  const int m = nb * e.rows_per_sample + y * gq.OW + x;
  // shared-space byte addresses (see lds128): staging tile [128 rows x 128 B], bias [256] fp32, gate [128], row ids [2][128],
  // residual row ids [2][128]
  const uint32_t stg = smem_u32(stg8);
  const uint32_t s_bias = stg + 128 * 128;
  const uint32_t s_gate = s_bias + (256 + 128) * 4;
  const uint32_t s_m = s_gate + 128 * 4 + parity * 128 * 4;
  const uint32_t s_res = s_gate + 128 * 4 + 256 * 4 + parity * 128 * 4;
  // GEGLU (gemm_tc3 only; BN == 64): this call owns output columns [tn * 64, + 64); their value accumulators start at
  // taddr, their gate accumulators 128 TMEM columns further, and the packed weight rows are [128 value | 128 gate] per
  // 128 outputs.  Same arithmetic as staged_epilogue: (a + b_a) * gelu_erf_fast(g + b_g).
  const bool geglu = e.act == DCB_ACT_GEGLU;
  const int ncols_out = BN;
  const int ocol0 = tn * BN;
  const int wrow0 = geglu ? (ocol0 >> 7) * 256 + (ocol0 & 127) : tn * BN;
  const int up_a = (e.up_phase - 1) >> 1, up_b = (e.up_phase - 1) & 1;
  const int m_out = e.up_phase ? ((nb * 2 * gq.OH + 2 * y + up_a) * (2 * gq.OW) + 2 * x + up_b) : m;
  sts32(s_m + r * 4, row_ok ? m_out : -1);
  if (e.residual) sts32(s_res + r * 4, row_ok ? (e.res_idx ? e.res_idx[m] : (e.res_mod > 0 ? m % e.res_mod : m)) : -1);
  const bool rv_folded = gq.uniform && e.rowvec != nullptr;
  float my_bias = 0.f;     // s_bias[et] as this thread wrote it (BN <= 128: one entry per thread)
  for (int c = et; c < BN; c += 128) {
    my_bias = (e.bias && wrow0 + c < e.N) ? e.bias[wrow0 + c] : 0.f;
    sts32f(s_bias + c * 4, my_bias);
  }
  if (geglu && et < BN) sts32f(s_bias + (128 + et) * 4, e.bias ? e.bias[wrow0 + 128 + et] : 0.f);
  if (gq.uniform && (e.rowvec || e.gate) && et < ncols_out) {
    const int m0 = (tb * gq.bn) * e.rows_per_sample + (ty * gq.bh) * gq.OW + tx * gq.bw;
    const int grp0 = m0 / e.rows_per_group;
    const bool ok = ocol0 + et < e.n_out;
    if (e.rowvec)
      sts32f(s_bias + et * 4,
             my_bias + (ok ? e.rowvec[(int64_t)(e.rowvec_idx ? e.rowvec_idx[grp0] : grp0) * e.rowvec_ld + ocol0 + et] : 0.f));
    if (e.gate) sts32f(s_gate + et * 4, ok ? e.gate[(int64_t)grp0 * e.gate_ld + ocol0 + et] : 0.f);
  }
  epi_bar(bar_id);  // ids / constants visible; the previous tile's copy-out is complete
  const int cc = et & 7, rr0 = et >> 3;     // coalesced role: 16-byte chunk cc of rows rr0, rr0 + 16, ...
  const int grp = (!gq.uniform && e.rows_per_group > 0 && row_ok) ? m / e.rows_per_group : 0;
  bool waited = false;
  for (int h0 = 0; h0 < ncols_out; h0 += 64) {
    const bool cc_ok = h0 + cc * 8 < ncols_out && ocol0 + h0 + cc * 8 + 8 <= e.n_out;
    if (h0 > 0) epi_bar(bar_id);            // the previous half has been copied out (and its statistics scratch read)
    if (e.residual && cc_ok) {
      const __nv_bfloat16* rbase = (const __nv_bfloat16*)e.residual + ocol0 + h0 + cc * 8;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int row = rr0 + 16 * i;
        const int rrow = lds32(s_res + row * 4);
        if (rrow >= 0)
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(stg + row * 128 + ((cc ^ (row & 7)) << 4)),
                       "l"(rbase + (int64_t)rrow * e.res_ld)
                       : "memory");
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
    }
    if (!waited) {
      mbar_wait_long(full_bar, full_parity);
      tc_fence_after();
      waited = true;
    }
    if (e.residual) {
      asm volatile("cp.async.wait_group 0;" ::: "memory");
      epi_bar(bar_id);
    }
    float ss = 0.f;     // |row|^2 over this 64-column half: one attention head of q or k (EpiDev::attn_norms)
    auto process = [&](int c, const uint32_t (&ra)[16]) {     // c: column inside this half
      float v[16], bz[16];
      lds16f(s_bias + (h0 + c) * 4, bz);
      if (e.rowvec && !rv_folded && row_ok) {
        const float* rv = e.rowvec + (int64_t)(e.rowvec_idx ? e.rowvec_idx[grp] : grp) * e.rowvec_ld + ocol0 + h0 + c;
#pragma unroll
        for (int i = 0; i < 16; ++i)
          if (ocol0 + h0 + c + i < e.n_out) bz[i] += rv[i];
      }
#pragma unroll
      for (int i = 0; i < 16; i += 2)
        f2_unpack(f2_add(f2_pack(__uint_as_float(ra[i]), __uint_as_float(ra[i + 1])), f2_pack(bz[i], bz[i + 1])), v[i], v[i + 1]);
      if (e.act == DCB_ACT_SILU || e.act == DCB_ACT_GELU_TANH) {
#pragma unroll
        for (int i = 0; i < 16; i += 2) act_fast_pair(e.act, v[i], v[i + 1]);
      }
      if (e.gate) {
        if (gq.uniform) {
          float gvv[16];
          lds16f(s_gate + (h0 + c) * 4, gvv);
#pragma unroll
          for (int i = 0; i < 16; ++i) v[i] *= gvv[i];
        } else if (row_ok) {
          const float* gt = e.gate + (int64_t)grp * e.gate_ld + ocol0 + h0 + c;
#pragma unroll
          for (int i = 0; i < 16; ++i)
            if (ocol0 + h0 + c + i < e.n_out) v[i] *= gt[i];
        }
      }
      const uint32_t s0 = stg + r * 128 + ((((c >> 3)) ^ (r & 7)) << 4);
      const uint32_t s1 = stg + r * 128 + ((((c >> 3) + 1) ^ (r & 7)) << 4);
      if (e.residual && row_ok) {
        float f[8];
        unpack_bf16x8(lds128(s0), f);
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] += f[i];
        unpack_bf16x8(lds128(s1), f);
#pragma unroll
        for (int i = 0; i < 8; ++i) v[8 + i] += f[i];
      }
      if (e.act_post != DCB_ACT_NONE) {
#pragma unroll
        for (int i = 0; i < 16; i += 2) act_fast_pair(e.act_post, v[i], v[i + 1]);
      }
      if (e.attn_norms != nullptr) {
#pragma unroll
        for (int i = 0; i < 16; ++i) ss = fmaf(v[i], v[i], ss);
      }
      sts128(s0, pack_bf16x8(v));
      sts128(s1, pack_bf16x8(v + 8));
    };
    if (geglu) {
      for (int c = 0; c < 64; c += 16) {
        uint32_t ra[16], rg[16];
        tmem_ld16_nowait(taddr + (uint32_t)c, ra);
        tmem_ld16_nowait(taddr + (uint32_t)(128 + c), rg);
        tmem_ld_wait_dep(ra);
        float v[16], bz[16], bg[16];
        lds16f(s_bias + c * 4, bz);
        lds16f(s_bias + (128 + c) * 4, bg);
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          float a_ = __uint_as_float(ra[i]) + bz[i];
          a_ *= gelu_erf_fast_f(__uint_as_float(rg[i]) + bg[i]);
          v[i] = a_;
        }
        sts128(stg + r * 128 + ((((c >> 3)) ^ (r & 7)) << 4), pack_bf16x8(v));
        sts128(stg + r * 128 + ((((c >> 3) + 1) ^ (r & 7)) << 4), pack_bf16x8(v + 8));
      }
    } else {
      const int ncol_h = min(64, ncols_out - h0);
      uint32_t ra[16], rb[16];
      tmem_ld16_nowait(taddr + (uint32_t)h0, ra);
      for (int c = 0; c < ncol_h; c += 32) {
        tmem_ld_wait_dep(ra);
        const bool more1 = c + 16 < ncol_h;
        if (more1) tmem_ld16_nowait(taddr + (uint32_t)(h0 + c + 16), rb);
        process(c, ra);
        if (more1) {
          tmem_ld_wait_dep(rb);
          if (c + 32 < ncol_h) tmem_ld16_nowait(taddr + (uint32_t)(h0 + c + 32), ra);
          process(c + 16, rb);
        }
      }
    }
    if (h0 + 64 >= ncols_out) {     // accumulator fully drained: hand the TMEM stage back to the MMA warp
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (release_cta < 0) mbar_arrive(empty_bar);
        else mbar_arrive_cluster(empty_bar, (uint32_t)release_cta);    // CTA pair: the MMA warp lives in the leader CTA
      }
    }
    if (e.attn_norms != nullptr && ocol0 + h0 < 2 * e.attn_heads * 64) {
      // the 64 columns of this half are one head of q (columns below heads * 64) or k: max over the warp's 32 rows of
      // |row|^2, then one atomic per warp.  attn_tok % 128 == 0: the tile's rows belong to one sample.  Non-negative
      // floats order like their bit patterns, so an integer max does the reduction (attention_tc.cu reads it back).
      float mx = row_ok ? ss : 0.f;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
      const int m_any = __shfl_sync(0xffffffffu, m, 0);
      if (lane == 0 && row_ok) {   // rows % 128 == 0: a warp's rows exist together or not at all (odd tile count of a CTA pair)
        const int col = ocol0 + h0, which = col >= e.attn_heads * 64 ? 1 : 0;
        const int head = (col >> 6) - which * e.attn_heads;
        atomicMax(reinterpret_cast<int*>(e.attn_norms) + 2 + ((m_any / e.attn_tok) * e.attn_heads + head) * 2 + which,
                  __float_as_int(mx));
      }
    }
    epi_bar(bar_id);
    // ---- coalesced copy-out of this half + P[rr0] of the canonical statistics tree (see staged_epilogue) ----
    float cs[8], cq[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) cs[i] = cq[i] = 0.f;
    if (cc_ok) {
      __nv_bfloat16* obase = (__nv_bfloat16*)e.out + ocol0 + h0 + cc * 8;
      if (e.gn_part) {    // two loops: the (uniform) statistics test is not re-evaluated per row
        uint64_t cs2[4], cq2[4];   // two columns per packed add / fma: same order and rounding as the scalar form
#pragma unroll
        for (int j = 0; j < 4; ++j) cs2[j] = cq2[j] = f2_pack(0.f, 0.f);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int row = rr0 + 16 * i;
          const int mm = lds32(s_m + row * 4);
          if (mm >= 0) {
            const uint4 val = lds128(stg + row * 128 + ((cc ^ (row & 7)) << 4));
            *reinterpret_cast<uint4*>(obase + (int64_t)mm * e.out_ld) = val;
            float f[8];
            unpack_bf16x8(val, f);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const uint64_t f2 = f2_pack(f[2 * j], f[2 * j + 1]);
              cs2[j] = f2_add(cs2[j], f2);
              cq2[j] = f2_fma(f2, f2, cq2[j]);
            }
          }
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          f2_unpack(cs2[j], cs[2 * j], cs[2 * j + 1]);
          f2_unpack(cq2[j], cq[2 * j], cq[2 * j + 1]);
        }
      } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int row = rr0 + 16 * i;
          const int mm = lds32(s_m + row * 4);
          if (mm >= 0) *reinterpret_cast<uint4*>(obase + (int64_t)mm * e.out_ld) = lds128(stg + row * 128 + ((cc ^ (row & 7)) << 4));
        }
      }
    }
    if (e.gn_part) {
      epi_bar(bar_id);  // every thread has read its rows out of the staging tile
      const uint32_t red = stg;  // [16 row classes][64 columns][2] fp32
#pragma unroll
      for (int j = 0; j < 8; j += 2)
        sts128f(red + (((rr0 * 64) + cc * 8 + j) * 2) * 4, cs[j], cq[j], cs[j + 1], cq[j + 1]);
      epi_bar(bar_id);
      if (et < 64 && h0 + et < ncols_out && ocol0 + h0 + et < e.n_out && tb * gq.bn < gq.NB) {
        float S[4], Q4[4];
#pragma unroll
        for (int w = 0; w < 4; ++w) {
          const float2 p0 = lds64f(red + (((2 * w) * 64 + et) * 2) * 4), p8 = lds64f(red + (((2 * w + 8) * 64 + et) * 2) * 4);
          const float2 p1 = lds64f(red + (((2 * w + 1) * 64 + et) * 2) * 4), p9 = lds64f(red + (((2 * w + 9) * 64 + et) * 2) * 4);
          const float qa = p0.x + p8.x;
          const float qb = p1.x + p9.x;
          S[w] = qa + qb;
          const float ra_ = p0.y + p8.y;
          const float rb_ = p1.y + p9.y;
          Q4[w] = ra_ + rb_;
        }
        const float s4 = (S[0] + S[1]) + (S[2] + S[3]);
        const float q4 = (Q4[0] + Q4[1]) + (Q4[2] + Q4[3]);
        const int tps = gq.tiles_x * gq.tiles_y;
        const int64_t gtile = e.up_phase ? ((int64_t)(tb * 4 + e.up_phase - 1) * tps + (tm_lin - tb * tps)) : (int64_t)tm_lin;
        float* gp = e.gn_part + (gtile * e.n_out + ocol0 + h0 + et) * 2;
        gp[0] = s4;
        gp[1] = q4;
      }
    }
  }
}

}  // namespace dcb
