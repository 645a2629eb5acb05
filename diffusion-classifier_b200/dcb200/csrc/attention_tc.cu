// attention_tc.cu -- tcgen05 / TMEM flash attention (forward, no mask): the running-maximum kernel documented here
// (head dims 32 / 64 / 96 / 128, described for 64; see AtGeo for the others) and the single-pass kernel further down
// (head dim 64, which the launcher prefers for N >= 1024): softmax(Q K^T * scale) V per
// (batch, head), replacing F.scaled_dot_product_attention behind diffusers' AttnProcessor2_0
// for the DiT blocks (12 heads x 64, N = 4096 tokens) and the U-Net Transformer2D blocks with C/8 = 64.
//
// One CTA = 256 queries (two 128-row tiles) of one (batch, head); keys/values stream in 128-row blocks.
//   warp 0      TMA producer: Q tiles once, then K_j / V_j boxes [128 tokens x 64 ch] (3-D maps: channel, token, batch;
//               out-of-range tokens are zero filled) into 128B-swizzled rings
//   warp 1      MMA issuer (warp-uniform loop, one elected lane):
//                 S_t = Q_t K_j^T      tcgen05.mma SS, M=128 N=128 K=64   -> TMEM S_t (fp32)
//                 O_t = P_t V_j        tcgen05.mma TS: A = P_t straight from TMEM (bf16 pairs), B = V_j as an MN-major
//                                      128B-swizzled smem operand (the [token][channel] box TMA delivers), N=64 K=128
//   warps 2-5 / 6-9   softmax group of tile 0 / 1, one TMEM lane = one query row per thread (no shuffles):
//                 pass 1 row max of S, pass 2 P = exp2(S*c - m*c) -> bf16 -> tcgen05.st into TMEM; the PV product of the
//                 previous block is folded into fp32 register accumulators (O = O*alpha + O_j) while the tensor pipe
//                 already runs the next QK^T, so nothing is ever rescaled in TMEM and the N x N scores never exist in HBM.
// TMEM columns: S0 [0,128) S1 [128,256) P0 [256,320) P1 [320,384) O0 [384,448) O1 [448,512).
//
// Roofline at d = 64: 128 x 128 exponentials per key block and tile = 1024 cycles of MUFU (16 ex2/clk/SM) against 768
// cycles of tensor pipe (QK^T 256 + PV 512 at the N = 64 half rate), and ~4 issue slots per score on the 4 schedulers
// (~1000 cycles): the kernel is bound by the softmax warps, not by the tensor pipe or by L2.  Measured (ncu, DiT-B/4
// shape): MUFU pipe 55 %, issue slots 41 %, tensor pipe 26 % busy -> 614 TF/s.  Variants tried and measured NOT faster
// (tools/attn_micro.py, git history): single-pass softmax with a lagged running maximum; 2 or 4 threads per query row
// (needs a per-block max exchange, which locks the warps of a scheduler in step); one tile per CTA with S, P and O all
// double buffered; K/V shared by 2 / 4 CTAs through TMA multicast (L2 -> SMEM is not the limit: 9.7 TB/s measured).
#include <cudaTypedefs.h>

#include <mutex>
#include <type_traits>

#include "tc_common.cuh"

namespace dcb {

constexpr int AT_THREADS = 320;
constexpr int AT_KV_STAGES = 4;
constexpr int AT_TILE_BYTES = 128 * 64 * 2;  // one [128 tokens x 64 ch] bf16 box

PFN_cuTensorMapEncodeTiled_v12000 tc_encode_fn();

__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem desc]
__device__ __forceinline__ void umma_f16_ts(uint32_t tmem_d, uint32_t tmem_a, uint32_t blo, uint32_t bhi, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 db;\n\t"
      "mov.b64 db, {%2, %3};\n\t"
      "setp.ne.b32 p, %5, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], db, %4, p;\n\t}"
      ::"r"(tmem_d), "r"(tmem_a), "r"(blo), "r"(bhi), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tmem_ld32_nowait(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t* r) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
        "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ float ex2f(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__device__ __forceinline__ void add_f32x2(float& a0, float& a1, float b0, float b1) {   // {a0, a1} += {b0, b1}
  f2_unpack(f2_add(f2_pack(a0, a1), f2_pack(b0, b1)), a0, a1);
}

// 2^x for |x| <= 126 on the FMA / ALU pipes (no MUFU op), two scores at a time: round-to-nearest split x = n + f through
// the 1.5 * 2^23 magic constant (n sits in the low mantissa bits of t), cubic minimax 2^f on [-0.5, 0.5] (max relative
// error 7.5e-5, far below the bf16 rounding of P), exponent add.  Packed: 6 f32x2 slots + 2 integer ops per PAIR of scores
// (4 per score), against 1 MUFU slot that occupies the 4-lane XU pipe for 8 clocks.
// History: the scalar form (8 slots per score) was a loss when the loop still carried a subtract, a multiply and the
// ragged-N compare + select on every score (issue slots 55 %: never 652 TF/s, one in eight 656, one in four 580).  With
// the loop down to ex2 + 1/2 add.f32x2 + 1/2 cvt per score the XU pipe is the only thing left to relieve.
// DCB_ATTN_POLY: 0 = never; 8 / 4 / 2 = one score in eight / four / two goes to the FMA pipe.
// MEASURED (tools/attn_micro.py, one box, B = 16 x 12 heads x 4096 x 64, incl. the norm pre-pass, query pre-scaled):
// never 772 TF/s, one in eight 821, one in four 912, one in two 961 -> the product uses one in two.
#ifndef DCB_ATTN_POLY
#define DCB_ATTN_POLY 2
#endif
__device__ __forceinline__ void ex2_poly2(float x0, float x1, float& p0, float& p1) {
  const uint64_t magic = f2_pack(12582912.f, 12582912.f), nmagic = f2_pack(-12582912.f, -12582912.f);
  const uint64_t neg1 = f2_pack(-1.f, -1.f);
  const uint64_t c3 = f2_pack(0.0551716685295105f, 0.0551716685295105f), c2 = f2_pack(0.2426111251115799f, 0.2426111251115799f);
  const uint64_t c1 = f2_pack(0.6932609677314758f, 0.6932609677314758f), c0 = f2_pack(0.9999280571937561f, 0.9999280571937561f);
  const uint64_t x = f2_pack(x0, x1);
  const uint64_t t = f2_add(x, magic);
  const uint64_t f = f2_fma(f2_add(t, nmagic), neg1, x);      // x - n
  uint64_t pl = f2_fma(c3, f, c2);
  pl = f2_fma(pl, f, c1);
  pl = f2_fma(pl, f, c0);
  float t0, t1, q0, q1;
  f2_unpack(t, t0, t1);
  f2_unpack(pl, q0, q1);
  p0 = __int_as_float(__float_as_int(q0) + (__float_as_int(t0) << 23));
  p1 = __int_as_float(__float_as_int(q1) + (__float_as_int(t1) << 23));
}

// the single-pass kernel is exact iff nothing underflows: max|q| max|k| c <= 50 bounds every (M_i - s_ij) c by 100 < 126
// Decided per (batch, head) -- blockIdx.z, blockIdx.y -- from that head's own max|q|^2, max|k|^2, so which of the two kernels
// scores a sample never depends on the other samples of the launch (results stay independent of the batch composition).
__device__ __forceinline__ bool at_fast_ok(const float* norms, float sc, int b, int h) {
  const float* n = norms + 2 + ((int64_t)b * gridDim.y + h) * 2;
  return sqrtf(n[0] * n[1]) * sc <= 50.f;
}

struct AtParams {
  int B;              // samples (the running-maximum kernel strides over them when it is only the fallback)
  int N, nblk;        // tokens per (batch, head); ceil(N / 128)
  float sc;           // softmax scale * log2(e)
  int out_ld;
  const float* norms;   // [2 + B*heads*2] fp32: global max|q|^2, max|k|^2, then the same per (batch, head); NULL = none
};

// Head dims other than 64 (template D = 32 / 96 / 128, the U-Net levels whose channels / heads is not 64): the same program
// with NB = ceil(D / 64) boxes of [128 tokens x 64 ch] per Q / K / V tile (a box always carries 64 channels -- for D = 32
// and 96 the upper half of the last box is the next head's data or TMA zero fill and is never multiplied in Q K^T: the
// product runs D / 16 k-steps), P V computed NB * 64 channels wide (the surplus columns are finite and never stored), the
// K / V rings two deep for NB = 2, and -- D > 64 only, where 2 x (128 S + 64 P + NB * 64 O) columns do not fit TMEM --
// P_t written over the first 64 columns of S_t: pass 2 reads 32 score columns and then stores 16 packed columns that lie at
// or below what it has read, and the MMA warp issues P_t V_j BEFORE Q_t K_{j+1}^T so the tensor pipe (in order) has consumed
// P_t when S_t is overwritten.
template <int D>
struct AtGeo {
  static constexpr int NB = (D + 63) / 64;                 // 64-channel boxes per tile
  static constexpr int STAGES = NB == 1 ? AT_KV_STAGES : 2;
  static constexpr int TILE = NB * AT_TILE_BYTES;
  static constexpr int PVN = NB * 64;                      // channels the P V product computes
  static constexpr bool ALIAS = D > 64;
  static constexpr int P_COL0 = ALIAS ? 0 : 256, P_STRIDE = ALIAS ? 128 : 64;
  static constexpr int O_COL0 = ALIAS ? 256 : 384, O_STRIDE = PVN;
  static constexpr size_t SMEM = 1024 + (size_t)(2 + 2 * STAGES) * TILE + 256 + 2048;
};

template <int D>
__global__ void __launch_bounds__(AT_THREADS, 1)
flash_attn_tc_kernel(const __grid_constant__ CUtensorMap mapQ, const __grid_constant__ CUtensorMap mapK,
                     const __grid_constant__ CUtensorMap mapV, const __grid_constant__ AtParams p,
                     __nv_bfloat16* __restrict__ out) {
  using G = AtGeo<D>;
  constexpr int NB = G::NB, STAGES = G::STAGES, TILE = G::TILE;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* q_s = smem;                                    // 2 tiles
  uint8_t* k_s = q_s + 2 * TILE;                          // ring
  uint8_t* v_s = k_s + STAGES * TILE;                     // ring
  uint64_t* bars = reinterpret_cast<uint64_t*>(v_s + STAGES * TILE);
  uint64_t* q_full = bars;                       // [1]
  uint64_t* k_full = bars + 1;                   // [stages]
  uint64_t* k_empty = k_full + STAGES;
  uint64_t* v_full = k_empty + STAGES;
  uint64_t* v_empty = v_full + STAGES;
  uint64_t* s_full = v_empty + STAGES;           // [2]  MMA -> softmax: S_t(j) complete
  uint64_t* p_full = s_full + 2;                 // [2]  softmax -> MMA: P_t(j) written, S_t and O_t free again
  uint64_t* o_full = p_full + 2;                 // [2]  MMA -> softmax: O_t(j) = P_t(j) V_j complete
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_full + 2);

  // A CTA serves samples blockIdx.z, blockIdx.z + gridDim.z, ...  Without a norm workspace gridDim.z == B (one sample).
  // With one this kernel is only the FALLBACK of the single-pass kernel -- for every (batch, head) that one can take, this
  // one has nothing to do -- and is launched with a few CTAs per (query tile, head) that stride over the samples: when
  // nothing falls back (the normal case) the launch costs 8 instead of B CTA start-ups per (tile, head) (1.1 % of the
  // DiT-B/4 step at B = 250), and when everything does, there are still >= 148 CTAs.
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  const int h = blockIdx.y;
  const int q0 = blockIdx.x * 256;
  {
    bool any = p.norms == nullptr;
    for (int bb = blockIdx.z; bb < p.B && !any; bb += gridDim.z) any = !at_fast_ok(p.norms, p.sc, bb, h);
    if (!any) return;
  }
  if (warp == 0 && lane == 0) {
    prefetch_tmap(&mapQ);
    prefetch_tmap(&mapK);
    prefetch_tmap(&mapV);
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  const uint32_t k_full0 = smem_u32(k_full), k_empty0 = smem_u32(k_empty), v_full0 = smem_u32(v_full),
                 v_empty0 = smem_u32(v_empty);
  const uint32_t s_full0 = smem_u32(s_full), p_full0 = smem_u32(p_full), o_full0 = smem_u32(o_full);
  bool first = true;
  uint32_t tmem_base = 0;
  for (int b = blockIdx.z; b < p.B; b += gridDim.z) {
  if (p.norms != nullptr && at_fast_ok(p.norms, p.sc, b, h)) continue;     // CTA-uniform
  if (warp == 0 && lane == 0) {
    constexpr int NBARS = 1 + 4 * STAGES + 6;
    if (!first)
      for (int i = 0; i < NBARS; ++i) asm volatile("mbarrier.inval.shared::cta.b64 [%0];" ::"r"(smem_u32(bars + i)) : "memory");
    mbar_init(smem_u32(q_full), 1);
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(smem_u32(&k_full[i]), 1);
      mbar_init(smem_u32(&k_empty[i]), 1);
      mbar_init(smem_u32(&v_full[i]), 1);
      mbar_init(smem_u32(&v_empty[i]), 1);
    }
    for (int t = 0; t < 2; ++t) {
      mbar_init(smem_u32(&s_full[t]), 1);
      mbar_init(smem_u32(&p_full[t]), 4);  // one arrive per softmax warp of the group
      mbar_init(smem_u32(&o_full[t]), 1);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  first = false;
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (elect_one()) {
      const uint32_t fq = smem_u32(q_full);
      mbar_expect_tx(fq, 2u * TILE);
#pragma unroll
      for (int x = 0; x < NB; ++x) {
        tma_load_3d(smem_u32(q_s + x * AT_TILE_BYTES), &mapQ, fq, h * D + x * 64, q0, b);
        tma_load_3d(smem_u32(q_s + TILE + x * AT_TILE_BYTES), &mapQ, fq, h * D + x * 64, q0 + 128, b);
      }
    }
    __syncwarp();
    int st = 0;
    uint32_t ph = 0;
    for (int j = 0; j < p.nblk; ++j) {
      mbar_wait(k_empty0 + st * 8, ph ^ 1);
      if (elect_one()) {
        mbar_expect_tx(k_full0 + st * 8, (uint32_t)TILE);
#pragma unroll
        for (int x = 0; x < NB; ++x)
          tma_load_3d(smem_u32(k_s + st * TILE + x * AT_TILE_BYTES), &mapK, k_full0 + st * 8, h * D + x * 64, j * 128, b);
      }
      __syncwarp();
      mbar_wait(v_empty0 + st * 8, ph ^ 1);
      if (elect_one()) {
        mbar_expect_tx(v_full0 + st * 8, (uint32_t)TILE);
#pragma unroll
        for (int x = 0; x < NB; ++x)
          tma_load_3d(smem_u32(v_s + st * TILE + x * AT_TILE_BYTES), &mapV, v_full0 + st * 8, h * D + x * 64, j * 128, b);
      }
      __syncwarp();
      if (++st == STAGES) { st = 0; ph ^= 1; }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    // instruction descriptors: D fp32, A/B bf16; QK^T: N=128, both K-major; PV: N = NB * 64, B (=V) MN-major
    const uint32_t idesc_qk = (1u << 4) | (1u << 7) | (1u << 10) | ((128u >> 3) << 17) | ((128u >> 4) << 24);
    const uint32_t idesc_pv =
        (1u << 4) | (1u << 7) | (1u << 10) | (1u << 16) | (((uint32_t)G::PVN >> 3) << 17) | ((128u >> 4) << 24);
    const uint32_t hi_k = (uint32_t)(make_kmajor_sw128_desc(0) >> 32);  // SBO 1024, version 1, SWIZZLE_128B
    const uint32_t hi_v = hi_k;                                         // MN-major: SBO = 1024 B between 8-token groups
    const uint32_t q_lo = ((smem_u32(q_s) & 0x3FFFFu) >> 4) | (1u << 16);
    const uint32_t k_lo0 = ((smem_u32(k_s) & 0x3FFFFu) >> 4) | (1u << 16);
    // V: leading byte offset = distance between the 64-channel atoms along N (the next box); unused for NB = 1
    const uint32_t v_lo0 = ((smem_u32(v_s) & 0x3FFFFu) >> 4) | ((NB == 1 ? 1u : (uint32_t)(AT_TILE_BYTES >> 4)) << 16);
    auto issue_qk = [&](int t, int st) {
      const uint32_t a = q_lo + (uint32_t)t * (TILE >> 4), bq = k_lo0 + (uint32_t)st * (TILE >> 4);
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < D / 16; ++k) {   // 16 channels per step: 32 B inside the 128-byte row of box k / 4
          const uint32_t off = (uint32_t)((k >> 2) * (AT_TILE_BYTES >> 4) + 2 * (k & 3));
          umma_f16_lohi(tmem_base + (uint32_t)(t * 128), a + off, bq + off, hi_k, idesc_qk, k ? 1u : 0u);
        }
        umma_commit(s_full0 + t * 8);
      }
      __syncwarp();
    };
    auto issue_pv = [&](int t, int st) {
      const uint32_t bv = v_lo0 + (uint32_t)st * (TILE >> 4);
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < 8; ++k)   // 16 keys per step: 8 TMEM columns of packed bf16 pairs, 2 x 1024 B of V rows
          umma_f16_ts(tmem_base + (uint32_t)(G::O_COL0 + t * G::O_STRIDE),
                      tmem_base + (uint32_t)(G::P_COL0 + t * G::P_STRIDE + k * 8), bv + (uint32_t)k * (2048 >> 4), hi_v,
                      idesc_pv, k ? 1u : 0u);
        umma_commit(o_full0 + t * 8);
      }
      __syncwarp();
    };
    mbar_wait(smem_u32(q_full), 0);
    mbar_wait(k_full0, 0);
    tc_fence_after();
    issue_qk(0, 0);
    issue_qk(1, 0);
    if (elect_one()) umma_commit(k_empty0);
    __syncwarp();
    int st = 0, stn = 1 % STAGES;
    uint32_t ph = 0, phn = (STAGES == 1) ? 1u : 0u;
    for (int j = 0; j < p.nblk; ++j) {
      const bool more = j + 1 < p.nblk;
      for (int t = 0; t < 2; ++t) {
        mbar_wait(p_full0 + t * 8, (uint32_t)(j & 1));
        tc_fence_after();
        if (G::ALIAS) {   // P_t lives in S_t: the product that reads it goes first
          if (t == 0) { mbar_wait(v_full0 + st * 8, ph); tc_fence_after(); }
          issue_pv(t, st);
        }
        if (more) {
          if (t == 0) { mbar_wait(k_full0 + stn * 8, phn); tc_fence_after(); }
          issue_qk(t, stn);
        }
        if (!G::ALIAS) {
          if (t == 0) { mbar_wait(v_full0 + st * 8, ph); tc_fence_after(); }
          issue_pv(t, st);
        }
      }
      if (elect_one()) {
        if (more) umma_commit(k_empty0 + stn * 8);
        umma_commit(v_empty0 + st * 8);
      }
      __syncwarp();
      st = stn; ph = phn;
      if (++stn == STAGES) { stn = 0; phn ^= 1; }
    }
  } else {
    // ===================== softmax groups =====================
    const int t = (warp - 2) >> 2;          // query tile of this group
    const int qd = warp & 3;                // TMEM lane quarter
    const int r = qd * 32 + lane;           // query row inside the tile == TMEM lane
    const uint32_t lane_addr = tmem_base + ((uint32_t)(qd * 32) << 16);
    const uint32_t s_addr = lane_addr + (uint32_t)(t * 128);
    const uint32_t p_addr = lane_addr + (uint32_t)(G::P_COL0 + t * G::P_STRIDE);
    const uint32_t o_addr = lane_addr + (uint32_t)(G::O_COL0 + t * G::O_STRIDE);
    float o[D];
#pragma unroll
    for (int i = 0; i < D; ++i) o[i] = 0.f;
    float m = -1e30f, l = 0.f, alpha_prev = 1.f;
    const float sc = p.sc;
    const uint64_t sc2 = f2_pack(sc, sc);
    // MASKED: only the last key block of a ragged N compares column indices (written as a branch inside the unrolled loop
    // the compiler turns it into a compare + select on every score of every block)
    auto key_block = [&](int j, auto mask_tag) {
      constexpr bool MASKED = decltype(mask_tag)::value;
      mbar_wait(s_full0 + t * 8, (uint32_t)(j & 1));
      tc_fence_after();
      const int valid = p.N - j * 128;       // keys of this block that exist (>= 128: all)
      // ---- pass 1: row max ----
      float mx = -1e30f;
#pragma unroll
      for (int c = 0; c < 128; c += 32) {
        uint32_t sv[32];
        tmem_ld32_nowait(s_addr + (uint32_t)c, sv);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i)
          if (!MASKED || c + i < valid) mx = fmaxf(mx, __uint_as_float(sv[i]));
      }
      const float m_new = fmaxf(m, mx);
      const float alpha = ex2f((m - m_new) * sc);
      const float msc = m_new * sc;
      const uint64_t nmsc2 = f2_pack(-msc, -msc);
      m = m_new;
      // ---- fold the previous block's P V into the register accumulators (its MMAs finished long ago) ----
      if (j > 0) {
        mbar_wait(o_full0 + t * 8, (uint32_t)((j - 1) & 1));
        tc_fence_after();
#pragma unroll
        for (int c = 0; c < D; c += 32) {
          uint32_t ov[32];
          tmem_ld32_nowait(o_addr + (uint32_t)c, ov);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) o[c + i] = fmaf(o[c + i], alpha_prev, __uint_as_float(ov[i]));
        }
      }
      alpha_prev = alpha;
      // ---- pass 2: P = exp2(S*c - m*c) -> bf16 pairs -> TMEM; exponents <= 0, so half of them may take the FMA-pipe
      //      form as in the single-pass kernel (clamped at -120: the polynomial has no flush to zero of its own) ----
      float r0 = 0.f, r1 = 0.f;
#pragma unroll
      for (int c = 0; c < 128; c += 32) {
        uint32_t sv[32];
        tmem_ld32_nowait(s_addr + (uint32_t)c, sv);
        tmem_ld_wait();
        uint32_t pk[16];
#pragma unroll
        for (int i = 0; i < 32; i += 4) {
          float x[4], e[4];
#pragma unroll
          for (int u = 0; u < 4; u += 2)
            f2_unpack(f2_fma(f2_pack(__uint_as_float(sv[i + u]), __uint_as_float(sv[i + u + 1])), sc2, nmsc2), x[u], x[u + 1]);
          e[0] = ex2f(x[0]);
          e[1] = ex2f(x[1]);
          // (not at D = 128: the 128 fp32 accumulators of O leave no registers for it -- measured 949 -> 922 TF/s with it,
          //  against 330 -> 392 at D = 32 and 788 -> 831 at D = 96)
          if (D <= 96 && (DCB_ATTN_POLY == 2 || (DCB_ATTN_POLY == 4 && (i & 4) == 0) || (DCB_ATTN_POLY == 8 && (i & 12) == 0))) {
            ex2_poly2(fmaxf(x[2], -120.f), fmaxf(x[3], -120.f), e[2], e[3]);
          } else {
            e[2] = ex2f(x[2]);
            e[3] = ex2f(x[3]);
          }
          if (MASKED) {
#pragma unroll
            for (int u = 0; u < 4; ++u)
              if (c + i + u >= valid) e[u] = 0.f;
          }
          add_f32x2(r0, r1, e[0], e[1]);
          add_f32x2(r0, r1, e[2], e[3]);
          pk[i >> 1] = pack_bf16x2(e[0], e[1]);
          pk[(i >> 1) + 1] = pack_bf16x2(e[2], e[3]);
        }
        tmem_st16(p_addr + (uint32_t)(c >> 1), pk);
      }
      l = fmaf(l, alpha, r0 + r1);
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(p_full0 + t * 8);
    };
    {
      const bool ragged = (p.N & 127) != 0;
      const int nfull = ragged ? p.nblk - 1 : p.nblk;
      for (int j = 0; j < nfull; ++j) key_block(j, std::false_type{});
      if (ragged) key_block(p.nblk - 1, std::true_type{});
    }
    // last block's product
    mbar_wait(o_full0 + t * 8, (uint32_t)((p.nblk - 1) & 1));
    tc_fence_after();
    const float inv = 1.f / l;
    const int qrow = q0 + t * 128 + r;
    __nv_bfloat16* orow = out + ((int64_t)b * p.N + qrow) * p.out_ld + h * D;
#pragma unroll
    for (int c = 0; c < D; c += 32) {
      uint32_t ov[32];
      tmem_ld32_nowait(o_addr + (uint32_t)c, ov);
      tmem_ld_wait();
      float v[32];
#pragma unroll
      for (int i = 0; i < 32; ++i) v[i] = fmaf(o[c + i], alpha_prev, __uint_as_float(ov[i])) * inv;
      if (qrow < p.N) {
#pragma unroll
        for (int i = 0; i < 32; i += 8) *reinterpret_cast<uint4*>(orow + c + i) = pack_bf16x8(v + i);
      }
    }
  }

  tc_fence_before();
  __syncthreads();      // every role is done with this sample: the barriers may be re-initialised
  }  // samples of this CTA
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
  }
}


// =====================================================================================================================
// Single-pass kernel.  Softmax needs a reference value only to keep exp() in range.  Here the range is known before the
// first key block: |s_ij| c <= |q_i| |k_j| c <= max|q| max|k| c (Cauchy-Schwarz; the two maxima per (batch, head) come from
// the epilogue of the projection that wrote q and k, or from attn_norms_rows_kernel), and the launch takes this kernel only
// where that bound is <= 50.  Then P' = 2^(s c) lies in [2^-50, 2^50], every partial sum fits fp32 / bf16 with room to
// spare, and softmax = P' / sum P' needs NO running maximum, no shift, no extra pass over S, no rescale and no per-block
// fold of O: one pass over S, and the tensor core accumulates O in TMEM over all key blocks.  Without a running maximum the two halves
// of a score row are independent, so each query row is shared by TWO threads (warps w and w+8: keys 0-63 / 64-127, O
// channels 0-31 / 32-63) with no exchange until the final row sum -- 16 softmax warps, four per scheduler, keep the MUFU
// pipe fed.  Exact as long as nothing underflows; the launch falls back to flash_attn_tc_kernel by itself otherwise
// (at_fast_ok, evaluated on the device per (batch, head): both kernels are always enqueued and, for every head, the CTAs of
// one of them return at once).
// =====================================================================================================================
constexpr int ATF_THREADS = 64 + 512;   // TMA warp, MMA warp, 2 tiles x 2 halves x 4 softmax warps

__global__ void __launch_bounds__(ATF_THREADS, 1)
flash_attn_tc_fast_kernel(const __grid_constant__ CUtensorMap mapQ, const __grid_constant__ CUtensorMap mapK,
                          const __grid_constant__ CUtensorMap mapV, const __grid_constant__ AtParams p,
                          __nv_bfloat16* __restrict__ out) {
  if (!at_fast_ok(p.norms, p.sc, blockIdx.z, blockIdx.y)) return;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* q_s = smem;                                    // 2 tiles
  uint8_t* k_s = q_s + 2 * AT_TILE_BYTES;                 // ring
  uint8_t* v_s = k_s + AT_KV_STAGES * AT_TILE_BYTES;      // ring
  uint64_t* bars = reinterpret_cast<uint64_t*>(v_s + AT_KV_STAGES * AT_TILE_BYTES);
  uint64_t* q_full = bars;
  uint64_t* k_full = bars + 1;
  uint64_t* k_empty = k_full + AT_KV_STAGES;
  uint64_t* v_full = k_empty + AT_KV_STAGES;
  uint64_t* v_empty = v_full + AT_KV_STAGES;
  uint64_t* s_full = v_empty + AT_KV_STAGES;     // [2]  MMA -> softmax: S_t(j) complete
  uint64_t* s_free = s_full + 2;                 // [2]  softmax -> MMA: S_t(j) is in registers
  uint64_t* p_full = s_free + 2;                 // [2]  softmax -> MMA: P_t(j) written
  uint64_t* o_full = p_full + 2;                 // [2]  MMA -> softmax: P_t(j) V_j accumulated (P_t free again)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_full + 2);
  float* lsum = reinterpret_cast<float*>(bars + 32);     // [2 tiles][2 halves][128 rows]

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  const int h = blockIdx.y, b = blockIdx.z;
  const int q0 = blockIdx.x * 256;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&mapQ);
    prefetch_tmap(&mapK);
    prefetch_tmap(&mapV);
    mbar_init(smem_u32(q_full), 1);
    for (int i = 0; i < AT_KV_STAGES; ++i) {
      mbar_init(smem_u32(&k_full[i]), 1);
      mbar_init(smem_u32(&k_empty[i]), 1);
      mbar_init(smem_u32(&v_full[i]), 1);
      mbar_init(smem_u32(&v_empty[i]), 1);
    }
    for (int t = 0; t < 2; ++t) {
      mbar_init(smem_u32(&s_full[t]), 1);
      mbar_init(smem_u32(&s_free[t]), 8);   // one arrive per softmax warp of the tile
      mbar_init(smem_u32(&p_full[t]), 8);
      mbar_init(smem_u32(&o_full[t]), 1);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t k_full0 = smem_u32(k_full), k_empty0 = smem_u32(k_empty), v_full0 = smem_u32(v_full),
                 v_empty0 = smem_u32(v_empty);
  const uint32_t s_full0 = smem_u32(s_full), s_free0 = smem_u32(s_free), p_full0 = smem_u32(p_full),
                 o_full0 = smem_u32(o_full);

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (elect_one()) {
      const uint32_t fq = smem_u32(q_full);
      mbar_expect_tx(fq, 2u * AT_TILE_BYTES);
      tma_load_3d(smem_u32(q_s), &mapQ, fq, h * 64, q0, b);
      tma_load_3d(smem_u32(q_s + AT_TILE_BYTES), &mapQ, fq, h * 64, q0 + 128, b);
    }
    __syncwarp();
    int st = 0;
    uint32_t ph = 0;
    for (int j = 0; j < p.nblk; ++j) {
      mbar_wait(k_empty0 + st * 8, ph ^ 1);
      if (elect_one()) {
        mbar_expect_tx(k_full0 + st * 8, (uint32_t)AT_TILE_BYTES);
        tma_load_3d(smem_u32(k_s + st * AT_TILE_BYTES), &mapK, k_full0 + st * 8, h * 64, j * 128, b);
      }
      __syncwarp();
      mbar_wait(v_empty0 + st * 8, ph ^ 1);
      if (elect_one()) {
        mbar_expect_tx(v_full0 + st * 8, (uint32_t)AT_TILE_BYTES);
        tma_load_3d(smem_u32(v_s + st * AT_TILE_BYTES), &mapV, v_full0 + st * 8, h * 64, j * 128, b);
      }
      __syncwarp();
      if (++st == AT_KV_STAGES) { st = 0; ph ^= 1; }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    const uint32_t idesc_qk = (1u << 4) | (1u << 7) | (1u << 10) | ((128u >> 3) << 17) | ((128u >> 4) << 24);
    const uint32_t idesc_pv = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 16) | ((64u >> 3) << 17) | ((128u >> 4) << 24);
    const uint32_t hi_k = (uint32_t)(make_kmajor_sw128_desc(0) >> 32);
    const uint32_t q_lo = ((smem_u32(q_s) & 0x3FFFFu) >> 4) | (1u << 16);
    const uint32_t k_lo0 = ((smem_u32(k_s) & 0x3FFFFu) >> 4) | (1u << 16);
    const uint32_t v_lo0 = ((smem_u32(v_s) & 0x3FFFFu) >> 4) | (1u << 16);
    auto issue_qk = [&](int t, int st_) {
      const uint32_t a = q_lo + (uint32_t)t * (AT_TILE_BYTES >> 4), bq = k_lo0 + (uint32_t)st_ * (AT_TILE_BYTES >> 4);
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_f16_lohi(tmem_base + (uint32_t)(t * 128), a + 2 * k, bq + 2 * k, hi_k, idesc_qk, k ? 1u : 0u);
        umma_commit(s_full0 + t * 8);
      }
      __syncwarp();
    };
    auto issue_pv = [&](int t, int st_, uint32_t acc0) {
      const uint32_t bv = v_lo0 + (uint32_t)st_ * (AT_TILE_BYTES >> 4);
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < 8; ++k)   // 16 keys per step: 8 TMEM columns of packed bf16 pairs, 2 x 1024 B of V rows
          umma_f16_ts(tmem_base + (uint32_t)(384 + t * 64), tmem_base + (uint32_t)(256 + t * 64 + k * 8),
                      bv + (uint32_t)k * (2048 >> 4), hi_k, idesc_pv, k ? 1u : acc0);
        umma_commit(o_full0 + t * 8);
      }
      __syncwarp();
    };
    mbar_wait(smem_u32(q_full), 0);
    mbar_wait(k_full0, 0);
    tc_fence_after();
    issue_qk(0, 0);
    issue_qk(1, 0);
    if (elect_one()) umma_commit(k_empty0);
    __syncwarp();
    int st = 0, stn = 1 % AT_KV_STAGES;
    uint32_t ph = 0, phn = (AT_KV_STAGES == 1) ? 1u : 0u;
    for (int j = 0; j < p.nblk; ++j) {
      if (j + 1 < p.nblk) {
        mbar_wait(k_full0 + stn * 8, phn);
        for (int t = 0; t < 2; ++t) {
          mbar_wait(s_free0 + t * 8, (uint32_t)(j & 1));   // S_t(j) is in the softmax warps' registers
          tc_fence_after();
          issue_qk(t, stn);
        }
        if (elect_one()) umma_commit(k_empty0 + stn * 8);
        __syncwarp();
      }
      mbar_wait(v_full0 + st * 8, ph);
      for (int t = 0; t < 2; ++t) {
        mbar_wait(p_full0 + t * 8, (uint32_t)(j & 1));
        tc_fence_after();
        issue_pv(t, st, j ? 1u : 0u);                      // O_t accumulates in TMEM over all key blocks
      }
      if (elect_one()) umma_commit(v_empty0 + st * 8);
      __syncwarp();
      st = stn; ph = phn;
      if (++stn == AT_KV_STAGES) { stn = 0; phn ^= 1; }
    }
  } else {
    // ===================== softmax warps =====================
    const int t = (warp - 2) >> 3;          // query tile
    const int hf = ((warp - 2) >> 2) & 1;   // keys [64 hf, +64) of every block, O channels [32 hf, +32)
    const int qd = warp & 3;                // TMEM lane quarter
    const int r = qd * 32 + lane;           // query row inside the tile == TMEM lane
    const uint32_t lane_addr = tmem_base + ((uint32_t)(qd * 32) << 16);
    const uint32_t s_addr = lane_addr + (uint32_t)(t * 128 + hf * 64);
    const uint32_t p_addr = lane_addr + (uint32_t)(256 + t * 64 + hf * 32);
    const uint32_t o_addr = lane_addr + (uint32_t)(384 + t * 64 + hf * 32);
    // No shift at all: at_fast_ok bounds every |s_ij| c by 50 (Cauchy-Schwarz on this head's max |q|, max |k|), so
    // P' = 2^(s c) lies in [2^-50, 2^50] -- inside the range of bf16 / fp32 with room for the 4096-term sums -- and
    // softmax = P' / sum P' needs no reference value: one instruction less per score than exp2(s c - M_i c).  When the
    // caller has folded c into the query projection (sc == 1: dcb200's DiT does, at pack time) the score IS the exponent.
    const float sc = p.sc;
    const bool prescaled = fabsf(sc - 1.0f) < 1e-6f;
    const uint64_t sc2 = f2_pack(sc, sc);
    float l0 = 0.f, l1 = 0.f;                  // even / odd columns: one packed add per pair of scores
    // PRE: the multiply is gone from the loop, not selected in it.  MASKED: only the last key block of a ragged N
    // compares column indices; everywhere else a score costs ex2 + half an add.f32x2 + half a cvt.bf16x2
    auto key_block = [&](int j, auto pre_tag, auto mask_tag) {
      constexpr bool PRE = decltype(pre_tag)::value;
      constexpr bool MASKED = decltype(mask_tag)::value;
      mbar_wait(s_full0 + t * 8, (uint32_t)(j & 1));
      tc_fence_after();
      const int valid = p.N - j * 128 - hf * 64;   // keys of this half block that exist (>= 64: all)
      uint32_t pk[32];
#pragma unroll
      for (int c = 0; c < 64; c += 32) {
        uint32_t sv[32];
        tmem_ld32_nowait(s_addr + (uint32_t)c, sv);
        tmem_ld_wait();
        if (c == 32) {   // this thread's half row of S is in registers
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(s_free0 + t * 8);
        }
#pragma unroll
        for (int i = 0; i < 32; i += 4) {
          float x[4], e[4];
#pragma unroll
          for (int u = 0; u < 4; u += 2) {   // |x| <= 50 (at_fast_ok)
            x[u] = __uint_as_float(sv[i + u]);
            x[u + 1] = __uint_as_float(sv[i + u + 1]);
            if (!PRE) f2_unpack(f2_mul(f2_pack(x[u], x[u + 1]), sc2), x[u], x[u + 1]);
          }
          e[0] = ex2f(x[0]);
          e[1] = ex2f(x[1]);
          const bool on_fma = DCB_ATTN_POLY == 2 || (DCB_ATTN_POLY == 4 && (i & 4) == 0) || (DCB_ATTN_POLY == 8 && (i & 12) == 0);
          if (on_fma) {
            ex2_poly2(x[2], x[3], e[2], e[3]);
          } else {
            e[2] = ex2f(x[2]);
            e[3] = ex2f(x[3]);
          }
          if (MASKED) {
#pragma unroll
            for (int u = 0; u < 4; ++u)
              if (c + i + u >= valid) e[u] = 0.f;
          }
          add_f32x2(l0, l1, e[0], e[1]);
          add_f32x2(l0, l1, e[2], e[3]);
          pk[(c + i) >> 1] = pack_bf16x2(e[0], e[1]);
          pk[((c + i) >> 1) + 1] = pack_bf16x2(e[2], e[3]);
        }
      }
      if (j > 0) {   // P_t is free once the previous block's P V has been consumed
        mbar_wait(o_full0 + t * 8, (uint32_t)((j - 1) & 1));
        tc_fence_after();
      }
      tmem_st16(p_addr, pk);
      tmem_st16(p_addr + 16u, pk + 16);
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(p_full0 + t * 8);
    };
    auto key_blocks = [&](auto pre_tag) {
      const bool ragged = (p.N & 127) != 0;
      const int nfull = ragged ? p.nblk - 1 : p.nblk;
      for (int j = 0; j < nfull; ++j) key_block(j, pre_tag, std::false_type{});
      if (ragged) key_block(p.nblk - 1, pre_tag, std::true_type{});
    };
    if (prescaled) key_blocks(std::true_type{});
    else key_blocks(std::false_type{});
    float l = l0 + l1;
    // the other half's share of the row sum, then O / l
    lsum[(t * 2 + hf) * 128 + r] = l;
    asm volatile("bar.sync %0, 256;" ::"r"(1 + t) : "memory");
    l += lsum[(t * 2 + (hf ^ 1)) * 128 + r];
    mbar_wait(o_full0 + t * 8, (uint32_t)((p.nblk - 1) & 1));
    tc_fence_after();
    const float inv = 1.f / l;
    const int qrow_i = q0 + t * 128 + r;
    __nv_bfloat16* orow = out + ((int64_t)b * p.N + qrow_i) * p.out_ld + h * 64 + hf * 32;
    {
      uint32_t ov[32];
      tmem_ld32_nowait(o_addr, ov);
      tmem_ld_wait();
      float v[32];
#pragma unroll
      for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(ov[i]) * inv;
      if (qrow_i < p.N) {
#pragma unroll
        for (int i = 0; i < 32; i += 8) *reinterpret_cast<uint4*>(orow + i) = pack_bf16x8(v + i);
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
  }
}

// norms[0..1]: max over everything; norms[2 + (b*heads + h)*2 + {0,1}]: max_i |q_i|^2, max_j |k_j|^2 of that (batch, head).
// bf16 values, fp32 sums; non-negative floats order like their bit patterns, so integer atomicMax does the reduction.
__global__ void __launch_bounds__(256) attn_norms_kernel(const __nv_bfloat16* __restrict__ q, const __nv_bfloat16* __restrict__ k,
                                                         int ld, int N, float* __restrict__ out) {
  // 8 threads per token (one 16-byte chunk of the 128-byte head row each): a warp reads 4 x 128 contiguous bytes
  const int h = blockIdx.y, b = blockIdx.z, c = threadIdx.x & 7;
  float qmax = 0.f, kmax = 0.f;
  for (int tkn = blockIdx.x * 256 + (threadIdx.x >> 3); tkn < min(N, (int)(blockIdx.x + 1) * 256); tkn += 32) {
    const int64_t off = ((int64_t)b * N + tkn) * ld + h * 64 + c * 8;
    float f[8], qn = 0.f, kn = 0.f;
    unpack_bf16x8(*reinterpret_cast<const uint4*>(q + off), f);
#pragma unroll
    for (int i = 0; i < 8; ++i) qn = fmaf(f[i], f[i], qn);
    unpack_bf16x8(*reinterpret_cast<const uint4*>(k + off), f);
#pragma unroll
    for (int i = 0; i < 8; ++i) kn = fmaf(f[i], f[i], kn);
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) {   // the 8 chunks of a token are 8 consecutive lanes
      qn += __shfl_xor_sync(0xffffffffu, qn, o);
      kn += __shfl_xor_sync(0xffffffffu, kn, o);
    }
    qmax = fmaxf(qmax, qn);
    kmax = fmaxf(kmax, kn);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    qmax = fmaxf(qmax, __shfl_xor_sync(0xffffffffu, qmax, o));
    kmax = fmaxf(kmax, __shfl_xor_sync(0xffffffffu, kmax, o));
  }
  if ((threadIdx.x & 31) == 0) {
    int* o2 = reinterpret_cast<int*>(out);
    const int bh = (b * gridDim.y + h) * 2;
    atomicMax(o2 + 2 + bh, __float_as_int(qmax));
    atomicMax(o2 + 3 + bh, __float_as_int(kmax));
    atomicMax(o2, __float_as_int(qmax));
    atomicMax(o2 + 1, __float_as_int(kmax));
  }
}

// Row-contiguous version (heads <= 16): a warp reads WHOLE token rows of q (blockIdx.y = 0) or k (= 1) -- heads x 128
// contiguous bytes, 16 bytes per lane and step -- instead of one 128-byte head slice out of every 4.6 KB row as above
// (measured on the DiT-B/4 step: 1.5 ms per launch = 2.1 TB/s there).  Eight consecutive lanes hold one head of a token.
__global__ void __launch_bounds__(256) attn_norms_rows_kernel(const __nv_bfloat16* __restrict__ q,
                                                              const __nv_bfloat16* __restrict__ k, int ld, int N, int heads,
                                                              float* __restrict__ out) {
  __shared__ float s_max[8][16];
  const int which = blockIdx.y, b = blockIdx.z;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const __nv_bfloat16* src = which ? k : q;
  const int chunks = heads * 8;                 // 16-byte chunks per token row
  float mx[4] = {0.f, 0.f, 0.f, 0.f};           // running maxima of the heads this lane group sees: chunk slot j -> head 4 j + lane / 8
  const int t_end = min(N, (int)(blockIdx.x + 1) * 64);
  for (int tkn = blockIdx.x * 64 + warp; tkn < t_end; tkn += 8) {
    const __nv_bfloat16* row = src + ((int64_t)b * N + tkn) * ld;
    uint4 raw[4];
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (lane + 32 * j < chunks) raw[j] = *reinterpret_cast<const uint4*>(row + (lane + 32 * j) * 8);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float n2 = 0.f;
      if (lane + 32 * j < chunks) {
        float f[8];
        unpack_bf16x8(raw[j], f);
#pragma unroll
        for (int i = 0; i < 8; ++i) n2 = fmaf(f[i], f[i], n2);
      }
#pragma unroll
      for (int o = 4; o > 0; o >>= 1) n2 += __shfl_xor_sync(0xffffffffu, n2, o);
      mx[j] = fmaxf(mx[j], n2);
    }
  }
  if ((lane & 7) == 0) {
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (4 * j + (lane >> 3) < heads) s_max[warp][4 * j + (lane >> 3)] = mx[j];
  }
  __syncthreads();
  if (threadIdx.x < heads) {
    float m = s_max[0][threadIdx.x];
#pragma unroll
    for (int w = 1; w < 8; ++w) m = fmaxf(m, s_max[w][threadIdx.x]);
    int* o2 = reinterpret_cast<int*>(out);
    atomicMax(o2 + 2 + (b * heads + threadIdx.x) * 2 + which, __float_as_int(m));
    atomicMax(o2 + which, __float_as_int(m));
  }
}

static int encode_tok_map(CUtensorMap* map, const void* base, int ld, int N, int B, int width) {
  auto enc = tc_encode_fn();
  DCB_REQUIRE(enc != nullptr, "cuTensorMapEncodeTiled entry point not available");
  cuuint64_t dims[3] = {(cuuint64_t)width, (cuuint64_t)N, (cuuint64_t)B};
  cuuint64_t strides[2] = {(cuuint64_t)ld * 2, (cuuint64_t)N * ld * 2};
  cuuint32_t box[3] = {64, 128, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  DCB_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(attention) failed: %d (ld=%d N=%d B=%d)", (int)r, ld, N, B);
  return DCB_OK;
}

// head dim D, bf16, N >= 128; q/k/v: [B, N, heads, D] views with row stride ld.
// norms_ws: [2 + B*heads*2] fp32 workspace or NULL.  With it (D = 64 and enough key blocks for the norm pre-pass to pay)
// the single-pass kernel is enqueued ahead of the running-maximum kernel and the device decides which of the two runs.
int launch_attn_norms(const void* q, const void* k, int ld, int N, int B, int heads, float* norms, cudaStream_t st) {
  if (heads <= 16)
    attn_norms_rows_kernel<<<dim3((N + 63) / 64, 2, B), 256, 0, st>>>((const __nv_bfloat16*)q, (const __nv_bfloat16*)k, ld, N,
                                                                       heads, norms);
  else
    attn_norms_kernel<<<dim3((N + 255) / 256, heads, B), 256, 0, st>>>((const __nv_bfloat16*)q, (const __nv_bfloat16*)k, ld, N,
                                                                        norms);
  DCB_CHECK_LAUNCH("attn_norms");
  return DCB_OK;
}

template <int D>
static int launch_flash_tc_d(const void* q, const void* k, const void* v, int ld, int B, int N, int heads, float scale,
                             void* out, int out_ld, float* norms_ws, bool norms_ready, cudaStream_t st) {
  DCB_REQUIRE(out_ld % 8 == 0 && ((uintptr_t)out & 15) == 0, "attention: out rows must be 16-byte aligned");
  CUtensorMap mq, mk, mv;
  int rc;
  if ((rc = encode_tok_map(&mq, q, ld, N, B, heads * D))) return rc;
  if ((rc = encode_tok_map(&mk, k, ld, N, B, heads * D))) return rc;
  if ((rc = encode_tok_map(&mv, v, ld, N, B, heads * D))) return rc;
  AtParams p;
  p.B = B;
  p.N = N;
  p.nblk = (N + 127) / 128;
  p.sc = scale * 1.4426950408889634f;
  p.out_ld = out_ld;
  const bool fast_on = !(knobs() & DCB_KNOB_ATTN_NO_FAST);
  const bool use_fast = D == 64 && norms_ws != nullptr && fast_on && N >= 1024;   // short sequences: two extra launches do not pay
  p.norms = use_fast ? norms_ws : nullptr;
  const size_t smem = AtGeo<D>::SMEM;
  static std::once_flag once;
  std::call_once(once, [] {
    cudaFuncSetAttribute(flash_attn_tc_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_LIMIT);
    if (D == 64) cudaFuncSetAttribute(flash_attn_tc_fast_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_LIMIT);
  });
  dim3 grid((N + 255) / 256, heads, B);
  if (use_fast) {
    if (!norms_ready) {   // otherwise the projection that wrote q and k left them (dcb_gemm_desc.attn_norms)
      cudaMemsetAsync(norms_ws, 0, sizeof(float) * (2 + 2 * B * heads), st);
      int rcn = launch_attn_norms(q, k, ld, N, B, heads, norms_ws, st);
      if (rcn) return rcn;
    }
    flash_attn_tc_fast_kernel<<<grid, ATF_THREADS, smem, st>>>(mq, mk, mv, p, (__nv_bfloat16*)out);
    DCB_CHECK_LAUNCH("flash_attn_tc_fast");
  }
  const dim3 grid_rm((N + 255) / 256, heads, use_fast ? (B < 8 ? B : 8) : B);   // as the fallback: strided over samples
  flash_attn_tc_kernel<D><<<grid_rm, AT_THREADS, smem, st>>>(mq, mk, mv, p, (__nv_bfloat16*)out);
  DCB_CHECK_LAUNCH("flash_attn_tc");
  return DCB_OK;
}

int launch_flash_tc(const void* q, const void* k, const void* v, int ld, int B, int N, int heads, int d, float scale,
                    void* out, int out_ld, float* norms_ws, bool norms_ready, cudaStream_t st) {
  switch (d) {
    case 32: return launch_flash_tc_d<32>(q, k, v, ld, B, N, heads, scale, out, out_ld, norms_ws, norms_ready, st);
    case 64: return launch_flash_tc_d<64>(q, k, v, ld, B, N, heads, scale, out, out_ld, norms_ws, norms_ready, st);
    case 96: return launch_flash_tc_d<96>(q, k, v, ld, B, N, heads, scale, out, out_ld, norms_ws, norms_ready, st);
    case 128: return launch_flash_tc_d<128>(q, k, v, ld, B, N, heads, scale, out, out_ld, norms_ws, norms_ready, st);
  }
  set_error("attention: head dim %d not in {32,64,96,128}", d);
  return DCB_EINVAL;
}

}  // namespace dcb
