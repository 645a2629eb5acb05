// attention_tc.cu -- tcgen05 / TMEM flash attention (forward, no mask) for head dim 64: softmax(Q K^T * scale) V per
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
//   warps 2-9 / 10-17 softmax group of tile 0 / 1: one TMEM lane = one query row, shared by TWO threads (warps w, w+4:
//                 keys 0-63 / 64-127 of the block, channels 0-31 / 32-63 of O) so that 4 warps per scheduler overlap
//                 their MUFU, FMA and tcgen05.ld phases; the halves exchange their block maximum through smem:
//                 one pass over S per key block: P = exp2((S - R) c) against the lagged running maximum R -> bf16 pairs ->
//                 tcgen05.st into TMEM; the PV product of the previous block is folded into fp32 register accumulators
//                 while the tensor pipe already runs the next QK^T, so nothing is ever rescaled in TMEM and the N x N
//                 scores never exist in HBM.
// TMEM columns: S0 [0,128) S1 [128,256) P0 [256,320) P1 [320,384) O0 [384,448) O1 [448,512).
#include <cudaTypedefs.h>

#include <mutex>

#include "tc_common.cuh"

namespace dcb {

constexpr int AT_THREADS = 576;  // TMA warp, MMA warp, 2 query tiles x 2 column halves x 4 softmax warps
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
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t* r) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(taddr), "r"(r[0]), "r"(r[1]),
               "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ float ex2f(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

struct AtParams {
  int N, nblk;        // tokens per (batch, head); ceil(N / 128)
  float sc;           // softmax scale * log2(e)
  int out_ld;
};

__global__ void __launch_bounds__(AT_THREADS, 1)
flash_attn_tc_kernel(const __grid_constant__ CUtensorMap mapQ, const __grid_constant__ CUtensorMap mapK,
                     const __grid_constant__ CUtensorMap mapV, const __grid_constant__ AtParams p,
                     __nv_bfloat16* __restrict__ out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* q_s = smem;                                    // 2 tiles
  uint8_t* k_s = q_s + 2 * AT_TILE_BYTES;                 // ring
  uint8_t* v_s = k_s + AT_KV_STAGES * AT_TILE_BYTES;      // ring
  uint64_t* bars = reinterpret_cast<uint64_t*>(v_s + AT_KV_STAGES * AT_TILE_BYTES);
  uint64_t* q_full = bars;                       // [1]
  uint64_t* k_full = bars + 1;                   // [stages]
  uint64_t* k_empty = k_full + AT_KV_STAGES;
  uint64_t* v_full = k_empty + AT_KV_STAGES;
  uint64_t* v_empty = v_full + AT_KV_STAGES;
  uint64_t* s_full = v_empty + AT_KV_STAGES;     // [2]  MMA -> softmax: S_t(j) complete
  uint64_t* p_full = s_full + 2;                 // [2]  softmax -> MMA: P_t(j) written, S_t and O_t free again
  uint64_t* o_full = p_full + 2;                 // [2]  MMA -> softmax: O_t(j) = P_t(j) V_j complete
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_full + 2);
  float* xch = reinterpret_cast<float*>(bars + 32);   // [2 parities + 1][2 tiles][2 halves][128 rows] block maxima / row sums

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
      mbar_init(smem_u32(&p_full[t]), 8);  // one arrive per softmax warp of the group
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
  const uint32_t s_full0 = smem_u32(s_full), p_full0 = smem_u32(p_full), o_full0 = smem_u32(o_full);

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
    // instruction descriptors: D fp32, A/B bf16; QK^T: N=128, both K-major; PV: N=64, B (=V) MN-major
    const uint32_t idesc_qk = (1u << 4) | (1u << 7) | (1u << 10) | ((128u >> 3) << 17) | ((128u >> 4) << 24);
    const uint32_t idesc_pv = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 16) | ((64u >> 3) << 17) | ((128u >> 4) << 24);
    const uint32_t hi_k = (uint32_t)(make_kmajor_sw128_desc(0) >> 32);  // SBO 1024, version 1, SWIZZLE_128B
    const uint32_t hi_v = hi_k;                                         // MN-major: SBO = 1024 B between 8-token groups
    const uint32_t q_lo = ((smem_u32(q_s) & 0x3FFFFu) >> 4) | (1u << 16);
    const uint32_t k_lo0 = ((smem_u32(k_s) & 0x3FFFFu) >> 4) | (1u << 16);
    const uint32_t v_lo0 = ((smem_u32(v_s) & 0x3FFFFu) >> 4) | (1u << 16);
    auto issue_qk = [&](int t, int st) {
      const uint32_t a = q_lo + (uint32_t)t * (AT_TILE_BYTES >> 4), bq = k_lo0 + (uint32_t)st * (AT_TILE_BYTES >> 4);
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_f16_lohi(tmem_base + (uint32_t)(t * 128), a + 2 * k, bq + 2 * k, hi_k, idesc_qk, k ? 1u : 0u);
        umma_commit(s_full0 + t * 8);
      }
      __syncwarp();
    };
    auto issue_pv = [&](int t, int st) {
      const uint32_t bv = v_lo0 + (uint32_t)st * (AT_TILE_BYTES >> 4);
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < 8; ++k)   // 16 keys per step: 8 TMEM columns of packed bf16 pairs, 2 x 1024 B of V rows
          umma_f16_ts(tmem_base + (uint32_t)(384 + t * 64), tmem_base + (uint32_t)(256 + t * 64 + k * 8),
                      bv + (uint32_t)k * (2048 >> 4), hi_v, idesc_pv, k ? 1u : 0u);
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
      const bool more = j + 1 < p.nblk;
      for (int t = 0; t < 2; ++t) {
        mbar_wait(p_full0 + t * 8, (uint32_t)(j & 1));
        tc_fence_after();
        if (more) {
          if (t == 0) { mbar_wait(k_full0 + stn * 8, phn); tc_fence_after(); }
          issue_qk(t, stn);
        }
        if (t == 0) { mbar_wait(v_full0 + st * 8, ph); tc_fence_after(); }
        issue_pv(t, st);
      }
      if (elect_one()) {
        if (more) umma_commit(k_empty0 + stn * 8);
        umma_commit(v_empty0 + st * 8);
      }
      __syncwarp();
      st = stn; ph = phn;
      if (++stn == AT_KV_STAGES) { stn = 0; phn ^= 1; }
    }
  } else {
    // ===================== softmax groups =====================
    const int t = (warp - 2) >> 3;          // query tile of this group
    const int hf = ((warp - 2) >> 2) & 1;   // column half: keys [64 hf, 64 hf + 64) of a block, channels [32 hf, 32 hf + 32)
    const int qd = warp & 3;                // TMEM lane quarter
    const int r = qd * 32 + lane;           // query row inside the tile == TMEM lane
    const int bar_id = 1 + t;               // named barrier of the group (256 threads)
    const uint32_t lane_addr = tmem_base + ((uint32_t)(qd * 32) << 16);
    const uint32_t s_addr = lane_addr + (uint32_t)(t * 128 + hf * 64);
    const uint32_t p_addr = lane_addr + (uint32_t)(256 + t * 64 + hf * 32);
    const uint32_t o_addr = lane_addr + (uint32_t)(384 + t * 64 + hf * 32);
    float* my_x = xch + (t * 2 + hf) * 128 + r;            // this thread's exchange slot
    const float* other_x = xch + (t * 2 + (hf ^ 1)) * 128 + r;
    float o[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) o[i] = 0.f;
    // Single pass per key block with a LAGGED reference maximum R (the running max of the previous blocks):
    //   P_j = exp2((S_j - R_j) c),  O_j = P_j V_j,  then  R_{j+1} = max(R_j, max S_j),  beta_j = exp2((R_j - R_{j+1}) c)
    //   l <- (l + rowsum P_j) beta_j      o <- (o + O_{j-1}) beta_{j-1}   (O is folded one block late, see the header)
    // which is the online softmax with every quantity of block j expressed relative to R_j instead of R_{j+1}; P may
    // exceed 1 by 2^(max S_j - R_j) -- harmless in bf16/fp32 below 2^40, and if a row would exceed that (always the case
    // for the first block, where R = -inf) both halves first move R up to the block maximum and redo the block.
    float R = -1e30f, l = 0.f, beta_prev = 1.f;
    const float sc = p.sc;
    for (int j = 0; j < p.nblk; ++j) {
      mbar_wait(s_full0 + t * 8, (uint32_t)(j & 1));
      tc_fence_after();
      const int valid = p.N - j * 128 - hf * 64;   // keys of this half block that exist (>= 64: all)
      uint32_t sv[16];
      // ---- fold the previous block's P V into the register accumulators; also frees P_t and O_t for this block ----
      if (j > 0) {
        mbar_wait(o_full0 + t * 8, (uint32_t)((j - 1) & 1));
        tc_fence_after();
#pragma unroll
        for (int c = 0; c < 32; c += 16) {
          tmem_ld16_nowait(o_addr + (uint32_t)c, sv);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 16; ++i) o[c + i] = (o[c + i] + __uint_as_float(sv[i])) * beta_prev;
        }
      }
      float rs = 0.f, mx = -1e30f;
      auto pass = [&](float rsc) {
        rs = 0.f;
        mx = -1e30f;
#pragma unroll
        for (int c = 0; c < 64; c += 16) {
          tmem_ld16_nowait(s_addr + (uint32_t)c, sv);
          tmem_ld_wait();
          uint32_t pk[8];
#pragma unroll
          for (int i = 0; i < 16; i += 2) {
            const float s0 = __uint_as_float(sv[i]), s1 = __uint_as_float(sv[i + 1]);
            float p0 = ex2f(fmaf(s0, sc, -rsc));
            float p1 = ex2f(fmaf(s1, sc, -rsc));
            if (valid >= 64) {
              mx = fmaxf(mx, fmaxf(s0, s1));
            } else {
              if (c + i < valid) mx = fmaxf(mx, s0); else p0 = 0.f;
              if (c + i + 1 < valid) mx = fmaxf(mx, s1); else p1 = 0.f;
            }
            rs += p0 + p1;
            pk[i >> 1] = pack_bf16x2(p0, p1);
          }
          tmem_st8(p_addr + (uint32_t)(c >> 1), pk);
        }
      };
      pass(R * sc);
      // the two halves of a row agree on the block maximum (and therefore on R and on the redo decision)
      my_x[(j & 1) * 512] = mx;
      asm volatile("bar.sync %0, 256;" ::"r"(bar_id) : "memory");
      float bm = fmaxf(mx, other_x[(j & 1) * 512]);
      if (__any_sync(0xffffffffu, (bm - R) * sc > 40.f)) {
        const float Rn = fmaxf(R, bm);
        const float g = ex2f((R - Rn) * sc);
        l *= g;
#pragma unroll
        for (int i = 0; i < 32; ++i) o[i] *= g;
        R = Rn;
        tmem_st_wait();
        pass(R * sc);
      }
      const float Rn = fmaxf(R, bm);
      const float beta = ex2f((R - Rn) * sc);
      l = (l + rs) * beta;
      beta_prev = beta;
      R = Rn;
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(p_full0 + t * 8);
    }
    // last block's product, and the other half's share of the row sum
    my_x[1024] = l;
    asm volatile("bar.sync %0, 256;" ::"r"(bar_id) : "memory");
    l += other_x[1024];
    mbar_wait(o_full0 + t * 8, (uint32_t)((p.nblk - 1) & 1));
    tc_fence_after();
    const float inv = 1.f / l;
    const int qrow = q0 + t * 128 + r;
    __nv_bfloat16* orow = out + ((int64_t)b * p.N + qrow) * p.out_ld + h * 64 + hf * 32;
#pragma unroll
    for (int c = 0; c < 32; c += 16) {
      uint32_t ov[16];
      tmem_ld16_nowait(o_addr + (uint32_t)c, ov);
      tmem_ld_wait();
      float v[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) v[i] = (o[c + i] + __uint_as_float(ov[i])) * beta_prev * inv;
      if (qrow < p.N) {
        *reinterpret_cast<uint4*>(orow + c) = pack_bf16x8(v);
        *reinterpret_cast<uint4*>(orow + c + 8) = pack_bf16x8(v + 8);
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

// head dim 64, bf16, N >= 128; q/k/v: [B, N, heads, 64] views with row stride ld
int launch_flash_tc(const void* q, const void* k, const void* v, int ld, int B, int N, int heads, float scale, void* out,
                    int out_ld, cudaStream_t st) {
  DCB_REQUIRE(out_ld % 8 == 0 && ((uintptr_t)out & 15) == 0, "attention: out rows must be 16-byte aligned");
  CUtensorMap mq, mk, mv;
  int rc;
  if ((rc = encode_tok_map(&mq, q, ld, N, B, heads * 64))) return rc;
  if ((rc = encode_tok_map(&mk, k, ld, N, B, heads * 64))) return rc;
  if ((rc = encode_tok_map(&mv, v, ld, N, B, heads * 64))) return rc;
  AtParams p;
  p.N = N;
  p.nblk = (N + 127) / 128;
  p.sc = scale * 1.4426950408889634f;
  p.out_ld = out_ld;
  const size_t smem = 1024 + (2 + 2 * AT_KV_STAGES) * AT_TILE_BYTES + 256 + 3 * 512 * 4;
  static std::once_flag once;
  std::call_once(once, [] {
    cudaFuncSetAttribute(flash_attn_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_LIMIT);
  });
  dim3 grid((N + 255) / 256, heads, B);
  flash_attn_tc_kernel<<<grid, AT_THREADS, smem, st>>>(mq, mk, mv, p, (__nv_bfloat16*)out);
  DCB_CHECK_LAUNCH("flash_attn_tc");
  return DCB_OK;
}

}  // namespace dcb
