// gemm_tc3.cu -- CTA-PAIR (tcgen05 cta_group::2) GEMM for the plain linear layers with N a multiple of 256: DiT QKV / attention
// out / MLP, the U-Net's attention projections.  D[M, N] = A[M, K] W[N, K]^T + the staged epilogue.
//
// Why: with one CTA per tile these K <= 3072 projections are bound by operand supply L2 -> SMEM (48 KB per 64-deep K block
// for a 256 x 128 tile, 94 B/clk/SM at full MMA rate).  A pair of CTAs on the two SMs of a TPC computes a
// 256 x 256 tile with ONE MMA stream (M = 256: each CTA's tensor core takes its own 128 rows of A from its own shared memory
// and ALL 256 columns of W -- the half it loaded itself and the half its peer loaded), so each CTA lands only
// [128 x 64] of A + [128 x 64] of W per K block: 32 KB per 128 x 256 outputs, a third less per FLOP, and five stages fit.
//
//   both CTAs   warp 0: TMA producer (own A rows, own half of W; cp.async.bulk.tensor ... cta_group::2 completes the
//               transaction on the LEADER's full barrier, the peer adds a remote arrive)
//   leader      warp 1: MMA issuer, tcgen05.mma.cta_group::2.kind::f16 M=256 N=256 K=16 into 2 x 256 TMEM columns (both
//               CTAs' tensor memory); tcgen05.commit ... multicast::cluster releases the stage / publishes the accumulator in
//               BOTH CTAs
//   both CTAs   warps 2..9: two epilogue groups, group g drains accumulator columns [128 g, 128 g + 128) of this CTA's 128
//               rows through a half-width staging tile (staged_epilogue_half) and releases the stage on the leader's barrier
#include <cudaTypedefs.h>

#include <mutex>

#include "tc_common.cuh"

namespace dcb {

constexpr int T3_THREADS = 64 + 512;    // TMA warp, MMA warp, up to four epilogue groups of four warps
constexpr int T3_MAX_STAGES = 8;
constexpr int T3_STAGE_BYTES = 2 * TC_A_BYTES;     // A [128 x 64] + W half [128 x 64]

struct T3Params {
  int M, nkb, m_pairs, n_tiles, total_tiles, stages, uniform, c_off;
  int geglu;        // GEGLU epilogue: group g owns output columns [64 g, + 64) of the tile's 128
  int ngroups;      // epilogue groups per CTA: 2 (128 columns each) or 4 (64 columns each: layers whose epilogue -- an
                    // activation over 128 x 256 outputs per CTA -- is longer than a K <= 1024 main loop)
  uint32_t idesc;
};

template <int NGROUPS>     // 2: 320 threads (the epilogue keeps its registers); 4: 576 threads
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(64 + 128 * NGROUPS, 1)
gemm_tc3_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB,
                const __grid_constant__ T3Params p, const __grid_constant__ EpiDev e) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // (dynamic shared memory starts at the same offset in both CTAs, so this rounding is identical in the pair)
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)p.stages * T3_STAGE_BYTES);
  uint64_t* full_bar = bars;                          // leader's is the one that counts: 2 arrivals + both CTAs' bytes
  uint64_t* empty_bar = bars + T3_MAX_STAGES;         // per CTA, 1 arrival (multicast commit)
  uint64_t* tfull_bar = bars + 2 * T3_MAX_STAGES;     // [2] per CTA, multicast commit
  uint64_t* tempty_bar = tfull_bar + 2;               // [2] leader's: 16 arrivals (8 epilogue warps of each CTA)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);
  uint8_t* stg8 = reinterpret_cast<uint8_t*>(bars) + 512;

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  const uint32_t rank = cluster_cta_rank();
  const int pair = blockIdx.x >> 1, n_pairs = gridDim.x >> 1;

  cluster_sync_all();
  if (warp == 0 && lane == 0) {
    prefetch_tmap(&mapA);
    prefetch_tmap(&mapB);
    for (int i = 0; i < p.stages; ++i) { mbar_init(smem_u32(&full_bar[i]), 2); mbar_init(smem_u32(&empty_bar[i]), 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(smem_u32(&tfull_bar[i]), 1); mbar_init(smem_u32(&tempty_bar[i]), 8 * p.ngroups); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {   // one warp of EACH CTA, same shared-memory slot: allocates the same columns in both tensor memories
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t smem_base = smem_u32(smem);
  const uint32_t full0 = smem_u32(full_bar), empty0 = smem_u32(empty_bar);

  if (warp == 0) {
    // ===================== TMA producer (both CTAs) =====================
    int stage = 0;
    uint32_t phase = 0;
    for (int tile = pair; tile < p.total_tiles; tile += n_pairs) {
      const int tn = tile % p.n_tiles, tm = tile / p.n_tiles;
      const int row0 = tm * 256 + (int)rank * 128, wrow0 = tn * 256 + (int)rank * 128;
      for (int kb = 0; kb < p.nkb; ++kb) {
        mbar_wait_long(empty0 + stage * 8, phase ^ 1);
        if (elect_one()) {
          const uint32_t fb = full0 + stage * 8;
          const uint32_t sa = smem_base + (uint32_t)(stage * T3_STAGE_BYTES);
          if (rank == 0) mbar_expect_tx(fb, 2u * T3_STAGE_BYTES);      // this CTA's bytes and the peer's
          else mbar_arrive_cluster(fb, 0);
          tma_load_2d_2sm(sa, &mapA, fb, p.c_off + kb * TC_BK, row0);
          tma_load_2d_2sm(sa + TC_A_BYTES, &mapB, fb, kb * TC_BK, wrow0);
        }
        __syncwarp();
        if (++stage == p.stages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    if (rank == 0) {
      // ===================== MMA issuer (leader CTA only) =====================
      int stage = 0, as = 0;
      uint32_t phase = 0, aphase = 0;
      const uint32_t desc_hi = (uint32_t)(make_kmajor_sw128_desc(0) >> 32);
      const uint32_t lo_base = ((smem_base & 0x3FFFFu) >> 4) | (1u << 16), lo_step = (uint32_t)T3_STAGE_BYTES >> 4;
      uint32_t alo = lo_base, fb = full0, eb = empty0;
      for (int tile = pair; tile < p.total_tiles; tile += n_pairs) {
        mbar_wait_long(smem_u32(&tempty_bar[as]), aphase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(as * 256);
        for (int kb = 0; kb < p.nkb; ++kb) {
          mbar_wait(fb, phase);
          tc_fence_after();
          const uint32_t blo = alo + (TC_A_BYTES >> 4);
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < TC_BK / 16; ++k)
              umma_f16_2sm(d_tmem, alo + 2 * k, blo + 2 * k, desc_hi, p.idesc, (kb | k) ? 1u : 0u);
            umma_commit_2sm(eb);
            if (kb == p.nkb - 1) umma_commit_2sm(smem_u32(&tfull_bar[as]));
          }
          __syncwarp();
          alo += lo_step; fb += 8; eb += 8;
          if (++stage == p.stages) { stage = 0; phase ^= 1; alo = lo_base; fb = full0; eb = empty0; }
        }
        if (++as == 2) { as = 0; aphase ^= 1; }
      }
    }
  } else if (((warp - 2) >> 2) < p.ngroups) {
    // ===================== epilogue (both CTAs): group g drains columns [w g, + w) of this CTA's 128 rows, w = 256 / ngroups ==========
    const int q = warp & 3;
    const int grp = (warp - 2) >> 2;
    const int gw = p.geglu ? 64 : 256 / p.ngroups;
    const int tiles_x = (p.M + 127) / 128;
    EpiGeom gq{tiles_x, 1, 128, 1, 1, p.M, 1, 1, p.uniform};
    int as = 0;
    uint32_t aphase = 0;
    for (int tile = pair, it = 0; tile < p.total_tiles; tile += n_pairs, ++it) {
      const int tn = tile % p.n_tiles, tm = tile / p.n_tiles;
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * 256 + grp * gw);
      staged_epilogue_half(gq, e, stg8 + grp * TC_EPI_HALF_BYTES, it & 1, tm * 2 + (int)rank, tn * p.ngroups + grp, gw, taddr,
                           smem_u32(&tfull_bar[as]), aphase, smem_u32(&tempty_bar[as]), 1 + grp, rank == 0 ? -1 : 0);
      if (++as == 2) { as = 0; aphase ^= 1; }
    }
  }

  tc_fence_before();
  cluster_sync_all();        // no CTA may leave while its peer can still arrive on its barriers or read its shared memory
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
  }
}

// ---- host side ----------------------------------------------------------------------------------------------
PFN_cuTensorMapEncodeTiled_v12000 tc_encode_fn();

static int encode_2d(CUtensorMap* map, const void* base, int64_t rows, int64_t cols, int64_t ld, int box_rows) {
  auto enc = tc_encode_fn();
  DCB_REQUIRE(enc != nullptr, "cuTensorMapEncodeTiled entry point not available");
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {(cuuint32_t)TC_BK, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  DCB_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(tc3) failed: %d", (int)r);
  return DCB_OK;
}

// returns DCB_EUNSUPPORTED when the descriptor is not a plain linear layer this kernel covers
int launch_gemm_tc3(const GemmDev& g, cudaStream_t st, int uniform) {
  const EpiDev& e = g.epi;
  const SegDev& s = g.seg[0];
  if (g.nseg != 1 || g.OH != 1 || g.NB != 1 || s.H != 1 || s.stride != 1 || s.dx != 0 || s.dy != 0 || s.nb_div > 1 ||
      e.N % 256 != 0 || g.K % TC_BK != 0 || e.gn_part != nullptr || e.mse_part != nullptr || e.up_phase != 0 || s.W != g.OW)
    return DCB_EUNSUPPORTED;
  // GEGLU: [128 value | 128 gate] weight rows per tile, two epilogue groups of 64 OUTPUT columns; plain form only
  const bool geglu = e.act == DCB_ACT_GEGLU;
  if (geglu && (e.rowvec != nullptr || e.gate != nullptr || e.residual != nullptr || e.act_post != DCB_ACT_NONE ||
                e.attn_norms != nullptr))
    return DCB_EUNSUPPORTED;
  const int M = g.OW;
  T3Params p;
  memset(&p, 0, sizeof(p));
  p.M = M;
  p.nkb = g.K / TC_BK;
  p.c_off = s.c_off;
  p.m_pairs = (M + 255) / 256;
  p.n_tiles = e.N / 256;
  p.total_tiles = p.m_pairs * p.n_tiles;
  p.uniform = uniform;
  const int pairs = num_sms() / 2;
  if (p.total_tiles < 2 * pairs) return DCB_EUNSUPPORTED;     // too small to fill the pairs twice: the one-CTA kernels do better
  CUtensorMap mapA, mapB;
  int rc;
  if ((rc = encode_2d(&mapA, s.src, M, s.C, s.C, 128))) return rc;
  if ((rc = encode_2d(&mapB, g.W, e.N, g.K, g.K, 128))) return rc;
#ifndef DCB_TC3_RULE
#define DCB_TC3_RULE 2
#endif
  // four 64-column epilogue groups (each issues its residual prefetch BEFORE it waits for the accumulator, and none has a
  // second half whose prefetch would be exposed) when the main loop is short and the epilogue is not a plain store
  const bool heavy = e.act != DCB_ACT_NONE || (DCB_TC3_RULE >= 1 && (e.residual != nullptr || e.gate != nullptr)) ||
                     (DCB_TC3_RULE >= 2 && e.attn_norms != nullptr);
  p.ngroups = (heavy && g.K <= 1024 && !geglu) ? 4 : 2;
  p.geglu = geglu ? 1 : 0;
  int stages = (TC_SMEM_LIMIT - 1024 - 512 - p.ngroups * TC_EPI_HALF_BYTES) / T3_STAGE_BYTES;
  if (stages > T3_MAX_STAGES) stages = T3_MAX_STAGES;
  if (stages > p.nkb) stages = p.nkb < 2 ? 2 : p.nkb;
  p.stages = stages;
  // instruction descriptor: D = f32, A = B = bf16, both K-major, N = 256, M = 256 (the pair)
  p.idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((256u >> 3) << 17) | ((256u >> 4) << 24);
  const size_t smem = (size_t)stages * T3_STAGE_BYTES + 1024 + 512 + p.ngroups * TC_EPI_HALF_BYTES;
  static std::once_flag attr_once;
  std::call_once(attr_once, [] {
    cudaFuncSetAttribute(gemm_tc3_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_LIMIT);
    cudaFuncSetAttribute(gemm_tc3_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_LIMIT);
  });
  const int grid = 2 * (p.total_tiles < pairs ? p.total_tiles : pairs);
  if (p.ngroups == 4) gemm_tc3_kernel<4><<<grid, 64 + 128 * 4, smem, st>>>(mapA, mapB, p, e);
  else gemm_tc3_kernel<2><<<grid, 64 + 128 * 2, smem, st>>>(mapA, mapB, p, e);
  DCB_CHECK_LAUNCH("gemm_tc3");
  if (g.epi.attn_norms != nullptr) g_attn_norms_written = true;   // staged_epilogue_half filled them
  return DCB_OK;
}

}  // namespace dcb
