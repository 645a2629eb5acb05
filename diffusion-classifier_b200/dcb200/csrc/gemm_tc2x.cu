// gemm_tc2x.cu -- stride-1 3x3 convolution over rows of >= 128 pixels with the preceding GroupNorm(+SiLU) applied to the
// operand ON THE FLY (dcb_gemm_desc.xf_a): the normalised tensor -- one HBM write + one HBM read of every full-resolution
// activation, the separate gn_apply pass -- never exists.
//
// A CTA tile is TWO VERTICALLY ADJACENT output rows of 128 pixels (two TMEM accumulators).  Per 64-channel block the four
// input rows y0-1 .. y0+2 they need land as four [130 px x 64 ch] "row boxes" (TMA, 128B swizzle, zero fill = padding) in
// a ring; eight transform warps rewrite each box ONCE, in place -- y = a[n,c] x + b[n,c], SiLU: the arithmetic and
// rounding of gn_apply_kernel, so the result is bit-identical to the unfused path -- and all nine taps of both output rows
// are served from them: output row s, tap (ky, kx) reads box s + ky starting kx smem rows (128 B) in.  K blocks are
// therefore issued (channel block, ky, kx); every other tcgen05 path uses the same order for these convs
// (conv9_kb_outer in gemm_tc.cu).  Compared with per-(ky) halo boxes (gemm_tc2's x-halo mode) the activation traffic
// L2 -> SMEM and the transform work drop by a third (4 instead of 6 boxes per channel block).
//
//   warp 0        TMA producer of the activations (row boxes, plain tap tiles of trailing 1x1 segments)
//   warp 19       TMA producer of the weight blocks -- its own warp: with one in-order producer the weight ring's depth
//                 (5 blocks = 2.5 k cycles) also capped how far ahead of the MMAs a row box could be requested, and a box
//                 needs TMA latency + the transform before it is usable (measured: 737 -> see DESIGN.md)
//   warps 1, 2    MMA issuers, one per output row (accumulator); tcgen05.commit releases boxes / weight blocks
//   warps 3..10   two epilogue groups, each with a HALF-width staging tile (staged_epilogue_half: 64 columns at a time; two
//                 full tiles would cost two row boxes, one shared tile serialises the groups: +0.5 ms on a K = 1152 conv,
//                 profiles/r02_gn_fusion.md), or the direct eps-MSE epilogue of conv_out
//   warps 11..18  transform: thread t owns logical 16-byte chunk t & 7 (8 channels: 16 coefficients in registers for the
//                 whole channel block) of box rows t >> 3, + 32, ...
#include <cudaTypedefs.h>

#include <mutex>

#include "tc_common.cuh"

namespace dcb {

constexpr int TX_THREADS = 640;
constexpr int TX_MAX_SLOTS = 12;
constexpr int TX_BAR_BYTES = 1024;
constexpr int TX_BOX = 17 * 1024;      // 130 rows x 128 B = 16640 B, padded to the 1024-B swizzle repeat
constexpr int TX_XF_WARPS = 8;

struct TxParams {
  int ntap;                    // plain tap segments after the conv (1x1 shortcut over raw tensors), stride 1
  TcSeg tap[DCB_MAX_SEGS];
  int nkb_conv, nkb0;          // channel blocks of the (concatenated) conv input; blocks >= nkb0 come from source 1 (map 1)
  int div0, div1;              // sample divisors of the two raw sources
  int tiles_x, OH, OW, NB, BN, n_tiles, total_tiles, m_tiles;
  int num_P;                   // row-pair tiles (m_tiles / 2); total_tiles counts the launch's work items: row-pair tiles x N tiles,
                               // or (PAIR) CTA-pair items = ceil(num_P / 2) x N tiles, two row-pair tiles each
  int nbox, b_slots;
  uint32_t idesc;
  int uniform, staged, silu, xf_C;
  int dbg;                     // -DDCB_PROBES builds only: 1 = no transform arithmetic, 2 = no epilogue work
  const float* xf_a;           // [NB][xf_C]
  const float* xf_b;
};

struct TxPair {
  int x0, y0, nb, tm0;         // tile origin (pixel, first output row), sample, linear 128-row tile index of output row y0
};
__device__ __forceinline__ TxPair decode_pair(const TxParams& p, int P) {
  TxPair t;
  if (P >= p.num_P) {          // the odd tail of a CTA pair: boxes land as zeros (sample out of range), every row is masked
    t.x0 = 0; t.y0 = 0; t.nb = p.NB; t.tm0 = p.m_tiles;
    return t;
  }
  const int tx = P % p.tiles_x;
  P /= p.tiles_x;
  const int hp = p.OH >> 1;
  t.nb = P / hp;
  t.y0 = 2 * (P - t.nb * hp);
  t.x0 = tx * 128;
  t.tm0 = (t.nb * p.OH + t.y0) * p.tiles_x + tx;     // row-major tile numbering of gemm_tc / gemm_tc2 (gn_part, mse_part)
  return t;
}

// PAIR: two CTAs on the SMs of one TPC (cluster of 2) run their row-pair tiles in lockstep and share every weight block:
// each lands only HALF of it (N / 2 rows), the leader's two MMA warps issue tcgen05.mma.cta_group::2 with M = 256 -- output
// row s of BOTH CTAs at once, all N columns of W, half from each CTA's shared memory -- and its commits are multicast to both
// CTAs' barriers; transform / epilogue warps of the peer signal the leader's barriers (mapa + remote arrive).  Weight traffic
// L2 -> SMEM and the weight ring halve (4 slots = 32 KB instead of 64 KB: two more row boxes).
template <bool PAIR>
__global__ void __launch_bounds__(TX_THREADS, 1)
gemm_tc2x_kernel(const __grid_constant__ CUtensorMap mapA0, const __grid_constant__ CUtensorMap mapA1,
                 const __grid_constant__ CUtensorMap mapA2, const __grid_constant__ CUtensorMap mapB,
                 const __grid_constant__ TxParams p, const __grid_constant__ EpiDev e) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int b_rows = PAIR ? p.BN / 2 : p.BN;       // weight rows this CTA lands per block
  const int b_bytes = b_rows * TC_BK * 2;
  uint8_t* a_ring = smem;
  uint8_t* b_ring = a_ring + (size_t)p.nbox * TX_BOX;
  uint64_t* bars = reinterpret_cast<uint64_t*>(b_ring + (size_t)p.b_slots * b_bytes);
  uint64_t* a_full = bars;                          // TMA landed            (count 1 + tx bytes)
  uint64_t* a_ready = bars + TX_MAX_SLOTS;          // transformed           (count 8: one per transform warp)
  uint64_t* a_empty = bars + 2 * TX_MAX_SLOTS;      // both MMA warps done   (count 2)
  uint64_t* b_full = bars + 3 * TX_MAX_SLOTS;
  uint64_t* b_empty = bars + 4 * TX_MAX_SLOTS;      // count 2
  uint64_t* tfull_bar = bars + 5 * TX_MAX_SLOTS;    // [2 TMEM stages][2 output rows]
  uint64_t* tempty_bar = tfull_bar + 4;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 4);
  float* mse_smem = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bars) + 640);  // [2 groups][4 warps]
  uint8_t* stg8 = reinterpret_cast<uint8_t*>(bars) + TX_BAR_BYTES;

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  const uint32_t rank = PAIR ? cluster_cta_rank() : 0u;
  // work items of this CTA (pair): item -> (row-pair tile P, N tile tn); both CTAs of a pair walk the same items
  const int item0 = PAIR ? (int)(blockIdx.x >> 1) : (int)blockIdx.x, item_step = PAIR ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  auto tile_P = [&](int item) { return PAIR ? 2 * (item / p.n_tiles) + (int)rank : item / p.n_tiles; };
  if (PAIR) cluster_sync_all();
  if (warp == 0 && lane == 0) {
    prefetch_tmap(&mapA0);
    prefetch_tmap(&mapB);
    for (int i = 0; i < p.nbox; ++i) {
      mbar_init(smem_u32(&a_full[i]), 1);
      mbar_init(smem_u32(&a_ready[i]), PAIR ? 2 * TX_XF_WARPS : TX_XF_WARPS);   // PAIR: the leader's counts both CTAs' warps
      mbar_init(smem_u32(&a_empty[i]), 2);
    }
    for (int i = 0; i < p.b_slots; ++i) { mbar_init(smem_u32(&b_full[i]), PAIR ? 2 : 1); mbar_init(smem_u32(&b_empty[i]), 2); }
    for (int i = 0; i < 4; ++i) { mbar_init(smem_u32(&tfull_bar[i]), 1); mbar_init(smem_u32(&tempty_bar[i]), PAIR ? 8 : 4); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    if (PAIR) {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512)
                   : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512)
                   : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
  }
  tc_fence_before();
  if (PAIR) cluster_sync_all(); else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
#ifdef DCB_PROBES
  const int dbg = p.dbg;
#else
  constexpr int dbg = 0;
#endif

  const uint32_t a_ring0 = smem_u32(a_ring), b_ring0 = smem_u32(b_ring);
  const uint32_t a_full0 = smem_u32(a_full), a_ready0 = smem_u32(a_ready), a_empty0 = smem_u32(a_empty);
  const uint32_t b_full0 = smem_u32(b_full), b_empty0 = smem_u32(b_empty);
  int tap_items = 0;
  for (int s = 0; s < p.ntap; ++s) tap_items += p.tap[s].nkb;

  if (warp == 0) {
    // ===================== TMA producer =====================
    int ai = 0;
    uint32_t aph = 0;
    for (int tile = item0; tile < p.total_tiles; tile += item_step) {
      const TxPair t = decode_pair(p, tile_P(tile));
      for (int kb = 0; kb < p.nkb_conv; ++kb) {
        const bool second = kb >= p.nkb0;
        const CUtensorMap* mp = second ? &mapA1 : &mapA0;
        const int c = (second ? kb - p.nkb0 : kb) * TC_BK;
        const int dv = second ? p.div1 : p.div0;
        const int smp = dv > 1 ? t.nb / dv : t.nb;
        auto load_box = [&](int j) {
          mbar_wait_long(a_empty0 + ai * 8, aph ^ 1);
          if (elect_one()) {
            const uint32_t fa = a_full0 + ai * 8;
            mbar_expect_tx(fa, 130u * 128u);
            tma_load_5d(a_ring0 + (uint32_t)(ai * TX_BOX), mp, fa, c, t.x0 - 1, 0, t.y0 - 1 + j, smp);
          }
          __syncwarp();
          if (++ai == p.nbox) { ai = 0; aph ^= 1; }
        };
        for (int j = 0; j < 4; ++j) load_box(j);
      }
      for (int s = 0; s < p.ntap; ++s) {
        const TcSeg sg = p.tap[s];
        const CUtensorMap* mp = sg.map == 0 ? &mapA0 : (sg.map == 1 ? &mapA1 : &mapA2);
        const int smp = sg.div > 1 ? t.nb / sg.div : t.nb;
        for (int kb = 0; kb < sg.nkb; ++kb) {
          for (int sub = 0; sub < 2; ++sub) {       // one [128 px x 64 ch] tile per output row, in two consecutive slots
            mbar_wait_long(a_empty0 + ai * 8, aph ^ 1);
            if (elect_one()) {
              const uint32_t fa = a_full0 + ai * 8;
              mbar_expect_tx(fa, (uint32_t)TC_A_BYTES);
              tma_load_5d(a_ring0 + (uint32_t)(ai * TX_BOX), mp, fa, sg.c0 + kb * TC_BK, t.x0 + sg.dx, 0, t.y0 + sub + sg.dy, smp);
            }
            __syncwarp();
            if (++ai == p.nbox) { ai = 0; aph ^= 1; }
          }
        }
      }
    }
  } else if (warp == 19) {
    // ===================== TMA producer of the weight blocks, in the MMAs' (channel block, ky, kx) order =====================
    int bi = 0;
    uint32_t bph = 0;
    prefetch_tmap(&mapB);
    for (int tile = item0; tile < p.total_tiles; tile += item_step) {
      const int tn = tile % p.n_tiles;
      const int conv_blocks = 9 * p.nkb_conv;
      for (int i = 0; i < conv_blocks + tap_items; ++i) {
        const int kb_glob = i < conv_blocks ? (i % 9) * p.nkb_conv + i / 9 : i;
        mbar_wait_long(b_empty0 + bi * 8, bph ^ 1);
        if (elect_one()) {
          const uint32_t fb = b_full0 + bi * 8;
          if (PAIR) {      // this CTA's half of the block; both halves are counted on the leader's barrier
            if (rank == 0) mbar_expect_tx(fb, 2u * (uint32_t)b_bytes);
            else mbar_arrive_cluster(fb, 0);
            tma_load_2d_2sm(b_ring0 + (uint32_t)(bi * b_bytes), &mapB, fb, kb_glob * TC_BK, tn * p.BN + (int)rank * b_rows);
          } else {
            mbar_expect_tx(fb, (uint32_t)b_bytes);
            tma_load_2d(b_ring0 + (uint32_t)(bi * b_bytes), &mapB, fb, kb_glob * TC_BK, tn * p.BN);
          }
        }
        __syncwarp();
        if (++bi == p.b_slots) { bi = 0; bph ^= 1; }
      }
    }
  } else if (warp <= 2) {
   if (!PAIR || rank == 0) {
    // ===================== MMA issuers: warp 1 -> output row y0, warp 2 -> output row y0 + 1 =====================
    // (PAIR: of the leader CTA only, each MMA covering that output row of both CTAs)
    const int sub = warp - 1;
    int ai = 0, bi = 0, as = 0;
    uint32_t aph = 0, bph = 0, aphase = 0;
    const uint32_t desc_hi = (uint32_t)(make_kmajor_sw128_desc(0) >> 32);
    const uint32_t b_step = (uint32_t)b_bytes >> 4;
    const uint32_t b_lo_base = ((b_ring0 & 0x3FFFFu) >> 4) | (1u << 16);
    uint32_t b_lo = b_lo_base, b_fb = b_full0, b_eb = b_empty0;
    // slot j positions after the current ring position: index and the parity its barriers are in
    auto slot = [&](int j, uint32_t& par) {
      int idx = ai + j;
      par = aph;
      if (idx >= p.nbox) { idx -= p.nbox; par ^= 1; }
      return idx;
    };
    auto advance_a = [&](int n) {
      ai += n;
      if (ai >= p.nbox) { ai -= p.nbox; aph ^= 1; }
    };
    // a box this warp does not read: wait until it is THIS use's box (so the arrival lands in the right phase), release
    auto skip_box = [&](int j) {
      uint32_t par;
      const int idx = slot(j, par);
      mbar_wait(a_ready0 + idx * 8, par);
      if (elect_one()) {
        mbar_arrive(a_empty0 + idx * 8);
        if (PAIR) mbar_arrive_cluster(a_empty0 + idx * 8, 1);
      }
      __syncwarp();
    };
    auto mma_block = [&](uint32_t alo, uint32_t accumulate, uint32_t a_release, bool last) {
      mbar_wait(b_fb, bph);
      tc_fence_after();
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < TC_BK / 16; ++k) {
          if (PAIR) umma_f16_2sm(tmem_base + (uint32_t)(as * 256 + sub * 128), alo + 2 * k, b_lo + 2 * k, desc_hi, p.idesc,
                                 k == 0 ? accumulate : 1u);
          else umma_f16_lohi2(tmem_base + (uint32_t)(as * 256 + sub * 128), alo + 2 * k, desc_hi, b_lo + 2 * k, desc_hi, p.idesc,
                              k == 0 ? accumulate : 1u);
        }
        if (PAIR) {
          umma_commit_2sm(b_eb);
          if (a_release) umma_commit_2sm(a_release);
          if (last) umma_commit_2sm(smem_u32(&tfull_bar[as * 2 + sub]));
        } else {
          umma_commit(b_eb);
          if (a_release) umma_commit(a_release);
          if (last) umma_commit(smem_u32(&tfull_bar[as * 2 + sub]));
        }
      }
      __syncwarp();
      b_lo += b_step; b_fb += 8; b_eb += 8;
      if (++bi == p.b_slots) { bi = 0; bph ^= 1; b_lo = b_lo_base; b_fb = b_full0; b_eb = b_empty0; }
    };
    for (int tile = item0; tile < p.total_tiles; tile += item_step) {
      mbar_wait_long(smem_u32(&tempty_bar[as * 2 + sub]), aphase ^ 1);
      tc_fence_after();
      uint32_t accumulate = 0;
      for (int kb = 0; kb < p.nkb_conv; ++kb) {
        if (sub == 1) skip_box(0);
        for (int ky = 0; ky < 3; ++ky) {
          uint32_t par;
          const int idx = slot(ky + sub, par);
          mbar_wait(a_ready0 + idx * 8, par);
          const uint32_t alo = (((a_ring0 + (uint32_t)(idx * TX_BOX)) & 0x3FFFFu) >> 4) | (1u << 16);
          for (int kx = 0; kx < 3; ++kx) {
            // tap kx of output pixel x reads box row x + kx: start the operand kx rows (128 B) in
            const bool last = tap_items == 0 && kb == p.nkb_conv - 1 && ky == 2 && kx == 2;
            mma_block(alo + (uint32_t)kx * 8u, accumulate, kx == 2 ? a_empty0 + idx * 8 : 0u, last);
            accumulate = 1;
          }
        }
        if (sub == 0) skip_box(3);
        advance_a(4);
      }
      for (int k = 0; k < tap_items; ++k) {
        uint32_t par;
        skip_box(1 - sub);
        const int idx = slot(sub, par);
        mbar_wait(a_ready0 + idx * 8, par);
        const uint32_t alo = (((a_ring0 + (uint32_t)(idx * TX_BOX)) & 0x3FFFFu) >> 4) | (1u << 16);
        mma_block(alo, accumulate, a_empty0 + idx * 8, k == tap_items - 1);
        accumulate = 1;
        advance_a(2);
      }
      if (++as == 2) { as = 0; aphase ^= 1; }
    }
   }
  } else if (warp <= 10) {
    // ===================== epilogue: warps 3..6 drain output row y0, warps 7..10 row y0 + 1 =====================
    const int q = warp & 3;             // TMEM lane quarter this warp may access
    const int grp = (warp - 3) >> 2;
    EpiGeom gq{p.tiles_x, p.OH, 128, 1, 1, p.OW, p.OH, p.NB, p.uniform};
    int as = 0;
    uint32_t aphase = 0;
    const int release_cta = (PAIR && rank != 0) ? 0 : -1;      // the accumulator is handed back on the leader's barrier
    for (int tile = item0, it = 0; tile < p.total_tiles; tile += item_step, ++it) {
      const int tn = tile % p.n_tiles;
      const TxPair t = decode_pair(p, tile_P(tile));
      const int tm_lin = t.tm0 + grp * p.tiles_x;        // output row y0 + grp
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * 256 + grp * 128);
      if (dbg & 2) {
        mbar_wait_long(smem_u32(&tfull_bar[as * 2 + grp]), aphase);
        tc_fence_after();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if (release_cta < 0) mbar_arrive(smem_u32(&tempty_bar[as * 2 + grp]));
          else mbar_arrive_cluster(smem_u32(&tempty_bar[as * 2 + grp]), 0);
        }
      } else if (p.staged) {
        // each group drains its accumulator through its own HALF-width staging tile (64 columns at a time)
        staged_epilogue_half(gq, e, stg8 + grp * TC_EPI_HALF_BYTES, it & 1, tm_lin, tn, p.BN, taddr,
                             smem_u32(&tfull_bar[as * 2 + grp]), aphase, smem_u32(&tempty_bar[as * 2 + grp]), 1 + grp, release_cta);
      } else {
        // direct epilogue (same arithmetic and summation order as gemm_tc_kernel's): fused eps-MSE, one partial per row tile
        const int rr = q * 32 + lane;
        const int x = t.x0 + rr, y = t.y0 + grp;
        const bool row_ok = x < p.OW && t.nb < p.NB;
        const int pix = y * p.OW + x;
        const int m = t.nb * e.rows_per_sample + pix;
        mbar_wait_long(smem_u32(&tfull_bar[as * 2 + grp]), aphase);
        tc_fence_after();
        float mse_acc = 0.f;
        for (int c = 0; c < p.BN; c += 16) {
          const int n0 = tn * p.BN + c;
          if (n0 >= e.N) break;  // warp-uniform
          float v[16];
          tmem_ld16(taddr + (uint32_t)c, v);
          if (row_ok) {
            epi_add_bias_rowvec16(e, m, n0, v);
            epi_store16(e, m, n0, v, mse_acc, t.nb, pix, false);
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if (release_cta < 0) mbar_arrive(smem_u32(&tempty_bar[as * 2 + grp]));
          else mbar_arrive_cluster(smem_u32(&tempty_bar[as * 2 + grp]), 0);
        }
        if (e.mse_part) {
          float* my_mse = mse_smem + grp * 4;
          mse_acc = warp_sum(mse_acc);
          if (lane == 0) my_mse[q] = mse_acc;
          epi_bar(1 + grp);
          if (q == 0 && lane == 0 && t.nb < p.NB)
            e.mse_part[(int64_t)tm_lin * p.n_tiles + tn] = (my_mse[0] + my_mse[1]) + (my_mse[2] + my_mse[3]);
          epi_bar(1 + grp);
        }
      }
      if (++as == 2) { as = 0; aphase ^= 1; }
    }
  } else if (warp <= 18) {
    // ===================== GroupNorm(+SiLU) transform of the landed row boxes, in place =====================
    const int t = ((warp - 11) << 5) + lane;      // 0..255
    const int l = t & 7, r0 = t >> 3;             // logical 16-byte chunk, first box row
    int ai = 0;
    uint32_t aph = 0;
    const bool remote = PAIR && rank != 0;                       // the MMA warps that wait for the boxes live in the leader CTA
    for (int tile = item0; tile < p.total_tiles; tile += item_step) {
      const TxPair tp = decode_pair(p, tile_P(tile));
      const bool tile_ok = tp.nb < p.NB;
      for (int kb = 0; kb < p.nkb_conv; ++kb) {
        // y = a x + b; with SiLU the canonical form is h = (a/2) x + (b/2), out = h tanh(h) + h  (= y sigmoid(y)): see
        // gn_apply_kernel, which computes exactly this
        float ca[8], cb[8];
        {
          const int nbc = tile_ok ? tp.nb : 0;
          const float4* pa = reinterpret_cast<const float4*>(p.xf_a + (int64_t)nbc * p.xf_C + kb * TC_BK + l * 8);
          const float4* pb = reinterpret_cast<const float4*>(p.xf_b + (int64_t)nbc * p.xf_C + kb * TC_BK + l * 8);
          const float4 a0 = __ldg(pa), a1 = __ldg(pa + 1), b0 = __ldg(pb), b1 = __ldg(pb + 1);
          ca[0] = a0.x; ca[1] = a0.y; ca[2] = a0.z; ca[3] = a0.w; ca[4] = a1.x; ca[5] = a1.y; ca[6] = a1.z; ca[7] = a1.w;
          cb[0] = b0.x; cb[1] = b0.y; cb[2] = b0.z; cb[3] = b0.w; cb[4] = b1.x; cb[5] = b1.y; cb[6] = b1.z; cb[7] = b1.w;
          if (p.silu) {
#pragma unroll
            for (int i = 0; i < 8; ++i) { ca[i] *= 0.5f; cb[i] *= 0.5f; }
          }
        }
        for (int j = 0; j < 4; ++j) {
          const int y = tp.y0 - 1 + j;
          mbar_wait_long(a_full0 + ai * 8, aph);
          if (tile_ok && y >= 0 && y < p.OH && !(dbg & 1)) {       // rows outside the image stay zero (TMA fill) = the conv padding
            const uint32_t base = a_ring0 + (uint32_t)(ai * TX_BOX);
#pragma unroll
            for (int rr = 0; rr < 5; ++rr) {
              const int r = r0 + 32 * rr;
              const int x = tp.x0 - 1 + r;
              if (r < 130 && x >= 0 && x < p.OW) {
                const uint32_t addr = base + (uint32_t)(r * 128) + (uint32_t)((l ^ (r & 7)) << 4);
                uint4 v;
                asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
                // bf16 -> fp32 is a shift (low half) or a mask (high half): one ALU op per element
                const uint32_t wv[4] = {v.x, v.y, v.z, v.w};
                float f[8];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                  f[2 * i] = __uint_as_float(wv[i] << 16);
                  f[2 * i + 1] = __uint_as_float(wv[i] & 0xffff0000u);
                }
#pragma unroll
                for (int i = 0; i < 8; ++i) f[i] = gn_act_bf16(fmaf(f[i], ca[i], cb[i]), p.silu);
                v = pack_bf16x8(f);
                asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
              }
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> visible to the MMA's reads
          }
          __syncwarp();
          if (lane == 0) {
            if (remote) mbar_arrive_cluster(a_ready0 + ai * 8, 0);
            else mbar_arrive(a_ready0 + ai * 8);
          }
          if (++ai == p.nbox) { ai = 0; aph ^= 1; }
        }
      }
      for (int k = 0; k < 2 * tap_items; ++k) {    // plain tap tiles (raw tensors): pass through
        mbar_wait_long(a_full0 + ai * 8, aph);
        __syncwarp();
        if (lane == 0) {
          if (remote) mbar_arrive_cluster(a_ready0 + ai * 8, 0);
          else mbar_arrive(a_ready0 + ai * 8);
        }
        if (++ai == p.nbox) { ai = 0; aph ^= 1; }
      }
    }
  }

  tc_fence_before();
  if (PAIR) cluster_sync_all(); else __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    if (PAIR) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
    else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
  }
}

// ---- host side ----------------------------------------------------------------------------------------------
PFN_cuTensorMapEncodeTiled_v12000 tc_encode_fn();
bool is_conv9(const GemmDev& g);

static int encode_box_map(CUtensorMap* map, const void* src, int C, int H, int W, int NBsrc, int box_px) {
  auto enc = tc_encode_fn();
  DCB_REQUIRE(enc != nullptr, "cuTensorMapEncodeTiled entry point not available");
  const cuuint64_t es = 2;
  cuuint64_t dims[5] = {(cuuint64_t)C, (cuuint64_t)W, 1, (cuuint64_t)H, (cuuint64_t)NBsrc};
  cuuint64_t strides[4] = {(cuuint64_t)C * es, (cuuint64_t)W * C * es, (cuuint64_t)W * C * es, (cuuint64_t)H * W * C * es};
  cuuint32_t box[5] = {(cuuint32_t)TC_BK, (cuuint32_t)box_px, 1, 1, 1};
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(src), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  DCB_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(A/tc2x) failed: %d", (int)r);
  return DCB_OK;
}

// returns DCB_EUNSUPPORTED when the descriptor does not fit this kernel
int launch_gemm_tc2x(const GemmDev& g, cudaStream_t st, int tiles_x, int BN, int uniform, int staged, bool dry_run) {
  const EpiDev& e = g.epi;
  if (g.xf_a == nullptr || !is_conv9(g) || g.OW % 128 != 0 || g.OH % 2 != 0 || BN > 128) return DCB_EUNSUPPORTED;
  const SegDev& s0 = g.seg[0];
  if (s0.stride != 1 || s0.H != g.OH || s0.W != g.OW || s0.c_off != 0 || s0.kc != s0.C || s0.C % TC_BK != 0)
    return DCB_EUNSUPPORTED;
  TxParams p;
  memset(&p, 0, sizeof(p));
  p.nkb0 = s0.C / TC_BK;
  p.nkb_conv = p.nkb0 + g.xf_c1 / TC_BK;
  p.div0 = s0.nb_div > 1 ? s0.nb_div : 1;
  p.div1 = g.xf_div1 > 1 ? g.xf_div1 : 1;
  p.tiles_x = tiles_x; p.OH = g.OH; p.OW = g.OW; p.NB = g.NB; p.BN = BN;
  p.m_tiles = tiles_x * g.OH * g.NB;
  p.n_tiles = (e.N + BN - 1) / BN;
  p.num_P = p.m_tiles / 2;
  // CTA pairs (cta_group::2) whenever the tile width allows it: UMMA M = 256 needs N % 16 == 0, each CTA lands N / 2 weight rows
  const bool pair = BN % 16 == 0 && BN >= 16 && !(knobs() & DCB_KNOB_TC2X_NO_PAIR);
  p.total_tiles = (pair ? (p.num_P + 1) / 2 : p.num_P) * p.n_tiles;
  p.uniform = uniform; p.staged = staged; p.silu = g.xf_silu;
  p.xf_C = p.nkb_conv * TC_BK;
  p.xf_a = g.xf_a; p.xf_b = g.xf_b;

  CUtensorMap maps[3];
  memset(maps, 0, sizeof(maps));
  int nmaps = 0, rc;
  if ((rc = encode_box_map(&maps[nmaps++], s0.src, s0.C, s0.H, s0.W, (g.NB + p.div0 - 1) / p.div0, 130))) return rc;
  // map 1 is the second raw source of a concatenated GroupNorm when there is one (conv1 of the up-path resnets: no tap
  // segments), else the maps after map 0 serve the tap segments (conv2: one halo source + the 1x1 shortcut over [h | skip])
  if (g.xf_c1 > 0) {
    if ((rc = encode_box_map(&maps[nmaps++], g.xf_src1, g.xf_c1, s0.H, s0.W, (g.NB + p.div1 - 1) / p.div1, 130))) return rc;
  }
  const int first_tap_map = nmaps;
  SegDev map_key[3];
  for (int i = 9; i < g.nseg; ++i) {
    const SegDev& s = g.seg[i];
    if (s.stride != 1 || s.H != g.OH || s.W != g.OW || s.kc % TC_BK != 0) return DCB_EUNSUPPORTED;
    int mi = -1;
    for (int j = first_tap_map; j < nmaps; ++j)
      if (map_key[j].src == s.src && map_key[j].C == s.C && map_key[j].nb_div == s.nb_div) mi = j;
    if (mi < 0) {
      if (nmaps >= 3) return DCB_EUNSUPPORTED;
      mi = nmaps++;
      map_key[mi] = s;
      const int dv = s.nb_div > 1 ? s.nb_div : 1;
      if ((rc = encode_box_map(&maps[mi], s.src, s.C, s.H, s.W, (g.NB + dv - 1) / dv, 128))) return rc;
    }
    TcSeg& ts = p.tap[p.ntap++];
    ts.map = mi; ts.div = s.nb_div > 1 ? s.nb_div : 1; ts.nkb = s.kc / TC_BK;
    ts.c0 = s.c_off; ts.dx = s.dx; ts.p = 0; ts.dy = s.dy;
  }
  for (int j = nmaps; j < 3; ++j) maps[j] = maps[0];

  CUtensorMap mapB;
  {
    auto enc = tc_encode_fn();
    cuuint64_t dims[2] = {(cuuint64_t)g.K, (cuuint64_t)e.N};
    cuuint64_t strides[1] = {(cuuint64_t)g.K * 2};
    cuuint32_t box[2] = {(cuuint32_t)TC_BK, (cuuint32_t)(pair ? BN / 2 : BN)};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(&mapB, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(g.W), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    DCB_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(W/tc2x) failed: %d", (int)r);
  }

  // rings: two half-width staging tiles (staged epilogue) or none; 6 row boxes (1.5 channel blocks in flight), the rest weights
  const int b_bytes = (pair ? BN / 2 : BN) * TC_BK * 2;      // per CTA
#ifdef DCB_PROBES
  p.dbg = getenv("DCB_TX_DBG") ? atoi(getenv("DCB_TX_DBG")) : 0;
#endif
  const int fixed = 1024 + TX_BAR_BYTES + (staged ? 2 * TC_EPI_HALF_BYTES : 0);
  // (measured, tools/xf_micro.py: 7 boxes + 4 weight slots beat 6 + 5 on the residual / shortcut convs by 10 % and 8 + 8 on
  //  conv_out by 25 %; a ring that is a multiple of the 4 boxes per channel block does worst)
  int nbox = 7;                   // (PAIR: the halved weight blocks double the weight ring to 8 slots; 9 boxes + 4 slots measured equal)
#ifdef DCB_PROBES
  if (getenv("DCB_TX_NBOX")) nbox = atoi(getenv("DCB_TX_NBOX"));
#endif
  int b_slots = (TC_SMEM_LIMIT - fixed - nbox * TX_BOX) / b_bytes;
  if (b_slots > TX_MAX_SLOTS) b_slots = TX_MAX_SLOTS;
  if (b_slots < 4) return DCB_EUNSUPPORTED;
  p.nbox = nbox; p.b_slots = b_slots;
  p.idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)((pair ? 256 : TC_BM) >> 4) << 24);
  if (dry_run) return DCB_OK;

  const size_t smem = (size_t)fixed + (size_t)nbox * TX_BOX + (size_t)b_slots * b_bytes;
  static std::once_flag attr_once;
  std::call_once(attr_once, [] {
    cudaFuncSetAttribute(gemm_tc2x_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_LIMIT);
    cudaFuncSetAttribute(gemm_tc2x_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_LIMIT);
  });
  if (pair) {
    const int pairs = num_sms() / 2;
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(2 * (p.total_tiles < pairs ? p.total_tiles : pairs));
    cfg.blockDim = dim3(TX_THREADS);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    cudaError_t le = cudaLaunchKernelEx(&cfg, gemm_tc2x_kernel<true>, maps[0], maps[1], maps[2], mapB, p, e);
    if (le != cudaSuccess) {
      set_error("gemm_tc2x<pair> launch failed: %s", cudaGetErrorString(le));
      return (int)le;
    }
  } else {
    const int grid = p.total_tiles < num_sms() ? p.total_tiles : num_sms();
    gemm_tc2x_kernel<false><<<grid, TX_THREADS, smem, st>>>(maps[0], maps[1], maps[2], mapB, p, e);
  }
  DCB_CHECK_LAUNCH("gemm_tc2x");
  return DCB_OK;
}

}  // namespace dcb
