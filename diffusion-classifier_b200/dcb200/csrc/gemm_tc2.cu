// gemm_tc2.cu -- tcgen05 implicit GEMM, 256 output pixels per CTA (two 128-row TMEM accumulators sharing every weight
// block) with SPLIT activation / weight smem rings and an x-halo mode for stride-1 3x3 convolutions.
//
// Why: with 128x128 tiles the dominant N=128 full-resolution convs of the U-Nets are bound by L2->SMEM traffic
// (16 KB A + 16 KB B per 64-deep K block per 128 pixels ~ 590 KB/tile -> ~850 TF/s measured), not by the tensor pipe.
//   * 256 pixels per CTA: each weight block feeds two MMAs (two accumulators)           -> B traffic / 2
//   * x-halo mode: one [130 px x 64 ch] activation box per (ky, channel block) serves the three kx taps; the tap
//     shift is a 128-byte row offset of the UMMA descriptor start address                -> A traffic / 3
//   => ~244 KB per 128 pixels instead of 590 KB; the kernel becomes MMA / SMEM-bandwidth bound.
// Warp roles, barriers, TMEM double buffering and the staged epilogue are as in gemm_tc.cu (see tc_common.cuh).
#include <cudaTypedefs.h>

#include <mutex>

#include "tc_common.cuh"

namespace dcb {

constexpr int T2_THREADS = 352;  // TMA warp, 2 MMA warps (one per sub-tile), 2 epilogue groups x 4 warps (group g drains sub-tile g)
constexpr int T2_MAX_SLOTS = 8;
constexpr int T2_HALO_SUB = 17 * 1024;  // 130 rows x 128 B = 16640 B, padded to the 1024-B swizzle repeat

struct Tc2Params {
  int nseg;
  TcSeg seg[DCB_MAX_SEGS];
  int halo;      // segments 0..8 are a stride-1 3x3 conv served by halo boxes (map 0); segments 9.. are plain taps
                 // 1: x-halo, one [130 px x 64 ch] box per sub-tile and (ky, channel block); tap kx = row offset kx
                 // 2: y-halo (OW < 128), one [OW px x (2 bh + 2) rows x 64 ch] box per (kx, channel block) for BOTH sub-tiles;
                 //    tap ky of sub-tile s = row offset (s bh + ky) OW
  int halo_jstep;   // descriptor-word step between the three taps served by one box: 8 (one 128-B row) or OW * 8
  int halo_bytes;   // bytes of one halo item (expect_tx)
  int kx_outer;     // conv9 issued (kx, channel block, ky) (see conv9_kx_outer in gemm_tc.cu)
  int kb_outer;     // conv9 issued (channel block, ky, kx) (see conv9_kb_outer in gemm_tc.cu)
  int nkb_conv;  // K blocks per tap of that conv (Cin / 64)
  int halo_div;  // nb_div of the halo source
  int conv9;     // segments 0..8 are one 3x3 conv (halo or not): K blocks are issued in (ky, channel block, kx) order
  int tiles_x, tiles_y, tiles_nb, m_tiles, n_tiles, total_tiles;  // m_tiles counts 128-row sub-tiles; a CTA tile = 2
  int bw, bh, bn, OW, OH, NB, BN;
  int a_slots, b_slots, a_slot_bytes;
  uint32_t idesc;
  int uniform, base_off_mode;
  int acc_stages, acc_stage_cols, acc_sub_cols;   // TMEM accumulators: 2 stages x 2 sub-tiles x 128 columns, or (wide
                                                  // mode, BN = 256) 1 stage x 2 sub-tiles x 256 columns
  int staged;    // 0: direct epilogue (fused eps-MSE of conv_out: nothing is written but per-tile partial sums)
  int dbg;  // experiments: 1 = no TMA traffic (barriers only), 2 = no epilogue work, 4 = no MMA issue
};

struct SubTile {
  int x0, y0, nb0;
};
__device__ __forceinline__ SubTile decode_sub(const Tc2Params& p, int tm_lin) {
  SubTile s;
  if (tm_lin >= p.m_tiles) {  // odd tail: a fully out-of-range box (zero filled), rows masked by the epilogue
    s.x0 = 0; s.y0 = 0; s.nb0 = p.tiles_nb * p.bn;
    return s;
  }
  const int tx = tm_lin % p.tiles_x;
  tm_lin /= p.tiles_x;
  s.x0 = tx * p.bw;
  s.y0 = (tm_lin % p.tiles_y) * p.bh;
  s.nb0 = (tm_lin / p.tiles_y) * p.bn;
  return s;
}

__global__ void __launch_bounds__(T2_THREADS, 1)
gemm_tc2_kernel(const __grid_constant__ CUtensorMap mapA0, const __grid_constant__ CUtensorMap mapA1,
                const __grid_constant__ CUtensorMap mapA2, const __grid_constant__ CUtensorMap mapB,
                const __grid_constant__ Tc2Params p, const __grid_constant__ EpiDev e) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int b_bytes = p.BN * TC_BK * 2;
  uint8_t* a_ring = smem;
  uint8_t* b_ring = a_ring + (size_t)p.a_slots * p.a_slot_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(b_ring + (size_t)p.b_slots * b_bytes);
  uint64_t* a_full = bars;
  uint64_t* a_empty = bars + T2_MAX_SLOTS;
  uint64_t* b_full = bars + 2 * T2_MAX_SLOTS;
  uint64_t* b_empty = bars + 3 * T2_MAX_SLOTS;
  uint64_t* tfull_bar = bars + 4 * T2_MAX_SLOTS;   // [2 TMEM stages][2 sub-tiles]
  uint64_t* tempty_bar = tfull_bar + 4;            // [2 TMEM stages][2 sub-tiles]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 4);
  float* mse_smem = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bars) + 384);  // [2 groups][4 warps]
  uint8_t* stg8 = reinterpret_cast<uint8_t*>(bars) + 512;

  // warp-uniform role dispatch (see gemm_tc.cu): loop state and descriptors stay in uniform registers
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    prefetch_tmap(&mapA0);
    prefetch_tmap(&mapB);
    // every smem slot is read by BOTH MMA warps (one commit each); accumulators are per (stage, sub-tile)
    for (int i = 0; i < p.a_slots; ++i) { mbar_init(smem_u32(&a_full[i]), 1); mbar_init(smem_u32(&a_empty[i]), 2); }
    for (int i = 0; i < p.b_slots; ++i) { mbar_init(smem_u32(&b_full[i]), 1); mbar_init(smem_u32(&b_empty[i]), 2); }
    for (int i = 0; i < 4; ++i) { mbar_init(smem_u32(&tfull_bar[i]), 1); mbar_init(smem_u32(&tempty_bar[i]), 4); }
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
  const int first_tap_seg = p.halo ? 9 : 0;
  // issue-loop experiments (tools/gemm_micro.py, profiles/r01_mma_issue_probes.txt) exist only in -DDCB_PROBES builds:
  // the product kernel carries none of their branches
#ifdef DCB_PROBES
  const int dbg = p.dbg;
#else
  constexpr int dbg = 0;
#endif

  const uint32_t a_ring0 = smem_u32(a_ring), b_ring0 = smem_u32(b_ring);
  const uint32_t a_full0 = smem_u32(a_full), a_empty0 = smem_u32(a_empty);
  const uint32_t b_full0 = smem_u32(b_full), b_empty0 = smem_u32(b_empty);
  if (warp == 0) {
    // ===================== TMA producer (warp-uniform bookkeeping, one elected lane issues) =====================
    int ai = 0, bi = 0;
    uint32_t aph = 0, bph = 0;
    const bool no_tma = (dbg & 1) != 0;
    for (int tile = blockIdx.x; tile < ((dbg & 8) ? 0 : p.total_tiles); tile += gridDim.x) {
      const int tn = tile % p.n_tiles, tp = tile / p.n_tiles;
      const SubTile s0 = decode_sub(p, 2 * tp), s1 = decode_sub(p, 2 * tp + 1);
      auto load_b = [&](int kb_glob) {
        if (dbg & 64) mbar_spin(b_empty0 + bi * 8, bph ^ 1); else mbar_wait_long(b_empty0 + bi * 8, bph ^ 1);
        if (elect_one()) {
          const uint32_t fb = b_full0 + bi * 8;
          if (no_tma) mbar_arrive(fb);
          else {
            mbar_expect_tx(fb, (uint32_t)b_bytes);
            tma_load_2d(b_ring0 + (uint32_t)(bi * b_bytes), &mapB, fb, kb_glob * TC_BK, tn * p.BN);
          }
        }
        __syncwarp();
        if (++bi == p.b_slots) { bi = 0; bph ^= 1; }
      };
      if (p.halo == 2) {
        const int h0 = p.halo_div > 1 ? s0.nb0 / p.halo_div : s0.nb0;
        for (int kx = 0; kx < 3; ++kx)
          for (int kb = 0; kb < p.nkb_conv; ++kb) {
            mbar_wait_long(a_empty0 + ai * 8, aph ^ 1);
            if (elect_one()) {
              const uint32_t fa = a_full0 + ai * 8;
              if (no_tma) mbar_arrive(fa);
              else {
                mbar_expect_tx(fa, (uint32_t)p.halo_bytes);
                tma_load_5d(a_ring0 + (uint32_t)(ai * p.a_slot_bytes), &mapA0, fa, kb * TC_BK, kx - 1, 0, s0.y0 - 1, h0);
              }
            }
            __syncwarp();
            if (++ai == p.a_slots) { ai = 0; aph ^= 1; }
            for (int ky = 0; ky < 3; ++ky) load_b((ky * 3 + kx) * p.nkb_conv + kb);
          }
      } else if (p.halo) {
        const int h0 = p.halo_div > 1 ? s0.nb0 / p.halo_div : s0.nb0, h1 = p.halo_div > 1 ? s1.nb0 / p.halo_div : s1.nb0;
        // items in (channel block, ky) order when kb_outer (stride 1, OW >= 128: always true here), else (ky, channel block)
        const int n_items = 3 * p.nkb_conv;
        for (int it = 0; it < n_items; ++it) {
            const int ky = p.kb_outer ? it % 3 : it / p.nkb_conv, kb = p.kb_outer ? it / 3 : it % p.nkb_conv;
            if (dbg & 64) mbar_spin(a_empty0 + ai * 8, aph ^ 1); else mbar_wait_long(a_empty0 + ai * 8, aph ^ 1);
            if (elect_one()) {
              const uint32_t fa = a_full0 + ai * 8;
              const uint32_t sa = a_ring0 + (uint32_t)(ai * p.a_slot_bytes);
              if (no_tma) mbar_arrive(fa);
              else {
                mbar_expect_tx(fa, 2u * 130u * 128u);
                tma_load_5d(sa, &mapA0, fa, kb * TC_BK, s0.x0 - 1, 0, s0.y0 + ky - 1, h0);
                tma_load_5d(sa + T2_HALO_SUB, &mapA0, fa, kb * TC_BK, s1.x0 - 1, 0, s1.y0 + ky - 1, h1);
              }
            }
            __syncwarp();
            if (++ai == p.a_slots) { ai = 0; aph ^= 1; }
            for (int kx = 0; kx < 3; ++kx) load_b((ky * 3 + kx) * p.nkb_conv + kb);
          }
      }
      auto issue_tap = [&](const TcSeg& sg, int kb, int kb_glob) {
        const CUtensorMap* mp = sg.map == 0 ? &mapA0 : (sg.map == 1 ? &mapA1 : &mapA2);
        const int n0s = sg.div > 1 ? s0.nb0 / sg.div : s0.nb0, n1s = sg.div > 1 ? s1.nb0 / sg.div : s1.nb0;
        mbar_wait_long(a_empty0 + ai * 8, aph ^ 1);
        if (elect_one()) {
          const uint32_t fa = a_full0 + ai * 8;
          const uint32_t sa = a_ring0 + (uint32_t)(ai * p.a_slot_bytes);
          if (no_tma) mbar_arrive(fa);
          else {
            mbar_expect_tx(fa, 2u * TC_A_BYTES);
            tma_load_5d(sa, mp, fa, sg.c0 + kb * TC_BK, s0.x0 + sg.dx, sg.p, s0.y0 + sg.dy, n0s);
            tma_load_5d(sa + TC_A_BYTES, mp, fa, sg.c0 + kb * TC_BK, s1.x0 + sg.dx, sg.p, s1.y0 + sg.dy, n1s);
          }
        }
        __syncwarp();
        if (++ai == p.a_slots) { ai = 0; aph ^= 1; }
        load_b(kb_glob);
      };
      int kb_glob = p.halo ? 9 * p.nkb_conv : 0;
      int s_first = first_tap_seg;
      if (!p.halo && p.conv9) {   // same accumulation order as the halo modes and gemm_tc
        if (p.kx_outer) {
          for (int kx = 0; kx < 3; ++kx)
            for (int kb = 0; kb < p.nkb_conv; ++kb)
              for (int ky = 0; ky < 3; ++ky) issue_tap(p.seg[ky * 3 + kx], kb, (ky * 3 + kx) * p.nkb_conv + kb);
        } else if (p.kb_outer) {
          for (int kb = 0; kb < p.nkb_conv; ++kb)
            for (int ky = 0; ky < 3; ++ky)
              for (int kx = 0; kx < 3; ++kx) issue_tap(p.seg[ky * 3 + kx], kb, (ky * 3 + kx) * p.nkb_conv + kb);
        } else {
          for (int ky = 0; ky < 3; ++ky)
            for (int kb = 0; kb < p.nkb_conv; ++kb)
              for (int kx = 0; kx < 3; ++kx) issue_tap(p.seg[ky * 3 + kx], kb, (ky * 3 + kx) * p.nkb_conv + kb);
        }
        s_first = 9;
        kb_glob = 9 * p.nkb_conv;
      }
      for (int s = s_first; s < p.nseg; ++s) {
        const TcSeg sg = p.seg[s];
        for (int kb = 0; kb < sg.nkb; ++kb, ++kb_glob) issue_tap(sg, kb, kb_glob);
      }
    }
  } else if (warp <= 2) {
    // ===================== two MMA issuers: warp 1 -> sub-tile 0, warp 2 -> sub-tile 1 =====================
    // One issuing thread cannot hide its per-K-block bookkeeping (two try_waits, three commits, descriptor words) behind
    // four or eight 64-cycle MMAs -- the MMA queue is ~2 deep (profiles/r01_mma_issue_probes.txt: one warp 75-98 %,
    // two warps 100 % of the pipe).  With one warp per accumulator the bookkeeping of one overlaps the MMAs of the other.
    const int sub = warp - 1;
    int ai = 0, bi = 0, as = 0;
    uint32_t aph = 0, bph = 0, aphase = 0;
    const int halo_items = p.halo ? 3 * p.nkb_conv : 0;
    int tap_items = 0;
    for (int s = first_tap_seg; s < p.nseg; ++s) tap_items += p.seg[s].nkb;
    const int items = halo_items + tap_items;
    const uint32_t desc_hi = (uint32_t)(make_kmajor_sw128_desc(0) >> 32);
    const bool no_mma = (dbg & 4) != 0;
    const bool no_ring = (dbg & 8) != 0;   // experiment: no smem-ring handshakes at all (pure MMA issue + tile handshake)
    const bool prof = (dbg & 128) != 0;    // experiment: cycle accounting of this warp's waits (block 0 prints)
    long long w_te = 0, w_a = 0, w_b = 0, w_iss = 0, w_com = 0, t_all = prof ? clock64() : 0;
    // running descriptor words / barrier addresses of the current slots (no multiplications in the K loop)
    const uint32_t a_step = (uint32_t)p.a_slot_bytes >> 4, b_step = (uint32_t)b_bytes >> 4;
    const uint32_t a_lo_base = (((a_ring0 + (uint32_t)sub * (p.halo == 1 ? (uint32_t)T2_HALO_SUB : (uint32_t)TC_A_BYTES)) & 0x3FFFFu) >> 4) | (1u << 16);
    const uint32_t jstep = (uint32_t)p.halo_jstep;
    const uint32_t a_lo_tap_base = (((a_ring0 + (uint32_t)sub * (uint32_t)TC_A_BYTES) & 0x3FFFFu) >> 4) | (1u << 16);
    const uint32_t b_lo_base = ((b_ring0 & 0x3FFFFu) >> 4) | (1u << 16);
    uint32_t a_off = 0, b_lo = b_lo_base;           // a_off: (slot index * slot bytes) >> 4
    uint32_t a_fb = a_full0, a_eb = a_empty0, b_fb = b_full0, b_eb = b_empty0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
      { long long c0 = prof ? clock64() : 0; mbar_wait_long(smem_u32(&tempty_bar[as * 2 + sub]), aphase ^ 1); if (prof) w_te += clock64() - c0; }
      const uint32_t d_tmem = tmem_base + (uint32_t)(as * p.acc_stage_cols + sub * p.acc_sub_cols);
      const uint32_t tfull_addr = smem_u32(&tfull_bar[as * 2 + sub]);
      uint32_t accumulate = 0;
      for (int item = 0; item < items; ++item) {
        const bool is_halo = item < halo_items;
        const int nb_blocks = is_halo ? 3 : 1;
        if (!no_ring) { long long c0 = prof ? clock64() : 0; mbar_wait(a_fb, aph); if (prof) w_a += clock64() - c0; }
        const uint32_t alo_item = (is_halo ? a_lo_base : a_lo_tap_base) + a_off;
        const bool last_item = item == items - 1;
        for (int j = 0; j < nb_blocks; ++j) {
          if (!no_ring) { long long c0 = prof ? clock64() : 0; mbar_wait(b_fb, bph); if (prof) w_b += clock64() - c0; }
          tc_fence_after();
          // halo box: output pixel xl with tap kx=j reads smem row xl + j  ->  start the operand j rows (128 B) in
          // (y-halo box: tap ky=j of this sub-tile starts j image rows = j * OW smem rows further in)
          const uint32_t alo = alo_item + (uint32_t)j * jstep;
          const bool last_j = j == nb_blocks - 1;
          if (elect_one()) {
            const long long c_i0 = prof ? clock64() : 0;
            if (!no_mma) {
#pragma unroll
              for (int k = 0; k < TC_BK / 16; ++k)
                umma_f16_lohi2(d_tmem, alo + 2 * k, desc_hi, b_lo + 2 * k, desc_hi, p.idesc, k == 0 ? accumulate : 1u);
            }
            const long long c_i1 = prof ? clock64() : 0;
            if (prof) w_iss += c_i1 - c_i0;
            if (dbg & 32) {          // experiment: release the slots at ISSUE time (plain arrive), not at MMA completion
              mbar_arrive(b_eb);
              if (last_j) mbar_arrive(a_eb);
            } else if (!no_ring) {
              umma_commit(b_eb);
              if (last_j) umma_commit(a_eb);
            }
            if (last_item && last_j) umma_commit(tfull_addr);
            if (prof) w_com += clock64() - c_i1;
          }
          __syncwarp();
          accumulate = 1;
          b_lo += b_step; b_fb += 8; b_eb += 8;
          if (++bi == p.b_slots) { bi = 0; bph ^= 1; b_lo = b_lo_base; b_fb = b_full0; b_eb = b_empty0; }
        }
        a_off += a_step; a_fb += 8; a_eb += 8;
        if (++ai == p.a_slots) { ai = 0; aph ^= 1; a_off = 0; a_fb = a_full0; a_eb = a_empty0; }
      }
      if (++as == p.acc_stages) { as = 0; aphase ^= 1; }
    }
    if (prof && (blockIdx.x % 21 == 0) && lane == 0)
      printf("tc2 block %d mma warp %d: total %lld cycles, waits: tempty %lld a_full %lld b_full %lld; mma issue %lld commits %lld (tiles %d, k-blocks/tile %d)\n", (int)blockIdx.x, sub,
             clock64() - t_all, w_te, w_a, w_b, w_iss, w_com, (p.total_tiles + (int)gridDim.x - 1) / (int)gridDim.x, items + 2 * halo_items);
  } else {
    // ===================== epilogue: warps 3..6 drain sub-tile 0, warps 7..10 sub-tile 1, concurrently =====================
    const int q = warp & 3;             // TMEM lane quarter this warp may access
    const int grp = (warp - 3) >> 2;    // epilogue group == sub-tile == accumulator half
    uint8_t* my_stg = stg8 + grp * TC_EPI_HALF_BYTES;
    EpiGeom gq{p.tiles_x, p.tiles_y, p.bw, p.bh, p.bn, p.OW, p.OH, p.NB, p.uniform};
    int as = 0;
    uint32_t aphase = 0;
    for (int tile = blockIdx.x, it = 0; tile < p.total_tiles; tile += gridDim.x, ++it) {
      const int tn = tile % p.n_tiles, tp = tile / p.n_tiles;
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * p.acc_stage_cols + grp * p.acc_sub_cols);
      if (dbg & 2) {
        mbar_wait_long(smem_u32(&tfull_bar[as * 2 + grp]), aphase);
        tc_fence_after();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(&tempty_bar[as * 2 + grp]));
        if (++as == p.acc_stages) { as = 0; aphase ^= 1; }
        continue;
      }
      if (p.staged) {
        // half-width staging tiles (64 columns at a time): the 33 KB this saves over two full tiles go to the operand rings
        staged_epilogue_half(gq, e, my_stg, it & 1, 2 * tp + grp, tn, p.BN, taddr, smem_u32(&tfull_bar[as * 2 + grp]), aphase,
                             smem_u32(&tempty_bar[as * 2 + grp]), 1 + grp);
      } else {
        // direct epilogue (same arithmetic and summation order as gemm_tc_kernel's): thread-per-row over the BN columns,
        // fused eps-MSE -> one partial per 128-row sub-tile
        const int tm_lin = 2 * tp + grp;
        const SubTile st = decode_sub(p, tm_lin);
        const int rr = q * 32 + lane;
        const int xl = rr % p.bw, yl = (rr / p.bw) % p.bh, nl = rr / (p.bw * p.bh);
        const int x = st.x0 + xl, y = st.y0 + yl, nb = st.nb0 + nl;
        const bool row_ok = tm_lin < p.m_tiles && x < p.OW && y < p.OH && nb < p.NB;
        const int pix = y * p.OW + x;
        const int m = nb * e.rows_per_sample + pix;
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
            epi_store16(e, m, n0, v, mse_acc, nb, pix, false);
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(&tempty_bar[as * 2 + grp]));
        if (e.mse_part) {
          float* my_mse = mse_smem + grp * 4;
          mse_acc = warp_sum(mse_acc);
          if (lane == 0) my_mse[q] = mse_acc;
          epi_bar(1 + grp);
          if (q == 0 && lane == 0 && tm_lin < p.m_tiles)
            e.mse_part[(int64_t)tm_lin * p.n_tiles + tn] = (my_mse[0] + my_mse[1]) + (my_mse[2] + my_mse[3]);
          epi_bar(1 + grp);
        }
      }
      if (++as == p.acc_stages) { as = 0; aphase ^= 1; }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
  }
}

// ---- host side ----------------------------------------------------------------------------------------------
PFN_cuTensorMapEncodeTiled_v12000 tc_encode_fn();
bool is_conv9(const GemmDev& g);
bool conv9_kx_outer(const GemmDev& g);
bool conv9_kb_outer(const GemmDev& g);

static int encode_map(CUtensorMap* map, const SegDev& s, int NBsrc, int bx, int by, int bnb) {
  auto enc = tc_encode_fn();
  DCB_REQUIRE(enc != nullptr, "cuTensorMapEncodeTiled entry point not available");
  cuuint64_t dims[5], strides[4];
  const cuuint64_t es = 2, C = s.C, H = s.H, W = s.W;
  if (s.stride == 1) {
    dims[0] = C; dims[1] = W; dims[2] = 1; dims[3] = H; dims[4] = NBsrc;
    strides[0] = C * es; strides[1] = W * C * es; strides[2] = W * C * es; strides[3] = H * W * C * es;
  } else {
    dims[0] = 2 * C; dims[1] = W / 2; dims[2] = 2; dims[3] = H / 2; dims[4] = NBsrc;
    strides[0] = 2 * C * es; strides[1] = W * C * es; strides[2] = 2 * W * C * es; strides[3] = H * W * C * es;
  }
  cuuint32_t box[5] = {(cuuint32_t)TC_BK, (cuuint32_t)bx, 1, (cuuint32_t)by, (cuuint32_t)bnb};
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(s.src), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  DCB_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(A/tc2) failed: %d", (int)r);
  return DCB_OK;
}

// returns DCB_EUNSUPPORTED when the descriptor does not fit this kernel (the caller falls back to gemm_tc_kernel)
int launch_gemm_tc2(const GemmDev& g, cudaStream_t st, int bw, int bh, int bn, int tiles_x, int tiles_y, int tiles_nb,
                    int BN, int uniform, int staged) {
  const EpiDev& e = g.epi;
  Tc2Params p;
  memset(&p, 0, sizeof(p));
  p.staged = staged;
  // wide mode (BN = 256): two 128 x 256 accumulators fill the 512 TMEM columns, so there is one accumulator stage -- the
  // MMAs of the next tile wait for the epilogue -- in exchange for 64 instead of 94 B/clk/SM of operand traffic
  DCB_REQUIRE(BN <= 128 || (BN == 256 && !staged), "gemm_tc2: BN = 256 needs the direct epilogue");
  p.acc_stages = BN == 256 ? 1 : 2;
  p.acc_stage_cols = BN == 256 ? 0 : 256;
  p.acc_sub_cols = BN == 256 ? 256 : 128;
  p.bw = bw; p.bh = bh; p.bn = bn; p.tiles_x = tiles_x; p.tiles_y = tiles_y; p.tiles_nb = tiles_nb;
  p.m_tiles = tiles_x * tiles_y * tiles_nb;
  p.OW = g.OW; p.OH = g.OH; p.NB = g.NB; p.BN = BN;
  p.n_tiles = (e.N + BN - 1) / BN;
  p.total_tiles = ((p.m_tiles + 1) / 2) * p.n_tiles;
  p.uniform = uniform;
  // measured on B200: the UMMA 128B swizzle is a pure function of the absolute smem address, so an operand that
  // starts j rows into a 1024-B-aligned swizzled box needs base_offset = 0 (setting it per the "(addr>>7)&7" rule breaks)
  p.base_off_mode = 0;
#ifdef DCB_PROBES
  p.dbg = getenv("DCB_TC2_DBG") ? atoi(getenv("DCB_TC2_DBG")) : 0;
#endif
  p.nseg = g.nseg;

  // x-halo mode: the first 9 segments are a stride-1 3x3 conv over one source in (ky,kx) order, full 128-px rows
  bool halo = g.nseg >= 9 && bw == 128 && bh == 1 && bn == 1 && g.OW % 128 == 0 && !(knobs() & DCB_KNOB_TC2_NO_HALO);
  for (int i = 0; halo && i < 9; ++i) {
    const SegDev& s = g.seg[i];
    halo = s.src == g.seg[0].src && s.C == g.seg[0].C && s.H == g.OH && s.W == g.OW && s.stride == 1 && s.c_off == 0 &&
           s.kc == s.C && s.dy == i / 3 - 1 && s.dx == i % 3 - 1 && s.nb_div == g.seg[0].nb_div;
  }
  for (int i = 9; halo && i < g.nseg; ++i) halo = g.seg[i].src != g.seg[0].src;
  // y-halo mode: rows narrower than a tile (a sub-tile = bh full rows of one sample), sub-tile pairs stacked vertically
  bool yhalo = !halo && g.nseg >= 9 && g.OW < 128 && bw == g.OW && bn == 1 && tiles_x == 1 && tiles_y % 2 == 0 &&
               bh * bw == TC_BM && !(knobs() & DCB_KNOB_TC2_NO_YHALO);
  for (int i = 0; yhalo && i < 9; ++i) {
    const SegDev& s = g.seg[i];
    yhalo = s.src == g.seg[0].src && s.C == g.seg[0].C && s.H == g.OH && s.W == g.OW && s.stride == 1 && s.c_off == 0 &&
            s.kc == s.C && s.dy == i / 3 - 1 && s.dx == i % 3 - 1 && s.nb_div == g.seg[0].nb_div;
  }
  for (int i = 9; yhalo && i < g.nseg; ++i) yhalo = g.seg[i].src != g.seg[0].src;
  p.halo = halo ? 1 : (yhalo ? 2 : 0);
  p.halo_jstep = yhalo ? g.OW * 8 : 8;
  p.halo_bytes = yhalo ? (2 * bh + 2) * g.OW * 128 : 2 * 130 * 128;
  halo = halo || yhalo;
  p.kx_outer = conv9_kx_outer(g);
  p.kb_outer = conv9_kb_outer(g);
  p.conv9 = is_conv9(g);
  p.nkb_conv = (halo || p.conv9) ? g.seg[0].kc / TC_BK : 0;
  p.halo_div = halo ? g.seg[0].nb_div : 1;

  CUtensorMap maps[3];
  memset(maps, 0, sizeof(maps));
  SegDev map_key[3];
  int nmaps = 0;
  int rc;
  if (halo) {
    rc = yhalo ? encode_map(&maps[0], g.seg[0], (g.NB + g.seg[0].nb_div - 1) / g.seg[0].nb_div, g.OW, 2 * bh + 2, 1)
               : encode_map(&maps[0], g.seg[0], (g.NB + g.seg[0].nb_div - 1) / g.seg[0].nb_div, 130, 1, 1);
    if (rc) return rc;
    map_key[0] = g.seg[0];
    map_key[0].src = nullptr;  // never matches a tap segment: the halo map has a different box
    nmaps = 1;
  }
  for (int i = halo ? 9 : 0; i < g.nseg; ++i) {
    const SegDev& s = g.seg[i];
    int mi = -1;
    for (int j = 0; j < nmaps; ++j)
      if (map_key[j].src == s.src && map_key[j].C == s.C && map_key[j].H == s.H && map_key[j].W == s.W &&
          map_key[j].stride == s.stride && map_key[j].nb_div == s.nb_div)
        mi = j;
    if (s.nb_div > 1 && bn != 1) return DCB_EUNSUPPORTED;
    if (mi < 0) {
      if (nmaps >= 3) return DCB_EUNSUPPORTED;
      mi = nmaps++;
      map_key[mi] = s;
      rc = encode_map(&maps[mi], s, (g.NB + s.nb_div - 1) / s.nb_div, bw, bh, bn);
      if (rc) return rc;
    }
    TcSeg& ts = p.seg[i];
    ts.map = mi;
    ts.div = s.nb_div;
    ts.nkb = s.kc / TC_BK;
    if (s.stride == 1) {
      ts.c0 = s.c_off; ts.dx = s.dx; ts.p = 0; ts.dy = s.dy;
    } else {
      const int px = s.dx & 1, py = s.dy & 1;
      ts.c0 = px * s.C + s.c_off;
      ts.dx = (s.dx - px) / 2;
      ts.p = py;
      ts.dy = (s.dy - py) / 2;
    }
  }
  for (int j = nmaps; j < 3; ++j) maps[j] = maps[0];

  CUtensorMap mapB;
  {
    auto enc = tc_encode_fn();
    cuuint64_t dims[2] = {(cuuint64_t)g.K, (cuuint64_t)e.N};
    cuuint64_t strides[1] = {(cuuint64_t)g.K * 2};
    cuuint32_t box[2] = {(cuuint32_t)TC_BK, (cuuint32_t)BN};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(&mapB, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(g.W), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    DCB_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(W/tc2) failed: %d", (int)r);
  }

  const int b_bytes = BN * TC_BK * 2;
  p.a_slot_bytes = yhalo ? p.halo_bytes : (halo ? 2 * T2_HALO_SUB : 2 * TC_A_BYTES);
  if (yhalo && p.a_slot_bytes < 2 * TC_A_BYTES) return DCB_EUNSUPPORTED;   // tap segments share the ring slots
  // Ring depths: L2 -> SMEM throughput of a CTA is (bytes in flight) / (TMA latency), so what counts is how many K blocks
  // BOTH rings can hold in flight.  A halo slot feeds three K blocks, a tap slot one: x-halo is balanced at 2 A slots
  // (6 K blocks) + 5 B slots, but tap-mode GEMMs (the K <= 768 projections) were A-ring bound with 2 slots -- with 3 + 3
  // the same kernel moves 20 % more (measured: 782 -> 941 TF/s at M=204800, K=N=768).
  const int kb_per_a = halo ? 3 : 1;
  const int epi_bytes = staged ? 2 * TC_EPI_HALF_BYTES : 0;   // the direct epilogues need no staging tiles: deeper rings
  const int ring_bytes = TC_SMEM_LIMIT - (1024 + 512 + epi_bytes);
  int best_a = 2, best_score = -1;
  for (int a = 2; a <= 4; ++a) {
    int b = (ring_bytes - a * p.a_slot_bytes) / b_bytes;
    if (b > T2_MAX_SLOTS) b = T2_MAX_SLOTS;
    if (b < 3) break;
    const int score = a * kb_per_a < b ? a * kb_per_a : b;
    if (score > best_score) { best_score = score; best_a = a; }
  }
  p.a_slots = best_a;
#ifdef DCB_PROBES
  if (getenv("DCB_TC2_ASLOTS")) p.a_slots = atoi(getenv("DCB_TC2_ASLOTS"));
#endif
  const int fixed = 1024 + 512 + epi_bytes + p.a_slots * p.a_slot_bytes;
  int b_slots = (TC_SMEM_LIMIT - fixed) / b_bytes;
  if (b_slots > T2_MAX_SLOTS) b_slots = T2_MAX_SLOTS;
#ifdef DCB_PROBES
  if (getenv("DCB_TC2_BSLOTS") && atoi(getenv("DCB_TC2_BSLOTS")) < b_slots) b_slots = atoi(getenv("DCB_TC2_BSLOTS"));
#endif
  if (b_slots < 3) return DCB_EUNSUPPORTED;
  p.b_slots = b_slots;
  p.idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);

  const size_t smem = (size_t)fixed + (size_t)b_slots * b_bytes;
  static std::once_flag attr_once;
  std::call_once(attr_once, [] {
    cudaFuncSetAttribute(gemm_tc2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_LIMIT);
  });
  const int grid = p.total_tiles < num_sms() ? p.total_tiles : num_sms();
  gemm_tc2_kernel<<<grid, T2_THREADS, smem, st>>>(maps[0], maps[1], maps[2], mapB, p, e);
  DCB_CHECK_LAUNCH("gemm_tc2");
  return DCB_OK;
}

}  // namespace dcb
