// gemm_tc.cu -- tcgen05 / TMEM / TMA implicit-GEMM engine (bf16 in, fp32 accumulate) for sm_100a.
//
// One persistent, warp-specialised kernel covers every dense contraction on the denoiser path
// (diffusers ResnetBlock2D conv1/conv2/conv_shortcut, Down/Upsample2D convs, Transformer2D proj_in/out,
// attention projections, GEGLU/GELU feed-forwards, DiT patch-embed / QKV / MLP / adaLN linears,
// conv_out / proj_out_2 with the fused eps-MSE epilogue):
//
//   D[m, n] = sum over K-segments  A_seg[pixel(m) + tap_seg, c] * W[n, k]        m = (sample, y, x) output pixel
//
//   warp 0      TMA producer: for each K block (64 channels of one segment/tap) one 5-D tiled TMA load pulls the
//               [128 pixels x 64 ch] activation box straight out of the NHWC tensor -- the box is shifted by the
//               tap offset and TMA's out-of-bounds zero fill *is* the conv padding; stride-2 convs use a
//               (2C, W/2, 2, H/2, N) view of the same memory -- plus one 2-D load of the [BN x 64] weight box.
//               Both land 128B-swizzled, i.e. directly in the canonical K-major UMMA layout.
//   warp 1      MMA issuer: one elected lane issues tcgen05.mma.cta_group::1.kind::f16 (M=128, N=BN, K=16) x4 per
//               stage into a TMEM accumulator; tcgen05.commit releases the smem stage / publishes the accumulator.
//   warps 2-5   epilogue: tcgen05.ld 32x32b.x16 (one TMEM lane = one output pixel per thread), fused
//               bias / time-embedding row vector / activation / GEGLU / adaLN gate / residual / eps-MSE, bf16 store.
//   TMEM is double buffered (2 x BN <= 512 columns) so the epilogue of tile i overlaps the mainloop of tile i+1.
#include <cuda.h>
#include <cudaTypedefs.h>

#include <mutex>

#include "tc_common.cuh"

namespace dcb {

constexpr int TC_THREADS = 320;  // TMA warp, MMA warp, 2 epilogue groups x 4 warps (group g owns TMEM stage g)
constexpr int TC_MAX_STAGES = 8;

struct TcParams {
  int nseg;
  TcSeg seg[DCB_MAX_SEGS];
  int tiles_x, tiles_y, tiles_nb, n_tiles, total_tiles;
  int bw, bh, bn;  // pixel box of one M tile: bw*bh*bn == 128
  int OW, OH, NB;
  int BN, stages, total_kb;
  uint32_t idesc;
  int conv9, nkb_conv;  // segments 0..8 are the taps of one 3x3 conv (any stride): issue them (ky, channel block, kx)
  int kx_outer;         // ... or (kx, channel block, ky): stride-1 convs with OW < 128, the order of gemm_tc2's y-halo mode
  int kb_outer;         // ... or (channel block, ky, kx): stride-1 convs with OW >= 128, the order of gemm_tc2x's row boxes
  int staged;   // epilogue through the swizzled smem staging tile + coalesced second pass
  int uniform;  // every row of an M tile belongs to one rowvec/gate group (tile-constant vectors live in smem)
};

__global__ void __launch_bounds__(TC_THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap mapA0, const __grid_constant__ CUtensorMap mapA1,
               const __grid_constant__ CUtensorMap mapA2, const __grid_constant__ CUtensorMap mapB,
               const __grid_constant__ TcParams p, const __grid_constant__ EpiDev e) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // carve: [stages x (A tile | B tile)] then barriers
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int b_bytes = p.BN * TC_BK * 2;
  const int stage_bytes = TC_A_BYTES + b_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)p.stages * stage_bytes);
  uint64_t* full_bar = bars;                          // [stages]
  uint64_t* empty_bar = bars + TC_MAX_STAGES;         // [stages]
  uint64_t* tfull_bar = bars + 2 * TC_MAX_STAGES;     // [2]
  uint64_t* tempty_bar = bars + 2 * TC_MAX_STAGES + 2;  // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * TC_MAX_STAGES + 4);
  float* mse_smem = reinterpret_cast<float*>(tmem_slot + 2);  // [2 groups][4]
  float* stg = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bars) + TC_BAR_BYTES);  // staged epilogue region

  // warp index through a shuffle: ptxas then knows every role branch below is warp-uniform and keeps loop state, smem
  // addresses and descriptors in uniform registers.  (With a per-thread `lane == 0` region every UTCHMMA / UTMALDG is
  // wrapped in an R2UR + BRA.U.ANY waterfall loop and the MMA thread cannot keep the pipe fed: 75 % -> 98 % of the pipe
  // in profiles/r01_mma_issue_probes.txt.)
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&mapA0);
    prefetch_tmap(&mapB);
    for (int i = 0; i < p.stages; ++i) {
      mbar_init(smem_u32(&full_bar[i]), 1);
      mbar_init(smem_u32(&empty_bar[i]), 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(smem_u32(&tfull_bar[i]), 1);
      mbar_init(smem_u32(&tempty_bar[i]), 4);  // one arrive per epilogue warp
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {  // TMEM allocation (whole 512 columns: 1 CTA per SM by construction)
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"(512)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t smem_base = smem_u32(smem);
  const uint32_t full0 = smem_u32(full_bar), empty0 = smem_u32(empty_bar);

  if (warp == 0) {
    // ===================== TMA producer (warp-uniform bookkeeping, one elected lane issues) =====================
    int stage = 0;
    uint32_t phase = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
      const int tn = tile % p.n_tiles;
      int tm = tile / p.n_tiles;
      const int tx = tm % p.tiles_x;
      tm /= p.tiles_x;
      const int ty = tm % p.tiles_y;
      const int tb = tm / p.tiles_y;
      const int x0 = tx * p.bw, y0 = ty * p.bh, nb0 = tb * p.bn;
      auto issue = [&](const TcSeg& sg, int nbs, int kb, int kb_glob) {
        const CUtensorMap* mp = sg.map == 0 ? &mapA0 : (sg.map == 1 ? &mapA1 : &mapA2);
        mbar_wait_long(empty0 + stage * 8, phase ^ 1);
        if (elect_one()) {
          const uint32_t fb = full0 + stage * 8;
          mbar_expect_tx(fb, (uint32_t)stage_bytes);
          const uint32_t sa = smem_base + (uint32_t)(stage * stage_bytes);
          tma_load_5d(sa, mp, fb, sg.c0 + kb * TC_BK, x0 + sg.dx, sg.p, y0 + sg.dy, nbs);
          tma_load_2d(sa + TC_A_BYTES, &mapB, fb, kb_glob * TC_BK, tn * p.BN);
        }
        __syncwarp();
        if (++stage == p.stages) { stage = 0; phase ^= 1; }
      };
      int s_first = 0, kb_glob = 0;
      if (p.conv9) {
        // 3x3 taps in (ky, channel block, kx) order -- the accumulation order of gemm_tc2's x-halo mode, so a layer's
        // result does not depend on which of the two kernels the tile count selects (bit-stable across batch sizes)
        const int nbs = p.seg[0].div > 1 ? nb0 / p.seg[0].div : nb0;
        if (p.kx_outer) {
          for (int kx = 0; kx < 3; ++kx)
            for (int kb = 0; kb < p.nkb_conv; ++kb)
              for (int ky = 0; ky < 3; ++ky) issue(p.seg[ky * 3 + kx], nbs, kb, (ky * 3 + kx) * p.nkb_conv + kb);
        } else if (p.kb_outer) {
          for (int kb = 0; kb < p.nkb_conv; ++kb)
            for (int ky = 0; ky < 3; ++ky)
              for (int kx = 0; kx < 3; ++kx) issue(p.seg[ky * 3 + kx], nbs, kb, (ky * 3 + kx) * p.nkb_conv + kb);
        } else {
          for (int ky = 0; ky < 3; ++ky)
            for (int kb = 0; kb < p.nkb_conv; ++kb)
              for (int kx = 0; kx < 3; ++kx) issue(p.seg[ky * 3 + kx], nbs, kb, (ky * 3 + kx) * p.nkb_conv + kb);
        }
        s_first = 9;
        kb_glob = 9 * p.nkb_conv;
      }
      for (int s = s_first; s < p.nseg; ++s) {
        const TcSeg sg = p.seg[s];
        const int nbs = sg.div > 1 ? nb0 / sg.div : nb0;
        for (int kb = 0; kb < sg.nkb; ++kb, ++kb_glob) issue(sg, nbs, kb, kb_glob);
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (warp-uniform bookkeeping, one elected lane issues) =====================
    int stage = 0;
    uint32_t phase = 0;
    int as = 0;
    uint32_t aphase = 0;
    const uint32_t desc_hi = (uint32_t)(make_kmajor_sw128_desc(0) >> 32);
    // running descriptor word / barrier addresses of the current stage (no multiplications in the K loop)
    const uint32_t lo_base = ((smem_base & 0x3FFFFu) >> 4) | (1u << 16), lo_step = (uint32_t)stage_bytes >> 4;
    uint32_t alo = lo_base, fb = full0, eb = empty0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
      mbar_wait_long(smem_u32(&tempty_bar[as]), aphase ^ 1);
      const uint32_t d_tmem = tmem_base + (uint32_t)(as * 256);
      const uint32_t tfull_addr = smem_u32(&tfull_bar[as]);
      for (int kb = 0; kb < p.total_kb; ++kb) {
        mbar_wait(fb, phase);
        tc_fence_after();
        // descriptor low words: (address >> 4); +2 advances 16 elements (32 B) along K inside the swizzle atom
        const uint32_t blo = alo + (TC_A_BYTES >> 4);
        if (elect_one()) {
          umma_f16_lohi(d_tmem, alo, blo, desc_hi, p.idesc, (uint32_t)kb);
#pragma unroll
          for (int k = 1; k < TC_BK / 16; ++k) umma_f16_lohi(d_tmem, alo + 2 * k, blo + 2 * k, desc_hi, p.idesc, 1u);
          umma_commit(eb);                                   // frees the smem stage when the MMAs retire
          if (kb == p.total_kb - 1) umma_commit(tfull_addr);  // accumulator complete
        }
        __syncwarp();
        alo += lo_step; fb += 8; eb += 8;
        if (++stage == p.stages) { stage = 0; phase ^= 1; alo = lo_base; fb = full0; eb = empty0; }
      }
      if (++as == 2) { as = 0; aphase ^= 1; }
    }
  } else {
    // ===================== epilogue: two groups of 4 warps; group g drains the tiles that use TMEM stage g ==============
    // (tile it of this CTA -> stage it & 1), so the epilogue of tile i overlaps the epilogue of tile i+1 and the mainloop
    // of whichever tile owns the other stage
    const int q = warp & 3;  // TMEM lane quarter this warp may access
    const int grp = (warp - 2) >> 2;
    const int bar_id = 1 + grp;
    const int r = q * 32 + lane;  // accumulator row == pixel index inside the tile
    const bool geglu = e.act == DCB_ACT_GEGLU;
    uint8_t* my_stg = reinterpret_cast<uint8_t*>(stg) + grp * TC_EPI_BYTES;
    float* my_mse = mse_smem + grp * 4;
    const int as = grp;
    uint32_t aphase = 0;
    for (int tile = blockIdx.x + grp * gridDim.x, it = 0; tile < p.total_tiles; tile += 2 * gridDim.x, ++it, aphase ^= 1) {
      const int tn = tile % p.n_tiles;
      int tm = tile / p.n_tiles;
      const int tm_lin = tm;
      const int tx = tm % p.tiles_x;
      tm /= p.tiles_x;
      const int ty = tm % p.tiles_y;
      const int tb = tm / p.tiles_y;
      const int xl = r % p.bw, yl = (r / p.bw) % p.bh, nl = r / (p.bw * p.bh);
      const int x = tx * p.bw + xl, y = ty * p.bh + yl, nb = tb * p.bn + nl;
      const bool row_ok = x < p.OW && y < p.OH && nb < p.NB;
      const int pix = y * p.OW + x;
      const int m = nb * e.rows_per_sample + pix;

      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * 256);
      if (p.staged) {
        EpiGeom gq{p.tiles_x, p.tiles_y, p.bw, p.bh, p.bn, p.OW, p.OH, p.NB, p.uniform};
        staged_epilogue(gq, e, my_stg, it & 1, tm_lin, tn, p.BN, taddr, smem_u32(&tfull_bar[as]), aphase, true,
                        smem_u32(&tempty_bar[as]), true, bar_id);
        continue;
      }
      mbar_wait_long(smem_u32(&tfull_bar[as]), aphase);
      tc_fence_after();
      float mse_acc = 0.f;
      if (!geglu) {
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
      } else {
        // tile = [128 value rows | 128 gate rows] of the packed GEGLU weight: out = (a + ba) * gelu_erf(g + bg)
        for (int c = 0; c < 128; c += 16) {
          float a[16], g[16];
          tmem_ld16(taddr + (uint32_t)c, a);
          tmem_ld16(taddr + (uint32_t)(128 + c), g);
          if (row_ok) {
            const int wr = tn * 256 + c;
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              float av = a[i] + (e.bias ? e.bias[wr + i] : 0.f);
              float gv = g[i] + (e.bias ? e.bias[wr + 128 + i] : 0.f);
              a[i] = av * gelu_erf_f(gv);
            }
            epi_store16(e, m, tn * 128 + c, a, mse_acc, nb, pix);
          }
        }
      }
      // release the accumulator stage
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&tempty_bar[as]));
      if (e.mse_part) {
        mse_acc = warp_sum(mse_acc);
        if (lane == 0) my_mse[q] = mse_acc;
        epi_bar(bar_id);
        if (q == 0 && lane == 0)
          e.mse_part[(int64_t)tm_lin * p.n_tiles + tn] = (my_mse[0] + my_mse[1]) + (my_mse[2] + my_mse[3]);
        epi_bar(bar_id);
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

// ---- host side ----------------------------------------------------------------------------------------------
PFN_cuTensorMapEncodeTiled_v12000 tc_encode_fn() {
  static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(ptr);
  });
  return fn;
}

// segments 0..8 = the (ky, kx)-ordered taps of one 3x3 convolution over a single source
bool is_conv9(const GemmDev& g) {
  if (g.nseg < 9) return false;
  const SegDev& a = g.seg[0];
  for (int i = 0; i < 9; ++i) {
    const SegDev& s = g.seg[i];
    if (s.src != a.src || s.C != a.C || s.H != a.H || s.W != a.W || s.stride != a.stride || s.c_off != a.c_off ||
        s.kc != a.kc || s.nb_div != a.nb_div || s.dy != i / 3 - 1 || s.dx != i % 3 - 1)
      return false;
  }
  return true;
}

// K-block order of such a conv: (kx, channel block, ky) when the rows are narrower than a tile (stride 1, OW < 128) --
// gemm_tc2 then serves the three ky taps of a (kx, channel block) from one y-halo box -- else (ky, channel block, kx), the
// order of its x-halo boxes.  Every tcgen05 code path follows the same rule, so results do not depend on the kernel chosen.
bool conv9_kx_outer(const GemmDev& g) { return is_conv9(g) && g.seg[0].stride == 1 && g.OW < 128; }
// ... and (channel block, ky, kx) when the rows fill whole tiles (stride 1, OW >= 128): gemm_tc2x then keeps the four input
// rows of a channel block that two vertically adjacent output rows need in shared memory and serves all nine taps from them
bool conv9_kb_outer(const GemmDev& g) { return is_conv9(g) && g.seg[0].stride == 1 && g.OW >= 128; }

struct TcGeom {
  int bw, bh, bn, tiles_x, tiles_y, tiles_nb, BN, n_tiles;
};

static int choose_geometry(const GemmDev& g, TcGeom* t) {
  const int OW = g.OW, OH = g.OH, NB = g.NB;
  if (OH == 1 && NB == 1) {
    t->bw = 128; t->bh = 1; t->bn = 1;
  } else if (OW >= 128) {
    t->bw = 128; t->bh = 1; t->bn = 1;
  } else {
    DCB_REQUIRE(128 % OW == 0, "tcgen05 engine needs OW (%d) to divide 128 or be >= 128", OW);
    t->bw = OW;
    const int rem = 128 / OW;
    if (OH >= rem) { t->bh = rem; t->bn = 1; }
    else {
      DCB_REQUIRE(rem % OH == 0, "tcgen05 engine needs OH (%d) to divide %d", OH, rem);
      t->bh = OH; t->bn = rem / OH;
    }
  }
  t->tiles_x = (OW + t->bw - 1) / t->bw;
  t->tiles_y = (OH + t->bh - 1) / t->bh;
  t->tiles_nb = (NB + t->bn - 1) / t->bn;
  const int m_tiles = t->tiles_x * t->tiles_y * t->tiles_nb;
  const int N = g.epi.N;
  if (g.epi.act == DCB_ACT_GEGLU) {
    DCB_REQUIRE(N % 256 == 0, "GEGLU needs N %% 256 == 0 (got %d)", N);
    t->BN = 256;
  } else if (N <= 256) {
    t->BN = (N + 15) / 16 * 16;
    // shallow K: the epilogue paces the tile, so prefer 128-wide tiles (deeper smem ring + staged epilogue)
    if (t->BN == 256 && g.K <= 2304) t->BN = 128;
    // small-M layers: split N so that more SMs get a tile (smem-bandwidth cost is acceptable below one wave)
    while (t->BN >= 128 && t->BN % 32 == 0 && m_tiles * ((N + t->BN - 1) / t->BN) < num_sms() / 2) t->BN /= 2;
  } else {
    t->BN = g.K <= 2304 ? 128 : 256;
    while (t->BN > 64 && m_tiles * ((N + t->BN - 1) / t->BN) < num_sms()) t->BN /= 2;
  }
  t->n_tiles = (N + t->BN - 1) / t->BN;
  return DCB_OK;
}

// what any staged (coalesced, bf16) epilogue needs of the output / residual, whatever the tile width
static bool staged_layout_ok(const GemmDev& g) {
  const EpiDev& e = g.epi;
  if (knobs() & DCB_KNOB_TC_DIRECT_EPILOGUE) return false;
  return e.out != nullptr && e.out_dtype == DCB_BF16 && e.mse_part == nullptr && e.n_out % 8 == 0 && e.out_ld % 8 == 0 &&
         ((uintptr_t)e.out % 16) == 0 &&
         (e.residual == nullptr || (e.res_dtype == DCB_BF16 && e.res_ld % 8 == 0 && ((uintptr_t)e.residual % 16) == 0));
}
// bf16 outputs with <= 128 output columns per tile (BN <= 128, or GEGLU's 256 -> 128); 16-byte aligned rows
static bool staged_for(const GemmDev& g, const TcGeom& t) {
  return staged_layout_ok(g) && (t.BN <= 128 || g.epi.act == DCB_ACT_GEGLU);
}

bool tc_staged(const GemmDev& g) {
  TcGeom t;
  if (g.dtype != DCB_BF16 || choose_geometry(g, &t)) return false;
  return staged_for(g, t);
}

int tc_geometry(const GemmDev& g, int* m_tiles, int* n_tiles, int* BN) {
  TcGeom t;
  int rc = choose_geometry(g, &t);
  if (rc) return rc;
  *m_tiles = t.tiles_x * t.tiles_y * t.tiles_nb;
  *n_tiles = t.n_tiles;
  *BN = t.BN;
  return DCB_OK;
}

static int encode_a_map(CUtensorMap* map, const SegDev& s, int NBsrc, const TcGeom& t) {
  auto enc = tc_encode_fn();
  DCB_REQUIRE(enc != nullptr, "cuTensorMapEncodeTiled entry point not available");
  cuuint64_t dims[5], strides[4];
  const cuuint64_t es = 2, C = s.C, H = s.H, W = s.W;
  if (s.stride == 1) {
    dims[0] = C; dims[1] = W; dims[2] = 1; dims[3] = H; dims[4] = NBsrc;
    strides[0] = C * es; strides[1] = W * C * es; strides[2] = W * C * es; strides[3] = H * W * C * es;
  } else {
    DCB_REQUIRE(s.stride == 2 && s.H % 2 == 0 && s.W % 2 == 0, "stride-2 segment needs even H, W");
    dims[0] = 2 * C; dims[1] = W / 2; dims[2] = 2; dims[3] = H / 2; dims[4] = NBsrc;
    strides[0] = 2 * C * es; strides[1] = W * C * es; strides[2] = 2 * W * C * es; strides[3] = H * W * C * es;
  }
  cuuint32_t box[5] = {(cuuint32_t)TC_BK, (cuuint32_t)t.bw, 1, (cuuint32_t)t.bh, (cuuint32_t)t.bn};
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(s.src), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  DCB_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(A) failed: %d (C=%d H=%d W=%d NB=%d stride=%d box=%d,%d,%d)",
              (int)r, s.C, s.H, s.W, NBsrc, s.stride, t.bw, t.bh, t.bn);
  return DCB_OK;
}

int launch_gemm_tc2(const GemmDev& g, cudaStream_t st, int bw, int bh, int bn, int tiles_x, int tiles_y, int tiles_nb,
                    int BN, int uniform, int staged);
int launch_gemm_tc2x(const GemmDev& g, cudaStream_t st, int tiles_x, int BN, int uniform, int staged, bool dry_run);
int launch_gemm_tc3(const GemmDev& g, cudaStream_t st, int uniform);

// dry_run: everything but the launch (dcb_gemm_xf_layout: would this descriptor run on the kernel that can apply a fused
// GroupNorm transform?)
int launch_gemm_tc(const GemmDev& g, cudaStream_t st, bool dry_run) {
  DCB_REQUIRE(g.dtype == DCB_BF16, "tcgen05 engine is bf16 only");
  DCB_REQUIRE(g.K % TC_BK == 0, "tcgen05 engine needs K %% 64 == 0 (K=%d)", g.K);
  DCB_REQUIRE(((uintptr_t)g.W & 15) == 0, "weights must be 16-byte aligned");
  TcGeom t;
  int rc = choose_geometry(g, &t);
  if (rc) return rc;

  TcParams p;
  memset(&p, 0, sizeof(p));
  CUtensorMap maps[3];
  memset(maps, 0, sizeof(maps));
  SegDev map_key[3];
  int nmaps = 0;
  p.nseg = g.nseg;
  p.total_kb = 0;
  for (int i = 0; i < g.nseg; ++i) {
    const SegDev& s = g.seg[i];
    DCB_REQUIRE(s.kc % TC_BK == 0 && s.c_off % 8 == 0 && s.C % 8 == 0, "segment %d: kc %% 64, c_off %% 8, C %% 8", i);
    DCB_REQUIRE(((uintptr_t)s.src & 15) == 0, "segment %d: src must be 16-byte aligned", i);
    int mi = -1;
    for (int j = 0; j < nmaps; ++j)
      if (map_key[j].src == s.src && map_key[j].C == s.C && map_key[j].H == s.H && map_key[j].W == s.W &&
          map_key[j].stride == s.stride && map_key[j].nb_div == s.nb_div)
        mi = j;
    DCB_REQUIRE(s.nb_div == 1 || t.bn == 1, "segment %d: nb_div needs tiles that lie inside one sample (OH*OW >= 128)", i);
    if (mi < 0) {
      DCB_REQUIRE(nmaps < 3, "at most 3 distinct A sources per GEMM");
      mi = nmaps++;
      map_key[mi] = s;
      rc = encode_a_map(&maps[mi], s, (g.NB + s.nb_div - 1) / s.nb_div, t);
      if (rc) return rc;
    }
    TcSeg& ts = p.seg[i];
    ts.map = mi;
    ts.div = s.nb_div;
    ts.nkb = s.kc / TC_BK;
    if (s.stride == 1) {
      ts.c0 = s.c_off; ts.dx = s.dx; ts.p = 0; ts.dy = s.dy;
    } else {
      // input x = 2*ox + dx = 2*(ox + floor(dx/2)) + (dx & 1)
      const int px = s.dx & 1, py = s.dy & 1;
      ts.c0 = px * s.C + s.c_off;
      ts.dx = (s.dx - px) / 2;
      ts.p = py;
      ts.dy = (s.dy - py) / 2;
    }
    p.total_kb += ts.nkb;
  }
  for (int j = nmaps; j < 3; ++j) maps[j] = maps[0];
  p.conv9 = is_conv9(g);
  p.kx_outer = conv9_kx_outer(g);
  p.kb_outer = conv9_kb_outer(g);
  p.nkb_conv = p.conv9 ? g.seg[0].kc / TC_BK : 0;

  CUtensorMap mapB;
  {
    auto enc = tc_encode_fn();
    cuuint64_t dims[2] = {(cuuint64_t)g.K, (cuuint64_t)g.epi.N};
    cuuint64_t strides[1] = {(cuuint64_t)g.K * 2};
    cuuint32_t box[2] = {(cuuint32_t)TC_BK, (cuuint32_t)t.BN};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(&mapB, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(g.W), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    DCB_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(W) failed: %d (K=%d N=%d BN=%d)", (int)r, g.K, g.epi.N, t.BN);
  }

  p.tiles_x = t.tiles_x; p.tiles_y = t.tiles_y; p.tiles_nb = t.tiles_nb; p.n_tiles = t.n_tiles;
  p.total_tiles = t.tiles_x * t.tiles_y * t.tiles_nb * t.n_tiles;
  p.bw = t.bw; p.bh = t.bh; p.bn = t.bn;
  p.OW = g.OW; p.OH = g.OH; p.NB = g.NB;
  p.BN = t.BN;
  const int stage_bytes = TC_A_BYTES + t.BN * TC_BK * 2;
  {
    const EpiDev& e = g.epi;
    // bf16 outputs with <= 128 output columns per tile (BN <= 128, or GEGLU's 256 -> 128); 16-byte aligned rows
    p.staged = staged_for(g, t);
    DCB_REQUIRE(e.gn_part == nullptr || p.staged, "gn_part needs the staged bf16 epilogue (ask dcb_gemm_gn_layout first)");
    // all rows of an M tile fall into one rowvec/gate group?
    if (e.rows_per_group <= 0) p.uniform = 1;
    else if (g.OH == 1 && g.NB == 1) p.uniform = e.rows_per_group % TC_BM == 0;
    else p.uniform = t.bn == 1 && e.rows_per_group % e.rows_per_sample == 0;
  }
  // big N<=128-wide problems: 256-pixel CTAs with split A/B rings (+ x-halo reuse for 3x3 convs), see gemm_tc2.cu
  // wide mode of gemm_tc2 (256 x 256 CTA tiles: a third less L2 -> SMEM operand traffic per FLOP for the K <= 3072
  // projections) -- EXPERIMENT, opt-in with DCB_KNOB_TC2_WIDE.  Measured on B200 (M = 204800): K=N=768 438 vs 950 TF/s,
  // N=2304 451 vs 1005, K=3072/N=768 952 vs 1138, K=512/N=1536 310 vs 614: with both 128 x 256 accumulators filling TMEM
  // there is no second accumulator stage, and the thread-per-row direct epilogue (~10k cycles per tile) serialises with
  // the MMAs.  It needs a coalesced epilogue and cta_group::2 (256 columns per CTA) to pay; kept bit-identical to the
  // default path so that it can be A/B'd (test_tc2_wide_256x256_tiles).
  {
    const int m_tiles = t.tiles_x * t.tiles_y * t.tiles_nb;
    const EpiDev& e = g.epi;
    const bool wide = (knobs() & DCB_KNOB_TC2_WIDE) && !(knobs() & DCB_KNOB_NO_TC2) && e.N % 256 == 0 && e.N >= 512 && g.K >= 512 &&
                      e.act != DCB_ACT_GEGLU && e.gn_part == nullptr && e.mse_part == nullptr && e.out != nullptr &&
                      e.out_dtype == DCB_BF16 && e.out_ld % 8 == 0 && ((uintptr_t)e.out % 16) == 0 && !is_conv9(g) &&
                      t.bn == 1 && (m_tiles / 2) * (e.N / 256) >= 2 * num_sms();
    if (wide && g.xf_a == nullptr && !dry_run) {
      rc = launch_gemm_tc2(g, st, t.bw, t.bh, t.bn, t.tiles_x, t.tiles_y, t.tiles_nb, 256, p.uniform, 0);
      if (rc != DCB_EUNSUPPORTED) return rc;
    }
  }
  // plain linear layers with N a multiple of 256 and enough rows: CTA pairs, 256 x 256 tiles (gemm_tc3.cu)
  if (staged_layout_ok(g) && g.xf_a == nullptr && !dry_run && !(knobs() & DCB_KNOB_NO_TC3)) {
    rc = launch_gemm_tc3(g, st, p.uniform);
    if (rc != DCB_EUNSUPPORTED) return rc;
  }
  // (also the fused eps-MSE of conv_out -- N = out_channels, nothing stored: with 9 separately loaded taps it is bound by
  //  L2->SMEM traffic, the x-halo boxes cut that 3x)
  const bool mse_only = g.epi.mse_part != nullptr && g.epi.out == nullptr && g.epi.residual == nullptr && t.bn == 1 &&
                        (g.OH * g.OW) % TC_BM == 0 && !(knobs() & DCB_KNOB_NO_TC2_MSE);
  if ((p.staged || mse_only) && t.BN <= 128 && g.epi.act != DCB_ACT_GEGLU && !(knobs() & DCB_KNOB_NO_TC2) &&
      t.tiles_x * t.tiles_y * t.tiles_nb * t.n_tiles >= 4 * num_sms()) {
    if (g.xf_a != nullptr) {     // fused GroupNorm: the row-box kernel (gemm_tc2x.cu) or nothing
      if (t.bw == 128 && t.bh == 1 && t.bn == 1) {
        rc = launch_gemm_tc2x(g, st, t.tiles_x, t.BN, p.uniform, p.staged, dry_run);
        if (rc != DCB_EUNSUPPORTED) return rc;
      }
    } else if (!dry_run) {
      rc = launch_gemm_tc2(g, st, t.bw, t.bh, t.bn, t.tiles_x, t.tiles_y, t.tiles_nb, t.BN, p.uniform, p.staged);
      if (rc != DCB_EUNSUPPORTED) return rc;
    }
  }
  if (g.xf_a != nullptr) {
    if (!dry_run) set_error("gemm: xf_a (fused GroupNorm) needs the row-box kernel -- ask dcb_gemm_xf_layout first");
    return DCB_EUNSUPPORTED;
  }
  if (dry_run) return DCB_OK;
  const int epi_bytes = p.staged ? 2 * TC_EPI_BYTES : 0;
  int stages = (TC_SMEM_LIMIT - 1024 - TC_BAR_BYTES - epi_bytes) / stage_bytes;
  if (stages > TC_MAX_STAGES) stages = TC_MAX_STAGES;
  if (stages > p.total_kb) stages = p.total_kb < 2 ? 2 : p.total_kb;
  p.stages = stages;
  // instruction descriptor (cute::UMMA::InstrDescriptor): D=f32, A=B=bf16, both K-major, N>>3 @17, M>>4 @24
  p.idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(t.BN >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);

  if (g.epi.mse_part) {
    DCB_REQUIRE(t.bn == 1 && (g.OH * g.OW) % TC_BM == 0, "fused MSE needs OH*OW %% 128 == 0");
  }
  // always ask for more than half of the SM's shared memory so exactly one CTA (one TMEM owner) is resident
  size_t smem = (size_t)stages * stage_bytes + 1024 + TC_BAR_BYTES + epi_bytes;
  if (smem < 120 * 1024) smem = 120 * 1024;
  static std::once_flag attr_once;
  std::call_once(attr_once, [] {
    cudaFuncSetAttribute(gemm_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  });
  int grid = p.total_tiles < num_sms() ? p.total_tiles : num_sms();
  gemm_tc_kernel<<<grid, TC_THREADS, smem, st>>>(maps[0], maps[1], maps[2], mapB, p, g.epi);
  DCB_CHECK_LAUNCH("gemm_tc");
  return DCB_OK;
}

}  // namespace dcb
