// api.cu -- C-ABI glue: error text, launch accounting, descriptor validation and engine selection for dcb_gemm.
#include <stdarg.h>

#include <atomic>
#include <mutex>

#include "common.cuh"

namespace dcb {

static thread_local char g_err[512] = "";
static std::atomic<int64_t> g_launches{0};
static std::atomic<uint32_t> g_knobs{0};
uint32_t knobs() { return g_knobs.load(std::memory_order_relaxed); }

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

int num_sms() {
  static int sms = 0;
  static std::once_flag once;
  std::call_once(once, [] {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess ||
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0)
      sms = 148;
  });
  return sms;
}

static int to_dev(const dcb_gemm_desc* d, GemmDev* g) {
  DCB_REQUIRE(d != nullptr, "gemm: null descriptor");
  DCB_REQUIRE(d->dtype == DCB_F32 || d->dtype == DCB_BF16, "gemm: bad dtype %d", d->dtype);
  DCB_REQUIRE(d->nseg >= 1 && d->nseg <= DCB_MAX_SEGS, "gemm: nseg %d out of range", d->nseg);
  DCB_REQUIRE(d->NB >= 1 && d->OH >= 1 && d->OW >= 1 && d->N >= 1, "gemm: bad output geometry");
  DCB_REQUIRE(d->W != nullptr, "gemm: null weights");
  memset(g, 0, sizeof(*g));
  g->dtype = d->dtype;
  g->NB = d->NB; g->OH = d->OH; g->OW = d->OW;
  g->nseg = d->nseg;
  g->W = d->W;
  int K = 0;
  for (int i = 0; i < d->nseg; ++i) {
    const dcb_seg& s = d->seg[i];
    DCB_REQUIRE(s.src != nullptr && s.kc > 0 && s.c_off >= 0 && s.c_off + s.kc <= s.C, "gemm: segment %d channel range", i);
    DCB_REQUIRE(s.stride == 1 || s.stride == 2, "gemm: segment %d stride must be 1 or 2", i);
    SegDev& o = g->seg[i];
    o.src = s.src; o.C = s.C; o.H = s.H; o.W = s.W; o.c_off = s.c_off; o.kc = s.kc; o.dy = s.dy; o.dx = s.dx;
    o.stride = s.stride;
    o.nb_div = s.nb_div > 1 ? s.nb_div : 1;
    K += s.kc;
  }
  g->xf_a = d->xf_a; g->xf_b = d->xf_b; g->xf_src1 = d->xf_src1;
  g->xf_c1 = d->xf_c1; g->xf_div1 = d->xf_div1; g->xf_silu = d->xf_silu;
  if (d->xf_a != nullptr) {
    DCB_REQUIRE(d->xf_b != nullptr && d->dtype == DCB_BF16 && d->nseg >= 9, "gemm: xf_a needs xf_b, bf16 and a 3x3 conv in segments 0..8");
    DCB_REQUIRE(d->xf_c1 >= 0 && d->xf_c1 % 64 == 0 && (d->xf_c1 == 0) == (d->xf_src1 == nullptr), "gemm: xf_src1 / xf_c1 mismatch");
    DCB_REQUIRE((((uintptr_t)d->xf_a | (uintptr_t)d->xf_b) & 15) == 0, "gemm: xf_a / xf_b must be 16-byte aligned");
    K += 9 * d->xf_c1;
  }
  g->K = K;
  EpiDev& e = g->epi;
  e.M = d->NB * d->OH * d->OW;
  e.N = d->N;
  e.n_out = d->act == DCB_ACT_GEGLU ? d->N / 2 : d->N;
  e.rows_per_sample = d->OH * d->OW;
  e.bias = d->bias; e.rowvec = d->rowvec; e.rowvec_idx = d->rowvec_idx; e.gate = d->gate; e.residual = d->residual; e.res_idx = d->res_idx;
  e.out = d->out; e.mse_target = d->mse_target; e.mse_scale = d->mse_scale; e.mse_part = d->mse_part; e.gn_part = d->gn_part;
  e.rowvec_ld = d->rowvec_ld; e.gate_ld = d->gate_ld; e.rows_per_group = d->rows_per_group;
  e.act = d->act; e.act_post = d->act_post;
  e.res_ld = d->res_ld; e.res_mod = d->res_mod; e.res_dtype = d->res_dtype;
  e.out_ld = d->out_ld; e.out_dtype = d->out_dtype;
  e.mse_div = d->mse_div > 0 ? d->mse_div : 1; e.mse_ld = d->mse_ld;
  e.up_phase = d->up_phase; e.OH = d->OH; e.OW = d->OW;
  DCB_REQUIRE(e.up_phase >= 0 && e.up_phase <= 4, "gemm: up_phase must be 0..4");
  DCB_REQUIRE(e.up_phase == 0 || (e.mse_part == nullptr && e.residual == nullptr), "gemm: up_phase excludes residual / fused MSE");
  DCB_REQUIRE(e.out != nullptr || e.mse_part != nullptr, "gemm: nothing to produce (out and mse_part both NULL)");
  DCB_REQUIRE((e.rowvec == nullptr && e.gate == nullptr) || e.rows_per_group > 0, "gemm: rows_per_group needed");
  DCB_REQUIRE(e.mse_part == nullptr || e.mse_target != nullptr, "gemm: mse_part without mse_target");
  DCB_REQUIRE(d->act != DCB_ACT_GEGLU || d->N % 256 == 0, "gemm: GEGLU needs N %% 256 == 0");
  e.attn_norms = d->attn_norms; e.attn_heads = d->attn_heads; e.attn_tok = d->attn_tok;
  if (d->attn_norms != nullptr) {
    DCB_REQUIRE(d->dtype == DCB_BF16 && d->out_dtype == DCB_BF16 && e.out != nullptr && d->act == DCB_ACT_NONE &&
                    d->act_post == DCB_ACT_NONE && d->up_phase == 0,
                "gemm: attn_norms needs a plain bf16 projection");
    DCB_REQUIRE(d->attn_heads >= 1 && d->N >= 2 * d->attn_heads * 64 && d->attn_tok >= 128 && d->attn_tok % 128 == 0 &&
                    e.M % d->attn_tok == 0 && d->out_ld % 8 == 0 && ((uintptr_t)d->out & 15) == 0,
                "gemm: attn_norms needs N >= 2 * heads * 64, attn_tok %% 128 == 0, rows %% attn_tok == 0, 16-byte rows");
  }
  return DCB_OK;
}

thread_local bool g_attn_norms_written = false;

static int pick_engine(const dcb_gemm_desc* d) {
  if (d->engine == DCB_ENGINE_SIMT || d->engine == DCB_ENGINE_TCGEN05) return d->engine;
  return d->dtype == DCB_BF16 ? DCB_ENGINE_TCGEN05 : DCB_ENGINE_SIMT;
}

}  // namespace dcb

using namespace dcb;

extern "C" int dcb_version(void) { return 111; }
extern "C" const char* dcb_last_error(void) { return g_err; }
extern "C" int64_t dcb_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }
extern "C" void dcb_note_graph_replay(int64_t n_kernels) { g_launches.fetch_add(n_kernels, std::memory_order_relaxed); }

extern "C" void dcb_set_knobs(uint32_t mask) { g_knobs.store(mask, std::memory_order_relaxed); }
extern "C" uint32_t dcb_get_knobs(void) { return knobs(); }

extern "C" int dcb_struct_size(int which) { return which == 0 ? (int)sizeof(dcb_seg) : (int)sizeof(dcb_gemm_desc); }

extern "C" int dcb_gemm(const dcb_gemm_desc* d, dcb_stream stream) {
  GemmDev g;
  int rc = to_dev(d, &g);
  if (rc) return rc;
  const int eng = pick_engine(d);
  cudaStream_t st = (cudaStream_t)stream;
  const EpiDev& e = g.epi;
  if (e.attn_norms != nullptr) {
    cudaMemsetAsync(e.attn_norms, 0, sizeof(float) * (2 + 2 * (size_t)(e.M / e.attn_tok) * e.attn_heads), st);
    g_attn_norms_written = false;
  }
  rc = eng == DCB_ENGINE_TCGEN05 ? launch_gemm_tc(g, st) : launch_gemm_simt(g, st);
  if (rc == DCB_OK && e.attn_norms != nullptr && !g_attn_norms_written)   // this launch's kernel has no norm epilogue
    rc = launch_attn_norms(e.out, (const __nv_bfloat16*)e.out + e.attn_heads * 64, e.out_ld, e.attn_tok, e.M / e.attn_tok,
                           e.attn_heads, e.attn_norms, st);
  return rc;
}

extern "C" int dcb_gemm_gn_layout(const dcb_gemm_desc* d, int32_t* supported) {
  GemmDev g;
  int rc = to_dev(d, &g);
  if (rc) return rc;
  *supported = (pick_engine(d) == DCB_ENGINE_TCGEN05 && g.K % 64 == 0 && tc_staged(g)) ? 1 : 0;
  return DCB_OK;
}

extern "C" int dcb_gemm_xf_layout(const dcb_gemm_desc* d, int32_t* supported) {
  GemmDev g;
  int rc = to_dev(d, &g);
  if (rc) return rc;
  *supported = 0;
  if (pick_engine(d) != DCB_ENGINE_TCGEN05 || g.xf_a == nullptr || g.K % 64 != 0) return DCB_OK;
  rc = launch_gemm_tc(g, nullptr, true);
  if (rc == DCB_OK) *supported = 1;
  return rc == DCB_EUNSUPPORTED ? DCB_OK : rc;
}

extern "C" int dcb_gemm_mse_layout(const dcb_gemm_desc* d, int32_t* rows_per_part, int32_t* n_tiles) {
  GemmDev g;
  int rc = to_dev(d, &g);
  if (rc) return rc;
  if (pick_engine(d) != DCB_ENGINE_TCGEN05) {
    set_error("fused MSE epilogue exists only in the tcgen05 engine");
    return DCB_EUNSUPPORTED;
  }
  int mt, nt, bn;
  rc = tc_geometry(g, &mt, &nt, &bn);
  if (rc) return rc;
  *rows_per_part = 128;
  *n_tiles = nt;
  return DCB_OK;
}
