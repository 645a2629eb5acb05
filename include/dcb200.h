/* dcb200.h -- C ABI of libdcb200.so: hand-written sm_100a kernels for the ELBO-classification hot path
 * of faverogian/diffusion-classifier (diffusion/diffusion_classifier.py:657-725 and the diffusers 0.31.0
 * denoisers behind nets/unet.py:186-195, nets/dit.py:49-51).
 *
 * The reference has no FFI of its own (pure Python over torch/diffusers); the boundary a maintainer binds is
 * this header, called from Python via ctypes (INTEGRATION.md shows the stub).  Conventions:
 *   - plain pointers and sizes only; every pointer is a DEVICE pointer unless named host_*;
 *   - the caller owns every buffer (inputs, outputs, workspaces); the library never allocates device memory;
 *   - every entry point is stream-ordered on `stream` (a cudaStream_t passed as void*), re-entrant, and
 *     returns 0 on success, a negative DCB_E* code or a positive cudaError_t; dcb_last_error() has the text;
 *   - activations are NHWC ("pixels x channels"), dtype DCB_BF16 (fast path) or DCB_F32 (fp32-verify path);
 *   - weights are [N][K] K-major in the activation dtype, packed once by the host (see docs in DESIGN.md).
 */
#ifndef DCB200_H
#define DCB200_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* dcb_stream; /* cudaStream_t */

enum { DCB_F32 = 0, DCB_BF16 = 1 };
enum { DCB_ACT_NONE = 0, DCB_ACT_SILU = 1, DCB_ACT_GELU_TANH = 2, DCB_ACT_GEGLU = 3 };
enum { DCB_ENGINE_AUTO = 0, DCB_ENGINE_SIMT = 1, DCB_ENGINE_TCGEN05 = 2 };
enum { DCB_OK = 0, DCB_EINVAL = -1, DCB_EUNSUPPORTED = -2, DCB_EDRIVER = -3 };

#define DCB_MAX_SEGS 12

/* ---- library ------------------------------------------------------------------------------------------ */
int dcb_version(void);
const char* dcb_last_error(void);
/* number of kernel launches issued by this library since load (bench.py's "gpu_launches") */
int64_t dcb_launch_count(void);
/* the host replayed a CUDA graph that holds n_kernels of this library's launches (keeps dcb_launch_count honest) */
void dcb_note_graph_replay(int64_t n_kernels);

/* kernel-selection switches used by the A/B parity tests (every alternative computes bit-identical results; the default,
 * mask 0, is the product path).  Process-wide; takes effect from the next launch. */
enum {
  DCB_KNOB_NO_TC2 = 1,              /* 256-pixel CTA kernel (gemm_tc2) off: everything runs on gemm_tc_kernel */
  DCB_KNOB_TC2_NO_HALO = 2,         /* gemm_tc2 without x-halo boxes (nine separately loaded taps) */
  DCB_KNOB_TC2_NO_YHALO = 4,        /* gemm_tc2 without y-halo boxes */
  DCB_KNOB_NO_TC2_MSE = 8,          /* fused eps-MSE of conv_out on gemm_tc_kernel */
  DCB_KNOB_TC2_WIDE = 16,           /* experimental single-CTA 256 x 256 tiles (measured slower; kept as an A/B) */
  DCB_KNOB_TC_DIRECT_EPILOGUE = 32, /* thread-per-row epilogue instead of the staged (coalesced) one */
  DCB_KNOB_ATTN_NO_TC = 64,         /* head-dim-64 attention on the mma.sync kernel */
  DCB_KNOB_ATTN_NO_FAST = 128,      /* tcgen05 attention always with the running maximum (no single-pass kernel) */
  DCB_KNOB_NO_TC3 = 256,            /* no CTA-pair (cta_group::2) GEMM for the N % 256 == 0 linear layers */
  DCB_KNOB_TC2X_NO_PAIR = 512       /* fused-GroupNorm conv on single CTAs (no cta_group::2 weight sharing) */
};
void dcb_set_knobs(uint32_t mask);
uint32_t dcb_get_knobs(void);

/* ---- (1) prologue: q_sample fused with the denoiser's input staging ------------------------------------
 * replaces DiffusionClassifier.diffuse (diffusion_classifier.py:100-117) + the first-layer unfold:
 *   eps = predrawn[u] or Philox(seed, unit_id0+u);  z = alpha[u]*x[img[u]] + sigma[u]*eps
 *   a_out[(u*rep + r) * rows + row][k] = im2col3x3(z) (mode 0, U-Net conv_in, K order (ky,kx,c)) or
 *                                        patchify(z)  (mode 1, DiT PatchEmbed, K order (py,px,c)), zero padded to kpad
 *   target[u*rows + row][c'] = eps - (v_param ? sigma[u]*z : 0)     (layout NHWC / token-major (py,px,c))
 * x: [n_img,C,H,W] fp32 NCHW; eps_predrawn: [U,C,H,W] fp32 NCHW or NULL; img: [U] int32 or NULL (identity);
 * sigma == NULL means "no noise" (z = alpha*x; plain forward); z_ws: [U,H,W,C] fp32 scratch holding z (NHWC). */
int dcb_prologue(int mode, int dtype, const float* x, const float* eps_predrawn, uint64_t seed, int64_t unit_id0,
                 const float* alpha, const float* sigma, const int32_t* img, int U, int rep, int C, int H, int W,
                 int patch, int kpad, float* z_ws, void* a_out, float* target, int v_param, dcb_stream stream);

/* ---- (1b) DDPM ancestral sampler step with classifier-free guidance (next-row f2) -------------------------
 * replaces DiffusionClassifier.ddpm_sampler_step + the z_s draw (diffusion_classifier.py:176-207, 263-264, 280-284):
 *   pred = (1+w)*cond - w*uncond;  x = clip(v_param ? a_t*z - s_t*pred : (z - s_t*pred)/a_t);
 *   mu = a_s*(z*(1-c)/a_t + c*x);  z_out = final_step ? clip(mu) : mu + sqrt(var)*noise
 * z_t / z_out / noise_predrawn: [B,C,H,W] fp32 NCHW (noise NULL -> Philox(seed, unit_id0+b)); pred: fp32 output of the
 * denoiser's last GEMM for samples b*rep (+1 = unconditional when rep == 2), NHWC rows (patch == 0) or DiT token-major
 * (py,px,c) columns (patch > 0); coef: DEVICE [8] = {c, a_t, a_s, s_t, s_s, sqrt(var), w, step index i as int32 bits};
 * the Philox unit of image b is unit_id0 + i*B + b (so a captured CUDA graph of the step draws fresh noise per replay). */
int dcb_ddpm_step(const float* z_t, const float* pred, int rep, int patch, const float* coef, int v_param,
                  int final_step, const float* noise_predrawn, uint64_t seed, int64_t unit_id0, int B, int C, int H,
                  int W, float* z_out, dcb_stream stream);

/* sinusoidal embedding of the noise label (diffusers get_timestep_embedding):
 * out[s][0:half] = cos(t*w_k), out[s][half:] = sin(t*w_k), w_k = exp(-ln(max_period)*k/(half-shift)); s = u*rep+r */
int dcb_timestep_embed(int dtype, const float* t, int U, int rep, int dim, float shift, float max_period, void* out,
                       dcb_stream stream);

/* ---- (2) implicit GEMM: 3x3 / 1x1 / strided convolutions and linear layers ------------------------------ */
typedef struct dcb_seg {
  const void* src;  /* NHWC activations [NB,H,W,C] */
  int32_t C, H, W;  /* src channel count / spatial dims */
  int32_t c_off;    /* first src channel consumed by this segment */
  int32_t kc;       /* channels consumed (K extent); multiple of 64 (tcgen05) / 16 (SIMT) */
  int32_t dy, dx;   /* input pixel = (oy*stride + dy, ox*stride + dx); out-of-range reads as 0 (padding) */
  int32_t stride;   /* 1 or 2 */
  int32_t nb_div;   /* 0/1: src sample = n;  d > 1: src sample = n / d -- a tensor computed once per (image, timestep)
                       unit is read by all d class-conditional samples of that unit (shared class-independent prefix) */
  int32_t _r1;
} dcb_seg;

typedef struct dcb_gemm_desc {
  int32_t dtype, engine;
  int32_t NB, OH, OW; /* output rows m = (n*OH + y)*OW + x */
  int32_t N;          /* weight rows (GEMM N); with DCB_ACT_GEGLU the output has N/2 columns */
  int32_t nseg, _r0;
  dcb_seg seg[DCB_MAX_SEGS]; /* K = concatenation of the segments, in order */
  const void* W;             /* [N][K] */
  const float* bias;         /* [N] or NULL */
  const float* rowvec;       /* v += rowvec[g*rowvec_ld + n], g = m / rows_per_group, or rowvec_idx[g] when given
                                (time-embedding projection / collapsed single-token cross-attention bias per class) */
  const int32_t* rowvec_idx;
  const float* gate;         /* v *= gate[(m / rows_per_group)*gate_ld + n]       (adaLN-Zero gates) */
  const void* residual;      /* v += residual[row][n], row = res_idx ? res_idx[m] : (res_mod ? m % res_mod : m) */
  const int32_t* res_idx;
  void* out;                 /* [M][out_ld] or NULL when only the MSE epilogue is wanted */
  const float* mse_target;   /* fused eps-MSE: part[tile] = sum_{rows,cols}(mse_scale[s]*v - target[(s/mse_div)*rps + pix][n])^2 */
  const float* mse_scale;    /* [NB] or NULL (=1) */
  float* mse_part;           /* [m_tiles * n_tiles] partial sums, reduced by dcb_mse_finalize */
  float* gn_part;            /* optional [ceil(M/128)][n_out][2]: per-128-row-tile, per-column (sum, sum of squares) of the
                                bf16 values just written -- the GroupNorm statistics of the NEXT layer, so its stats
                                pass (a full re-read of this tensor) disappears; dcb_gemm_gn_layout says if supported */
  int32_t rowvec_ld, gate_ld, rows_per_group;
  int32_t act, act_post;     /* act before gate/residual, act_post after */
  int32_t res_ld, res_mod, res_dtype;
  int32_t out_ld, out_dtype;
  int32_t mse_div, mse_ld;
  int32_t up_phase;          /* 0: out row = m.  1 + 2a + b: this GEMM computes phase (a, b) of a 2x nearest-upsampled
                                output (Upsample2D + conv folded into four 2x2-tap convs over the LOW-resolution input):
                                row m = (n, y, x) of the (OH, OW) grid is stored at row (n, 2y + a, 2x + b) of the
                                (2 OH, 2 OW) output; gn_part tiles are laid out [n][phase][tile] */
  int32_t xf_silu;           /* (see xf_a) apply SiLU after the affine */
  /* fused GroupNorm(+SiLU) of the conv's INPUT: when xf_a != NULL, segments 0..8 (a stride-1 3x3 conv) name the RAW,
   * pre-normalisation tensor(s) -- channels [0, seg.C) from seg[0].src and, when xf_c1 > 0, channels [seg.C, seg.C + xf_c1)
   * from xf_src1 ([ceil(NB / xf_div1), H, W, xf_c1]: the up-path skip of a concatenated GroupNorm) -- and the kernel
   * applies y = xf_a[n][c] * x + xf_b[n][c] (then SiLU) to the operand on the fly, bit-identical to
   * dcb_groupnorm_apply.  xf_a / xf_b: fp32 [NB][seg.C + xf_c1] from dcb_groupnorm_coef_from_tiles.  Weights span
   * K = 9 (seg.C + xf_c1).  Only some launches support it: ask dcb_gemm_xf_layout. */
  const float* xf_a;
  const float* xf_b;
  const void* xf_src1;
  int32_t xf_c1, xf_div1;
  /* attention pre-pass folded into the projection that PRODUCES q and k (a bf16 linear over [rows, >= 2*attn_heads*64]
   * output columns, q heads first, then k heads, head dim 64): when attn_norms != NULL the call also leaves
   * max_i |q_i|^2 and max_j |k_j|^2 per (sample, head) in attn_norms[2 + (sample * attn_heads + head) * 2 + {0, 1}]
   * (fp32 [2 + 2 * samples * attn_heads]; sample = row / attn_tok; the buffer is zeroed by this call), computed from the
   * epilogue's own registers where the launch runs on the CTA-pair kernel and by a pass over the written rows otherwise.
   * Hand the buffer to dcb_attention_ws with DCB_ATTN_NORMS_READY or-ed into dtype: that launch then skips its own
   * pre-pass over q and k (1.7 % of the DiT-B/4 step). */
  float* attn_norms;
  int32_t attn_heads, attn_tok;
} dcb_gemm_desc;

int dcb_gemm(const dcb_gemm_desc* d, dcb_stream stream);
/* sizeof(dcb_seg) (which = 0) / sizeof(dcb_gemm_desc) (which = 1) as compiled: lets a binding check its struct layout */
int dcb_struct_size(int which);
/* *supported = 1 when dcb_gemm(d) would fill d->gn_part (tcgen05 engine, staged bf16 epilogue); rows per tile = 128 */
int dcb_gemm_gn_layout(const dcb_gemm_desc* d, int32_t* supported);
/* *supported = 1 when dcb_gemm(d) can apply the fused GroupNorm transform d->xf_a asks for (tcgen05 engine, 256-pixel
 * CTAs with x-halo boxes: stride-1 3x3 conv, OW a multiple of 128, at least 4 tiles per SM) */
int dcb_gemm_xf_layout(const dcb_gemm_desc* d, int32_t* supported);
/* rows covered by one mse_part entry for descriptor d (128 for tcgen05, 64 for SIMT), and n-tile count */
int dcb_gemm_mse_layout(const dcb_gemm_desc* d, int32_t* rows_per_part, int32_t* n_tiles);
/* err[s] (+)= sum of parts_per_sample consecutive partials (fixed order => deterministic) */
int dcb_mse_finalize(const float* part, int parts_per_sample, int S, float* err, int err_stride, dcb_stream stream);
/* unfused eps-MSE (a7: diffusion_classifier.py:706-711): err[s] = sum_k (scale[s]*pred[s][k] - target[s/div][k])^2 */
int dcb_eps_mse(int dtype, const void* pred, const float* target, const float* scale, int S, int div, int64_t K,
                float* err, int err_stride, dcb_stream stream);

/* ---- (3) normalisation ------------------------------------------------------------------------------ */
/* GroupNorm (+optional SiLU) over the channel-concatenation of x0 [NB,HW,C0] and x1 [NB,HW,C1] (x1 may be NULL):
 * stats pass writes part[NB][chunks][G][2] (sum, sumsq); apply pass reduces them in fixed order. */
int dcb_groupnorm_stats(int dtype, const void* x0, int C0, const void* x1, int C1, int NB, int HW, int G, int chunks,
                        float* part, dcb_stream stream);
int dcb_groupnorm_apply(int dtype, const void* x0, int C0, const void* x1, int C1, int NB, int HW, int G, int chunks,
                        const float* part, const float* gamma, const float* beta, float eps, int silu, void* out,
                        dcb_stream stream);
/* statistics + apply in ONE launch, one thread block per sample: for samples too small to feed a GEMM tile of their own
 * (HW < 128: no producer-side tile statistics), where two launches and a second HBM read of the tensor dominate.
 * Needs (C0 + C1) / G to be a multiple of the vector width (8 bf16 / 4 fp32) and to divide C0. */
int dcb_groupnorm_fused(int dtype, const void* x0, int C0, int div0, const void* x1, int C1, int div1, int NB, int HW,
                        int G, const float* gamma, const float* beta, float eps, int silu, void* out, dcb_stream stream);
/* statistics pass replaced by a reduction of producer-written tile partials (dcb_gemm_desc.gn_part):
 * part0/part1: [n_src*tiles_per_sample][C][2] per source (part1 NULL when C1 == 0), sample n reads source sample n/div;
 * writes part_out[NB][1][G][2] = (sum, sumsq) per group, i.e. the `part` of dcb_groupnorm_apply with chunks = 1 */
int dcb_groupnorm_stats_from_tiles(const float* part0, int C0, int div0, const float* part1, int C1, int div1, int NB,
                                   int tiles_per_sample, int G, float* part_out, dcb_stream stream);
/* the same reduction, finished into the per-(sample, channel) affine of the normalisation: coef_a[n][c] = rstd * gamma[c],
 * coef_b[n][c] = beta[c] - mean * coef_a[n][c]  (fp32 [NB][C0 + C1] each) -- what dcb_gemm_desc.xf_a / xf_b consume */
int dcb_groupnorm_coef_from_tiles(const float* part0, int C0, int div0, const float* part1, int C1, int div1, int NB,
                                  int tiles_per_sample, int G, const float* gamma, const float* beta, float eps,
                                  float* coef_a, float* coef_b, dcb_stream stream);
/* same, with per-source sample divisors: sample n reads x0[n / div0] and x1[n / div1] (see dcb_seg.nb_div) */
int dcb_groupnorm_stats_div(int dtype, const void* x0, int C0, int div0, const void* x1, int C1, int div1, int NB, int HW,
                            int G, int chunks, float* part, dcb_stream stream);
int dcb_groupnorm_apply_div(int dtype, const void* x0, int C0, int div0, const void* x1, int C1, int div1, int NB, int HW,
                            int G, int chunks, const float* part, const float* gamma, const float* beta, float eps,
                            int silu, void* out, dcb_stream stream);
/* LayerNorm over C with optional affine (gamma,beta) and optional adaLN modulation
 * y = LN(x)*(1 + scale[g]) + shift[g], g = row / rows_per_group */
int dcb_layernorm(int dtype, const void* x, int64_t rows, int C, const float* gamma, const float* beta, float eps,
                  const float* scale, const float* shift, int mod_ld, int rows_per_group, void* out, dcb_stream stream);

/* ---- (4) self-attention: softmax(Q K^T * scale) V per (batch, head) ------------------------------------
 * q/k/v: [B, Ntok, heads, d] views with row stride ld (elements); out [B, Ntok, heads*d] row stride out_ld */
int dcb_attention(int dtype, const void* q, const void* k, const void* v, int ld, int B, int Ntok, int heads, int d,
                  float scale, void* out, int out_ld, dcb_stream stream);
/* same with a [2 + B*heads*2] fp32 workspace: lets the tcgen05 kernel (d = 64) take its single-pass path, which bounds
 * every score by max|q| max|k| per (batch, head) and so needs no running maximum (exact; falls back by itself).
 * dtype | DCB_ATTN_NORMS_READY: the workspace already holds this launch's maxima (dcb_gemm_desc.attn_norms). */
#define DCB_ATTN_NORMS_READY 0x200
int dcb_attention_ws(int dtype, const void* q, const void* k, const void* v, int ld, int B, int Ntok, int heads, int d,
                     float scale, void* out, int out_ld, float* ws, dcb_stream stream);

/* ---- (5) Haar DWT / IDWT (utils/wavelet.py:4-35 / 37-67), batched NCHW fp32 -------------------------------
 * dwt: [B,C,H,W] -> [B,4C,H/2,W/2] channel order 4i+{0,1,2,3} = cA,cH,cV,cD; out *= post_scale */
int dcb_haar_dwt(const float* x, int B, int C, int H, int W, float post_scale, float* out, dcb_stream stream);
int dcb_haar_idwt(const float* w, int B, int C4, int h, int wd, float pre_scale, float* out, dcb_stream stream);

/* ---- data movement helpers ------------------------------------------------------------------------------ */
int dcb_upsample2x(int dtype, const void* x, int NB, int H, int W, int C, void* out, dcb_stream stream);
/* out[n] = x[n / div]: materialised class expansion of a per-unit tensor [NB/div, HW, C] -> [NB, HW, C] (only for
 * geometries where a GEMM tile spans several samples and dcb_seg.nb_div cannot be used) */
int dcb_expand_samples(int dtype, const void* x, int NB, int div, int64_t elems_per_sample, void* out, dcb_stream stream);
int dcb_nhwc_to_nchw(int dtype, const void* x, int NB, int HW, int C, int ld, float* out, dcb_stream stream);
/* DiT unpatchify: tok [B, g*g, p*p*C] (dtype) -> [B,C,g*p,g*p] fp32  ("nhwpqc->nchpwq") */
int dcb_unpatchify(int dtype, const void* tok, int B, int g, int p, int C, int ld, float* out, dcb_stream stream);
/* dst[i] = (dtype) src_f32[i] */
int dcb_cast_f32(int dtype, const float* src, int64_t n, void* dst, dcb_stream stream);

/* ---- weight packing (once per parameter version): checkpoint tensors (diffusers layout, fp32, DEVICE) -> the operand
 * layouts dcb_gemm consumes, in `dtype`.  Sums are formed in fp32 and rounded once (as dcb_cast_f32). ---------------------- */
/* conv weight [Cout][Cin][kh][kw] -> out [Cout][kpad]: column (ky*kw + kx)*Cin + ci, zero padded to kpad (3x3 / 1x1 convs,
 * conv_in with K padded to 64, DiT patch embedding with K order (py, px, c)) */
int dcb_pack_conv(int dtype, const float* w_oihw, int Cout, int Cin, int kh, int kw, int kpad, void* out, dcb_stream stream);
/* GEGLU projection w [2*inner][C] (rows [0,inner) value, [inner,2*inner) gate: diffusers' chunk order), bias [2*inner] ->
 * rows interleaved per 128 so that a 256-wide GEMM tile holds [128 value | 128 gate] of the same outputs (DCB_ACT_GEGLU) */
int dcb_pack_geglu(int dtype, const float* w, const float* bias, int inner, int C, void* w_out, float* bias_out,
                   dcb_stream stream);
/* Upsample2D (nearest 2x) + conv3x3 [Cout][Cin][3][3] -> out [4][Cout][4*Cin]: phase 2a + b, K order (ty, tx, cin); the 3x3
 * taps that land on the same low-resolution pixel are summed (dcb_gemm_desc.up_phase consumes phase 2a + b) */
int dcb_pack_upsample(int dtype, const float* w_oihw, int Cout, int Cin, void* out, dcb_stream stream);
/* dst[dst_row0 + r][dst_col0 + c] = (dtype) src[r][c], r < rows, c < cols: builds row- or column-concatenated operands (fused
 * QKV [3C][C], all time-embedding projections [sum Cout][tdim], [conv2 | 1x1 shortcut] along K) */
int dcb_pack_rows(int dtype, const float* src, int src_ld, int rows, int cols, void* dst, int dst_ld, int dst_row0,
                  int dst_col0, dcb_stream stream);

#ifdef __cplusplus
}
#endif
#endif /* DCB200_H */
