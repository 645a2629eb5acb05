// common.cuh -- shared host/device helpers for libdcb200 (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../../include/dcb200.h"

namespace dcb {

// ---- host side error / launch bookkeeping (api.cu owns the storage) ------------------------------------
void set_error(const char* fmt, ...);
void count_launch(int n = 1);
int num_sms();
// kernel-selection switches for A/B parity tests (dcb_set_knobs; include/dcb200.h DCB_KNOB_*): one relaxed atomic load per
// launch instead of getenv() calls on the launch path
uint32_t knobs();

#define DCB_REQUIRE(cond, ...)            \
  do {                                    \
    if (!(cond)) {                        \
      dcb::set_error(__VA_ARGS__);        \
      return DCB_EINVAL;                  \
    }                                     \
  } while (0)

#define DCB_CHECK_LAUNCH(name)                                            \
  do {                                                                    \
    cudaError_t e__ = cudaGetLastError();                                 \
    if (e__ != cudaSuccess) {                                             \
      dcb::set_error("%s launch failed: %s", name, cudaGetErrorString(e__)); \
      return (int)e__;                                                    \
    }                                                                     \
    dcb::count_launch();                                                  \
  } while (0)

// ---- device helpers --------------------------------------------------------------------------------------
__device__ __forceinline__ float silu_f(float x) { return x / (1.0f + __expf(-x)); }
__device__ __forceinline__ float gelu_tanh_f(float x) {
  // torch F.gelu(approximate='tanh'): 0.5 x (1 + tanh(sqrt(2/pi) (x + 0.044715 x^3)))
  const float k0 = 0.7978845608028654f, k1 = 0.044715f;
  float u = k0 * (x + k1 * x * x * x);
  return 0.5f * x * (1.0f + tanhf(u));
}
__device__ __forceinline__ float gelu_erf_f(float x) { return 0.5f * x * (1.0f + erff(x * 0.7071067811865476f)); }

__device__ __forceinline__ float apply_act(int act, float v) {
  if (act == DCB_ACT_SILU) return silu_f(v);
  if (act == DCB_ACT_GELU_TANH) return gelu_tanh_f(v);
  return v;
}

// packed fp32 pairs (sm_100 add / fma .f32x2: two lanes of arithmetic per issue slot)
__device__ __forceinline__ uint64_t f2_pack(float a, float b) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
  return r;
}
__device__ __forceinline__ void f2_unpack(uint64_t v, float& a, float& b) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v));
}
__device__ __forceinline__ uint64_t f2_add(uint64_t a, uint64_t b) {
  uint64_t r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ uint64_t f2_mul(uint64_t a, uint64_t b) {
  uint64_t r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ uint64_t f2_fma(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}
// ---- activations of the bf16 staged epilogue: one MUFU op each.  Their approximation error (tanh.approx: 2^-11
// relative; erf: Abramowitz-Stegun 7.1.26, 1.5e-7 absolute) is far below the bf16 rounding of the stored result; the
// fp32-verify engine keeps the exact forms above.
__device__ __forceinline__ float tanh_approx_f(float x) {
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(x));
  return t;
}
__device__ __forceinline__ float silu_fast_f(float x) { return x * fmaf(0.5f, tanh_approx_f(0.5f * x), 0.5f); }
__device__ __forceinline__ float gelu_tanh_fast_f(float x) {
  const float k0 = 0.7978845608028654f, k1 = 0.044715f * 0.7978845608028654f;
  const float u = x * fmaf(k1, x * x, k0);
  const float hx = 0.5f * x;
  return fmaf(hx, tanh_approx_f(u), hx);
}
// the same activations on two values at a time through the packed pipe forms: each packed operation rounds exactly like its
// scalar counterpart (IEEE fma / mul), so the results are bit-identical to silu_fast_f / gelu_tanh_fast_f -- the staged
// epilogues use these (their groups sit on the critical path of the short-K projections: FF1 + GELU ran 950 TF/s against
// 1 257 for the same shape without an activation)
__device__ __forceinline__ void act_fast_pair(int act, float& a, float& b) {
  if (act == DCB_ACT_GELU_TANH) {
    const float k0 = 0.7978845608028654f, k1 = 0.044715f * 0.7978845608028654f;
    const uint64_t x = f2_pack(a, b);
    const uint64_t w = f2_fma(f2_pack(k1, k1), f2_mul(x, x), f2_pack(k0, k0));
    float u0, u1;
    f2_unpack(f2_mul(x, w), u0, u1);
    const uint64_t hx = f2_mul(f2_pack(0.5f, 0.5f), x);
    f2_unpack(f2_fma(hx, f2_pack(tanh_approx_f(u0), tanh_approx_f(u1)), hx), a, b);
  } else if (act == DCB_ACT_SILU) {
    const uint64_t x = f2_pack(a, b), half = f2_pack(0.5f, 0.5f);
    float h0, h1;
    f2_unpack(f2_mul(half, x), h0, h1);
    f2_unpack(f2_mul(x, f2_fma(half, f2_pack(tanh_approx_f(h0), tanh_approx_f(h1)), half)), a, b);
  }
}
// Activation of the GroupNorm kernels on bf16 tensors.  The normalisation is y = a x + b; with SiLU the kernels carry the
// HALVED coefficients, h = (a/2) x + (b/2) = y/2 (exact), and SiLU(y) = y sigmoid(y) = y (0.5 tanh(y/2) + 0.5) = h tanh(h) + h:
// one FMA, one MUFU op, one FMA per element.  gn_apply_kernel and the fused transform of gemm_tc2x_kernel must produce the
// same bits: both call this on h.
__device__ __forceinline__ float gn_act_bf16(float h, int silu) {
  if (!silu) return h;
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(h));
  return fmaf(h, t, h);
}
__device__ __forceinline__ float rcp_approx_f(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float ex2_approx_f(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float gelu_erf_fast_f(float x) {
  // erf(|z|) = 1 - (a1 t + a2 t^2 + a3 t^3 + a4 t^4 + a5 t^5) exp(-z^2),  t = 1 / (1 + 0.3275911 |z|)   (two MUFU ops)
  const float z = fabsf(x) * 0.7071067811865476f;
  const float t = rcp_approx_f(fmaf(0.3275911f, z, 1.0f));
  float poly = fmaf(1.061405429f, t, -1.453152027f);
  poly = fmaf(poly, t, 1.421413741f);
  poly = fmaf(poly, t, -0.284496736f);
  poly = fmaf(poly, t, 0.254829592f);
  const float e = poly * t * ex2_approx_f(z * z * -1.4426950408889634f);   // 1 - erf(|z|)
  const float hx = 0.5f * x;
  return fmaf(hx, copysignf(1.0f - e, x), hx);
}
__device__ __forceinline__ float apply_act_fast(int act, float v) {
  if (act == DCB_ACT_SILU) return silu_fast_f(v);
  if (act == DCB_ACT_GELU_TANH) return gelu_tanh_fast_f(v);
  return v;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

template <typename T>
__device__ __forceinline__ float to_f(T v);
template <>
__device__ __forceinline__ float to_f<float>(float v) { return v; }
template <>
__device__ __forceinline__ float to_f<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }

template <typename T>
__device__ __forceinline__ T from_f(float v);
template <>
__device__ __forceinline__ float from_f<float>(float v) { return v; }
template <>
__device__ __forceinline__ __nv_bfloat16 from_f<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

__device__ __forceinline__ float load_as_f(const void* p, int dtype, int64_t i) {
  return dtype == DCB_BF16 ? __bfloat162float(((const __nv_bfloat16*)p)[i]) : ((const float*)p)[i];
}
__device__ __forceinline__ void store_from_f(void* p, int dtype, int64_t i, float v) {
  if (dtype == DCB_BF16) ((__nv_bfloat16*)p)[i] = __float2bfloat16_rn(v);
  else ((float*)p)[i] = v;
}

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ void unpack_bf16x8(const uint4& u, float* f) {
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float2 t = __bfloat1622float2(h[i]);
    f[2 * i] = t.x;
    f[2 * i + 1] = t.y;
  }
}
__device__ __forceinline__ uint4 pack_bf16x8(const float* f) {
  uint4 u;
  u.x = pack_bf16x2(f[0], f[1]);
  u.y = pack_bf16x2(f[2], f[3]);
  u.z = pack_bf16x2(f[4], f[5]);
  u.w = pack_bf16x2(f[6], f[7]);
  return u;
}

// ---- device-side GEMM description shared by the SIMT and tcgen05 engines ---------------------------------
struct SegDev {
  const void* src;
  int C, H, W, c_off, kc, dy, dx, stride;
  int nb_div;  // >= 1: source sample = output sample / nb_div
};

struct EpiDev {
  int M, N, n_out;  // n_out = N (or N/2 for GEGLU)
  int rows_per_sample;  // OH*OW
  const float* bias;
  const float* rowvec;
  const int* rowvec_idx;
  const float* gate;
  const void* residual;
  const int* res_idx;
  void* out;
  const float* mse_target;
  const float* mse_scale;
  float* mse_part;
  float* gn_part;  // per-tile per-column (sum, sumsq) of the written bf16 values, or null
  int rowvec_ld, gate_ld, rows_per_group;
  int act, act_post;
  int res_ld, res_mod, res_dtype;
  int out_ld, out_dtype;
  int mse_div, mse_ld;
  int up_phase;         // 0 or 1 + 2a + b (sub-pixel phase of a folded 2x upsample; see dcb200.h)
  int OH, OW;           // output grid of THIS launch (the low-resolution grid when up_phase != 0)
  float* attn_norms;    // dcb_gemm_desc.attn_norms: per (sample, head) max |q|^2, max |k|^2 of the written rows, or null
  int attn_heads, attn_tok;
};

struct GemmDev {
  int dtype;
  int NB, OH, OW;
  int nseg, K;
  SegDev seg[DCB_MAX_SEGS];
  const void* W;
  EpiDev epi;
  // fused GroupNorm of the 3x3 conv in segments 0..8 (dcb_gemm_desc.xf_*)
  const float* xf_a;
  const float* xf_b;
  const void* xf_src1;
  int xf_c1, xf_div1, xf_silu;
};

// output row of GEMM row m (identity, or the strided position of a folded-upsample phase)
__device__ __forceinline__ int64_t out_row_of(const EpiDev& e, int m) {
  if (e.up_phase == 0) return m;
  const int a = (e.up_phase - 1) >> 1, b = (e.up_phase - 1) & 1;
  const int x = m % e.OW, y = (m / e.OW) % e.OH, n = m / (e.OW * e.OH);
  return ((int64_t)n * 2 * e.OH + 2 * y + a) * (2 * e.OW) + 2 * x + b;
}

// scalar epilogue for one accumulator value (column index n is in OUTPUT space; bias handled by caller)
__device__ __forceinline__ float epi_scalar(const EpiDev& e, int m, int n, float v) {
  int grp = e.rows_per_group > 0 ? m / e.rows_per_group : 0;
  if (e.rowvec) v += e.rowvec[(int64_t)(e.rowvec_idx ? e.rowvec_idx[grp] : grp) * e.rowvec_ld + n];
  v = apply_act(e.act, v);
  if (e.gate) v *= e.gate[(int64_t)grp * e.gate_ld + n];
  if (e.residual) {
    int64_t r = e.res_idx ? e.res_idx[m] : (e.res_mod > 0 ? m % e.res_mod : m);
    v += load_as_f(e.residual, e.res_dtype, r * e.res_ld + n);
  }
  v = apply_act(e.act_post, v);
  return v;
}

// host entry points of the engines (gemm_simt.cu / gemm_tc.cu)
int launch_gemm_simt(const GemmDev& g, cudaStream_t st);
int launch_gemm_tc(const GemmDev& g, cudaStream_t st, bool dry_run = false);
// set by a launch whose epilogue filled EpiDev::attn_norms itself (dcb_gemm runs the row pass otherwise)
extern thread_local bool g_attn_norms_written;
int launch_attn_norms(const void* q, const void* k, int ld, int N, int B, int heads, float* norms, cudaStream_t st);
int tc_geometry(const GemmDev& g, int* m_tiles, int* n_tiles, int* BN);
bool tc_staged(const GemmDev& g);

}  // namespace dcb
