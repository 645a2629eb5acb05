// norm.cu -- HBM-bound normalisation kernels on NHWC activations:
//   GroupNorm(32 groups)+SiLU over an (optionally channel-concatenated) tensor   [diffusers ResnetBlock2D.norm1/2,
//        Transformer2DModel.norm, UNet conv_norm_out; the torch.cat of the up-path skip is folded in here]
//   LayerNorm with optional affine and optional adaLN modulation                 [BasicTransformerBlock.norm1/2/3,
//        AdaLayerNormZero, DiT norm_out]
// All reductions are fixed-order (no float atomics) so results are run-to-run deterministic.
// Vectorised 16-byte accesses; one read for statistics + one read/one write for the apply pass.
#include "common.cuh"

namespace dcb {

template <typename T>
struct Vec;
template <>
struct Vec<__nv_bfloat16> {
  static constexpr int N = 8;
  __device__ static void load(const __nv_bfloat16* p, float* f) { unpack_bf16x8(*reinterpret_cast<const uint4*>(p), f); }
  __device__ static void unpack(const uint4& u, float* f) { unpack_bf16x8(u, f); }
  __device__ static void store(__nv_bfloat16* p, const float* f) { *reinterpret_cast<uint4*>(p) = pack_bf16x8(f); }
};
template <>
struct Vec<float> {
  static constexpr int N = 4;
  __device__ static void load(const float* p, float* f) {
    float4 v = *reinterpret_cast<const float4*>(p);
    f[0] = v.x; f[1] = v.y; f[2] = v.z; f[3] = v.w;
  }
  __device__ static void store(float* p, const float* f) {
    *reinterpret_cast<float4*>(p) = make_float4(f[0], f[1], f[2], f[3]);
  }
  __device__ static void unpack(const uint4& u, float* f) {
    f[0] = __uint_as_float(u.x); f[1] = __uint_as_float(u.y); f[2] = __uint_as_float(u.z); f[3] = __uint_as_float(u.w);
  }
};

constexpr int GN_THREADS = 256;

// ---- GroupNorm statistics: part[((n*chunks + chunk)*G + g)*2 + {sum,sumsq}] -------------------------------
template <typename T>
__global__ void __launch_bounds__(GN_THREADS) gn_stats_kernel(const T* __restrict__ x0, int C0, const T* __restrict__ x1,
                                                              int C1, int HW, int G, int chunks, float* __restrict__ part,
                                                              int div0, int div1) {
  constexpr int VN = Vec<T>::N;
  __shared__ float s_acc[GN_THREADS][VN][2];
  const int n = blockIdx.y, chunk = blockIdx.x;
  const int C = C0 + C1, V = C / VN, cpg = C / G;
  const int ppc = (HW + chunks - 1) / chunks;
  const int p0 = chunk * ppc, p1 = min(HW, p0 + ppc);
  const int vslots = min(V, (int)blockDim.x);
  const int nrows = blockDim.x / vslots;
  const int prow = threadIdx.x / vslots, vs = threadIdx.x % vslots;
  float* out = part + ((int64_t)(n * chunks + chunk) * G) * 2;
  float gsum = 0.f, gsq = 0.f;  // owned by thread g < G

  for (int vb = 0; vb < V; vb += vslots) {
    const int v = vb + vs;
    float s[VN], q[VN];
#pragma unroll
    for (int i = 0; i < VN; ++i) s[i] = q[i] = 0.f;
    if (v < V && prow < nrows) {
      const int c = v * VN;
      const T* base;
      int cs, cc;
      if (c < C0) { base = x0 + (int64_t)(n / div0) * HW * C0; cs = C0; cc = c; }
      else { base = x1 + (int64_t)(n / div1) * HW * C1; cs = C1; cc = c - C0; }
      // 8 raw 16-byte loads in flight per thread (Little: ~66 KB/SM must be outstanding to saturate HBM3e)
      const T* bp = base + cc;
      for (int p = p0 + prow; p < p1; p += 8 * nrows) {
        uint4 raw[8];
#pragma unroll
        for (int u = 0; u < 8; ++u)
          if (p + u * nrows < p1) raw[u] = *reinterpret_cast<const uint4*>(bp + (int64_t)(p + u * nrows) * cs);
#pragma unroll
        for (int u = 0; u < 8; ++u)
          if (p + u * nrows < p1) {
            float f[VN];
            Vec<T>::unpack(raw[u], f);
#pragma unroll
            for (int i = 0; i < VN; ++i) { s[i] += f[i]; q[i] = fmaf(f[i], f[i], q[i]); }
          }
      }
    }
#pragma unroll
    for (int i = 0; i < VN; ++i) { s_acc[threadIdx.x][i][0] = s[i]; s_acc[threadIdx.x][i][1] = q[i]; }
    __syncthreads();
    if (threadIdx.x < G) {
      const int g = threadIdx.x;
      const int c_lo = max(g * cpg, vb * VN), c_hi = min((g + 1) * cpg, (vb + vslots) * VN);
      for (int c = c_lo; c < c_hi; ++c) {
        const int slot = c / VN - vb, i = c % VN;
        for (int r = 0; r < nrows; ++r) {
          gsum += s_acc[r * vslots + slot][i][0];
          gsq += s_acc[r * vslots + slot][i][1];
        }
      }
    }
    __syncthreads();
  }
  if (threadIdx.x < G) { out[threadIdx.x * 2] = gsum; out[threadIdx.x * 2 + 1] = gsq; }
}

// bf16 tensors: y = a x + b, then SiLU through gn_act_bf16 on HALVED coefficients (common.cuh; one MUFU op per element --
// the exp + rcp form needs two and made this kernel SFU-bound); fp32 (verify engine) keeps the exact form
template <typename T>
struct GnAct;
template <>
struct GnAct<float> {
  static constexpr bool HALVE = false;
  __device__ static float apply(float y, int silu) { return silu ? silu_f(y) : y; }
};
template <>
struct GnAct<__nv_bfloat16> {
  static constexpr bool HALVE = true;
  __device__ static float apply(float h, int silu) { return gn_act_bf16(h, silu); }
};

// ---- GroupNorm apply (+SiLU), writes the concatenated normalised tensor [NB,HW,C0+C1] ----------------------
// thread -> fixed 16-byte channel slot (coefficients live in registers), loops over pixels with 4 loads in flight
// FUSED is a template parameter so that the streaming kernel's code generation is untouched by the fused mode (sharing one
// kernel body through a run-time branch cost the big launches 30 %: 3.7 -> 4.9 ms on the 128^2 tensors, ncu launch lists)
template <typename T, bool FUSED>
__global__ void __launch_bounds__(GN_THREADS) gn_apply_kernel(const T* __restrict__ x0, int C0, const T* __restrict__ x1,
                                                              int C1, int HW, int G, int chunks,
                                                              const float* __restrict__ part, const float* __restrict__ gamma,
                                                              const float* __restrict__ beta, float eps, int silu,
                                                              T* __restrict__ out, int pix_per_block, int div0, int div1) {
  constexpr int VN = Vec<T>::N;
  __shared__ float s_mean[64], s_rstd[64];
  const int n = blockIdx.y;
  const int C = C0 + C1, V = C / VN, cpg = C / G;
  if (FUSED) {
    // fused mode (one block per sample, dcb_groupnorm_fused; V <= blockDim): the statistics pass runs here with the apply
    // pass's own thread -> (pixel row, 16-byte channel slot) mapping, so the whole sample is in flight at once; a thread's
    // slot lies in one group, partials are combined per group in fixed order.  The apply pass re-reads the sample from
    // L1/L2: HBM sees one read and one write, and there is one launch instead of two.
    __shared__ float s_ps[GN_THREADS], s_pq[GN_THREADS];
    const int nrows_f = blockDim.x / V, prow_f = threadIdx.x / V, v_f = threadIdx.x % V;
    const int c_f = v_f * VN;
    const T* base;
    int cs;
    if (c_f < C0) { base = x0 + (int64_t)(n / div0) * HW * C0 + c_f; cs = C0; }
    else { base = x1 + (int64_t)(n / div1) * HW * C1 + (c_f - C0); cs = C1; }
    float s = 0.f, q = 0.f;
    for (int p = prow_f; p < HW; p += 4 * nrows_f) {
      float f[4][VN];
#pragma unroll
      for (int u = 0; u < 4; ++u)
        if (p + u * nrows_f < HW) Vec<T>::load(base + (int64_t)(p + u * nrows_f) * cs, f[u]);
#pragma unroll
      for (int u = 0; u < 4; ++u)
        if (p + u * nrows_f < HW) {
#pragma unroll
          for (int i = 0; i < VN; ++i) { s += f[u][i]; q = fmaf(f[u][i], f[u][i], q); }
        }
    }
    s_ps[threadIdx.x] = s;
    s_pq[threadIdx.x] = q;
    __syncthreads();
    if (threadIdx.x < G) {
      const int vpg = cpg / VN;
      double ds = 0.0, dq = 0.0;
      for (int r = 0; r < nrows_f; ++r)
        for (int j = 0; j < vpg; ++j) {
          const int t = r * V + threadIdx.x * vpg + j;
          ds += s_ps[t];
          dq += s_pq[t];
        }
      const double cnt = (double)HW * cpg;
      const double mean = ds / cnt;
      double var = dq / cnt - mean * mean;
      if (var < 0.0) var = 0.0;
      s_mean[threadIdx.x] = (float)mean;
      s_rstd[threadIdx.x] = (float)(1.0 / sqrt(var + (double)eps));
    }
  } else if (threadIdx.x < G) {
    double s = 0.0, q = 0.0;
    const float* pp = part + ((int64_t)n * chunks * G + threadIdx.x) * 2;
    for (int k = 0; k < chunks; ++k) { s += pp[(int64_t)k * G * 2]; q += pp[(int64_t)k * G * 2 + 1]; }
    const double cnt = (double)HW * cpg;
    const double mean = s / cnt;
    double var = q / cnt - mean * mean;
    if (var < 0.0) var = 0.0;
    s_mean[threadIdx.x] = (float)mean;
    s_rstd[threadIdx.x] = (float)(1.0 / sqrt(var + (double)eps));
  }
  __syncthreads();
  const int vslots = min(V, (int)blockDim.x);
  const int nrows = blockDim.x / vslots;
  const int prow = threadIdx.x / vslots, vs = threadIdx.x % vslots;
  const int p0 = blockIdx.x * pix_per_block, p1 = min(HW, p0 + pix_per_block);
  T* ob = out + (int64_t)n * HW * C;
  for (int vb = 0; vb < V; vb += vslots) {
    const int v = vb + vs;
    if (v >= V) continue;
    const int c = v * VN;
    float ca[VN], cbias[VN];
#pragma unroll
    for (int i = 0; i < VN; ++i) {
      const int g = (c + i) / cpg;
      ca[i] = s_rstd[g] * gamma[c + i];
      cbias[i] = __fmaf_rn(-s_mean[g], ca[i], beta[c + i]);     // written out: gn_tiles_finalize_kernel must produce the same bits
      if (GnAct<T>::HALVE && silu) { ca[i] *= 0.5f; cbias[i] *= 0.5f; }
    }
    const T* base;
    int cs;
    if (c < C0) { base = x0 + (int64_t)(n / div0) * HW * C0 + c; cs = C0; }
    else { base = x1 + (int64_t)(n / div1) * HW * C1 + (c - C0); cs = C1; }
    T* obase = ob + c;
    for (int p = p0 + prow; p < p1; p += 4 * nrows) {
      float f[4][VN];
#pragma unroll
      for (int u = 0; u < 4; ++u)
        if (p + u * nrows < p1) Vec<T>::load(base + (int64_t)(p + u * nrows) * cs, f[u]);
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        if (p + u * nrows < p1) {
#pragma unroll
          for (int i = 0; i < VN; ++i) {
            f[u][i] = GnAct<T>::apply(fmaf(f[u][i], ca[i], cbias[i]), silu);
          }
          Vec<T>::store(obase + (int64_t)(p + u * nrows) * C, f[u]);
        }
      }
    }
  }
}

// ---- GroupNorm statistics from the producer GEMM's tile partials (dcb_gemm_desc.gn_part) -------------------
// one block per sample; thread (slice, g) walks every 8th tile of group g's channels, fixed-order combine
// coef_a != nullptr: also finish the statistics into the per-(sample, channel) affine y = a x + b exactly as gn_apply_kernel
// derives it from `out` (same double-precision mean / rstd from the fp32 group sums, same fp32 products)
__global__ void __launch_bounds__(512) gn_tiles_finalize_kernel(const float* __restrict__ p0, int C0, int div0,
                                                                const float* __restrict__ p1, int C1, int div1, int tps,
                                                                int G, float* __restrict__ out, int HW,
                                                                const float* __restrict__ gamma,
                                                                const float* __restrict__ beta, float eps,
                                                                float* __restrict__ coef_a, float* __restrict__ coef_b) {
  __shared__ double sh[8][64][2];
  __shared__ float s_mean[64], s_rstd[64];
  const int n = blockIdx.x, g = threadIdx.x % G, slice = threadIdx.x / G;
  const int cpg = (C0 + C1) / G;
  double s = 0.0, q = 0.0;
  for (int c = g * cpg; c < (g + 1) * cpg; ++c) {
    const float* src;
    int stride;
    if (c < C0) { src = p0 + ((int64_t)(n / div0) * tps * C0 + c) * 2; stride = C0 * 2; }
    else { src = p1 + ((int64_t)(n / div1) * tps * C1 + (c - C0)) * 2; stride = C1 * 2; }
    for (int t = slice; t < tps; t += 8) {
      const float2 v = *reinterpret_cast<const float2*>(src + (int64_t)t * stride);
      s += v.x;
      q += v.y;
    }
  }
  sh[slice][g][0] = s;
  sh[slice][g][1] = q;
  __syncthreads();
  if (slice == 0) {
    for (int k = 1; k < 8; ++k) { s += sh[k][g][0]; q += sh[k][g][1]; }
    if (out) {
      out[((int64_t)n * G + g) * 2] = (float)s;
      out[((int64_t)n * G + g) * 2 + 1] = (float)q;
    }
    if (coef_a) {      // gn_apply_kernel reads the fp32 sums back: round exactly as it would see them
      const double sf = (double)(float)s, qf = (double)(float)q;
      const double cnt = (double)HW * cpg;
      const double mean = sf / cnt;
      double var = qf / cnt - mean * mean;
      if (var < 0.0) var = 0.0;
      s_mean[g] = (float)mean;
      s_rstd[g] = (float)(1.0 / sqrt(var + (double)eps));
    }
  }
  if (coef_a) {
    __syncthreads();
    const int C = C0 + C1;
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
      const int gg = c / cpg;
      const float ca = s_rstd[gg] * gamma[c];
      coef_a[(int64_t)n * C + c] = ca;
      coef_b[(int64_t)n * C + c] = __fmaf_rn(-s_mean[gg], ca, beta[c]);
    }
  }
}

static int gn_threads(int V) { return V >= GN_THREADS ? GN_THREADS : (GN_THREADS / V) * V; }

template <typename T>
static int gn_stats_t(const void* x0, int C0, const void* x1, int C1, int NB, int HW, int G, int chunks, float* part,
                      cudaStream_t st, int div0, int div1) {
  const int V = (C0 + C1) / Vec<T>::N;
  dim3 grid(chunks, NB);
  gn_stats_kernel<T><<<grid, gn_threads(V), 0, st>>>((const T*)x0, C0, (const T*)x1, C1, HW, G, chunks, part, div0, div1);
  DCB_CHECK_LAUNCH("gn_stats");
  return DCB_OK;
}

template <typename T>
static int gn_apply_t(const void* x0, int C0, const void* x1, int C1, int NB, int HW, int G, int chunks, const float* part,
                      const float* gamma, const float* beta, float eps, int silu, void* out, cudaStream_t st, int div0,
                      int div1) {
  const int V = (C0 + C1) / Vec<T>::N;
  const int threads = gn_threads(V);
  const int nrows = threads / (V < threads ? V : threads);
  const int vwin = (V + threads - 1) / threads;           // channel windows a thread walks through
  int ppb = nrows * (32 / (vwin < 32 ? vwin : 32));      // ~32 vectors per thread
  if (ppb < nrows) ppb = nrows;
  dim3 grid((HW + ppb - 1) / ppb, NB);
  gn_apply_kernel<T, false><<<grid, threads, 0, st>>>((const T*)x0, C0, (const T*)x1, C1, HW, G, chunks, part, gamma, beta, eps,
                                               silu, (T*)out, ppb, div0, div1);
  DCB_CHECK_LAUNCH("gn_apply");
  return DCB_OK;
}

static int gn_check(int dtype, int C0, int C1, int G, const void* x1) {
  const int vn = dtype == DCB_BF16 ? 8 : 4;
  DCB_REQUIRE(G > 0 && G <= 64 && (C0 + C1) % G == 0, "groupnorm: C=%d not divisible by G=%d (G<=64)", C0 + C1, G);
  DCB_REQUIRE(C0 % vn == 0 && C1 % vn == 0, "groupnorm: channel counts must be multiples of %d", vn);
  DCB_REQUIRE((C1 == 0) == (x1 == nullptr), "groupnorm: x1/C1 mismatch");
  return DCB_OK;
}

// ---- LayerNorm: one warp per row, row cached in registers; per-channel coefficients by 16-byte loads -------------
// (measured: keeping the coefficients in registers across several rows per warp costs occupancy and is slower)
constexpr int LN_ROWS = 1;
template <typename T>
__global__ void __launch_bounds__(256) layernorm_kernel(const T* __restrict__ x, int64_t rows, int C,
                                                       const float* __restrict__ gamma, const float* __restrict__ beta,
                                                       float eps, const float* __restrict__ scale,
                                                       const float* __restrict__ shift, int mod_ld, int rows_per_group,
                                                       T* __restrict__ out) {
  constexpr int VN = Vec<T>::N;
  constexpr int MAXV = 32 / VN;  // per-lane vectors: C <= 32 lanes * 32 elements = 1024
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int V = C / VN;
  const float inv_c = 1.f / (float)C;
  const T* xr = x + row * C;
  float buf[MAXV][VN];
  float s = 0.f;
#pragma unroll
  for (int j = 0; j < MAXV; ++j) {
    const int v = lane + j * 32;
    if (v < V) {
      Vec<T>::load(xr + v * VN, buf[j]);
#pragma unroll
      for (int i = 0; i < VN; ++i) s += buf[j][i];
    }
  }
  const float mean = warp_sum(s) * inv_c;
  float q = 0.f;
#pragma unroll
  for (int j = 0; j < MAXV; ++j) {
    const int v = lane + j * 32;
    if (v < V) {
#pragma unroll
      for (int i = 0; i < VN; ++i) { float d = buf[j][i] - mean; q = fmaf(d, d, q); }
    }
  }
  const float rstd = rsqrtf(warp_sum(q) * inv_c + eps);
  const int64_t grp = rows_per_group > 0 ? row / rows_per_group : 0;
  const float* sp = scale ? scale + grp * mod_ld : nullptr;
  const float* hp = scale ? shift + grp * mod_ld : nullptr;
  T* orow = out + row * C;
#pragma unroll
  for (int j = 0; j < MAXV; ++j) {
    const int v = lane + j * 32;
    if (v < V) {
      const int c = v * VN;
#pragma unroll
      for (int i = 0; i < VN; i += 4) {
        float y[4] = {(buf[j][i] - mean) * rstd, (buf[j][i + 1] - mean) * rstd, (buf[j][i + 2] - mean) * rstd,
                      (buf[j][i + 3] - mean) * rstd};
        if (gamma) {
          const float4 g4 = *reinterpret_cast<const float4*>(gamma + c + i);
          y[0] *= g4.x; y[1] *= g4.y; y[2] *= g4.z; y[3] *= g4.w;
        }
        if (beta) {
          const float4 b4 = *reinterpret_cast<const float4*>(beta + c + i);
          y[0] += b4.x; y[1] += b4.y; y[2] += b4.z; y[3] += b4.w;
        }
        if (sp) {
          const float4 s4 = *reinterpret_cast<const float4*>(sp + c + i);
          const float4 h4 = *reinterpret_cast<const float4*>(hp + c + i);
          y[0] = fmaf(y[0], 1.f + s4.x, h4.x); y[1] = fmaf(y[1], 1.f + s4.y, h4.y);
          y[2] = fmaf(y[2], 1.f + s4.z, h4.z); y[3] = fmaf(y[3], 1.f + s4.w, h4.w);
        }
        buf[j][i] = y[0]; buf[j][i + 1] = y[1]; buf[j][i + 2] = y[2]; buf[j][i + 3] = y[3];
      }
      Vec<T>::store(orow + c, buf[j]);
    }
  }
}


// bf16 rows: the row stays PACKED in registers (16-byte vectors, unpacked again in each of the three passes -- a shift per
// element) so that a warp can hold R rows in flight at 4 registers per vector instead of 8: the one-row fp32 form needs 79
// registers (3 blocks / SM, 36 KB of loads in flight per SM, and only while a warp is in its load phase: 2.9 TB/s = 45 %
// of the copy peak on the DiT adaLN shape).  Arithmetic and its order per row are those of layernorm_kernel: bit-identical.
// MEASURED (tools/ln_micro.py, GB/s of read + write, one box): fp32-register form 3 291 (DiT adaLN 262 144 x 768) / 3 155
// (U-Net affine 204 800 x 512) / 3 381 (51 200 x 1024); packed R = 1: 3 988 / 4 550 / 4 458; R = 2: 4 156 / 4 118 / 3 455;
// R = 4: 3 431 / 4 215 / 3 313 -> one row per warp, two for the three-vector rows (C in (512, 768]: the DiT width).
#ifdef DCB_LN_R
#define DCB_LN_ROWS(MAXV) DCB_LN_R
#else
#define DCB_LN_ROWS(MAXV) ((MAXV) == 3 ? 2 : 1)
#endif
template <int MAXV, int R>
__global__ void __launch_bounds__(256) layernorm_bf16_kernel(const __nv_bfloat16* __restrict__ x, int64_t rows, int C,
                                                            const float* __restrict__ gamma, const float* __restrict__ beta,
                                                            float eps, const float* __restrict__ scale,
                                                            const float* __restrict__ shift, int mod_ld, int rows_per_group,
                                                            __nv_bfloat16* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int64_t row0 = ((int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * R;
  if (row0 >= rows) return;
  const int V = C >> 3;
  const float inv_c = 1.f / (float)C;
  uint4 raw[R][MAXV];
#pragma unroll
  for (int r = 0; r < R; ++r) {
#pragma unroll
    for (int j = 0; j < MAXV; ++j) {
      const int v = lane + j * 32;
      raw[r][j] = make_uint4(0u, 0u, 0u, 0u);
      if (v < V && row0 + r < rows) raw[r][j] = *reinterpret_cast<const uint4*>(x + (row0 + r) * C + v * 8);
    }
  }
#pragma unroll
  for (int r = 0; r < R; ++r) {
    const int64_t row = row0 + r;
    if (row >= rows) break;
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < MAXV; ++j) {
      if (lane + j * 32 < V) {
        float f[8];
        unpack_bf16x8(raw[r][j], f);
#pragma unroll
        for (int i = 0; i < 8; ++i) s += f[i];
      }
    }
    const float mean = warp_sum(s) * inv_c;
    float q = 0.f;
#pragma unroll
    for (int j = 0; j < MAXV; ++j) {
      if (lane + j * 32 < V) {
        float f[8];
        unpack_bf16x8(raw[r][j], f);
#pragma unroll
        for (int i = 0; i < 8; ++i) { float d = f[i] - mean; q = fmaf(d, d, q); }
      }
    }
    const float rstd = rsqrtf(warp_sum(q) * inv_c + eps);
    const int64_t grp = rows_per_group > 0 ? row / rows_per_group : 0;
    const float* sp = scale ? scale + grp * mod_ld : nullptr;
    const float* hp = scale ? shift + grp * mod_ld : nullptr;
    __nv_bfloat16* orow = out + row * C;
#pragma unroll
    for (int j = 0; j < MAXV; ++j) {
      const int v = lane + j * 32;
      if (v < V) {
        const int c = v * 8;
        float f[8];
        unpack_bf16x8(raw[r][j], f);
#pragma unroll
        for (int i = 0; i < 8; i += 4) {
          float y[4] = {(f[i] - mean) * rstd, (f[i + 1] - mean) * rstd, (f[i + 2] - mean) * rstd, (f[i + 3] - mean) * rstd};
          if (gamma) {
            const float4 g4 = *reinterpret_cast<const float4*>(gamma + c + i);
            y[0] *= g4.x; y[1] *= g4.y; y[2] *= g4.z; y[3] *= g4.w;
          }
          if (beta) {
            const float4 b4 = *reinterpret_cast<const float4*>(beta + c + i);
            y[0] += b4.x; y[1] += b4.y; y[2] += b4.z; y[3] += b4.w;
          }
          if (sp) {
            const float4 s4 = *reinterpret_cast<const float4*>(sp + c + i);
            const float4 h4 = *reinterpret_cast<const float4*>(hp + c + i);
            y[0] = fmaf(y[0], 1.f + s4.x, h4.x); y[1] = fmaf(y[1], 1.f + s4.y, h4.y);
            y[2] = fmaf(y[2], 1.f + s4.z, h4.z); y[3] = fmaf(y[3], 1.f + s4.w, h4.w);
          }
          f[i] = y[0]; f[i + 1] = y[1]; f[i + 2] = y[2]; f[i + 3] = y[3];
        }
        *reinterpret_cast<uint4*>(orow + c) = pack_bf16x8(f);
      }
    }
  }
}

template <int MAXV>
static void launch_ln_bf16(const void* x, int64_t rows, int C, const float* gamma, const float* beta, float eps,
                           const float* scale, const float* shift, int mod_ld, int rows_per_group, void* out, cudaStream_t st) {
  constexpr int R = DCB_LN_ROWS(MAXV), WPB = 8;
  const int64_t warps = (rows + R - 1) / R;
  layernorm_bf16_kernel<MAXV, R><<<(unsigned)((warps + WPB - 1) / WPB), WPB * 32, 0, st>>>(
      (const __nv_bfloat16*)x, rows, C, gamma, beta, eps, scale, shift, mod_ld, rows_per_group, (__nv_bfloat16*)out);
}

}  // namespace dcb

using namespace dcb;

extern "C" int dcb_groupnorm_stats_div(int dtype, const void* x0, int C0, int div0, const void* x1, int C1, int div1,
                                       int NB, int HW, int G, int chunks, float* part, dcb_stream stream) {
  int rc = gn_check(dtype, C0, C1, G, x1);
  if (rc) return rc;
  DCB_REQUIRE(chunks >= 1 && NB >= 1 && NB <= 65535, "groupnorm: bad chunks/NB");
  if (div0 < 1) div0 = 1;
  if (div1 < 1) div1 = 1;
  cudaStream_t st = (cudaStream_t)stream;
  return dtype == DCB_BF16 ? gn_stats_t<__nv_bfloat16>(x0, C0, x1, C1, NB, HW, G, chunks, part, st, div0, div1)
                           : gn_stats_t<float>(x0, C0, x1, C1, NB, HW, G, chunks, part, st, div0, div1);
}

extern "C" int dcb_groupnorm_apply_div(int dtype, const void* x0, int C0, int div0, const void* x1, int C1, int div1,
                                       int NB, int HW, int G, int chunks, const float* part, const float* gamma,
                                       const float* beta, float eps, int silu, void* out, dcb_stream stream) {
  int rc = gn_check(dtype, C0, C1, G, x1);
  if (rc) return rc;
  DCB_REQUIRE(chunks >= 1 && NB >= 1 && NB <= 65535, "groupnorm: bad chunks/NB");
  if (div0 < 1) div0 = 1;
  if (div1 < 1) div1 = 1;
  cudaStream_t st = (cudaStream_t)stream;
  return dtype == DCB_BF16 ? gn_apply_t<__nv_bfloat16>(x0, C0, x1, C1, NB, HW, G, chunks, part, gamma, beta, eps, silu,
                                                       out, st, div0, div1)
                           : gn_apply_t<float>(x0, C0, x1, C1, NB, HW, G, chunks, part, gamma, beta, eps, silu, out, st,
                                               div0, div1);
}

extern "C" int dcb_groupnorm_fused(int dtype, const void* x0, int C0, int div0, const void* x1, int C1, int div1, int NB,
                                   int HW, int G, const float* gamma, const float* beta, float eps, int silu, void* out,
                                   dcb_stream stream) {
  int rc = gn_check(dtype, C0, C1, G, x1);
  if (rc) return rc;
  const int vn = dtype == DCB_BF16 ? 8 : 4, cpg = (C0 + C1) / G;
  DCB_REQUIRE(cpg % vn == 0 && C0 % cpg == 0, "groupnorm_fused: channels per group (%d) must be a multiple of %d and "
              "divide C0 (%d)", cpg, vn, C0);
  DCB_REQUIRE((C0 + C1) / vn <= GN_THREADS, "groupnorm_fused: at most %d channels", GN_THREADS * vn);
  DCB_REQUIRE(NB >= 1 && NB <= 65535 && HW >= 1, "groupnorm_fused: bad NB/HW");
  if (div0 < 1) div0 = 1;
  if (div1 < 1) div1 = 1;
  cudaStream_t st = (cudaStream_t)stream;
  const int V = (C0 + C1) / vn, threads = gn_threads(V);
  dim3 grid(1, NB);
  if (dtype == DCB_BF16)
    gn_apply_kernel<__nv_bfloat16, true><<<grid, threads, 0, st>>>((const __nv_bfloat16*)x0, C0, (const __nv_bfloat16*)x1, C1, HW,
                                                             G, 1, nullptr, gamma, beta, eps, silu, (__nv_bfloat16*)out,
                                                             HW, div0, div1);
  else
    gn_apply_kernel<float, true><<<grid, threads, 0, st>>>((const float*)x0, C0, (const float*)x1, C1, HW, G, 1, nullptr, gamma,
                                                     beta, eps, silu, (float*)out, HW, div0, div1);
  DCB_CHECK_LAUNCH("gn_fused");
  return DCB_OK;
}

extern "C" int dcb_groupnorm_stats_from_tiles(const float* part0, int C0, int div0, const float* part1, int C1, int div1,
                                              int NB, int tiles_per_sample, int G, float* part_out, dcb_stream stream) {
  DCB_REQUIRE(G > 0 && G <= 64 && (C0 + C1) % G == 0, "groupnorm: C=%d not divisible by G=%d (G<=64)", C0 + C1, G);
  DCB_REQUIRE(part0 != nullptr && (C1 == 0) == (part1 == nullptr) && tiles_per_sample >= 1 && NB >= 1,
              "groupnorm_stats_from_tiles: bad arguments");
  gn_tiles_finalize_kernel<<<NB, 8 * G, 0, (cudaStream_t)stream>>>(part0, C0, div0 < 1 ? 1 : div0, part1, C1,
                                                                  div1 < 1 ? 1 : div1, tiles_per_sample, G, part_out, 0,
                                                                  nullptr, nullptr, 0.f, nullptr, nullptr);
  DCB_CHECK_LAUNCH("gn_tiles_finalize");
  return DCB_OK;
}

extern "C" int dcb_groupnorm_coef_from_tiles(const float* part0, int C0, int div0, const float* part1, int C1, int div1,
                                             int NB, int tiles_per_sample, int G, const float* gamma, const float* beta,
                                             float eps, float* coef_a, float* coef_b, dcb_stream stream) {
  DCB_REQUIRE(G > 0 && G <= 64 && (C0 + C1) % G == 0, "groupnorm: C=%d not divisible by G=%d (G<=64)", C0 + C1, G);
  DCB_REQUIRE(part0 != nullptr && (C1 == 0) == (part1 == nullptr) && tiles_per_sample >= 1 && NB >= 1 &&
                  gamma != nullptr && beta != nullptr && coef_a != nullptr && coef_b != nullptr,
              "groupnorm_coef_from_tiles: bad arguments");
  gn_tiles_finalize_kernel<<<NB, 8 * G, 0, (cudaStream_t)stream>>>(part0, C0, div0 < 1 ? 1 : div0, part1, C1,
                                                                  div1 < 1 ? 1 : div1, tiles_per_sample, G, nullptr,
                                                                  tiles_per_sample * 128, gamma, beta, eps, coef_a, coef_b);
  DCB_CHECK_LAUNCH("gn_tiles_finalize_coef");
  return DCB_OK;
}

extern "C" int dcb_groupnorm_stats(int dtype, const void* x0, int C0, const void* x1, int C1, int NB, int HW, int G,
                                   int chunks, float* part, dcb_stream stream) {
  return dcb_groupnorm_stats_div(dtype, x0, C0, 1, x1, C1, 1, NB, HW, G, chunks, part, stream);
}

extern "C" int dcb_groupnorm_apply(int dtype, const void* x0, int C0, const void* x1, int C1, int NB, int HW, int G,
                                   int chunks, const float* part, const float* gamma, const float* beta, float eps,
                                   int silu, void* out, dcb_stream stream) {
  return dcb_groupnorm_apply_div(dtype, x0, C0, 1, x1, C1, 1, NB, HW, G, chunks, part, gamma, beta, eps, silu, out, stream);
}

extern "C" int dcb_layernorm(int dtype, const void* x, int64_t rows, int C, const float* gamma, const float* beta,
                             float eps, const float* scale, const float* shift, int mod_ld, int rows_per_group, void* out,
                             dcb_stream stream) {
  const int vn = dtype == DCB_BF16 ? 8 : 4;
  DCB_REQUIRE(C % vn == 0 && C <= 1024, "layernorm: need C %% %d == 0 and C <= 1024 (C=%d)", vn, C);
  DCB_REQUIRE((scale == nullptr) == (shift == nullptr), "layernorm: scale/shift must come together");
  DCB_REQUIRE((((uintptr_t)gamma | (uintptr_t)beta | (uintptr_t)scale | (uintptr_t)shift) & 15) == 0 &&
                  (scale == nullptr || mod_ld % 4 == 0),
              "layernorm: gamma/beta/scale/shift must be 16-byte aligned (mod_ld %% 4 == 0)");
  cudaStream_t st = (cudaStream_t)stream;
  const int wpb = 8;
  const int64_t warps = (rows + LN_ROWS - 1) / LN_ROWS;
  const unsigned grid = (unsigned)((warps + wpb - 1) / wpb);
#ifdef DCB_LN_OLD   // A/B only: the one-row fp32-register form on bf16 data
  if (dtype == DCB_BF16) {
    layernorm_kernel<__nv_bfloat16><<<grid, wpb * 32, 0, st>>>((const __nv_bfloat16*)x, rows, C, gamma, beta, eps, scale,
                                                               shift, mod_ld, rows_per_group, (__nv_bfloat16*)out);
  } else
#endif
  if (dtype == DCB_BF16) {
    switch ((C / 8 + 31) / 32) {
      case 1: launch_ln_bf16<1>(x, rows, C, gamma, beta, eps, scale, shift, mod_ld, rows_per_group, out, st); break;
      case 2: launch_ln_bf16<2>(x, rows, C, gamma, beta, eps, scale, shift, mod_ld, rows_per_group, out, st); break;
      case 3: launch_ln_bf16<3>(x, rows, C, gamma, beta, eps, scale, shift, mod_ld, rows_per_group, out, st); break;
      default: launch_ln_bf16<4>(x, rows, C, gamma, beta, eps, scale, shift, mod_ld, rows_per_group, out, st); break;
    }
  } else
    layernorm_kernel<float><<<grid, wpb * 32, 0, st>>>((const float*)x, rows, C, gamma, beta, eps, scale, shift, mod_ld,
                                                       rows_per_group, (float*)out);
  DCB_CHECK_LAUNCH("layernorm");
  return DCB_OK;
}
