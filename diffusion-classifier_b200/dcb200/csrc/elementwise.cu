// elementwise.cu -- HBM-bound helpers around the denoiser: fused q_sample prologue, sinusoidal noise-label
// embedding, eps-MSE reductions, nearest-2x upsample, Haar DWT/IDWT, layout conversions.
#include "common.cuh"

namespace dcb {

// ---- Philox4x32-10 + Box-Muller (throughput mode: eps is generated in-kernel) -----------------------------
__device__ __forceinline__ void philox_round(uint32_t (&c)[4], uint32_t (&k)[2]) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u;
  uint32_t hi0 = __umulhi(M0, c[0]), lo0 = M0 * c[0];
  uint32_t hi1 = __umulhi(M1, c[2]), lo1 = M1 * c[2];
  uint32_t n0 = hi1 ^ c[1] ^ k[0], n1 = lo1, n2 = hi0 ^ c[3] ^ k[1], n3 = lo0;
  c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
  k[0] += 0x9E3779B9u; k[1] += 0xBB67AE85u;
}
__device__ __forceinline__ float philox_normal(uint64_t seed, uint64_t unit, uint64_t elem) {
  uint32_t c[4] = {(uint32_t)elem, (uint32_t)(elem >> 32), (uint32_t)unit, (uint32_t)(unit >> 32)};
  uint32_t k[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
#pragma unroll
  for (int i = 0; i < 10; ++i) philox_round(c, k);
  const float u1 = ((float)(c[0] >> 8) + 0.5f) * (1.0f / 16777216.0f);  // (0,1)
  const float u2 = ((float)(c[1] >> 8) + 0.5f) * (1.0f / 16777216.0f);
  return sqrtf(-2.0f * logf(u1)) * cospif(2.0f * u2);
}

// z = alpha*x + sigma*eps  (diffusion_classifier.py:113-115); target = eps - (v ? sigma*z : 0)
// x/eps NCHW fp32 -> z_ws/target NHWC fp32.  One thread per element, x fastest (coalesced NCHW reads).
__global__ void qsample_kernel(const float* __restrict__ x, const float* __restrict__ eps_pre, uint64_t seed,
                               int64_t unit_id0, const float* __restrict__ alpha, const float* __restrict__ sigma,
                               const int* __restrict__ img, int U, int C, int HW, int W, int patch,
                               float* __restrict__ z_ws, float* __restrict__ target, int v_param) {
  const int64_t total = (int64_t)U * C * HW;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int pix = (int)(i % HW);
    const int c = (int)((i / HW) % C);
    const int u = (int)(i / ((int64_t)HW * C));
    const int b = img ? img[u] : u;
    const float xv = x[((int64_t)b * C + c) * HW + pix];
    const float a = alpha ? alpha[u] : 1.f;
    float z, e = 0.f;
    if (sigma) {
      const int64_t ei = ((int64_t)u * C + c) * HW + pix;
      e = eps_pre ? eps_pre[ei] : philox_normal(seed, (uint64_t)(unit_id0 + u), (uint64_t)((int64_t)c * HW + pix));
      z = a * xv + sigma[u] * e;
    } else {
      z = a * xv;
    }
    const int64_t o = ((int64_t)u * HW + pix) * C + c;
    z_ws[o] = z;
    if (target) {
      int64_t to = o;
      if (patch > 0) {  // token-major (py,px,c) layout matching DiT proj_out_2's output columns
        const int y = pix / W, xx = pix % W, g = W / patch;
        const int64_t tok = (int64_t)(y / patch) * g + xx / patch;
        to = ((int64_t)u * (HW / (patch * patch)) + tok) * (patch * patch * C) + ((y % patch) * patch + xx % patch) * C + c;
      }
      target[to] = v_param ? e - sigma[u] * z : e;
    }
  }
}

// One ancestral DDPM step with classifier-free guidance (diffusion_classifier.py:176-207 + the z_s draw of :263-264):
//   pred = (1+w)*cond - w*uncond ; x = clip(v ? a_t z - s_t pred : (z - s_t pred)/a_t) ; mu = a_s (z (1-c)/a_t + c x)
//   z_out = final ? clip(mu) : mu + sqrt(var) * noise
// z NCHW fp32; pred is read in the layout the last GEMM of the denoiser writes (NHWC, or DiT token-major (py,px,c) when
// patch > 0), rows of sample b*rep (+1 = unconditional).  coef[8] = {c, a_t, a_s, s_t, s_s, sqrt(var), w, step index (int bits)}.
__global__ void ddpm_step_kernel(const float* __restrict__ z_t, const float* __restrict__ pred, int rep, int patch,
                                 const float* __restrict__ coef, int v_param, int final_step,
                                 const float* __restrict__ noise_pre, uint64_t seed, int64_t unit_id0, int B, int C, int HW,
                                 int W, float* __restrict__ z_out) {
  const float c = coef[0], a_t = coef[1], a_s = coef[2], s_t = coef[3], sd = coef[5], w = coef[6];
  // coef[7] carries the step index as an int bit pattern: the Philox unit of image b at step i is unit_id0 + i*B + b, so a
  // CUDA graph of the step (fixed kernel arguments) still draws fresh noise every replay
  const int64_t unit0 = unit_id0 + (int64_t)__float_as_int(coef[7]) * B;
  const int64_t total = (int64_t)B * C * HW;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int pix = (int)(i % HW);
    const int ch = (int)((i / HW) % C);
    const int b = (int)(i / ((int64_t)HW * C));
    int64_t off, per_sample;
    if (patch > 0) {
      const int y = pix / W, xx = pix % W, g = W / patch;
      const int64_t tok = (int64_t)(y / patch) * g + xx / patch;
      off = tok * (patch * patch * C) + ((y % patch) * patch + xx % patch) * C + ch;
      per_sample = (int64_t)HW * C;
    } else {
      off = (int64_t)pix * C + ch;
      per_sample = (int64_t)HW * C;
    }
    float p = pred[(int64_t)b * rep * per_sample + off];
    if (rep == 2) p = (1.f + w) * p - w * pred[((int64_t)b * rep + 1) * per_sample + off];
    const float z = z_t[i];
    float x = v_param ? a_t * z - s_t * p : (z - s_t * p) / a_t;
    x = fminf(fmaxf(x, -1.f), 1.f);
    const float mu = a_s * (z * (1.f - c) / a_t + c * x);
    float o;
    if (final_step) o = fminf(fmaxf(mu, -1.f), 1.f);
    else {
      const float n = noise_pre ? noise_pre[i] : philox_normal(seed, (uint64_t)(unit0 + b), (uint64_t)((int64_t)ch * HW + pix));
      o = mu + n * sd;
    }
    z_out[i] = o;
  }
}

// stage the first layer's A operand from z_ws (NHWC fp32): 3x3 unfold (mode 0) or p x p patchify (mode 1)
template <typename T>
__global__ void stage_kernel(int mode, const float* __restrict__ z, int U, int rep, int C, int H, int W, int patch,
                             int kpad, T* __restrict__ a_out) {
  const int g = mode == 1 ? W / patch : 0;
  const int rows = mode == 0 ? H * W : g * (H / patch);
  const int kv = kpad / 8;
  const int64_t total = (int64_t)U * rep * rows * kv;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int k0 = (int)(i % kv) * 8;
    const int64_t rr = i / kv;
    const int row = (int)(rr % rows);
    const int u = (int)(rr / rows / rep);
    float f[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int k = k0 + j;
      float v = 0.f;
      if (mode == 0) {
        if (k < 9 * C) {
          const int tap = k / C, c = k - tap * C;
          const int y = row / W + tap / 3 - 1, xx = row % W + tap % 3 - 1;
          if (y >= 0 && y < H && xx >= 0 && xx < W) v = z[(((int64_t)u * H + y) * W + xx) * C + c];
        }
      } else {
        if (k < patch * patch * C) {
          const int pq = k / C, c = k - pq * C;
          const int y = (row / g) * patch + pq / patch, xx = (row % g) * patch + pq % patch;
          v = z[(((int64_t)u * H + y) * W + xx) * C + c];
        }
      }
      f[j] = v;
    }
    T* o = a_out + rr * kpad + k0;
#pragma unroll
    for (int j = 0; j < 8; ++j) o[j] = from_f<T>(f[j]);
  }
}

template <typename T>
__global__ void timestep_embed_kernel(const float* __restrict__ t, int U, int rep, int dim, float shift,
                                      float max_period, T* __restrict__ out) {
  const int half = dim / 2;
  const int64_t total = (int64_t)U * rep * half;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int k = (int)(i % half);
    const int64_t s = i / half;
    const int u = (int)(s / rep);
    const float w = expf(-logf(max_period) * (float)k / ((float)half - shift));
    const float arg = t[u] * w;
    out[s * dim + k] = from_f<T>(cosf(arg));          // flip_sin_to_cos=True layout: [cos | sin]
    out[s * dim + half + k] = from_f<T>(sinf(arg));
  }
}

// err[s] (+)= sum of consecutive partials -- fixed order
__global__ void mse_finalize_kernel(const float* __restrict__ part, int pps, int S, float* __restrict__ err,
                                    int err_stride) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= S) return;
  float a = 0.f;
  for (int i = 0; i < pps; ++i) a += part[(int64_t)s * pps + i];
  err[(int64_t)s * err_stride] = a;
}

template <typename T>
__global__ void __launch_bounds__(1024) eps_mse_kernel(const T* __restrict__ pred, const float* __restrict__ target,
                                                      const float* __restrict__ scale, int div, int64_t K,
                                                      float* __restrict__ err, int err_stride) {
  __shared__ float red[32];
  const int s = blockIdx.x;
  const float sc = scale ? scale[s] : 1.f;
  const T* p = pred + (int64_t)s * K;
  const float* tg = target + (int64_t)(s / div) * K;
  float a = 0.f;
  for (int64_t i = threadIdx.x; i < K; i += blockDim.x) {
    const float d = sc * to_f<T>(p[i]) - tg[i];
    a = fmaf(d, d, a);
  }
  a = warp_sum(a);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = a;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t += red[i];
    err[(int64_t)s * err_stride] = t;
  }
}

template <typename T>
__global__ void upsample2x_kernel(const T* __restrict__ x, int NB, int H, int W, int C, T* __restrict__ out) {
  constexpr int VN = sizeof(T) == 2 ? 8 : 4;
  const int V = C / VN;
  const int64_t total = (int64_t)NB * (2 * H) * (2 * W) * V;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int v = (int)(i % V);
    int64_t r = i / V;
    const int ox = (int)(r % (2 * W));
    r /= 2 * W;
    const int oy = (int)(r % (2 * H));
    const int n = (int)(r / (2 * H));
    const uint4 val = *reinterpret_cast<const uint4*>(x + (((int64_t)n * H + oy / 2) * W + ox / 2) * C + v * VN);
    *reinterpret_cast<uint4*>(out + (((int64_t)n * 2 * H + oy) * 2 * W + ox) * C + v * VN) = val;
  }
}

// out[n] = x[n / div]: 16-byte vectors, grid-stride
__global__ void expand_samples_kernel(const uint4* __restrict__ x, int64_t vec_per_sample, int64_t total, int div,
                                      uint4* __restrict__ out) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t n = i / vec_per_sample, r = i - n * vec_per_sample;
    out[i] = x[(n / div) * vec_per_sample + r];
  }
}

// Haar analysis of one 2x2 block: cA=(a+b+c+d)/2, cH=(a+b-c-d)/2, cV=(a-b+c-d)/2, cD=(a-b-c+d)/2
__global__ void haar_dwt_kernel(const float* __restrict__ x, int B, int C, int H, int W, float post, float* __restrict__ out) {
  const int h = H / 2, w = W / 2;
  const int64_t total = (int64_t)B * C * h * w;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int xx = (int)(i % w);
    const int yy = (int)((i / w) % h);
    const int c = (int)((i / ((int64_t)w * h)) % C);
    const int b = (int)(i / ((int64_t)w * h * C));
    const float* p = x + (((int64_t)b * C + c) * H + 2 * yy) * W + 2 * xx;
    const float2 r0 = *reinterpret_cast<const float2*>(p);
    const float2 r1 = *reinterpret_cast<const float2*>(p + W);
    const float a = r0.x, bb = r0.y, cc = r1.x, d = r1.y;
    float* o = out + (((int64_t)b * 4 * C + 4 * c) * h + yy) * w + xx;
    const int64_t cs = (int64_t)h * w;
    const float hs = 0.5f * post;
    o[0] = (a + bb + cc + d) * hs;
    o[cs] = (a + bb - cc - d) * hs;
    o[2 * cs] = (a - bb + cc - d) * hs;
    o[3 * cs] = (a - bb - cc + d) * hs;
  }
}

__global__ void haar_idwt_kernel(const float* __restrict__ wv, int B, int C4, int h, int w, float pre, float* __restrict__ out) {
  const int C = C4 / 4;
  const int64_t total = (int64_t)B * C * h * w;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int xx = (int)(i % w);
    const int yy = (int)((i / w) % h);
    const int c = (int)((i / ((int64_t)w * h)) % C);
    const int b = (int)(i / ((int64_t)w * h * C));
    const int64_t cs = (int64_t)h * w;
    const float* p = wv + (((int64_t)b * C4 + 4 * c) * h + yy) * w + xx;
    const float hs = 0.5f * pre;
    const float A = p[0] * hs, Hh = p[cs] * hs, V = p[2 * cs] * hs, D = p[3 * cs] * hs;
    float* o = out + (((int64_t)b * C + c) * 2 * h + 2 * yy) * 2 * w + 2 * xx;
    *reinterpret_cast<float2*>(o) = make_float2(A + Hh + V + D, A + Hh - V - D);
    *reinterpret_cast<float2*>(o + 2 * w) = make_float2(A - Hh + V - D, A - Hh - V + D);
  }
}

template <typename T>
__global__ void nhwc_to_nchw_kernel(const T* __restrict__ x, int NB, int HW, int C, int ld, float* __restrict__ out) {
  const int64_t total = (int64_t)NB * C * HW;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int pix = (int)(i % HW);
    const int c = (int)((i / HW) % C);
    const int64_t n = i / ((int64_t)HW * C);
    out[i] = to_f<T>(x[(n * HW + pix) * ld + c]);
  }
}

// "nhwpqc->nchpwq": out[n][c][gy*p+py][gx*p+px] = tok[n][gy*g+gx][(py*p+px)*C + c]
template <typename T>
__global__ void unpatchify_kernel(const T* __restrict__ tok, int B, int g, int p, int C, int ld, float* __restrict__ out) {
  const int S = g * p;
  const int64_t total = (int64_t)B * C * S * S;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int xx = (int)(i % S);
    const int yy = (int)((i / S) % S);
    const int c = (int)((i / ((int64_t)S * S)) % C);
    const int64_t n = i / ((int64_t)S * S * C);
    const int64_t row = n * g * g + (int64_t)(yy / p) * g + xx / p;
    out[i] = to_f<T>(tok[row * ld + ((yy % p) * p + xx % p) * C + c]);
  }
}

template <typename T>
__global__ void cast_kernel(const float* __restrict__ src, int64_t n, T* __restrict__ dst) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    dst[i] = from_f<T>(src[i]);
}

static unsigned grid_for(int64_t total, int threads = 256) {
  int64_t b = (total + threads - 1) / threads;
  const int64_t cap = (int64_t)num_sms() * 16;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (unsigned)b;
}

}  // namespace dcb

using namespace dcb;

extern "C" int dcb_prologue(int mode, int dtype, const float* x, const float* eps_predrawn, uint64_t seed, int64_t unit_id0,
                            const float* alpha, const float* sigma, const int32_t* img, int U, int rep, int C, int H, int W,
                            int patch, int kpad, float* z_ws, void* a_out, float* target, int v_param, dcb_stream stream) {
  DCB_REQUIRE(mode == 0 || mode == 1, "prologue: mode must be 0 (3x3 unfold) or 1 (patchify)");
  DCB_REQUIRE(kpad % 8 == 0 && kpad >= (mode == 0 ? 9 * C : patch * patch * C), "prologue: kpad too small / not %%8");
  DCB_REQUIRE(mode == 0 || (H % patch == 0 && W % patch == 0), "prologue: H,W must be multiples of patch");
  DCB_REQUIRE(target == nullptr || sigma != nullptr, "prologue: target needs noise (sigma)");
  cudaStream_t st = (cudaStream_t)stream;
  const int HW = H * W;
  qsample_kernel<<<grid_for((int64_t)U * C * HW), 256, 0, st>>>(x, eps_predrawn, seed, unit_id0, alpha, sigma, img, U, C,
                                                                 HW, W, mode == 1 ? patch : 0, z_ws, target, v_param);
  DCB_CHECK_LAUNCH("qsample");
  if (a_out) {
    const int rows = mode == 0 ? HW : (H / patch) * (W / patch);
    const int64_t total = (int64_t)U * rep * rows * (kpad / 8);
    if (dtype == DCB_BF16)
      stage_kernel<__nv_bfloat16><<<grid_for(total), 256, 0, st>>>(mode, z_ws, U, rep, C, H, W, patch, kpad,
                                                                   (__nv_bfloat16*)a_out);
    else
      stage_kernel<float><<<grid_for(total), 256, 0, st>>>(mode, z_ws, U, rep, C, H, W, patch, kpad, (float*)a_out);
    DCB_CHECK_LAUNCH("stage");
  }
  return DCB_OK;
}

extern "C" int dcb_ddpm_step(const float* z_t, const float* pred, int rep, int patch, const float* coef, int v_param,
                             int final_step, const float* noise_predrawn, uint64_t seed, int64_t unit_id0, int B, int C,
                             int H, int W, float* z_out, dcb_stream stream) {
  DCB_REQUIRE(rep == 1 || rep == 2, "ddpm_step: rep must be 1 (no guidance pair) or 2 (cond, uncond)");
  DCB_REQUIRE(patch == 0 || (H % patch == 0 && W % patch == 0), "ddpm_step: H,W must be multiples of patch");
  DCB_REQUIRE(z_t && pred && coef && z_out, "ddpm_step: null pointer");
  ddpm_step_kernel<<<grid_for((int64_t)B * C * H * W), 256, 0, (cudaStream_t)stream>>>(
      z_t, pred, rep, patch, coef, v_param, final_step, noise_predrawn, seed, unit_id0, B, C, H * W, W, z_out);
  DCB_CHECK_LAUNCH("ddpm_step");
  return DCB_OK;
}

extern "C" int dcb_timestep_embed(int dtype, const float* t, int U, int rep, int dim, float shift, float max_period,
                                  void* out, dcb_stream stream) {
  DCB_REQUIRE(dim % 2 == 0, "timestep_embed: dim must be even");
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t total = (int64_t)U * rep * (dim / 2);
  if (dtype == DCB_BF16)
    timestep_embed_kernel<__nv_bfloat16><<<grid_for(total), 256, 0, st>>>(t, U, rep, dim, shift, max_period,
                                                                         (__nv_bfloat16*)out);
  else
    timestep_embed_kernel<float><<<grid_for(total), 256, 0, st>>>(t, U, rep, dim, shift, max_period, (float*)out);
  DCB_CHECK_LAUNCH("timestep_embed");
  return DCB_OK;
}

extern "C" int dcb_mse_finalize(const float* part, int parts_per_sample, int S, float* err, int err_stride,
                                dcb_stream stream) {
  mse_finalize_kernel<<<(S + 127) / 128, 128, 0, (cudaStream_t)stream>>>(part, parts_per_sample, S, err, err_stride);
  DCB_CHECK_LAUNCH("mse_finalize");
  return DCB_OK;
}

extern "C" int dcb_eps_mse(int dtype, const void* pred, const float* target, const float* scale, int S, int div, int64_t K,
                           float* err, int err_stride, dcb_stream stream) {
  DCB_REQUIRE(div >= 1, "eps_mse: div must be >= 1");
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == DCB_BF16)
    eps_mse_kernel<__nv_bfloat16><<<S, 1024, 0, st>>>((const __nv_bfloat16*)pred, target, scale, div, K, err, err_stride);
  else
    eps_mse_kernel<float><<<S, 1024, 0, st>>>((const float*)pred, target, scale, div, K, err, err_stride);
  DCB_CHECK_LAUNCH("eps_mse");
  return DCB_OK;
}

extern "C" int dcb_upsample2x(int dtype, const void* x, int NB, int H, int W, int C, void* out, dcb_stream stream) {
  const int vn = dtype == DCB_BF16 ? 8 : 4;
  DCB_REQUIRE(C % vn == 0, "upsample2x: C %% %d", vn);
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t total = (int64_t)NB * 4 * H * W * (C / vn);
  if (dtype == DCB_BF16)
    upsample2x_kernel<__nv_bfloat16><<<grid_for(total), 256, 0, st>>>((const __nv_bfloat16*)x, NB, H, W, C,
                                                                     (__nv_bfloat16*)out);
  else
    upsample2x_kernel<float><<<grid_for(total), 256, 0, st>>>((const float*)x, NB, H, W, C, (float*)out);
  DCB_CHECK_LAUNCH("upsample2x");
  return DCB_OK;
}

extern "C" int dcb_expand_samples(int dtype, const void* x, int NB, int div, int64_t elems_per_sample, void* out,
                                  dcb_stream stream) {
  const int64_t bytes = elems_per_sample * (dtype == DCB_BF16 ? 2 : 4);
  DCB_REQUIRE(bytes % 16 == 0 && div >= 1 && NB >= 1, "expand_samples: sample size must be a multiple of 16 bytes");
  const int64_t vps = bytes / 16, total = vps * NB;
  expand_samples_kernel<<<grid_for(total), 256, 0, (cudaStream_t)stream>>>((const uint4*)x, vps, total, div, (uint4*)out);
  DCB_CHECK_LAUNCH("expand_samples");
  return DCB_OK;
}

extern "C" int dcb_haar_dwt(const float* x, int B, int C, int H, int W, float post_scale, float* out, dcb_stream stream) {
  DCB_REQUIRE(H % 2 == 0 && W % 2 == 0, "haar_dwt: even H, W only (as every reference config)");
  haar_dwt_kernel<<<grid_for((int64_t)B * C * (H / 2) * (W / 2)), 256, 0, (cudaStream_t)stream>>>(x, B, C, H, W,
                                                                                                    post_scale, out);
  DCB_CHECK_LAUNCH("haar_dwt");
  return DCB_OK;
}

extern "C" int dcb_haar_idwt(const float* w, int B, int C4, int h, int wd, float pre_scale, float* out, dcb_stream stream) {
  DCB_REQUIRE(C4 % 4 == 0, "haar_idwt: channel count must be a multiple of 4");
  haar_idwt_kernel<<<grid_for((int64_t)B * (C4 / 4) * h * wd), 256, 0, (cudaStream_t)stream>>>(w, B, C4, h, wd, pre_scale,
                                                                                                 out);
  DCB_CHECK_LAUNCH("haar_idwt");
  return DCB_OK;
}

extern "C" int dcb_nhwc_to_nchw(int dtype, const void* x, int NB, int HW, int C, int ld, float* out, dcb_stream stream) {
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t total = (int64_t)NB * C * HW;
  if (dtype == DCB_BF16)
    nhwc_to_nchw_kernel<__nv_bfloat16><<<grid_for(total), 256, 0, st>>>((const __nv_bfloat16*)x, NB, HW, C, ld, out);
  else
    nhwc_to_nchw_kernel<float><<<grid_for(total), 256, 0, st>>>((const float*)x, NB, HW, C, ld, out);
  DCB_CHECK_LAUNCH("nhwc_to_nchw");
  return DCB_OK;
}

extern "C" int dcb_unpatchify(int dtype, const void* tok, int B, int g, int p, int C, int ld, float* out, dcb_stream stream) {
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t total = (int64_t)B * C * g * p * g * p;
  if (dtype == DCB_BF16)
    unpatchify_kernel<__nv_bfloat16><<<grid_for(total), 256, 0, st>>>((const __nv_bfloat16*)tok, B, g, p, C, ld, out);
  else
    unpatchify_kernel<float><<<grid_for(total), 256, 0, st>>>((const float*)tok, B, g, p, C, ld, out);
  DCB_CHECK_LAUNCH("unpatchify");
  return DCB_OK;
}

extern "C" int dcb_cast_f32(int dtype, const float* src, int64_t n, void* dst, dcb_stream stream) {
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == DCB_BF16) cast_kernel<__nv_bfloat16><<<grid_for(n), 256, 0, st>>>(src, n, (__nv_bfloat16*)dst);
  else cast_kernel<float><<<grid_for(n), 256, 0, st>>>(src, n, (float*)dst);
  DCB_CHECK_LAUNCH("cast_f32");
  return DCB_OK;
}
