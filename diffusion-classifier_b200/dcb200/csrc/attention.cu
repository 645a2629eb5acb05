// attention.cu -- flash-style self-attention softmax(Q K^T * scale) V for the U-Net Transformer2D blocks
// (8 heads, d = C/8 in {32,64,96,128}, N = H*W tokens) and the DiT blocks (12 heads x 64, N = 4096).
// diffusers AttnProcessor2_0 -> F.scaled_dot_product_attention.
//
//   bf16 path : one CTA = 64 queries x one (batch, head); K/V streamed in 64-key tiles through a cp.async double
//               buffer; QK^T and PV on tensor cores (mma.sync.m16n8k16 bf16, fp32 accumulate) with the online-softmax
//               rescale in registers; the N x N score matrix never exists in HBM.  [round-1 engine; the tcgen05/TMEM
//               version is the next step for the DiT N=4096 case, see DESIGN.md]
//   fp32 path : CUDA-core verify engine (4 threads per query), used by the fp32-verify mode and to check the bf16 one.
#include "common.cuh"

namespace dcb {

constexpr int FA_BM = 64, FA_BN = 64, FA_THREADS = 128;

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, int src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void mma_bf16(float* c, const uint32_t* a, uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// load a [64 rows x D] bf16 tile (rows = tokens t0..t0+63 of one head) into padded smem; rows >= N are zero-filled
template <int D>
__device__ __forceinline__ void load_tile(uint32_t smem_base, const __nv_bfloat16* g, int64_t ld, int t0, int N) {
  constexpr int CH = D / 8;           // 16-byte chunks per row
  constexpr int STRIDE = (D + 8) * 2; // padded row stride in bytes (conflict-free ldmatrix)
  for (int i = threadIdx.x; i < 64 * CH; i += FA_THREADS) {
    const int r = i / CH, c = i % CH;
    const int t = t0 + r;
    const bool ok = t < N;
    const __nv_bfloat16* src = g + (int64_t)(ok ? t : 0) * ld + c * 8;
    cp_async16(smem_base + r * STRIDE + c * 16, src, ok ? 16 : 0);
  }
}

template <int D>
__global__ void __launch_bounds__(FA_THREADS) flash_attn_bf16_kernel(const __nv_bfloat16* __restrict__ q,
                                                                    const __nv_bfloat16* __restrict__ k,
                                                                    const __nv_bfloat16* __restrict__ v, int ld, int N,
                                                                    float scale_log2e, __nv_bfloat16* __restrict__ out,
                                                                    int out_ld) {
  constexpr int STRIDE = (D + 8) * 2;
  constexpr int TILE = 64 * STRIDE;
  extern __shared__ __align__(16) uint8_t fa_smem[];
  const uint32_t sQ = (uint32_t)__cvta_generic_to_shared(fa_smem);
  const uint32_t sK = sQ + TILE;       // 2 buffers
  const uint32_t sV = sK + 2 * TILE;   // 2 buffers

  const int b = blockIdx.z, h = blockIdx.y, q0 = blockIdx.x * FA_BM;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t base = (int64_t)b * N * ld + (int64_t)h * D;
  const __nv_bfloat16* qg = q + base;
  const __nv_bfloat16* kg = k + base;
  const __nv_bfloat16* vg = v + base;

  load_tile<D>(sQ, qg, ld, q0, N);
  load_tile<D>(sK, kg, ld, 0, N);
  load_tile<D>(sV, vg, ld, 0, N);
  cp_async_commit();

  const int ntiles = (N + FA_BN - 1) / FA_BN;
  uint32_t qf[D / 16][4];
  float o[D / 8][4];
#pragma unroll
  for (int i = 0; i < D / 8; ++i) o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f;
  float m_run[2] = {-INFINITY, -INFINITY}, l_run[2] = {0.f, 0.f};

  const int lm = lane >> 3, lr = lane & 7;  // ldmatrix: matrix id / row within matrix

  for (int t = 0; t < ntiles; ++t) {
    const int buf = t & 1;
    if (t + 1 < ntiles) {
      load_tile<D>(sK + (buf ^ 1) * TILE, kg, ld, (t + 1) * FA_BN, N);
      load_tile<D>(sV + (buf ^ 1) * TILE, vg, ld, (t + 1) * FA_BN, N);
      cp_async_commit();
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    if (t == 0) {
#pragma unroll
      for (int kk = 0; kk < D / 16; ++kk) {
        const int row = warp * 16 + (lm & 1) * 8 + lr, col = kk * 16 + (lm >> 1) * 8;
        ldsm_x4(sQ + row * STRIDE + col * 2, qf[kk][0], qf[kk][1], qf[kk][2], qf[kk][3]);
      }
    }
    // ---- S = Q K^T (16 x 64 per warp) ----
    float s[8][4];
#pragma unroll
    for (int j = 0; j < 8; ++j) s[j][0] = s[j][1] = s[j][2] = s[j][3] = 0.f;
    const uint32_t kb = sK + buf * TILE;
#pragma unroll
    for (int kk = 0; kk < D / 16; ++kk) {
#pragma unroll
      for (int j = 0; j < 8; j += 2) {
        uint32_t b0, b1, b2, b3;
        const int row = j * 8 + (lm >> 1) * 8 + lr, col = kk * 16 + (lm & 1) * 8;
        ldsm_x4(kb + row * STRIDE + col * 2, b0, b1, b2, b3);
        mma_bf16(s[j], qf[kk], b0, b1);
        mma_bf16(s[j + 1], qf[kk], b2, b3);
      }
    }
    // ---- online softmax (rows g = lane/4 and g+8) ----
    const int key0 = t * FA_BN + (lane & 3) * 2;
    float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
    for (int j = 0; j < 8; ++j) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int key = key0 + j * 8 + (i & 1);
        float val = s[j][i] * scale_log2e;
        if (key >= N) val = -INFINITY;
        s[j][i] = val;
        mx[i >> 1] = fmaxf(mx[i >> 1], val);
      }
    }
    float corr[2], m_new[2];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 1));
      mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 2));
      m_new[r] = fmaxf(m_run[r], mx[r]);  // finite: every tile holds at least one valid key
      corr[r] = exp2f(m_run[r] - m_new[r]);
      m_run[r] = m_new[r];
    }
    float rs[2] = {0.f, 0.f};
    uint32_t pf[4][4];  // P as bf16 A fragments, one per 16-key step
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float p0 = exp2f(s[j][0] - m_new[0]), p1 = exp2f(s[j][1] - m_new[0]);
      const float p2 = exp2f(s[j][2] - m_new[1]), p3 = exp2f(s[j][3] - m_new[1]);
      rs[0] += p0 + p1;
      rs[1] += p2 + p3;
      pf[j >> 1][(j & 1) * 2 + 0] = pack_bf16x2(p0, p1);
      pf[j >> 1][(j & 1) * 2 + 1] = pack_bf16x2(p2, p3);
    }
#pragma unroll
    for (int r = 0; r < 2; ++r) l_run[r] = l_run[r] * corr[r] + rs[r];
#pragma unroll
    for (int i = 0; i < D / 8; ++i) {
      o[i][0] *= corr[0]; o[i][1] *= corr[0];
      o[i][2] *= corr[1]; o[i][3] *= corr[1];
    }
    // ---- O += P V ----
    const uint32_t vb = sV + buf * TILE;
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {      // 16-key steps
#pragma unroll
      for (int j = 0; j < D / 8; j += 2) {  // d n-tiles, two per ldmatrix.x4.trans
        uint32_t b0, b1, b2, b3;
        const int row = kk * 16 + (lm & 1) * 8 + lr, col = j * 8 + (lm >> 1) * 8;
        ldsm_x4_t(vb + row * STRIDE + col * 2, b0, b1, b2, b3);
        mma_bf16(o[j], pf[kk], b0, b1);
        mma_bf16(o[j + 1], pf[kk], b2, b3);
      }
    }
    __syncthreads();  // all warps done with this K/V buffer before it is refilled
  }

  // ---- finalize: row sums across the quad, normalise, store ----
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    l_run[r] += __shfl_xor_sync(0xffffffffu, l_run[r], 1);
    l_run[r] += __shfl_xor_sync(0xffffffffu, l_run[r], 2);
  }
  const int g = lane >> 2;
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    const int tok = q0 + warp * 16 + g + r * 8;
    if (tok >= N) continue;
    const float inv = 1.f / l_run[r];
    __nv_bfloat16* op = out + ((int64_t)b * N + tok) * out_ld + h * D + (lane & 3) * 2;
#pragma unroll
    for (int i = 0; i < D / 8; ++i)
      *reinterpret_cast<uint32_t*>(op + i * 8) = pack_bf16x2(o[i][r * 2] * inv, o[i][r * 2 + 1] * inv);
  }
}

// ---- fp32 verify engine: 4 threads per query, 32-key smem tiles ------------------------------------------
template <typename T, int D>
__global__ void __launch_bounds__(128) attn_simt_kernel(const T* __restrict__ q, const T* __restrict__ k,
                                                       const T* __restrict__ v, int ld, int N, float scale,
                                                       T* __restrict__ out, int out_ld) {
  constexpr int DP = D / 4;
  __shared__ float sK[32][D + 1];
  __shared__ float sV[32][D + 1];
  const int b = blockIdx.z, h = blockIdx.y;
  const int qi = blockIdx.x * 32 + (threadIdx.x >> 2), part = threadIdx.x & 3;
  const int64_t base = (int64_t)b * N * ld + (int64_t)h * D;
  float qv[DP], o[DP];
  const bool q_ok = qi < N;
#pragma unroll
  for (int i = 0; i < DP; ++i) {
    qv[i] = q_ok ? to_f<T>(q[base + (int64_t)qi * ld + part * DP + i]) * scale : 0.f;
    o[i] = 0.f;
  }
  float m = -INFINITY, l = 0.f;
  for (int t0 = 0; t0 < N; t0 += 32) {
    for (int i = threadIdx.x; i < 32 * D; i += 128) {
      const int r = i / D, c = i % D;
      const bool ok = t0 + r < N;
      sK[r][c] = ok ? to_f<T>(k[base + (int64_t)(t0 + r) * ld + c]) : 0.f;
      sV[r][c] = ok ? to_f<T>(v[base + (int64_t)(t0 + r) * ld + c]) : 0.f;
    }
    __syncthreads();
    const int kmax = min(32, N - t0);
    for (int j = 0; j < kmax; ++j) {
      float d = 0.f;
#pragma unroll
      for (int i = 0; i < DP; ++i) d = fmaf(qv[i], sK[j][part * DP + i], d);
      d += __shfl_xor_sync(0xffffffffu, d, 1);
      d += __shfl_xor_sync(0xffffffffu, d, 2);
      const float mn = fmaxf(m, d);
      const float c = __expf(m - mn), p = __expf(d - mn);
      l = l * c + p;
#pragma unroll
      for (int i = 0; i < DP; ++i) o[i] = fmaf(p, sV[j][part * DP + i], o[i] * c);
      m = mn;
    }
    __syncthreads();
  }
  if (q_ok) {
    const float inv = 1.f / l;
#pragma unroll
    for (int i = 0; i < DP; ++i)
      out[((int64_t)b * N + qi) * out_ld + h * D + part * DP + i] = from_f<T>(o[i] * inv);
  }
}

template <int D>
static int launch_flash(const void* q, const void* k, const void* v, int ld, int B, int N, int heads, float scale,
                        void* out, int out_ld, cudaStream_t st) {
  constexpr size_t smem = 5 * 64 * (D + 8) * 2;
  static bool attr_set = false;
  if (!attr_set) {
    cudaFuncSetAttribute(flash_attn_bf16_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    attr_set = true;
  }
  dim3 grid((N + FA_BM - 1) / FA_BM, heads, B);
  flash_attn_bf16_kernel<D><<<grid, FA_THREADS, smem, st>>>((const __nv_bfloat16*)q, (const __nv_bfloat16*)k,
                                                             (const __nv_bfloat16*)v, ld, N,
                                                             scale * 1.4426950408889634f, (__nv_bfloat16*)out, out_ld);
  DCB_CHECK_LAUNCH("flash_attn_bf16");
  return DCB_OK;
}

template <typename T, int D>
static int launch_simt(const void* q, const void* k, const void* v, int ld, int B, int N, int heads, float scale,
                       void* out, int out_ld, cudaStream_t st) {
  dim3 grid((N + 31) / 32, heads, B);
  attn_simt_kernel<T, D><<<grid, 128, 0, st>>>((const T*)q, (const T*)k, (const T*)v, ld, N, scale, (T*)out, out_ld);
  DCB_CHECK_LAUNCH("attn_simt");
  return DCB_OK;
}

int launch_flash_tc(const void* q, const void* k, const void* v, int ld, int B, int N, int heads, int d, float scale,
                    void* out, int out_ld, float* norms_ws, bool norms_ready, cudaStream_t st);

}  // namespace dcb

using namespace dcb;

// dtype DCB_BF16 -> tensor-core flash kernel; DCB_F32 -> fp32 verify kernel; (DCB_BF16 | 0x100) -> SIMT on bf16 data
extern "C" int dcb_attention_ws(int dtype, const void* q, const void* k, const void* v, int ld, int B, int Ntok, int heads,
                                int d, float scale, void* out, int out_ld, float* ws, dcb_stream stream) {
  cudaStream_t st = (cudaStream_t)stream;
  const bool norms_ready = (dtype & DCB_ATTN_NORMS_READY) != 0;
  dtype &= ~DCB_ATTN_NORMS_READY;
  DCB_REQUIRE(!norms_ready || (ws != nullptr && dtype == DCB_BF16 && d == 64), "attention: NORMS_READY needs bf16, d = 64, a workspace");
  DCB_REQUIRE(B >= 1 && B <= 65535 && heads >= 1 && Ntok >= 1, "attention: bad sizes");
  DCB_REQUIRE(d == 32 || d == 64 || d == 96 || d == 128, "attention: head dim %d not in {32,64,96,128}", d);
#define DCB_ATTN_DISPATCH(FN, ...)                                         \
  switch (d) {                                                             \
    case 32: return FN<__VA_ARGS__ 32>(q, k, v, ld, B, Ntok, heads, scale, out, out_ld, st);  \
    case 64: return FN<__VA_ARGS__ 64>(q, k, v, ld, B, Ntok, heads, scale, out, out_ld, st);  \
    case 96: return FN<__VA_ARGS__ 96>(q, k, v, ld, B, Ntok, heads, scale, out, out_ld, st);  \
    default: return FN<__VA_ARGS__ 128>(q, k, v, ld, B, Ntok, heads, scale, out, out_ld, st); \
  }
  if (dtype == DCB_BF16) {
    DCB_REQUIRE(ld % 8 == 0 && ((uintptr_t)q & 15) == 0 && ((uintptr_t)k & 15) == 0 && ((uintptr_t)v & 15) == 0,
                "attention: bf16 path needs 16-byte aligned rows");
    // at least one full key block: tcgen05 / TMEM kernel (attention_tc.cu, every head dim); the mma.sync kernel keeps
    // the sequences shorter than one 128-key block (the 8^2 levels)
    const bool tc_on = !(knobs() & DCB_KNOB_ATTN_NO_TC);
    if (Ntok >= 128 && tc_on && out_ld % 8 == 0 && ((uintptr_t)out & 15) == 0)
      return launch_flash_tc(q, k, v, ld, B, Ntok, heads, d, scale, out, out_ld, ws, norms_ready, st);
    DCB_ATTN_DISPATCH(launch_flash, )
  } else if (dtype == DCB_F32) {
    DCB_ATTN_DISPATCH(launch_simt, float, )
  } else if (dtype == (DCB_BF16 | 0x100)) {
    DCB_ATTN_DISPATCH(launch_simt, __nv_bfloat16, )
  }
#undef DCB_ATTN_DISPATCH
  set_error("attention: unknown dtype %d", dtype);
  return DCB_EINVAL;
}

extern "C" int dcb_attention(int dtype, const void* q, const void* k, const void* v, int ld, int B, int Ntok, int heads,
                             int d, float scale, void* out, int out_ld, dcb_stream stream) {
  return dcb_attention_ws(dtype, q, k, v, ld, B, Ntok, heads, d, scale, out, out_ld, nullptr, stream);
}
