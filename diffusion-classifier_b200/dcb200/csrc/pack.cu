// pack.cu -- weight packing behind the C ABI ("weights pre-packed once by a dcb_pack_* call"): a binder that is
// not Python can build every operand layout the GEMM kernels consume from the checkpoint's own tensors (diffusers layout:
// conv [Cout][Cin][kh][kw], linear [N][K], fp32), on the device, once per parameter version.
//   dcb_pack_conv        conv weight -> [Cout][(ky, kx, cin)] K-major rows, K zero-padded (conv_in / DiT patch embedding)
//   dcb_pack_geglu       GEGLU projection [value rows | gate rows] -> 128-row interleave, so that one 256-wide GEMM tile holds
//                        [128 value | 128 gate] columns of the same outputs (diffusers GEGLU: proj(x).chunk(2))
//   dcb_pack_upsample    nearest-2x + conv3x3 folded into four 2x2-tap phase convs over the low-resolution input
//   dcb_pack_rows        copy rows [r0, r0 + n) of an [*, K] matrix into rows of a wider / taller packed matrix at a column
//                        offset (fused QKV, concatenated time-embedding projections, [conv2 | 1x1 shortcut] along K)
// Sums are formed in fp32 and rounded once to the engine dtype (round-to-nearest-even, as dcb_cast_f32).
#include "common.cuh"

namespace dcb {

static unsigned pack_grid(int64_t total) {
  int64_t b = (total + 255) / 256;
  const int64_t cap = (int64_t)num_sms() * 16;
  return (unsigned)(b < 1 ? 1 : (b > cap ? cap : b));
}

template <typename T>
__global__ void pack_conv_kernel(const float* __restrict__ w, int Cout, int Cin, int kh, int kw, int kpad, T* __restrict__ out) {
  const int64_t total = (int64_t)Cout * kpad;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int co = (int)(i / kpad), k = (int)(i % kpad);
    float v = 0.f;
    if (k < kh * kw * Cin) {
      const int ci = k % Cin, t = k / Cin, kx = t % kw, ky = t / kw;
      v = w[(((int64_t)co * Cin + ci) * kh + ky) * kw + kx];
    }
    out[i] = from_f<T>(v);
  }
}

template <typename T>
__global__ void pack_geglu_kernel(const float* __restrict__ w, const float* __restrict__ bias, int inner, int C,
                                  T* __restrict__ w_out, float* __restrict__ b_out) {
  // packed row p = 256 blk + j: j < 128 -> value row 128 blk + j, else gate row inner + 128 blk + (j - 128)
  const int64_t total = (int64_t)2 * inner * C;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int p = (int)(i / C), c = (int)(i % C);
    const int blk = p >> 8, j = p & 255;
    const int src = j < 128 ? blk * 128 + j : inner + blk * 128 + (j - 128);
    w_out[i] = from_f<T>(w[(int64_t)src * C + c]);
    if (c == 0 && bias != nullptr) b_out[p] = bias[src];
  }
}

template <typename T>
__global__ void pack_upsample_kernel(const float* __restrict__ w, int Cout, int Cin, T* __restrict__ out) {
  // out[phase 2a + b][co][(ty, tx, ci)]; phase a = 0 reads source rows (y - 1, y) with taps ({ky0}, {ky1, ky2}),
  // a = 1 reads (y, y + 1) with ({ky0, ky1}, {ky2}); same along x (engine.fold_upsample_weights, summed in the same order)
  const int64_t total = (int64_t)4 * Cout * 4 * Cin;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int ci = (int)(i % Cin);
    int64_t r = i / Cin;
    const int tx = (int)(r & 1), ty = (int)((r >> 1) & 1);
    r >>= 2;
    const int co = (int)(r % Cout), ph = (int)(r / Cout);
    const int a = ph >> 1, b = ph & 1;
    const int ky0 = a == 0 ? (ty == 0 ? 0 : 1) : (ty == 0 ? 0 : 2), ky1 = a == 0 ? (ty == 0 ? 0 : 2) : (ty == 0 ? 1 : 2);
    const int kx0 = b == 0 ? (tx == 0 ? 0 : 1) : (tx == 0 ? 0 : 2), kx1 = b == 0 ? (tx == 0 ? 0 : 2) : (tx == 0 ? 1 : 2);
    const float* wp = w + ((int64_t)co * Cin + ci) * 9;
    float acc = 0.f;
    for (int ky = ky0; ky <= ky1; ++ky)
      for (int kx = kx0; kx <= kx1; ++kx) acc = acc + wp[ky * 3 + kx];
    out[i] = from_f<T>(acc);
  }
}

template <typename T>
__global__ void pack_rows_kernel(const float* __restrict__ src, int src_ld, int rows, int cols, T* __restrict__ dst, int dst_ld) {
  const int64_t total = (int64_t)rows * cols;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int r = (int)(i / cols), c = (int)(i % cols);
    dst[(int64_t)r * dst_ld + c] = from_f<T>(src[(int64_t)r * src_ld + c]);
  }
}

}  // namespace dcb

using namespace dcb;


extern "C" int dcb_pack_conv(int dtype, const float* w_oihw, int Cout, int Cin, int kh, int kw, int kpad, void* out,
                             dcb_stream stream) {
  DCB_REQUIRE(dtype == DCB_F32 || dtype == DCB_BF16, "pack_conv: bad dtype %d", dtype);
  DCB_REQUIRE(w_oihw != nullptr && out != nullptr && Cout >= 1 && Cin >= 1 && kh >= 1 && kw >= 1 && kpad >= kh * kw * Cin,
              "pack_conv: kpad (%d) must cover kh*kw*Cin (%d)", kpad, kh * kw * Cin);
  if (dtype == DCB_BF16) pack_conv_kernel<__nv_bfloat16><<<pack_grid((int64_t)Cout * kpad), 256, 0, (cudaStream_t)stream>>>(
        w_oihw, Cout, Cin, kh, kw, kpad, (__nv_bfloat16*)out);
  else pack_conv_kernel<float><<<pack_grid((int64_t)Cout * kpad), 256, 0, (cudaStream_t)stream>>>(w_oihw, Cout, Cin, kh, kw, kpad,
                                                                                               (float*)out);
  DCB_CHECK_LAUNCH("pack_conv");
  return DCB_OK;
}

extern "C" int dcb_pack_geglu(int dtype, const float* w, const float* bias, int inner, int C, void* w_out, float* bias_out,
                              dcb_stream stream) {
  DCB_REQUIRE(dtype == DCB_F32 || dtype == DCB_BF16, "pack_geglu: bad dtype %d", dtype);
  DCB_REQUIRE(w != nullptr && w_out != nullptr && inner >= 128 && inner % 128 == 0 && C >= 1 &&
                  (bias == nullptr) == (bias_out == nullptr),
              "pack_geglu: inner (%d) must be a multiple of 128; bias / bias_out come together", inner);
  if (dtype == DCB_BF16) pack_geglu_kernel<__nv_bfloat16><<<pack_grid((int64_t)2 * inner * C), 256, 0, (cudaStream_t)stream>>>(
        w, bias, inner, C, (__nv_bfloat16*)w_out, bias_out);
  else pack_geglu_kernel<float><<<pack_grid((int64_t)2 * inner * C), 256, 0, (cudaStream_t)stream>>>(w, bias, inner, C,
                                                                                                 (float*)w_out, bias_out);
  DCB_CHECK_LAUNCH("pack_geglu");
  return DCB_OK;
}

extern "C" int dcb_pack_upsample(int dtype, const float* w_oihw, int Cout, int Cin, void* out, dcb_stream stream) {
  DCB_REQUIRE(dtype == DCB_F32 || dtype == DCB_BF16, "pack_upsample: bad dtype %d", dtype);
  DCB_REQUIRE(w_oihw != nullptr && out != nullptr && Cout >= 1 && Cin >= 1, "pack_upsample: bad arguments");
  if (dtype == DCB_BF16) pack_upsample_kernel<__nv_bfloat16><<<pack_grid((int64_t)16 * Cout * Cin), 256, 0, (cudaStream_t)stream>>>(
        w_oihw, Cout, Cin, (__nv_bfloat16*)out);
  else pack_upsample_kernel<float><<<pack_grid((int64_t)16 * Cout * Cin), 256, 0, (cudaStream_t)stream>>>(w_oihw, Cout, Cin,
                                                                                                      (float*)out);
  DCB_CHECK_LAUNCH("pack_upsample");
  return DCB_OK;
}

extern "C" int dcb_pack_rows(int dtype, const float* src, int src_ld, int rows, int cols, void* dst, int dst_ld, int dst_row0,
                             int dst_col0, dcb_stream stream) {
  DCB_REQUIRE(dtype == DCB_F32 || dtype == DCB_BF16, "pack_rows: bad dtype %d", dtype);
  DCB_REQUIRE(src != nullptr && dst != nullptr && rows >= 1 && cols >= 1 && src_ld >= cols && dst_ld >= dst_col0 + cols &&
                  dst_row0 >= 0 && dst_col0 >= 0,
              "pack_rows: bad geometry");
  const int64_t off = (int64_t)dst_row0 * dst_ld + dst_col0;
  if (dtype == DCB_BF16) pack_rows_kernel<__nv_bfloat16><<<pack_grid((int64_t)rows * cols), 256, 0, (cudaStream_t)stream>>>(
        src, src_ld, rows, cols, (__nv_bfloat16*)dst + off, dst_ld);
  else pack_rows_kernel<float><<<pack_grid((int64_t)rows * cols), 256, 0, (cudaStream_t)stream>>>(src, src_ld, rows, cols,
                                                                                              (float*)dst + off, dst_ld);
  DCB_CHECK_LAUNCH("pack_rows");
  return DCB_OK;
}
