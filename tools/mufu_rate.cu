// mufu_rate.cu -- measurement tool: MUFU exp2 throughput per SM for f32, f16x2 and bf16x2 forms (is the packed form 2 results per op?)
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
template <int MODE>
__global__ void k(float* out, int iters) {
  float a0 = threadIdx.x * 1e-3f, a1 = a0 + 0.1f, a2 = a0 + 0.2f, a3 = a0 + 0.3f;
  uint32_t b0 = 0x3c003c00u + threadIdx.x, b1 = b0 + 1, b2 = b0 + 2, b3 = b0 + 3;
  long long c0 = clock64();
  for (int i = 0; i < iters; ++i) {
    if (MODE == 0) {
      asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a0)); asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a1));
      asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a2)); asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a3));
    } else if (MODE == 1) {
      asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(b0)); asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(b1));
      asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(b2)); asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(b3));
    } else if (MODE == 2) {
      asm volatile("ex2.approx.ftz.bf16x2 %0, %0;" : "+r"(b0)); asm volatile("ex2.approx.ftz.bf16x2 %0, %0;" : "+r"(b1));
      asm volatile("ex2.approx.ftz.bf16x2 %0, %0;" : "+r"(b2)); asm volatile("ex2.approx.ftz.bf16x2 %0, %0;" : "+r"(b3));
    } else {
      asm volatile("tanh.approx.f32 %0, %0;" : "+f"(a0)); asm volatile("tanh.approx.f32 %0, %0;" : "+f"(a1));
      asm volatile("tanh.approx.f32 %0, %0;" : "+f"(a2)); asm volatile("tanh.approx.f32 %0, %0;" : "+f"(a3));
    }
  }
  long long c1 = clock64();
  if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = (float)(c1 - c0);
  out[1 + blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + __uint_as_float(b0 ^ b1 ^ b2 ^ b3);
}
template <int MODE> void run(const char* name, float* d) {
  const int iters = 4096, threads = 1024;
  k<MODE><<<148, threads>>>(d, iters); cudaDeviceSynchronize();
  k<MODE><<<148, threads>>>(d, iters); cudaDeviceSynchronize();
  float cyc; cudaMemcpy(&cyc, d, 4, cudaMemcpyDeviceToHost);
  const double ops = 4.0 * iters * threads;   // MUFU instructions x lanes per SM
  printf("%-22s %.2f lane-ops/clk/SM  (%s results/clk/SM: %.1f)\n", name, ops / cyc, name, ops / cyc * (MODE == 1 || MODE == 2 ? 2 : 1));
}
int main() {
  float* d; cudaMalloc(&d, (1 + 148 * 1024) * 4);
  run<0>("ex2.approx.ftz.f32", d); run<1>("ex2.approx.f16x2", d); run<2>("ex2.approx.ftz.bf16x2", d); run<3>("tanh.approx.f32", d);
  return 0;
}
