// mma_lean.cu -- measurement tool: what an MMA-issue loop with per-K-block full/empty handshakes sustains (see mma_rate.cu)
#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  while (!done) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
  }
}
__device__ __forceinline__ void mbar_test_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  while (!done) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
  }
}
__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
template <int CG>
__device__ __forceinline__ void umma(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  if (CG == 1)
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d),
        "l"(a), "l"(b), "r"(idesc), "r"(acc)
        : "memory");
  else
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d),
        "l"(a), "l"(b), "r"(idesc), "r"(acc)
        : "memory");
}
template <int CG>
__device__ __forceinline__ void commit(uint32_t bar) {
  if (CG == 1)
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
  else
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
                 "h"((uint16_t)3)
                 : "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t cta_rank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}

struct Res {
  long long cycles;
  long long ns;
  long long copies;
};


__device__ __forceinline__ void umma_lohi(uint32_t d, uint32_t alo, uint32_t blo, uint32_t hi, uint32_t idesc) {
  asm volatile(
      "{\n\t.reg .b64 da, db;\n\tmov.b64 da, {%1, %3};\n\tmov.b64 db, {%2, %3};\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %4, 1;\n\t}" ::"r"(d), "r"(alo), "r"(blo), "r"(hi), "r"(idesc)
      : "memory");
}
// V: 0 = wait(full) | MMAs | commit(empty);  1 = no full-wait (MMAs + commit to a sink);  2 = wait, plain arrive instead of
// commit;  3 = software-pipelined: the wait for the NEXT stage sits in the middle of this stage's MMAs;  4 = like 3 and the
// commit of the PREVIOUS stage is issued after the first MMA pair of this stage
template <int N, int NM, int NACC, int V>
__global__ void __launch_bounds__(128, 1) lean_kernel(int kblocks, int stages, Res* out) {
  extern __shared__ __align__(1024) uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  constexpr int stage_bytes = (NM / 4) * 16384 + N * 128;
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + stages * stage_bytes);
  uint64_t* empty = full + 8;
  uint64_t* done = empty + 8;
  uint64_t* sink = done + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sink + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < stages * stage_bytes / 2; i += blockDim.x) {
    uint32_t h = (uint32_t)i * 2654435761u + blockIdx.x * 40503u;
    h ^= h >> 15;
    reinterpret_cast<uint16_t*>(smem)[i] = (uint16_t)(0x3c00u | (h & 0x807fu));
  }
  if (threadIdx.x == 0) {
    for (int i = 0; i < stages; ++i) { mbar_init(smem_u32(&full[i]), 1); mbar_init(smem_u32(&empty[i]), 1); }
    mbar_init(smem_u32(done), 1);
    mbar_init(smem_u32(sink), 0xfffff);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;
  constexpr uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  if (warp == 0 && lane == 0 && V != 1) {
    int stage = 0; uint32_t phase = 0;
    for (int kb = 0; kb < kblocks; ++kb) {
      mbar_wait(smem_u32(&empty[stage]), phase ^ 1);
      asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&full[stage])) : "memory");
      if (++stage == stages) { stage = 0; phase ^= 1; }
    }
  } else if (warp == 1 && lane == 0) {
    long long c0 = clock64();
    const uint32_t hi = (uint32_t)(make_desc(0) >> 32);
    const uint32_t lo0 = (uint32_t)make_desc(smem_u32(smem));
    const uint32_t full0 = smem_u32(full), empty0 = smem_u32(empty);
    int stage = 0; uint32_t phase = 0;
    int pstage = -1;
    if (V >= 3) { mbar_wait(full0, 0); asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
    for (int kb = 0; kb < kblocks; ++kb) {
      if (V == 0 || V == 2) { mbar_wait(full0 + stage * 8, phase); asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
      if (V == 5) { mbar_test_wait(full0 + stage * 8, phase); asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
      if (V == 6) { mbar_wait(full0 + stage * 8, phase); }
      if (V == 7) { mbar_test_wait(full0 + stage * 8, phase); }
      int nstage = stage + 1; uint32_t nphase = phase;
      if (nstage == stages) { nstage = 0; nphase ^= 1; }
      const uint32_t alo = lo0 + (uint32_t)stage * (stage_bytes >> 4);
      const uint32_t blo = alo + (NM / 4) * (16384 >> 4);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        if (V >= 3 && k == 2 && kb + 1 < kblocks) { mbar_wait(full0 + nstage * 8, nphase); asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
        if (V == 4 && k == 1 && pstage >= 0) commit<1>(empty0 + pstage * 8);
#pragma unroll
        for (int s = 0; s < NM / 4; ++s)
          umma_lohi(tmem_base + (uint32_t)((s % NACC) * N), alo + s * (16384 >> 4) + 2 * k, blo + 2 * k, hi, idesc);
      }
      if (V == 2) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(empty0 + stage * 8) : "memory");
      else if (V == 1) commit<1>(smem_u32(sink));
      else if (V != 4) commit<1>(empty0 + stage * 8);
      pstage = stage;
      stage = nstage; phase = nphase;
    }
    if (V == 4) commit<1>(empty0 + pstage * 8);
    commit<1>(smem_u32(done));
    mbar_wait(smem_u32(done), 0);
    long long c1 = clock64();
    out[blockIdx.x].cycles = c1 - c0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
  }
}
template <int N, int NM, int NACC, int V>
static void run_lean(int stages, Res* dres) {
  const int kblocks = 4096, grid = 148;
  const int stage_bytes = (NM / 4) * 16384 + N * 128;
  const size_t smem = (size_t)stages * stage_bytes + 512 + 1024;
  cudaFuncSetAttribute(lean_kernel<N, NM, NACC, V>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  for (int rep = 0; rep < 2; ++rep) {
    lean_kernel<N, NM, NACC, V><<<grid, 128, smem>>>(kblocks, stages, dres);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("lean kernel failed: %s\n", cudaGetErrorString(e)); exit(1); }
  }
  Res h[148];
  cudaMemcpy(h, dres, sizeof(Res) * 148, cudaMemcpyDeviceToHost);
  long long cyc = 0;
  for (int i = 0; i < grid; ++i) if (h[i].cycles > cyc) cyc = h[i].cycles;
  printf("lean V=%d N=%3d mma/kblock=%d accumulators=%d stages=%d  cyc/kblock=%7.1f  (pipe floor %5.0f)  -> %5.1f%% of pipe\n", V, N, NM, NACC,
         stages, (double)cyc / kblocks, NM * 128.0 * N / 256.0, 100.0 * NM * 128.0 * N / 256.0 / ((double)cyc / kblocks));
}

__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}
// ---- NW issuing warps per CTA: each K block = 2 A sub-tiles x 4 k-steps against one shared B tile; warp w issues the MMAs of
// sub-tile (w % 2) [and k-half (w / 2) when NW == 4] into its own accumulator.  full[] has many waiters, empty[] counts NW
// commits.  STYLE 1 = warp-uniform loop (all lanes run the bookkeeping, elect.sync picks the issuing lane) so that ptxas keeps
// descriptors in uniform registers; STYLE 0 = everything under lane == 0.
template <int NW, int STYLE>
__global__ void __launch_bounds__(192, 1) multi_kernel(int kblocks, int stages, Res* out) {
  extern __shared__ __align__(1024) uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  constexpr int N = 128;
  constexpr int stage_bytes = 2 * 16384 + N * 128;
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + stages * stage_bytes);
  uint64_t* empty = full + 8;
  uint64_t* done = empty + 8;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(done + 1);
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  const int lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < stages * stage_bytes / 2; i += blockDim.x) {
    uint32_t h = (uint32_t)i * 2654435761u + blockIdx.x * 40503u;
    h ^= h >> 15;
    reinterpret_cast<uint16_t*>(smem)[i] = (uint16_t)(0x3c00u | (h & 0x807fu));
  }
  if (threadIdx.x == 0) {
    for (int i = 0; i < stages; ++i) { mbar_init(smem_u32(&full[i]), 1); mbar_init(smem_u32(&empty[i]), NW); }
    mbar_init(smem_u32(done), NW);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;
  constexpr uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  if (warp == 0) {
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      for (int kb = 0; kb < kblocks; ++kb) {
        mbar_wait(smem_u32(&empty[stage]), phase ^ 1);
        asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&full[stage])) : "memory");
        if (++stage == stages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp >= 1 && warp <= NW && (STYLE == 1 || lane == 0)) {
    const int w = warp - 1;
    const int sub = w & 1, khalf = w >> 1;
    long long c0 = clock64();
    const uint32_t hi = (uint32_t)(make_desc(0) >> 32);
    const uint32_t lo0 = (uint32_t)make_desc(smem_u32(smem)) + sub * (16384 >> 4);
    const uint32_t full0 = smem_u32(full), empty0 = smem_u32(empty);
    const uint32_t d = tmem_base + (uint32_t)(w * 128);
    int stage = 0; uint32_t phase = 0;
    for (int kb = 0; kb < kblocks; ++kb) {
      mbar_wait(full0 + stage * 8, phase);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t alo = lo0 + (uint32_t)stage * (stage_bytes >> 4);
      const uint32_t blo = alo + (2 - sub) * (16384 >> 4);
      if (STYLE == 0 || elect_one()) {
        if (NW == 1) {
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            umma_lohi(d, alo + 2 * k, blo + 2 * k, hi, idesc);
            umma_lohi(d + 128, alo + (16384 >> 4) + 2 * k, blo + 2 * k, hi, idesc);
          }
        } else if (NW == 2) {
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_lohi(d, alo + 2 * k, blo + 2 * k, hi, idesc);
        } else {
#pragma unroll
          for (int k = 0; k < 2; ++k) umma_lohi(d, alo + 2 * (2 * khalf + k), blo + 2 * (2 * khalf + k), hi, idesc);
        }
        commit<1>(empty0 + stage * 8);
      }
      if (STYLE == 1) __syncwarp();
      if (++stage == stages) { stage = 0; phase ^= 1; }
    }
    if (STYLE == 0 || elect_one()) commit<1>(smem_u32(done));
    mbar_wait(smem_u32(done), 0);
    long long c1 = clock64();
    if (lane == 0 && w == 0) out[blockIdx.x].cycles = c1 - c0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
  }
}
template <int NW, int STYLE>
static void run_multi(int stages, Res* dres) {
  const int kblocks = 4096, grid = 148;
  const int stage_bytes = 2 * 16384 + 128 * 128;
  const size_t smem = (size_t)stages * stage_bytes + 512 + 1024;
  cudaFuncSetAttribute(multi_kernel<NW, STYLE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  for (int rep = 0; rep < 2; ++rep) {
    multi_kernel<NW, STYLE><<<grid, 192, smem>>>(kblocks, stages, dres);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("multi kernel failed: %s\n", cudaGetErrorString(e)); exit(1); }
  }
  Res h[148];
  cudaMemcpy(h, dres, sizeof(Res) * 148, cudaMemcpyDeviceToHost);
  long long cyc = 0;
  for (int i = 0; i < grid; ++i) if (h[i].cycles > cyc) cyc = h[i].cycles;
  printf("multi issuing-warps=%d style=%d N=128 8 mma/kblock stages=%d  cyc/kblock=%7.1f  (pipe floor 512)  -> %5.1f%% of pipe\n", NW, STYLE,
         stages, (double)cyc / kblocks, 100.0 * 512.0 / ((double)cyc / kblocks));
}
int main() {
  { Res* dres; cudaMalloc(&dres, sizeof(Res) * 148);
    run_multi<1, 0>(3, dres); run_multi<1, 1>(3, dres); run_multi<2, 0>(3, dres); run_multi<2, 1>(3, dres); run_multi<4, 0>(3, dres); run_multi<4, 1>(3, dres); return 0; }
  Res* dres;
  cudaMalloc(&dres, sizeof(Res) * 148);
  run_lean<128, 4, 1, 0>(4, dres);
  run_lean<128, 4, 1, 1>(4, dres);
  run_lean<128, 4, 1, 2>(4, dres);
  run_lean<128, 4, 1, 3>(4, dres);
  run_lean<128, 4, 1, 4>(4, dres);
  run_lean<128, 8, 2, 0>(3, dres);
  run_lean<128, 8, 2, 1>(3, dres);
  run_lean<128, 8, 2, 2>(3, dres);
  run_lean<128, 8, 2, 3>(3, dres);
  run_lean<128, 8, 2, 4>(3, dres);
  run_lean<128, 4, 1, 5>(4, dres);
  run_lean<128, 4, 1, 6>(4, dres);
  run_lean<128, 4, 1, 7>(4, dres);
  run_lean<128, 8, 2, 5>(3, dres);
  run_lean<128, 8, 2, 6>(3, dres);
  run_lean<128, 8, 2, 7>(3, dres);
  run_lean<128, 4, 1, 0>(6, dres);
  run_lean<128, 4, 1, 7>(6, dres);
  return 0;
}
