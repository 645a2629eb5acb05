// mma_rate.cu -- measurement tool (not product code): raw issue rate of tcgen05.mma kind::f16 (bf16 x bf16 -> fp32) with
// both operands in shared memory (SS), for cta_group::1 (M=128) and cta_group::2 (M=256 over a CTA pair), N in
// {64,128,256}, optionally with concurrent bulk-copy (TMA engine) writes into other shared-memory slots.  It answers
// "what bounds a 128 x N tile: the MMA pipe, or shared-memory bandwidth?" for DESIGN.md section 4.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/bin/mma_rate tools/mma_rate.cu
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

// smem: [slots x (A 16 KB | B rows_b x 128 B)] [scratch 2 x 32 KB for the copy engine] [barriers]
template <int CG>
__global__ void __launch_bounds__(128, 1) rate_kernel(int N, int kblocks, int slots, int traffic, int bytes_per_copy,
                                                      const uint8_t* gsrc, Res* out) {
  extern __shared__ __align__(1024) uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  const int rows_b = N / CG;  // each CTA of a pair holds half of B's rows
  const int slot_bytes = 16384 + rows_b * 128;
  uint8_t* scratch = smem + slots * slot_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(scratch + 2 * 32768);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);
  volatile int* stop = reinterpret_cast<volatile int*>(tmem_slot + 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // fill operands with small pseudo-random bf16 values (non-zero toggling, finite accumulators)
  for (int i = threadIdx.x; i < slots * slot_bytes / 2; i += blockDim.x) {
    uint32_t h = (uint32_t)i * 2654435761u + blockIdx.x * 40503u;
    h ^= h >> 15;
    reinterpret_cast<uint16_t*>(smem)[i] = (uint16_t)(0x3c00u | (h & 0x807fu));  // +-(1.0 .. 1.99) * 2^-7..: bf16 ~ 0.0078
  }
  if (threadIdx.x == 0) {
    mbar_init(smem_u32(&bars[0]), 1);
    mbar_init(smem_u32(&bars[1]), 1);
    mbar_init(smem_u32(&bars[2]), 1);
    *stop = 0;
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    if (CG == 1) {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (CG == 2) cluster_sync_all();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;
  const bool leader = CG == 1 || cta_rank() == 0;
  const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)((128 * CG) >> 4) << 24);

  if (warp == 1 && lane == 0) {
    long long c0 = clock64();
    unsigned long long t0, t1;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    if (leader) {
      int slot = 0;
      for (int kb = 0; kb < kblocks; ++kb) {
        const uint32_t sa = smem_u32(smem + slot * slot_bytes);
        const uint64_t ad = make_desc(sa), bd = make_desc(sa + 16384);
        const uint32_t d = tmem_base + (uint32_t)((kb & 1) * 256 % (512 - N + 1));
#pragma unroll
        for (int k = 0; k < 4; ++k) umma<CG>(d, ad + 2 * k, bd + 2 * k, idesc, 1);
        if (++slot == slots) slot = 0;
      }
      commit<CG>(smem_u32(&bars[0]));
    }
    mbar_wait(smem_u32(&bars[0]), 0);
    long long c1 = clock64();
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
    *stop = 1;
    out[blockIdx.x].cycles = c1 - c0;
    out[blockIdx.x].ns = (long long)(t1 - t0);
  } else if (warp == 2 && lane == 0 && traffic) {
    // copy engine: back-to-back bulk copies global -> scratch smem (two in flight), as much as it will take
    long long n = 0;
    uint32_t ph[2] = {0, 0};
    const int nsrc = 64;  // rotate over 64 source blocks (L2-resident)
    for (int i = 0; i < 2; ++i) {
      mbar_expect_tx(smem_u32(&bars[1 + i]), bytes_per_copy);
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                       smem_u32(scratch + i * 32768)),
                   "l"(gsrc + (size_t)((blockIdx.x + i) % nsrc) * 32768), "r"(bytes_per_copy), "r"(smem_u32(&bars[1 + i]))
                   : "memory");
    }
    int i = 0;
    while (!*stop) {
      mbar_wait(smem_u32(&bars[1 + i]), ph[i]);
      ph[i] ^= 1;
      ++n;
      mbar_expect_tx(smem_u32(&bars[1 + i]), bytes_per_copy);
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                       smem_u32(scratch + i * 32768)),
                   "l"(gsrc + (size_t)((blockIdx.x + n) % nsrc) * 32768), "r"(bytes_per_copy), "r"(smem_u32(&bars[1 + i]))
                   : "memory");
      i ^= 1;
    }
    mbar_wait(smem_u32(&bars[1]), ph[0]);
    mbar_wait(smem_u32(&bars[2]), ph[1]);
    out[blockIdx.x].copies = n;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (CG == 2) cluster_sync_all();
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (CG == 1)
      asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
    else
      asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
  }
}


// ---- pipelined skeleton: producer warp (bulk copies of A+B bytes per stage, or bare arrives) -> full[] -> MMA warp -> commit
// empty[] -> producer.  Reports cycles per K block (4 MMAs) as a function of ring depth: what a real mainloop can sustain.
__global__ void __launch_bounds__(128, 1) pipe_kernel(int N, int kblocks, int stages, int do_copy, int do_mma, int src_blocks, int variant,
                                                      const uint8_t* gsrc, Res* out) {
  extern __shared__ __align__(1024) uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  const int stage_bytes = 16384 + N * 128;
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + stages * stage_bytes);
  uint64_t* empty = full + 8;
  uint64_t* done = empty + 8;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(done + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < stages * stage_bytes / 2; i += blockDim.x) {
    uint32_t h = (uint32_t)i * 2654435761u + blockIdx.x * 40503u;
    h ^= h >> 15;
    reinterpret_cast<uint16_t*>(smem)[i] = (uint16_t)(0x3c00u | (h & 0x807fu));
  }
  if (threadIdx.x == 0) {
    for (int i = 0; i < stages; ++i) { mbar_init(smem_u32(&full[i]), 1); mbar_init(smem_u32(&empty[i]), 1); }
    mbar_init(smem_u32(done), 1);
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
  const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  if (warp == 0 && lane == 0) {
    int stage = 0; uint32_t phase = 0;
    for (int kb = 0; kb < kblocks; ++kb) {
      mbar_wait(smem_u32(&empty[stage]), phase ^ 1);
      const uint32_t fb = smem_u32(&full[stage]);
      if (do_copy) {
        mbar_expect_tx(fb, stage_bytes);
        const uint8_t* src = gsrc + (size_t)((blockIdx.x * 7 + kb) % src_blocks) * 49152;
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                         smem_u32(smem + stage * stage_bytes)), "l"(src), "r"(stage_bytes), "r"(fb) : "memory");
      } else {
        asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(fb) : "memory");
      }
      if (++stage == stages) { stage = 0; phase ^= 1; }
    }
  } else if (warp == 1 && (!(variant & 4) || lane == 0)) {
    long long c0 = clock64();
    unsigned long long t0, t1;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    int stage = 0; uint32_t phase = 0;
    for (int kb = 0; kb < kblocks; ++kb) {
      mbar_wait(smem_u32(&full[stage]), phase);
      if (!(variant & 2)) asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      if (lane == 0) {
        const uint32_t sa = smem_u32(smem + stage * stage_bytes);
        const uint64_t ad = make_desc(sa), bd = make_desc(sa + 16384);
        if (do_mma) {
#pragma unroll
          for (int k = 0; k < 4; ++k) umma<1>(tmem_base + (uint32_t)((kb >> 4) & 1) * 256, ad + 2 * k, bd + 2 * k, idesc, 1);
        }
        if (variant & 1) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&empty[stage])) : "memory");
        else commit<1>(smem_u32(&empty[stage]));
      }
      if (!(variant & 4)) __syncwarp();
      if (++stage == stages) { stage = 0; phase ^= 1; }
    }
    if (lane == 0) commit<1>(smem_u32(done));
    mbar_wait(smem_u32(done), 0);
    long long c1 = clock64();
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
    if (lane == 0) { out[blockIdx.x].cycles = c1 - c0; out[blockIdx.x].ns = (long long)(t1 - t0); }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
  }
}

static void run_pipe(int N, int stages, int do_copy, int do_mma, int src_blocks, const uint8_t* gsrc, Res* dres, int variant = 0) {
  const int kblocks = 4096, grid = 148;
  const int stage_bytes = 16384 + N * 128;
  const size_t smem = (size_t)stages * stage_bytes + 512 + 1024;
  if (smem > 227 * 1024) return;
  cudaFuncSetAttribute(pipe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  cudaMemset(dres, 0, sizeof(Res) * 148);
  for (int rep = 0; rep < 2; ++rep) {
    pipe_kernel<<<grid, 128, smem>>>(N, kblocks, stages, do_copy, do_mma, src_blocks, variant, gsrc, dres);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("pipe kernel failed: %s\n", cudaGetErrorString(e)); exit(1); }
  }
  Res h[148];
  cudaMemcpy(h, dres, sizeof(Res) * 148, cudaMemcpyDeviceToHost);
  long long cyc = 0, ns = 0;
  for (int i = 0; i < grid; ++i) if (h[i].cycles > cyc) { cyc = h[i].cycles; ns = h[i].ns; }
  const double flops = 4.0 * kblocks * 2.0 * 128 * N * 16 * grid;
  printf("pipe v=%d N=%3d stages=%d copy=%d mma=%d src=%5.0f MB  cyc/kblock=%7.1f (mma floor %5.1f)  clk=%.3f GHz  chip %7.1f TF/s  load %6.2f TB/s\n",
         variant, N, stages, do_copy, do_mma, src_blocks * 49152 / 1e6, (double)cyc / kblocks, 4 * 128.0 * N / 256.0, (double)cyc / ns,
         do_mma ? flops / ns / 1e3 : 0.0, do_copy ? (double)stage_bytes * kblocks * grid / ns / 1e3 : 0.0);
}

// ---- what is additive with MMA issue in the issuing thread?  One thread per CTA loops over "K blocks" of 4 MMAs plus
// optional extras.  bits: 1 = 4 MMAs, 2 = try_wait on an already-completed barrier phase, 4 = tcgen05.commit to a dummy
// barrier (count huge), 8 = plain mbarrier.arrive on the dummy, 16 = tcgen05.fence::after_thread_sync, 32 = N=256 MMAs
__global__ void __launch_bounds__(128, 1) issue_kernel(int kblocks, int mode, int period, Res* out) {
  extern __shared__ __align__(1024) uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  const int N = (mode & 32) ? 256 : 128;
  const int stage_bytes = 16384 + N * 128;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 4 * stage_bytes);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 4);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 4 * stage_bytes / 2; i += blockDim.x) {
    uint32_t h = (uint32_t)i * 2654435761u + blockIdx.x * 40503u;
    h ^= h >> 15;
    reinterpret_cast<uint16_t*>(smem)[i] = (uint16_t)(0x3c00u | (h & 0x807fu));
  }
  if (threadIdx.x == 0) {
    mbar_init(smem_u32(&bars[0]), 1);          // completed once below -> parity-0 waits always succeed
    mbar_init(smem_u32(&bars[1]), 0xfffff);    // dummy sink for commits / arrives
    mbar_init(smem_u32(&bars[2]), 1);          // final
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&bars[0])) : "memory");
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
  const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  if (warp == 1 && lane == 0) {
    long long c0 = clock64();
    unsigned long long t0, t1;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    int stage = 0;
    for (int kb = 0; kb < kblocks; ++kb) {
      if (mode & 2) mbar_wait(smem_u32(&bars[0]), 0);
      if (mode & 16) asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      if (mode & 1) {
        const uint32_t sa = smem_u32(smem + stage * stage_bytes);
        const uint64_t ad = make_desc(sa), bd = make_desc(sa + 16384);
#pragma unroll
        for (int k = 0; k < 4; ++k) umma<1>(tmem_base + (uint32_t)(((kb * 4 + k) / period) & 1) * 256, ad + 2 * k, bd + 2 * k, idesc, 1);
      }
      if (mode & 4) commit<1>(smem_u32(&bars[1]));
      if (mode & 8) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&bars[1])) : "memory");
      if (++stage == 4) stage = 0;
    }
    commit<1>(smem_u32(&bars[2]));
    mbar_wait(smem_u32(&bars[2]), 0);
    long long c1 = clock64();
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
    out[blockIdx.x].cycles = c1 - c0;
    out[blockIdx.x].ns = (long long)(t1 - t0);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
  }
}
static void run_issue(int mode, Res* dres, int period = 64) {
  const int kblocks = 4096, grid = 148;
  const size_t smem = 4 * (16384 + 256 * 128) + 512 + 1024;
  cudaFuncSetAttribute(issue_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  for (int rep = 0; rep < 2; ++rep) {
    issue_kernel<<<grid, 128, smem>>>(kblocks, mode, period, dres);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("issue kernel failed: %s\n", cudaGetErrorString(e)); exit(1); }
  }
  Res h[148];
  cudaMemcpy(h, dres, sizeof(Res) * 148, cudaMemcpyDeviceToHost);
  long long cyc = 0;
  for (int i = 0; i < grid; ++i) if (h[i].cycles > cyc) cyc = h[i].cycles;
  printf("issue period=%2d mode=%2d [%s%s%s%s%s%s]  cyc/iter=%7.1f\n", period, mode, mode & 1 ? "4mma " : "", mode & 32 ? "N256 " : "", mode & 2 ? "try_wait " : "",
         mode & 16 ? "fence " : "", mode & 4 ? "commit " : "", mode & 8 ? "arrive " : "", (double)cyc / kblocks);
}

// ---- lean mainloop: what a carefully written MMA-issue loop sustains.  Producer warp re-arms full[] when empty[] completes
// (no data movement); the MMA thread does try_wait(full) -> NM MMAs (alternating NACC accumulators) -> commit(empty).
// Descriptors are built from 32-bit low words (one IMAD per K block); no 64-bit arithmetic, no runtime mode flags.
__device__ __forceinline__ void umma_lohi(uint32_t d, uint32_t alo, uint32_t blo, uint32_t hi, uint32_t idesc) {
  asm volatile(
      "{\n\t.reg .b64 da, db;\n\tmov.b64 da, {%1, %3};\n\tmov.b64 db, {%2, %3};\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %4, 1;\n\t}" ::"r"(d), "r"(alo), "r"(blo), "r"(hi), "r"(idesc)
      : "memory");
}
template <int N, int NM, int NACC, int LANES>
__global__ void __launch_bounds__(128, 1) lean_kernel(int kblocks, int stages, Res* out) {
  extern __shared__ __align__(1024) uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  constexpr int stage_bytes = (NM / 4) * 16384 + N * 128;   // NM/4 A sub-tiles + one B tile per stage
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + stages * stage_bytes);
  uint64_t* empty = full + 8;
  uint64_t* done = empty + 8;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(done + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < stages * stage_bytes / 2; i += blockDim.x) {
    uint32_t h = (uint32_t)i * 2654435761u + blockIdx.x * 40503u;
    h ^= h >> 15;
    reinterpret_cast<uint16_t*>(smem)[i] = (uint16_t)(0x3c00u | (h & 0x807fu));
  }
  if (threadIdx.x == 0) {
    for (int i = 0; i < stages; ++i) { mbar_init(smem_u32(&full[i]), 1); mbar_init(smem_u32(&empty[i]), 1); }
    mbar_init(smem_u32(done), 1);
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
  if (warp == 0 && lane == 0) {
    int stage = 0; uint32_t phase = 0;
    for (int kb = 0; kb < kblocks; ++kb) {
      mbar_wait(smem_u32(&empty[stage]), phase ^ 1);
      asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&full[stage])) : "memory");
      if (++stage == stages) { stage = 0; phase ^= 1; }
    }
  } else if (warp == 1 && lane < LANES) {
    long long c0 = clock64();
    unsigned long long t0, t1;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    const uint32_t hi = (uint32_t)(make_desc(0) >> 32);
    const uint32_t lo0 = (uint32_t)make_desc(smem_u32(smem));
    const uint32_t full0 = smem_u32(full), empty0 = smem_u32(empty);
    int stage = 0; uint32_t phase = 0;
    for (int kb = 0; kb < kblocks; ++kb) {
      mbar_wait(full0 + stage * 8, phase);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      if (lane == 0) {
        const uint32_t alo = lo0 + (uint32_t)stage * (stage_bytes >> 4);
        const uint32_t blo = alo + (NM / 4) * (16384 >> 4);
#pragma unroll
        for (int k = 0; k < 4; ++k)
#pragma unroll
          for (int s = 0; s < NM / 4; ++s)
            umma_lohi(tmem_base + (uint32_t)((s % NACC) * N), alo + s * (16384 >> 4) + 2 * k, blo + 2 * k, hi, idesc);
        commit<1>(empty0 + stage * 8);
      }
      if (LANES > 1) __syncwarp();
      if (++stage == stages) { stage = 0; phase ^= 1; }
    }
    if (lane == 0) commit<1>(smem_u32(done));
    mbar_wait(smem_u32(done), 0);
    long long c1 = clock64();
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
    if (lane == 0) { out[blockIdx.x].cycles = c1 - c0; out[blockIdx.x].ns = (long long)(t1 - t0); }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
  }
}
template <int N, int NM, int NACC, int LANES>
static void run_lean(int stages, Res* dres) {
  const int kblocks = 4096, grid = 148;
  const int stage_bytes = (NM / 4) * 16384 + N * 128;
  const size_t smem = (size_t)stages * stage_bytes + 512 + 1024;
  cudaFuncSetAttribute(lean_kernel<N, NM, NACC, LANES>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  for (int rep = 0; rep < 2; ++rep) {
    lean_kernel<N, NM, NACC, LANES><<<grid, 128, smem>>>(kblocks, stages, dres);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("lean kernel failed: %s\n", cudaGetErrorString(e)); exit(1); }
  }
  Res h[148];
  cudaMemcpy(h, dres, sizeof(Res) * 148, cudaMemcpyDeviceToHost);
  long long cyc = 0;
  for (int i = 0; i < grid; ++i) if (h[i].cycles > cyc) cyc = h[i].cycles;
  printf("lean N=%3d mma/kblock=%d accumulators=%d lanes=%2d stages=%d  cyc/kblock=%7.1f  (pipe floor %5.0f)  -> %4.1f%% of pipe\n", N, NM, NACC, LANES,
         stages, (double)cyc / kblocks, NM * 128.0 * N / 256.0, 100.0 * NM * 128.0 * N / 256.0 / ((double)cyc / kblocks));
}

template <int CG>
static void run(int N, int grid, int traffic, int bytes_per_copy, const uint8_t* gsrc, Res* dres) {
  const int slots = 4, kblocks = 4096;
  const int slot_bytes = 16384 + (N / CG) * 128;
  const size_t smem = (size_t)slots * slot_bytes + 2 * 32768 + 256 + 1024;
  cudaFuncSetAttribute(rate_kernel<CG>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  cudaMemset(dres, 0, sizeof(Res) * 148);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(128);
  cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = CG;
  at[0].val.clusterDim.y = 1;
  at[0].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  for (int rep = 0; rep < 2; ++rep) {
    cudaError_t e = cudaLaunchKernelEx(&cfg, rate_kernel<CG>, N, kblocks, slots, traffic, bytes_per_copy, gsrc, dres);
    if (e != cudaSuccess) { printf("launch failed: %s\n", cudaGetErrorString(e)); return; }
    e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("kernel failed: %s\n", cudaGetErrorString(e)); exit(1); }
  }
  Res h[148];
  cudaMemcpy(h, dres, sizeof(Res) * 148, cudaMemcpyDeviceToHost);
  // report the slowest leader CTA
  long long cyc = 0, ns = 0, cp = 0;
  for (int i = 0; i < grid; i += CG) {
    if (h[i].cycles > cyc) { cyc = h[i].cycles; ns = h[i].ns; }
    cp += h[i].copies + (CG == 2 ? h[i + 1].copies : 0);
  }
  const double mmas = 4.0 * kblocks;
  const double flops = mmas * 2.0 * 128 * CG * N * 16 * (grid / CG);
  const double floor_cyc = 128.0 * N / 256.0;  // guide: max(M,128)*N/(256*cg) with M = 128*cg
  printf("cg=%d N=%3d grid=%3d traffic=%d(%5d B)  cyc/mma=%7.1f (floor %5.1f, x%4.2f)  clk=%.3f GHz  chip %7.1f TF/s  smem MMA-read %5.1f B/clk  copy-write %5.1f B/clk/SM\n",
         CG, N, grid, traffic, bytes_per_copy, cyc / mmas, floor_cyc, cyc / mmas / floor_cyc, (double)cyc / ns, flops / ns / 1e3,
         (4096.0 + (N / CG) * 32.0) / (cyc / mmas), (double)cp * bytes_per_copy / grid / cyc);
}

int main(int argc, char** argv) {
  uint8_t* gsrc;
  Res* dres;
  const int max_blocks = 40000;  // 40000 x 48 KB = 1.97 GB  (>> L2)
  cudaMalloc(&gsrc, (size_t)max_blocks * 49152 + 49152);
  cudaMemset(gsrc, 0x3c, (size_t)max_blocks * 49152 + 49152);
  cudaMalloc(&dres, sizeof(Res) * 148);
  if (argc > 1 && argv[1][0] == 'r') {
    const int Ns[3] = {64, 128, 256};
    for (int grid : {2, 148})
      for (int traffic : {0, 1})
        for (int N : Ns) {
          run<1>(N, grid, traffic, 32768, gsrc, dres);
          run<2>(N, grid, traffic, 32768, gsrc, dres);
        }
  }
  if (argc > 1 && argv[1][0] == 'p') {
    for (int N : {128, 256})
      for (int stages : {2, 3, 4, 5, 6}) {
        run_pipe(N, stages, 0, 0, 1, gsrc, dres);       // skeleton only
        run_pipe(N, stages, 0, 1, 1, gsrc, dres);       // + MMAs
        run_pipe(N, stages, 1, 0, 1000, gsrc, dres);    // copies only, L2-resident source (49 MB)
        run_pipe(N, stages, 1, 1, 1000, gsrc, dres);    // copies (L2) + MMAs
        run_pipe(N, stages, 1, 1, max_blocks, gsrc, dres);  // copies (HBM) + MMAs
      }
  }
  run_lean<128, 4, 1, 32>(4, dres);
  run_lean<128, 4, 1, 1>(4, dres);
  run_lean<128, 8, 2, 32>(3, dres);
  run_lean<128, 8, 2, 1>(3, dres);
  run_lean<128, 8, 1, 1>(3, dres);
  run_lean<256, 4, 1, 32>(3, dres);
  run_lean<256, 4, 1, 1>(3, dres);
  run_lean<64, 8, 2, 1>(4, dres);
  if (argc > 1 && argv[1][0] == 'i') for (int period : {1, 2, 4, 8, 16, 64}) for (int m : {1, 7, 33}) run_issue(m, dres, period);
  if (argc > 1 && argv[1][0] == 'v')
  // which part of the MMA warp's loop costs the ~288 cycles per K block?  variant bits: 1 = plain arrive instead of
  // tcgen05.commit, 2 = no tcgen05.fence::after_thread_sync, 4 = single-thread loop (no warp-wide wait / __syncwarp)
  for (int v = 0; v < 8; ++v) {
    run_pipe(128, 4, 0, 0, 1, gsrc, dres, v);
    run_pipe(128, 4, 0, 1, 1, gsrc, dres, v);
    run_pipe(128, 4, 1, 1, 1000, gsrc, dres, v);
  }
  return 0;
}
