// Microbenchmark: cycles per tcgen05.mma for the shapes / operand sources the kNN filter could use (development aid).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_rate mma_rate.cu ; run on a B200.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(c) : "memory"); }
__device__ __forceinline__ bool mbar_try(uint32_t bar, uint32_t par) {
  uint32_t ok;
  asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0,1,0,p;\n}\n" : "=r"(ok) : "r"(bar), "r"(par) : "memory");
  return ok;
}
__device__ __forceinline__ void commit(uint32_t bar) { asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ void expect_tx(uint32_t bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory"); }
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

template <int KIND>  // 0 tf32, 1 bf16
__device__ __forceinline__ void mma_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  if (KIND == 0)
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}\n" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
  else
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
template <int KIND>
__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  if (KIND == 0)
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n}\n" ::"r"(d), "r"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
  else
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n}\n" ::"r"(d), "r"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}

// descriptor: swz 0 = none (K-major interleaved, LBO = rows*16 B, SBO = 128 B), 1 = 128B swizzle (SBO = 1024 B)
__device__ __forceinline__ uint64_t desc(uint32_t addr, int rows, int swz) {
  if (swz == 0)
    return (uint64_t)((addr >> 4) & 0x3fff) | ((uint64_t)((rows * 16) >> 4) << 16) | ((uint64_t)(128 >> 4) << 32) | (1ull << 46);
  return (uint64_t)((addr >> 4) & 0x3fff) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}

__device__ __forceinline__ bool elect_one() {
  uint32_t p;
  asm volatile("{\n.reg .pred P;\nelect.sync _|P, 0xffffffff;\nselp.u32 %0,1,0,P;\n}\n" : "=r"(p));
  return p != 0;
}

struct Cfg { int kind, N, ts, swz, tma, iters; };

template <int KIND, int TS>
__global__ void __launch_bounds__(128, 1) k_rate(Cfg c, const unsigned char* src, long long* out) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ uint64_t bars[4];
  __shared__ uint32_t tslot;
  const int tid = threadIdx.x, warp = tid >> 5;
  // smem: A 32 KB (128 rows x 64 k x 4 B), B 64 KB (256 rows), TMA landing zone 64 KB
  unsigned char* sA = smem;
  unsigned char* sB = smem + 32768;
  unsigned char* sT = smem + 32768 + 65536;
  for (int i = tid; i < (32768 + 65536) / 4; i += 128) reinterpret_cast<float*>(smem)[i] = 0.001f * (float)((i * 37) % 101);
  if (tid == 0) {
    mbar_init(smem_u32(&bars[0]), 1);
    mbar_init(smem_u32(&bars[1]), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tslot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tm = tslot;
  const int esz = KIND == 0 ? 4 : 2;
  const int kper = 32 / esz;            // K per instruction
  const uint32_t idesc = (1u << 4) | ((KIND == 0 ? 2u : 1u) << 7) | ((KIND == 0 ? 2u : 1u) << 10) | ((uint32_t)(c.N >> 3) << 17) | (8u << 24);
  if (warp == 1) {
    const bool leader = elect_one();
    const long long t0 = clock64();
    uint32_t tma_phase = 0;
    const uint64_t da0 = desc(smem_u32(sA), 128, c.swz), db0 = desc(smem_u32(sB), c.N, c.swz);
    const uint32_t stepA = c.swz == 0 ? (2u * 128u * 16u) >> 4 : 32u >> 4;   // descriptor units per K step (first 4 steps)
    const uint32_t stepB = c.swz == 0 ? (2u * (uint32_t)c.N * 16u) >> 4 : 32u >> 4;
    const uint32_t ta0 = tm + 512 - 64;
    const int tiles = c.iters / 21;
    for (int t = 0; t < tiles; ++t) {
      if (c.tma) {
        if (t > 0) { while (!mbar_try(smem_u32(&bars[1]), tma_phase)) {} tma_phase ^= 1; }
        if (leader) {
          expect_tx(smem_u32(&bars[1]), 57344);
          const unsigned char* sp = src + ((size_t)((blockIdx.x * 7 + t) & 4095)) * 57344;
#pragma unroll
          for (int p = 0; p < 4; ++p) bulk_g2s(smem_u32(sT) + p * 14336, sp + p * 14336, 14336, smem_u32(&bars[1]));
        }
        __syncwarp();
      }
      const uint32_t dcol = tm + (t & 1) * (c.N);
#pragma unroll
      for (int g = 0; g < 3; ++g) {
#pragma unroll
        for (int k = 0; k < 7; ++k) {
          if (leader) {
            if (TS) mma_ts<KIND>(dcol, ta0 + k * 8, db0 + (uint64_t)(k * stepB), idesc, (g | k) != 0);
            else mma_ss<KIND>(dcol, da0 + (uint64_t)(k * stepA), db0 + (uint64_t)(k * stepB), idesc, (g | k) != 0);
          }
        }
      }
    }
    if (leader) commit(smem_u32(&bars[0]));
    __syncwarp();
    while (!mbar_try(smem_u32(&bars[0]), 0)) {}
    const long long t1 = clock64();
    if (c.tma) { while (!mbar_try(smem_u32(&bars[1]), tma_phase)) {} }
    if (leader) out[blockIdx.x] = t1 - t0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(512u) : "memory");
}

int main() {
  unsigned char* src;
  long long* out;
  cudaMalloc(&src, (size_t)4096 * 57344 + 65536);
  cudaMemset(src, 0, (size_t)4096 * 57344 + 65536);
  cudaMalloc(&out, 148 * sizeof(long long));
  const int smem = 32768 + 65536 + 65536 + 1024;
  cudaFuncSetAttribute(k_rate<0, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaFuncSetAttribute(k_rate<0, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaFuncSetAttribute(k_rate<1, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaFuncSetAttribute(k_rate<1, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const int iters = 21 * 2000;
  const Cfg cfgs[] = {
      {0, 128, 0, 0, 0, iters}, {0, 128, 0, 0, 1, iters}, {0, 256, 0, 0, 0, iters}, {0, 256, 0, 0, 1, iters},
      {0, 128, 1, 0, 0, iters}, {0, 128, 1, 0, 1, iters}, {0, 256, 1, 0, 0, iters},
      {0, 128, 0, 1, 0, iters}, {0, 128, 0, 1, 1, iters}, {0, 256, 0, 1, 0, iters}, {0, 256, 0, 1, 1, iters},
      {1, 128, 0, 0, 0, iters}, {1, 128, 0, 0, 1, iters}, {1, 256, 0, 0, 0, iters}, {1, 256, 0, 0, 1, iters},
      {1, 128, 0, 1, 0, iters}, {1, 256, 0, 1, 1, iters}, {1, 128, 1, 0, 0, iters}, {0, 64, 0, 0, 0, iters},
  };
  for (const Cfg& c : cfgs) {
    for (int grid : {148}) {
      cudaEvent_t e0, e1;
      cudaEventCreate(&e0); cudaEventCreate(&e1);
      cudaEventRecord(e0);
      if (c.kind == 0 && !c.ts) k_rate<0, 0><<<grid, 128, smem>>>(c, src, out);
      else if (c.kind == 0) k_rate<0, 1><<<grid, 128, smem>>>(c, src, out);
      else if (!c.ts) k_rate<1, 0><<<grid, 128, smem>>>(c, src, out);
      else k_rate<1, 1><<<grid, 128, smem>>>(c, src, out);
      cudaEventRecord(e1);
      cudaError_t err = cudaDeviceSynchronize();
      float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
      long long h[148];
      cudaMemcpy(h, out, grid * sizeof(long long), cudaMemcpyDeviceToHost);
      long long mx = 0; for (int i = 0; i < grid; ++i) mx = h[i] > mx ? h[i] : mx;
      const double macs = 128.0 * c.N * (c.kind == 0 ? 8 : 16);
      printf("%s N=%3d %s %s tma=%d grid=%3d : %.1f cycles/MMA  (%.0f MAC/clk/SM)  kernel %.3f ms -> %.2f GHz eff  %s\n", c.kind ? "bf16" : "tf32", c.N,
             c.ts ? "A=TMEM" : "A=smem", c.swz ? "sw128 " : "noswz ", c.tma, grid, (double)mx / c.iters, macs * c.iters / (double)mx, ms,
             (double)mx / (ms * 1e6), err == cudaSuccess ? "" : cudaGetErrorString(err));
    }
  }
  return 0;
}
