// Microbenchmark for the next step of the kNN shortlist kernel (knn_tc.cu): how fast can all SMs stream the same
// sequence of 32 KB reference tiles from L2 into shared memory,
//   (a) every CTA on its own (cp.async.bulk, what the kernel does today: 148 x 32 KB per tile step), versus
//   (b) CTA pairs (cluster of 2) where each CTA fetches half of every tile and multicasts it to both
//       (cp.async.bulk ... .multicast::cluster): half the L2 reads for the same shared-memory fill.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tile_stream tile_stream.cu ; run on a B200.
// Every wait is bounded (about 2 s) so a protocol mistake ends in an error line, not a hung GPU.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

constexpr int TILE = 32768;
constexpr int STAGES = 4;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(c) : "memory"); }
__device__ __forceinline__ void expect_tx(uint32_t bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory"); }
__device__ __forceinline__ bool try_wait(uint32_t bar, uint32_t par) {
  uint32_t ok;
  asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0,1,0,p;\n}\n" : "=r"(ok) : "r"(bar), "r"(par) : "memory");
  return ok;
}
__device__ __forceinline__ bool wait(uint32_t bar, uint32_t par) {
  const long long t0 = clock64();
  while (!try_wait(bar, par))
    if (clock64() - t0 > 4000000000ll) return false;
  return true;
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void bulk_g2s_mc(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar, uint16_t mask) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar), "h"(mask) : "memory");
}
__device__ __forceinline__ uint32_t cluster_rank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// arrive on the same barrier of the peer CTA (the consumer of a stage tells BOTH producers that its copy is free)
__device__ __forceinline__ void remote_arrive(uint32_t bar, uint32_t peer) {
  uint32_t raddr;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(raddr) : "r"(bar), "r"(peer));
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(raddr) : "memory");
}

// mode 0: unicast.  mode 1: pairs with multicast (launch with cluster dimension 2).
template <int MODE>
__global__ void __launch_bounds__(128, 1) k_stream(const unsigned char* __restrict__ src, int n_tiles, int iters,
                                                   long long* out, int* err) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ uint64_t full[STAGES], empty[STAGES];
  const int tid = threadIdx.x;
  const uint32_t rank = MODE ? cluster_rank() : 0u;
  if (tid == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(smem_u32(&full[s]), 1);
      mbar_init(smem_u32(&empty[s]), MODE ? 2 : 1);  // both CTAs of a pair must have released a stage
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (MODE) cluster_sync();
  long long t0 = 0, t1 = 0;
  bool ok = true;
  if (tid == 0) {
    // producer
    t0 = clock64();
    for (int i = 0; i < iters && ok; ++i) {
      const int s = i % STAGES;
      const uint32_t ph = (i / STAGES) & 1u;
      ok = wait(smem_u32(&empty[s]), ph ^ 1u);
      if (!ok) break;
      expect_tx(smem_u32(&full[s]), TILE);
      const unsigned char* t = src + (size_t)(i % n_tiles) * TILE;
      const uint32_t dst = smem_u32(smem) + s * TILE;
      if (MODE == 0) {
        for (int p = 0; p < 4; ++p) bulk_g2s(dst + p * (TILE / 4), t + p * (TILE / 4), TILE / 4, smem_u32(&full[s]));
      } else {
        // this CTA fetches its half of the tile and multicasts it into both CTAs (same offsets, same barrier offset)
        const uint32_t off = rank * (TILE / 2);
        for (int p = 0; p < 2; ++p)
          bulk_g2s_mc(dst + off + p * (TILE / 4), t + off + p * (TILE / 4), TILE / 4, smem_u32(&full[s]), (uint16_t)0x3);
      }
    }
  } else if (tid == 32) {
    // consumer: waits for the tile, "uses" it, releases the stage (to both producers in pair mode)
    for (int i = 0; i < iters && ok; ++i) {
      const int s = i % STAGES;
      const uint32_t ph = (i / STAGES) & 1u;
      ok = wait(smem_u32(&full[s]), ph);
      if (!ok) break;
      if (MODE == 0) {
        asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&empty[s])) : "memory");
      } else {
        remote_arrive(smem_u32(&empty[s]), 0);
        remote_arrive(smem_u32(&empty[s]), 1);
      }
    }
    t1 = clock64();
    out[blockIdx.x] = t1;
  }
  if (tid == 0) out[gridDim.x + blockIdx.x] = t0;
  if (!ok) atomicExch(err, 1);
  __syncthreads();
  if (MODE) cluster_sync();  // no CTA may exit while its peer can still write into its shared memory
}

int main() {
  const int n_tiles = 8192;  // 256 MB of tiles: larger than L2, as the 1M-cell reference set is (512 MB)
  unsigned char* src;
  long long* out;
  int* err;
  cudaMalloc(&src, (size_t)n_tiles * TILE);
  cudaMemset(src, 1, (size_t)n_tiles * TILE);
  cudaMalloc(&out, 2 * 148 * sizeof(long long));
  cudaMalloc(&err, sizeof(int));
  const int smem = STAGES * TILE + 1024;
  cudaFuncSetAttribute(k_stream<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaFuncSetAttribute(k_stream<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const int iters = 20000;
  for (int mode = 0; mode < 2; ++mode) {
    cudaMemset(err, 0, sizeof(int));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(148);
    cfg.blockDim = dim3(128);
    cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = mode ? 2 : 1;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    cudaEventRecord(e0);
    cudaError_t le = mode ? cudaLaunchKernelEx(&cfg, k_stream<1>, (const unsigned char*)src, n_tiles, iters, out, err)
                          : cudaLaunchKernelEx(&cfg, k_stream<0>, (const unsigned char*)src, n_tiles, iters, out, err);
    cudaEventRecord(e1);
    cudaError_t se = cudaDeviceSynchronize();
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    int herr = 0;
    cudaMemcpy(&herr, err, sizeof(int), cudaMemcpyDeviceToHost);
    const double filled = 148.0 * iters * TILE;           // bytes written into shared memory, all SMs
    const double l2 = filled / (mode ? 2.0 : 1.0);        // bytes read from L2
    printf("%s: %.3f ms  smem fill %.2f TB/s  L2 reads %.2f TB/s  (%s / %s / timeout flag %d)\n",
           mode ? "pair multicast" : "unicast       ", ms, filled / ms / 1e9, l2 / ms / 1e9, cudaGetErrorString(le),
           cudaGetErrorString(se), herr);
  }
  return 0;
}
