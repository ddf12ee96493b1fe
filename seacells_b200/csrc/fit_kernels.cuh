// Device kernels of the Frank-Wolfe fit loop (cpu.py:425-536 restated for compressed A / B).
//
// Arithmetic contract (SURVEY.md §7 hard part 3): the simplex step  A <- A + gamma (e - A)  is evaluated as three
// separately rounded fp64 operations  s = e - a;  p = gamma * s;  a = a + p  with gamma = 2.0 / (t + 2.0), exactly as
// scipy / numpy evaluate `A += 2.0 / (t + 2.0) * (e - A)` (cpu.py:456, 494).  No FMA contraction on that path.
#pragma once
#include "common.cuh"

__device__ __forceinline__ double fw_step(double a, double e, double gamma) {
  return __dadd_rn(a, __dmul_rn(gamma, __dsub_rn(e, a)));
}
__device__ __forceinline__ double fw_gamma(int t) { return __ddiv_rn(2.0, __dadd_rn((double)t, 2.0)); }

// lexicographic (value, index) minimum across a warp; all lanes get the result
__device__ __forceinline__ void warp_argmin(double& v, int& i) {
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) {
    double ov = __shfl_xor_sync(0xffffffffu, v, d);
    int oi = __shfl_xor_sync(0xffffffffu, i, d);
    if (ov < v || (ov == v && oi < i)) {
      v = ov;
      i = oi;
    }
  }
}

// fixed-order warp sum (tree over lanes), result on all lanes
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) v = __dadd_rn(v, __shfl_xor_sync(0xffffffffu, v, d));
  return v;
}

constexpr int FWA_NC = 128;     // max candidates per cell on the shared-memory path
constexpr int FWA_WARPS = 8;
constexpr int FW_SMAX = 64;     // max slots per column supported by the kernels (>= max_FW_iter)
constexpr double FW_INF = 1.0e300;
