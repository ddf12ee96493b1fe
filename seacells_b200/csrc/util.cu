// Scan / sort plumbing (CUB, shipped with the CUDA toolkit) and a fixed-order fp64 reduction.
#include <cub/cub.cuh>

#include "common.cuh"

int sc_exclusive_scan_i64(sc_ctx* ctx, const long long* in, long long* out, size_t n, DBuf<char>& tmp) {
  size_t bytes = 0;
  SC_CUDA(ctx, cub::DeviceScan::ExclusiveSum(nullptr, bytes, in, out, (long long)n, ctx->stream));
  SC_TRY(tmp.ensure(ctx, bytes));
  SC_CUDA(ctx, cub::DeviceScan::ExclusiveSum(tmp.p, bytes, in, out, (long long)n, ctx->stream));
  ctx->launches += 1;
  return SC_OK;
}

struct I32ToI64 {
  __host__ __device__ long long operator()(int v) const { return (long long)v; }
};

int sc_exclusive_scan_i32_to_i64(sc_ctx* ctx, const int* in, long long* out, size_t n, DBuf<char>& tmp) {
  cub::TransformInputIterator<long long, I32ToI64, const int*> it(in, I32ToI64());
  size_t bytes = 0;
  SC_CUDA(ctx, cub::DeviceScan::ExclusiveSum(nullptr, bytes, it, out, (long long)n, ctx->stream));
  SC_TRY(tmp.ensure(ctx, bytes));
  SC_CUDA(ctx, cub::DeviceScan::ExclusiveSum(tmp.p, bytes, it, out, (long long)n, ctx->stream));
  ctx->launches += 1;
  return SC_OK;
}

int sc_sort_pairs_u64_f64(sc_ctx* ctx, unsigned long long* keys_in, unsigned long long* keys_out, double* vals_in,
                          double* vals_out, size_t n, int end_bit, DBuf<char>& tmp, int begin_bit) {
  size_t bytes = 0;
  SC_CUDA(ctx, cub::DeviceRadixSort::SortPairs(nullptr, bytes, keys_in, keys_out, vals_in, vals_out, (long long)n,
                                               begin_bit, end_bit, ctx->stream));
  SC_TRY(tmp.ensure(ctx, bytes));
  SC_CUDA(ctx, cub::DeviceRadixSort::SortPairs(tmp.p, bytes, keys_in, keys_out, vals_in, vals_out, (long long)n,
                                               begin_bit, end_bit, ctx->stream));
  ctx->launches += 1;
  return SC_OK;
}

int sc_sort_pairs_u64_u32(sc_ctx* ctx, unsigned long long* keys_in, unsigned long long* keys_out, unsigned* vals_in,
                          unsigned* vals_out, size_t n, int end_bit, DBuf<char>& tmp) {
  size_t bytes = 0;
  SC_CUDA(ctx, cub::DeviceRadixSort::SortPairs(nullptr, bytes, keys_in, keys_out, vals_in, vals_out, (long long)n, 0,
                                               end_bit, ctx->stream));
  SC_TRY(tmp.ensure(ctx, bytes));
  SC_CUDA(ctx, cub::DeviceRadixSort::SortPairs(tmp.p, bytes, keys_in, keys_out, vals_in, vals_out, (long long)n, 0,
                                               end_bit, ctx->stream));
  ctx->launches += 1;
  return SC_OK;
}

int sc_sort_pairs_u32_u32(sc_ctx* ctx, unsigned* keys_in, unsigned* keys_out, unsigned* vals_in, unsigned* vals_out,
                          size_t n, int end_bit, DBuf<char>& tmp) {
  size_t bytes = 0;
  SC_CUDA(ctx, cub::DeviceRadixSort::SortPairs(nullptr, bytes, keys_in, keys_out, vals_in, vals_out, (long long)n, 0,
                                               end_bit, ctx->stream));
  SC_TRY(tmp.ensure(ctx, bytes));
  SC_CUDA(ctx, cub::DeviceRadixSort::SortPairs(tmp.p, bytes, keys_in, keys_out, vals_in, vals_out, (long long)n, 0,
                                               end_bit, ctx->stream));
  ctx->launches += 1;
  return SC_OK;
}

int sc_sort_keys_u64(sc_ctx* ctx, unsigned long long* keys_in, unsigned long long* keys_out, size_t n, int end_bit,
                     DBuf<char>& tmp) {
  size_t bytes = 0;
  SC_CUDA(ctx, cub::DeviceRadixSort::SortKeys(nullptr, bytes, keys_in, keys_out, (long long)n, 0, end_bit, ctx->stream));
  SC_TRY(tmp.ensure(ctx, bytes));
  SC_CUDA(ctx, cub::DeviceRadixSort::SortKeys(tmp.p, bytes, keys_in, keys_out, (long long)n, 0, end_bit, ctx->stream));
  ctx->launches += 1;
  return SC_OK;
}

// ---- fixed-order fp64 sum: per-thread strided partials -> block tree -> one final block. Bitwise reproducible.
constexpr int RED_BLOCK = 256;
constexpr int RED_GRID = 592;  // 4 x 148 SMs

__global__ void k_reduce_partial(const double* __restrict__ in, double* __restrict__ part, size_t n) {
  __shared__ double sh[RED_BLOCK];
  double acc = 0.0;
  for (size_t i = (size_t)blockIdx.x * RED_BLOCK + threadIdx.x; i < n; i += (size_t)RED_GRID * RED_BLOCK) acc += in[i];
  sh[threadIdx.x] = acc;
  __syncthreads();
  for (int s = RED_BLOCK / 2; s > 0; s >>= 1) {
    if (threadIdx.x < s) sh[threadIdx.x] += sh[threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x == 0) part[blockIdx.x] = sh[0];
}

__global__ void k_reduce_final(const double* __restrict__ part, double* __restrict__ out, int m) {
  __shared__ double sh[1024];
  double acc = 0.0;
  for (int i = threadIdx.x; i < m; i += 1024) acc += part[i];
  sh[threadIdx.x] = acc;
  __syncthreads();
  for (int s = 512; s > 0; s >>= 1) {
    if (threadIdx.x < s) sh[threadIdx.x] += sh[threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x == 0) out[0] = sh[0];
}

int sc_sum_f64(sc_ctx* ctx, const double* in, double* out, size_t n, DBuf<char>& tmp) {
  SC_TRY(tmp.ensure(ctx, RED_GRID * sizeof(double)));
  SC_LAUNCH(ctx, k_reduce_partial, RED_GRID, RED_BLOCK, 0, in, (double*)tmp.p, n);
  SC_LAUNCH(ctx, k_reduce_final, 1, 1024, 0, (const double*)tmp.p, out, RED_GRID);
  return SC_OK;
}
