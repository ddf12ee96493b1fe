// Shared declarations for the seacells_b200 CUDA library (sm_100a only).
//
// Vocabulary (follows the reference):
//   cell       - a row of the embedding / of the kernel matrix M (n of them)
//   archetype  - a SEACell (k of them); column of B, row of A
//   M          - kernel_matrix, n x n symmetric CSR (build_graph.py:188); K = M M^T is never formed by the engine
//   row lists  - per-cell sparse rows keyed by archetype id ("CSR with slack": ptr = reserved start, len = used)
//   slots      - the compressed A (per cell) / B (per archetype): <= S (index, weight) pairs per column
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>
#include <vector>

#define SC_OK 0
#define SC_ERR_CUDA 1
#define SC_ERR_ARG 2
#define SC_ERR_STATE 3
#define SC_ERR_CAPACITY 4
#define SC_ERR_NCCL 5

struct sc_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  std::string err;
  struct sc_comm* comm = nullptr;  // set when this context is one rank of a row-sharded job (comm.cuh); owned by the ctx
  int rank() const;
  int world() const;
  // launch counter: every kernel launched by the library goes through SC_LAUNCH and bumps this
  unsigned long long launches = 0;
  int last_knn_fallback = 0;  // queries of the last kNN call that failed the filter certificate (recomputed exactly)
  float last_knn_filter_ms = 0.f;  // device time of the last shortlist kernel (CUDA events on the library stream)
  int last_knn_filter_kind = 0;    // 0 none (all-fp64 kernel), 1 fp32 FFMA tiles, 2 tcgen05 split-fp16
  int last_knn_kp = 0;             // padded K of the tensor-core filter
};

#define SC_CUDA(ctx, call)                                                                  \
  do {                                                                                      \
    cudaError_t e__ = (call);                                                               \
    if (e__ != cudaSuccess) {                                                               \
      char b__[512];                                                                        \
      snprintf(b__, sizeof b__, "%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__)); \
      (ctx)->err = b__;                                                                     \
      return SC_ERR_CUDA;                                                                   \
    }                                                                                       \
  } while (0)

#define SC_TRY(expr)            \
  do {                          \
    int r__ = (expr);           \
    if (r__ != SC_OK) return r__; \
  } while (0)

#define SC_FAIL(ctx, code, msg) \
  do {                          \
    (ctx)->err = (msg);         \
    return (code);              \
  } while (0)

// Makes the context's device current for the duration of an extern "C" call and restores the caller's device on exit
// (the caller - torch, another model on another GPU - may have switched devices between calls).
struct DeviceGuard {
  int prev = -1;
  bool switched = false;
  explicit DeviceGuard(int device) {
    if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
    if (prev != device) switched = (cudaSetDevice(device) == cudaSuccess);
  }
  ~DeviceGuard() {
    if (switched && prev >= 0) cudaSetDevice(prev);
  }
  DeviceGuard(const DeviceGuard&) = delete;
  DeviceGuard& operator=(const DeviceGuard&) = delete;
};

// kernel launch + bookkeeping (+ immediate launch-error check)
#define SC_LAUNCH(ctx, kern, grid, block, smem, ...)                        \
  do {                                                                      \
    kern<<<(grid), (block), (smem), (ctx)->stream>>>(__VA_ARGS__);          \
    (ctx)->launches++;                                                      \
    SC_CUDA(ctx, cudaGetLastError());                                       \
  } while (0)

// Grow-only device buffer.
template <typename T>
struct DBuf {
  T* p = nullptr;
  size_t cap = 0;
  ~DBuf() { release(); }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
  }
  int ensure(sc_ctx* ctx, size_t n) {
    if (n <= cap) return SC_OK;
    // buffers may still be in use by queued kernels: drain the stream before freeing
    if (p) {
      SC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
      cudaFree(p);
      p = nullptr;
      cap = 0;
    }
    size_t want = n + n / 2 + 64;  // generous slack: list sizes creep up during the Frank-Wolfe loop
    SC_CUDA(ctx, cudaMalloc((void**)&p, want * sizeof(T)));
    cap = want;
    return SC_OK;
  }
};

// Per-cell (or per-archetype) sparse rows: entries of row r live at [ptr[r], ptr[r]+len[r]).
struct RowLists {
  const long long* ptr;
  const int* len;
  const int* key;
  const double* val;
};

struct RowListsBuf {
  DBuf<long long> ptr;
  DBuf<int> len;
  DBuf<int> key;
  DBuf<double> val;
  size_t rows = 0;
  RowLists view() const { return RowLists{ptr.p, len.p, key.p, val.p}; }
};

struct CsrView {
  int n;
  const int* indptr;
  const int* indices;
  const double* data;
};

static inline bool sc_is_device_ptr(const void* p) {
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  return a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged;
}

// copy that accepts host or device on either side
static inline cudaError_t sc_copy(void* dst, const void* src, size_t bytes, cudaStream_t s) {
  if (bytes == 0) return cudaSuccess;
  return cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDefault, s);
}

__device__ __forceinline__ int lane_id() { return threadIdx.x & 31; }

// scan / sort plumbing (CUB), defined in util.cu
int sc_exclusive_scan_i64(sc_ctx* ctx, const long long* in, long long* out, size_t n, DBuf<char>& tmp);
int sc_exclusive_scan_i32_to_i64(sc_ctx* ctx, const int* in, long long* out, size_t n, DBuf<char>& tmp);
int sc_sort_pairs_u64_f64(sc_ctx* ctx, unsigned long long* keys_in, unsigned long long* keys_out, double* vals_in,
                          double* vals_out, size_t n, int end_bit, DBuf<char>& tmp, int begin_bit = 0);
int sc_sort_pairs_u64_u32(sc_ctx* ctx, unsigned long long* keys_in, unsigned long long* keys_out, unsigned* vals_in,
                          unsigned* vals_out, size_t n, int end_bit, DBuf<char>& tmp);
int sc_sort_pairs_u32_u32(sc_ctx* ctx, unsigned* keys_in, unsigned* keys_out, unsigned* vals_in, unsigned* vals_out,
                          size_t n, int end_bit, DBuf<char>& tmp);
int sc_sort_keys_u64(sc_ctx* ctx, unsigned long long* keys_in, unsigned long long* keys_out, size_t n, int end_bit,
                     DBuf<char>& tmp);
int sc_sum_f64(sc_ctx* ctx, const double* in, double* out, size_t n, DBuf<char>& tmp);
