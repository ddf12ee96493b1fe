// extern "C" surface declared in include/seacells_b200.h
#include <algorithm>
#include <cstdlib>
#include <new>

#include "../../include/seacells_b200.h"
#include "comm.cuh"
#include "common.cuh"
#include "fit.cuh"
#include "kernel_build.cuh"

int sc_summarize_dev(sc_ctx* ctx, int n, int n_genes, const long long* x_ptr, const int* x_col, const void* x_val,
                     int val_is_f64, long long m, const int* group, const int* cell, const double* weight, int n_groups,
                     const double* divisor, double* out);
int sc_greedy_centers_impl(sc_ctx* ctx, int n, const int* indptr, const int* indices, const double* data, long long nnz,
                           int n_centers, long long* centers_host);

namespace {

// stage a host-or-device input on the device; `own` keeps the staging buffer alive
template <typename T>
int to_device(sc_ctx* ctx, const T* src, size_t count, DBuf<T>& own, const T** out) {
  if (sc_is_device_ptr(src)) {
    *out = src;
    return SC_OK;
  }
  SC_TRY(own.ensure(ctx, count ? count : 1));
  SC_CUDA(ctx, sc_copy(own.p, src, sizeof(T) * count, ctx->stream));
  *out = own.p;
  return SC_OK;
}

// ---- dense-RHS CSR SpMM: one warp per (row, 256-column chunk), 8-byte lanes vectorised as double2
__global__ void __launch_bounds__(256)
k_spmm_csr_dense(int r, int c, const int* __restrict__ indptr, const int* __restrict__ indices,
                 const double* __restrict__ data, const double* __restrict__ X, double* __restrict__ Y, int chunks) {
  const long long w = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (w >= (long long)r * chunks) return;
  const int row = (int)(w / chunks), ch = (int)(w - (long long)row * chunks);
  const int c0 = ch * 256 + lane * 2;
  const int s0 = indptr[row], s1 = indptr[row + 1];
  double acc[4][2];
#pragma unroll
  for (int u = 0; u < 4; ++u) acc[u][0] = acc[u][1] = 0.0;
  const bool vec_ok = (c % 2 == 0);
  for (int base = s0; base < s1; base += 32) {
    int q = 0;
    double m = 0.0;
    if (base + lane < s1) {
      q = indices[base + lane];
      m = data[base + lane];
    }
    const int cnt = min(32, s1 - base);
    for (int e = 0; e < cnt; ++e) {
      const int qq = __shfl_sync(0xffffffffu, q, e);
      const double mm = __shfl_sync(0xffffffffu, m, e);
      const double* xr = X + (size_t)qq * c;
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int cc = c0 + u * 64;
        if (vec_ok && cc + 1 < c) {
          const double2 xv = *reinterpret_cast<const double2*>(xr + cc);
          acc[u][0] = fma(mm, xv.x, acc[u][0]);
          acc[u][1] = fma(mm, xv.y, acc[u][1]);
        } else {
          if (cc < c) acc[u][0] = fma(mm, xr[cc], acc[u][0]);
          if (cc + 1 < c) acc[u][1] = fma(mm, xr[cc + 1], acc[u][1]);
        }
      }
    }
  }
  double* yr = Y + (size_t)row * c;
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    const int cc = c0 + u * 64;
    if (vec_ok && cc + 1 < c) {
      *reinterpret_cast<double2*>(yr + cc) = make_double2(acc[u][0], acc[u][1]);
    } else {
      if (cc < c) yr[cc] = acc[u][0];
      if (cc + 1 < c) yr[cc + 1] = acc[u][1];
    }
  }
}

// ---- panel-ordered variant.  With ~24 stored entries per row every row of X is gathered ~24 times; X (8 c m bytes, 1 GB
// at the config-2 shape) does not fit the 126 MB L2, so in the row-major kernel above most of those gathers go to HBM.
// Here the CTAs are ordered panel-major: all rows for columns [0, W), then [W, 2W), ...  One panel of X (8 W m bytes:
// 51 MB for W = 64 at m = 100k) plus the CSR arrays stay L2-resident while the panel is processed, so HBM sees every
// operand about once (the algorithmic bytes of SURVEY.md 8d) and the re-reads are served by L2.
// One warp per (row, panel); lane l owns columns c0 + l + 32 j: every load instruction of a warp reads 256 contiguous
// bytes of one X row (c is odd at the config shapes - 133, 1333, 13333 - so rows are only 8-byte aligned and a double2
// mapping would split sectors).
template <int W>
__global__ void __launch_bounds__(256)
k_spmm_csr_panel(int r, int c, const int* __restrict__ indptr, const int* __restrict__ indices,
                 const double* __restrict__ data, const double* __restrict__ X, double* __restrict__ Y, int row_blocks) {
  constexpr int ROWS_PER_CTA = 64, U = W / 32;
  const int panel = blockIdx.x / row_blocks, rb = blockIdx.x - panel * row_blocks;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int c0 = panel * W + lane;
  const int row_end = min(r, (rb + 1) * ROWS_PER_CTA);
  for (int row = rb * ROWS_PER_CTA + warp; row < row_end; row += 8) {
    const int s0 = indptr[row], s1 = indptr[row + 1];
    double acc[U];
#pragma unroll
    for (int u = 0; u < U; ++u) acc[u] = 0.0;
    for (int base = s0; base < s1; base += 32) {
      int q = 0;
      double m = 0.0;
      if (base + lane < s1) {
        q = indices[base + lane];
        m = data[base + lane];
      }
      const int cnt = min(32, s1 - base);
#pragma unroll 8
      for (int e = 0; e < cnt; ++e) {
        const int qq = __shfl_sync(0xffffffffu, q, e);
        const double mm = __shfl_sync(0xffffffffu, m, e);
        const double* xr = X + (size_t)qq * c + c0;
#pragma unroll
        for (int u = 0; u < U; ++u)
          if (c0 + u * 32 < c) acc[u] = fma(mm, __ldg(xr + u * 32), acc[u]);
      }
    }
    double* yr = Y + (size_t)row * c + c0;
#pragma unroll
    for (int u = 0; u < U; ++u)
      if (c0 + u * 32 < c) yr[u * 32] = acc[u];
  }
}

}  // namespace

extern "C" {

const char* sc_version(void) { return "seacells_b200 0.1 (sm_100a)"; }

int sc_ctx_create(int device, sc_ctx** out) {
  if (!out) return SC_ERR_ARG;
  *out = nullptr;
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess || device < 0 || device >= count) {
    cudaGetLastError();
    return SC_ERR_CUDA;  // no usable CUDA device: there is no CPU fallback
  }
  DeviceGuard dev_guard(device);
  if (dev_guard.prev != device && !dev_guard.switched) return SC_ERR_CUDA;
  sc_ctx* c = new (std::nothrow) sc_ctx();
  if (!c) return SC_ERR_CUDA;
  c->device = device;
  if (cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess) {
    delete c;
    return SC_ERR_CUDA;
  }
  *out = c;
  return SC_OK;
}

// ---- row-sharded jobs: one context per rank ----------------------------------------------------------
int sc_nccl_load_library(const char* path) {
  std::string err;
  return sc_nccl_load(path, &err);
}
int sc_nccl_unique_id(void* out128) {
  if (!out128) return SC_ERR_ARG;
  std::string err;
  return sc_nccl_get_unique_id(out128, &err);
}
int sc_ctx_init_nccl(sc_ctx* ctx, int rank, int world, const void* id128) {
  if (!ctx || !id128 || world < 1 || rank < 0 || rank >= world) return SC_ERR_ARG;
  if (ctx->comm) SC_FAIL(ctx, SC_ERR_STATE, "context already belongs to a sharded job");
  DeviceGuard dev_guard(ctx->device);
  std::string err;
  sc_comm* c = sc_nccl_comm_create(rank, world, id128, &err);
  if (!c) SC_FAIL(ctx, SC_ERR_NCCL, err);
  ctx->comm = c;
  return SC_OK;
}
int sc_local_group_new(int world, void** out) {
  if (!out || world < 1) return SC_ERR_ARG;
  *out = sc_local_group_create(world);
  return *out ? SC_OK : SC_ERR_ARG;
}
void sc_local_group_free(void* group) { sc_local_group_destroy((sc_comm_group*)group); }
void sc_local_group_abort(void* group) { sc_local_group_break((sc_comm_group*)group); }
int sc_ctx_attach_local(sc_ctx* ctx, void* group, int rank) {
  if (!ctx || !group) return SC_ERR_ARG;
  if (ctx->comm) SC_FAIL(ctx, SC_ERR_STATE, "context already belongs to a sharded job");
  sc_comm* c = sc_local_comm_create((sc_comm_group*)group, rank);
  if (!c) SC_FAIL(ctx, SC_ERR_ARG, "bad rank for this group");
  ctx->comm = c;
  return SC_OK;
}
int sc_ctx_rank(sc_ctx* ctx) { return ctx ? ctx->rank() : 0; }
int sc_ctx_world(sc_ctx* ctx) { return ctx ? ctx->world() : 1; }

void sc_ctx_destroy(sc_ctx* ctx) {
  if (!ctx) return;
  DeviceGuard dev_guard(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  delete ctx->comm;
  ctx->comm = nullptr;
  cudaStreamDestroy(ctx->stream);
  delete ctx;
}

const char* sc_last_error(sc_ctx* ctx) { return ctx ? ctx->err.c_str() : "null context"; }
unsigned long long sc_launch_count(sc_ctx* ctx) { return ctx ? ctx->launches : 0ull; }
int sc_knn_fallback_rows(sc_ctx* ctx) { return ctx ? ctx->last_knn_fallback : 0; }
int sc_knn_profile(sc_ctx* ctx, double* out4) {
  if (!ctx || !out4) return SC_ERR_ARG;
  out4[0] = (double)ctx->last_knn_filter_ms;
  out4[1] = (double)ctx->last_knn_filter_kind;
  out4[2] = (double)ctx->last_knn_kp;
  out4[3] = (double)ctx->last_knn_fallback;
  return SC_OK;
}
void* sc_stream(sc_ctx* ctx) { return ctx ? (void*)ctx->stream : nullptr; }

int sc_sync(sc_ctx* ctx) {
  if (!ctx) return SC_ERR_ARG;
  DeviceGuard dev_guard(ctx->device);
  SC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return SC_OK;
}

// ------------------------------------------------------------------------------------------------ hot path 1
int sc_knn(sc_ctx* ctx, const void* X, int x_is_f64, int n, int d, int n_neighbors, int row_begin, int row_end,
           int32_t* idx_out, double* d2_out) {
  if (!ctx || !X || !idx_out || !d2_out || n <= 0 || d <= 0) return SC_ERR_ARG;
  DeviceGuard dev_guard(ctx->device);
  const size_t esz = x_is_f64 ? 8 : 4;
  DBuf<char> xs;
  const char* Xd = nullptr;
  SC_TRY(to_device<char>(ctx, (const char*)X, (size_t)n * d * esz, xs, &Xd));
  const int kk = n_neighbors - 1;
  const size_t rows = (size_t)(row_end - row_begin);
  const bool out_dev = sc_is_device_ptr(idx_out);
  DBuf<int> di;
  DBuf<double> dd;
  int* pi = idx_out;
  double* pd = d2_out;
  if (!out_dev) {
    SC_TRY(di.ensure(ctx, rows * kk + 1));
    SC_TRY(dd.ensure(ctx, rows * kk + 1));
    pi = di.p;
    pd = dd.p;
  }
  SC_TRY(sc_knn_exact_dev(ctx, Xd, x_is_f64, n, d, n_neighbors, row_begin, row_end, pi, pd));
  if (!out_dev) {
    SC_CUDA(ctx, sc_copy(idx_out, pi, sizeof(int) * rows * kk, ctx->stream));
    SC_CUDA(ctx, sc_copy(d2_out, pd, sizeof(double) * rows * kk, ctx->stream));
  }
  SC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return SC_OK;
}

int sc_knn_tc_probe(sc_ctx* ctx, const void* X, int x_is_f64, int n, int d, int cols, float* s_out, float* xc_out,
                    float* norm2_out) {
  if (!ctx || !X || !s_out || !xc_out || !norm2_out || n <= 0 || d <= 0 || cols <= 0) return SC_ERR_ARG;
  DeviceGuard dev_guard(ctx->device);
  const size_t esz = x_is_f64 ? 8 : 4;
  DBuf<char> xs;
  const char* Xd = nullptr;
  SC_TRY(to_device<char>(ctx, (const char*)X, (size_t)n * d * esz, xs, &Xd));
  DBuf<float> s, xc, nn;
  SC_TRY(s.ensure(ctx, (size_t)128 * cols));
  SC_TRY(xc.ensure(ctx, (size_t)n * d));
  SC_TRY(nn.ensure(ctx, (size_t)n));
  SC_CUDA(ctx, cudaMemsetAsync(s.p, 0, sizeof(float) * 128 * (size_t)cols, ctx->stream));
  SC_TRY(sc_knn_tc_probe_dev(ctx, Xd, x_is_f64, n, d, cols, s.p, xc.p, nn.p));
  SC_CUDA(ctx, sc_copy(s_out, s.p, sizeof(float) * 128 * (size_t)cols, ctx->stream));
  SC_CUDA(ctx, sc_copy(xc_out, xc.p, sizeof(float) * (size_t)n * d, ctx->stream));
  SC_CUDA(ctx, sc_copy(norm2_out, nn.p, sizeof(float) * (size_t)n, ctx->stream));
  SC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return SC_OK;
}

double sc_knn_tc_error_factor(int d) { return sc_knn_tc_gamma(d); }

int sc_summarize(sc_ctx* ctx, int n, int n_genes, const int64_t* indptr, const int32_t* indices, const void* data,
                 int data_is_f64, long long m, const int32_t* group, const int32_t* cell, const double* weight,
                 int n_groups, const double* divisor, double* out) {
  if (!ctx || !indptr || !out || n <= 0 || n_genes <= 0 || n_groups <= 0 || m < 0) return SC_ERR_ARG;
  if (m > 0 && (!group || !cell || !weight || !indices || !data)) return SC_ERR_ARG;
  DeviceGuard dev_guard(ctx->device);
  long long nnz = 0;
  {
    // indptr may be on the host or on the device
    SC_CUDA(ctx, cudaMemcpyAsync(&nnz, indptr + n, sizeof(long long), cudaMemcpyDefault, ctx->stream));
    SC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  }
  const size_t esz = data_is_f64 ? 8 : 4;
  DBuf<long long> s_ptr;
  DBuf<int> s_col, s_grp, s_cell;
  DBuf<char> s_val;
  DBuf<double> s_w, s_div, s_out;
  const long long* d_ptr = nullptr;
  const int *d_col = nullptr, *d_grp = nullptr, *d_cell = nullptr;
  const char* d_val = nullptr;
  const double *d_w = nullptr, *d_div = nullptr;
  SC_TRY(to_device<long long>(ctx, (const long long*)indptr, (size_t)n + 1, s_ptr, &d_ptr));
  SC_TRY(to_device<int>(ctx, indices, (size_t)nnz, s_col, &d_col));
  SC_TRY(to_device<char>(ctx, (const char*)data, (size_t)nnz * esz, s_val, &d_val));
  SC_TRY(to_device<int>(ctx, group, (size_t)m, s_grp, &d_grp));
  SC_TRY(to_device<int>(ctx, cell, (size_t)m, s_cell, &d_cell));
  SC_TRY(to_device<double>(ctx, weight, (size_t)m, s_w, &d_w));
  if (divisor) SC_TRY(to_device<double>(ctx, divisor, (size_t)n_groups, s_div, &d_div));
  const bool out_dev = sc_is_device_ptr(out);
  double* d_out = out;
  if (!out_dev) {
    SC_TRY(s_out.ensure(ctx, (size_t)n_groups * n_genes));
    d_out = s_out.p;
  }
  SC_TRY(sc_summarize_dev(ctx, n, n_genes, d_ptr, d_col, d_val, data_is_f64, m, d_grp, d_cell, d_w, n_groups, d_div,
                          d_out));
  if (!out_dev) {
    SC_CUDA(ctx, sc_copy(out, d_out, sizeof(double) * (size_t)n_groups * n_genes, ctx->stream));
    SC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  }
  return SC_OK;
}

static int kernel_build_impl(sc_ctx* ctx, const void* X, int x_is_f64, int n, int d, int n_neighbors,
                             const int32_t* idx, const double* d2, int intersection, sc_kernel** out,
                             long long* nnz_out) {
  if (!ctx || !X || !out || n <= 0 || d <= 0) return SC_ERR_ARG;
  if (intersection != 0 && intersection != 1) SC_FAIL(ctx, SC_ERR_ARG, "graph_construction must be union or intersection");
  DeviceGuard dev_guard(ctx->device);
  *out = nullptr;
  const size_t esz = x_is_f64 ? 8 : 4;
  DBuf<char> xs;
  const char* Xd = nullptr;
  SC_TRY(to_device<char>(ctx, (const char*)X, (size_t)n * d * esz, xs, &Xd));
  sc_kernel_graph* km = new (std::nothrow) sc_kernel_graph();
  if (!km) SC_FAIL(ctx, SC_ERR_CUDA, "out of host memory");
  const int kk = n_neighbors - 1;
  km->kk = kk;
  // a rank of a sharded job searches the neighbours of its own slab of query cells; the slabs are all-gathered on the
  // device (112 MB at 1M cells) and every rank then assembles the same symmetric matrix (build_graph.py:125-147)
  const int world = ctx->world();
  const int slab = (n + world - 1) / world;
  int rc = km->knn_idx.ensure(ctx, (size_t)slab * world * kk + 1);
  if (rc == SC_OK) rc = km->knn_d2.ensure(ctx, (size_t)slab * world * kk + 1);
  if (rc == SC_OK) {
    if (idx && d2) {
      if (cudaMemcpyAsync(km->knn_idx.p, idx, sizeof(int) * (size_t)n * kk, cudaMemcpyDefault, ctx->stream) != cudaSuccess ||
          cudaMemcpyAsync(km->knn_d2.p, d2, sizeof(double) * (size_t)n * kk, cudaMemcpyDefault, ctx->stream) != cudaSuccess) {
        ctx->err = "kernel_build: copying the kNN graph failed";
        rc = SC_ERR_CUDA;
      }
    } else if (world == 1) {
      rc = sc_knn_exact_dev(ctx, Xd, x_is_f64, n, d, n_neighbors, 0, n, km->knn_idx.p, km->knn_d2.p);
    } else {
      const int r0 = std::min(n, ctx->rank() * slab), r1 = std::min(n, r0 + slab);
      if (r1 > r0)
        rc = sc_knn_exact_dev(ctx, Xd, x_is_f64, n, d, n_neighbors, r0, r1, km->knn_idx.p + (size_t)r0 * kk,
                              km->knn_d2.p + (size_t)r0 * kk);
      if (rc == SC_OK) rc = ctx->comm->allgather(km->knn_idx.p, sizeof(int) * (size_t)slab * kk, ctx->stream);
      if (rc == SC_OK) rc = ctx->comm->allgather(km->knn_d2.p, sizeof(double) * (size_t)slab * kk, ctx->stream);
      if (rc != SC_OK && ctx->err.empty()) ctx->err = ctx->comm->err;
    }
  }
  if (rc == SC_OK) rc = km->build(ctx, Xd, x_is_f64, n, d, n_neighbors, km->knn_idx.p, km->knn_d2.p, intersection);
  if (rc != SC_OK) {
    cudaStreamSynchronize(ctx->stream);
    delete km;
    return rc;
  }
  if (nnz_out) *nnz_out = km->nnz;
  *out = km;
  return SC_OK;
}

int sc_kernel_build(sc_ctx* ctx, const void* X, int x_is_f64, int n, int d, int n_neighbors, int intersection,
                    sc_kernel** out, long long* nnz_out) {
  return kernel_build_impl(ctx, X, x_is_f64, n, d, n_neighbors, nullptr, nullptr, intersection, out, nnz_out);
}

int sc_kernel_build_from_knn(sc_ctx* ctx, const void* X, int x_is_f64, int n, int d, int n_neighbors,
                             const int32_t* idx, const double* d2, int intersection, sc_kernel** out,
                             long long* nnz_out) {
  if (!idx || !d2) return SC_ERR_ARG;
  return kernel_build_impl(ctx, X, x_is_f64, n, d, n_neighbors, idx, d2, intersection, out, nnz_out);
}

int sc_kernel_export(sc_ctx* ctx, sc_kernel* km, int32_t* indptr, int32_t* indices, double* data, double* sigma,
                     int32_t* knn_idx, double* knn_d2) {
  if (!ctx || !km) return SC_ERR_ARG;
  DeviceGuard dev_guard(ctx->device);
  const size_t n = km->n, nnz = km->nnz, kk = km->kk;
  if (indptr) SC_CUDA(ctx, sc_copy(indptr, km->indptr.p, sizeof(int) * (n + 1), ctx->stream));
  if (indices) SC_CUDA(ctx, sc_copy(indices, km->indices.p, sizeof(int) * nnz, ctx->stream));
  if (data) SC_CUDA(ctx, sc_copy(data, km->data.p, sizeof(double) * nnz, ctx->stream));
  if (sigma) SC_CUDA(ctx, sc_copy(sigma, km->sigma.p, sizeof(double) * n, ctx->stream));
  if (knn_idx) SC_CUDA(ctx, sc_copy(knn_idx, km->knn_idx.p, sizeof(int) * n * kk, ctx->stream));
  if (knn_d2) SC_CUDA(ctx, sc_copy(knn_d2, km->knn_d2.p, sizeof(double) * n * kk, ctx->stream));
  SC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return SC_OK;
}

void sc_kernel_free(sc_kernel* km) { delete km; }

int sc_spmm_csr_f64(sc_ctx* ctx, int r, int m, int c, const int32_t* indptr, const int32_t* indices,
                    const double* data, const double* X, double* Y) {
  if (!ctx || r <= 0 || m <= 0 || c <= 0 || !indptr || !indices || !data || !X || !Y) return SC_ERR_ARG;
  DeviceGuard dev_guard(ctx->device);
  int h_nnz = 0;
  SC_CUDA(ctx, cudaMemcpyAsync(&h_nnz, indptr + r, sizeof(int), cudaMemcpyDefault, ctx->stream));
  SC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  DBuf<int> sp, sc;
  DBuf<double> sv, sx, sy;
  const int *dp, *dc;
  const double *dv, *dx;
  SC_TRY(to_device<int>(ctx, indptr, (size_t)r + 1, sp, &dp));
  SC_TRY(to_device<int>(ctx, indices, (size_t)h_nnz, sc, &dc));
  SC_TRY(to_device<double>(ctx, data, (size_t)h_nnz, sv, &dv));
  SC_TRY(to_device<double>(ctx, X, (size_t)m * c, sx, &dx));
  const bool y_dev = sc_is_device_ptr(Y);
  double* dy = Y;
  if (!y_dev) {
    SC_TRY(sy.ensure(ctx, (size_t)r * c));
    dy = sy.p;
  }
  // SC_SPMM_PANEL = 0: row-major kernel; 64 / 128: panel width of the L2-blocked kernel (default 64)
  int panel_w = 64;
  if (const char* e = std::getenv("SC_SPMM_PANEL")) panel_w = atoi(e);
  if (panel_w == 64 || panel_w == 128) {
    const int row_blocks = (r + 63) / 64;
    const int panels = (c + panel_w - 1) / panel_w;
    const unsigned grid = (unsigned)((long long)row_blocks * panels);
    if (panel_w == 64) SC_LAUNCH(ctx, k_spmm_csr_panel<64>, grid, 256, 0, r, c, dp, dc, dv, dx, dy, row_blocks);
    else SC_LAUNCH(ctx, k_spmm_csr_panel<128>, grid, 256, 0, r, c, dp, dc, dv, dx, dy, row_blocks);
  } else {
    const int chunks = (c + 255) / 256;
    const long long warps = (long long)r * chunks;
    SC_LAUNCH(ctx, k_spmm_csr_dense, (unsigned)((warps * 32 + 255) / 256), 256, 0, r, c, dp, dc, dv, dx, dy, chunks);
  }
  if (!y_dev) SC_CUDA(ctx, sc_copy(Y, dy, sizeof(double) * (size_t)r * c, ctx->stream));
  SC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return SC_OK;
}

int sc_greedy_centers(sc_ctx* ctx, int n, const int32_t* indptr, const int32_t* indices, const double* data,
                      long long nnz, int n_centers, int64_t* centers_out) {
  if (!ctx || !indptr || !indices || !data || !centers_out) return SC_ERR_ARG;
  DeviceGuard dev_guard(ctx->device);
  return sc_greedy_centers_impl(ctx, n, indptr, indices, data, nnz, n_centers, (long long*)centers_out);
}

// ------------------------------------------------------------------------------------------------ hot path 2
int sc_fit_create(sc_ctx* ctx, int n_cells, int n_archetypes, int max_FW_iter, sc_fit** out) {
  if (!ctx || !out) return SC_ERR_ARG;
  *out = nullptr;
  DeviceGuard dev_guard(ctx->device);
  sc_fit* f = new (std::nothrow) sc_fit();
  if (!f) SC_FAIL(ctx, SC_ERR_CUDA, "out of host memory");
  int rc = f->init(ctx, n_cells, n_archetypes, max_FW_iter);
  if (rc != SC_OK) {
    cudaStreamSynchronize(ctx->stream);
    delete f;
    return rc;
  }
  *out = f;
  return SC_OK;
}

void sc_fit_destroy(sc_fit* fit) {
  if (!fit) return;
  DeviceGuard dev_guard(fit->ctx->device);
  cudaStreamSynchronize(fit->ctx->stream);
  delete fit;
}

int sc_fit_set_kernel(sc_fit* fit, const int32_t* indptr, const int32_t* indices, const double* data, long long nnz) {
  if (!fit || !indptr || !indices || !data || nnz <= 0) return SC_ERR_ARG;
  DeviceGuard dev_guard(fit->ctx->device);
  return fit->set_kernel(indptr, indices, data, nnz);
}

int sc_fit_set_kernel_from(sc_fit* fit, sc_kernel* km) {
  if (!fit || !km) return SC_ERR_ARG;
  DeviceGuard dev_guard(fit->ctx->device);
  if (km->n != fit->n) SC_FAIL(fit->ctx, SC_ERR_ARG, "kernel matrix dimension does not match n_cells");
  return fit->set_kernel(km->indptr.p, km->indices.p, km->data.p, km->nnz);
}

int sc_fit_set_archetypes(sc_fit* fit, const int64_t* cell_positions) {
  if (!fit || !cell_positions) return SC_ERR_ARG;
  DeviceGuard dev_guard(fit->ctx->device);
  return fit->set_B_onehot((const long long*)cell_positions);
}

int sc_fit_set_A_csc(sc_fit* fit, const int64_t* colptr, const int32_t* rows, const double* vals) {
  if (!fit || !colptr || !rows || !vals) return SC_ERR_ARG;
  DeviceGuard dev_guard(fit->ctx->device);
  return fit->set_A_lists((const long long*)colptr, rows, vals);
}

int sc_fit_set_A_slots(sc_fit* fit, const int32_t* idx, const double* val, const int32_t* cnt) {
  if (!fit || !idx || !val || !cnt) return SC_ERR_ARG;
  DeviceGuard dev_guard(fit->ctx->device);
  return fit->set_A_slots(idx, val, cnt);
}

int sc_fit_set_B_slots(sc_fit* fit, const int32_t* idx, const double* val, const int32_t* cnt) {
  if (!fit || !idx || !val || !cnt) return SC_ERR_ARG;
  DeviceGuard dev_guard(fit->ctx->device);
  return fit->set_B_slots(idx, val, cnt);
}

static int with_trace(sc_fit* fit, int32_t* trace, size_t count, int (*fn)(sc_fit*, int*, double), double arg) {
  sc_ctx* ctx = fit->ctx;
  DBuf<int> td;
  int* tp = nullptr;
  if (trace) {
    SC_TRY(td.ensure(ctx, count));
    SC_CUDA(ctx, cudaMemsetAsync(td.p, 0xff, sizeof(int) * count, ctx->stream));
    tp = td.p;
  }
  SC_TRY(fn(fit, tp, arg));
  if (trace) SC_CUDA(ctx, sc_copy(trace, tp, sizeof(int) * count, ctx->stream));
  SC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return SC_OK;
}

int sc_fit_update_A(sc_fit* fit, double l2_penalty, int32_t* trace) {
  if (!fit) return SC_ERR_ARG;
  DeviceGuard dev_guard(fit->ctx->device);
  return with_trace(fit, trace, (size_t)fit->T * fit->n,
                    [](sc_fit* f, int* t, double l2) { return f->update_A(l2, t); }, l2_penalty);
}

int sc_fit_update_B(sc_fit* fit, int32_t* trace) {
  if (!fit) return SC_ERR_ARG;
  DeviceGuard dev_guard(fit->ctx->device);
  return with_trace(fit, trace, (size_t)fit->T * fit->k, [](sc_fit* f, int* t, double) { return f->update_B(t); }, 0.0);
}

int sc_fit_shard(sc_fit* fit, int* row_begin, int* row_end) {
  if (!fit || !row_begin || !row_end) return SC_ERR_ARG;
  *row_begin = fit->row_begin;
  *row_end = fit->row_end;
  return SC_OK;
}

int sc_fit_rss(sc_fit* fit, double* rss_out) {
  if (!fit || !rss_out) return SC_ERR_ARG;
  DeviceGuard dev_guard(fit->ctx->device);
  return fit->rss(rss_out);
}

int sc_fit_run(sc_fit* fit, int max_iter, int min_iter, double convergence_epsilon, double l2_penalty,
               double* threshold_inout, int* n_iter_out, int* converged_out, double* rss_out, int rss_cap,
               int* n_rss_out) {
  if (!fit || !threshold_inout || !n_iter_out || !converged_out || !rss_out || !n_rss_out) return SC_ERR_ARG;
  DeviceGuard dev_guard(fit->ctx->device);
  if (max_iter < min_iter) SC_FAIL(fit->ctx, SC_ERR_ARG, "max_iter < min_iter");
  return fit->run(max_iter, min_iter, convergence_epsilon, l2_penalty, threshold_inout, n_iter_out, converged_out,
                  rss_out, rss_cap, n_rss_out);
}

int sc_fit_get_KB(sc_fit* fit, long long* nnz_out, int64_t* indptr, int32_t* indices, double* data) {
  if (!fit) return SC_ERR_ARG;
  DeviceGuard dev_guard(fit->ctx->device);
  return fit->get_KB_csr(nnz_out, (long long*)indptr, indices, data);
}
int sc_fit_slots(sc_fit* fit) { return fit ? fit->S : 0; }
int sc_fit_get_A(sc_fit* fit, int32_t* idx, double* val, int32_t* cnt) {
  if (!fit || !idx || !val || !cnt) return SC_ERR_ARG;
  DeviceGuard dev_guard(fit->ctx->device);
  return fit->get_A(idx, val, cnt);
}
int sc_fit_get_B(sc_fit* fit, int32_t* idx, double* val, int32_t* cnt) {
  if (!fit || !idx || !val || !cnt) return SC_ERR_ARG;
  DeviceGuard dev_guard(fit->ctx->device);
  return fit->get_B(idx, val, cnt);
}
int sc_fit_hard_assignments(sc_fit* fit, int32_t* labels) {
  if (!fit || !labels) return SC_ERR_ARG;
  DeviceGuard dev_guard(fit->ctx->device);
  return fit->hard_assignments(labels);
}
int sc_fit_hard_archetypes(sc_fit* fit, int32_t* cells) {
  if (!fit || !cells) return SC_ERR_ARG;
  DeviceGuard dev_guard(fit->ctx->device);
  return fit->hard_archetypes(cells);
}
int sc_fit_soft_assignments(sc_fit* fit, int top, int32_t* idx, double* w) {
  if (!fit || !idx || !w) return SC_ERR_ARG;
  DeviceGuard dev_guard(fit->ctx->device);
  return fit->soft_assignments(top, idx, w);
}
int sc_fit_stats(sc_fit* fit, long long* out4) {
  if (!fit || !out4) return SC_ERR_ARG;
  out4[4] = fit->last_evaluated;
  out4[5] = fit->eval_steps;
  out4[6] = (long long)(fit->eval_ms_total * 1e6);  /* ns */
  out4[7] = fit->last_kb_nnz;
  out4[0] = fit->last_slow_cells;
  out4[1] = fit->last_t2_nnz;
  out4[2] = fit->last_hub_block;
  out4[3] = fit->sp.total_big_rows;
  out4[8] = fit->hub_steps;
  out4[9] = (long long)(fit->hub_ms_total * 1e6);
  out4[10] = fit->hub_launches_counted ? (long long)(fit->hub_bytes_total / (double)fit->hub_launches_counted) : 0;
  out4[11] = fit->last_pruned_chunks;
  out4[12] = fit->last_chunk_launches;
  out4[13] = (long long)fit->last_updA_diag[0];
  out4[14] = (long long)fit->last_updA_diag[1];
  out4[15] = 0;
  return SC_OK;
}
int sc_fit_profile(sc_fit* fit, double* out16) {
  if (!fit || !out16) return SC_ERR_ARG;
  for (int i = 0; i < 16; ++i) out16[i] = fit->prof[i];
  return SC_OK;
}

}  // extern "C"
