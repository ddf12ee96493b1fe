// Greedy adaptive column-subset selection on K = M M^T  (cpu.py:342-423, `_get_greedy_centers`).
// "Next" row of the scope table: the in-tree half of initialize_archetypes.
//
// Per pick j (all fp64, fixed-order reductions):
//   p      = argmax_i f[i] / g[i]                      (first index on ties, np.argmax)
//   delta  = K[:, p] - sum_{r<j} omega[r, p] * omega[r, :]  ; delta[p] = max(0, delta[p])
//   o      = delta / max(sqrt(delta[p]), 1e-6)
//   pl     = sum_{r<j} (omega_r . o) * omega_r
//   f     += -2 o * (K o - pl) + ||o||^2 * o * o ;  g += o * o ;  omega[j] = o
// K is formed explicitly once (row lists, ~400 entries per cell) with the same sparse product kernel the fit uses.
// omega is a dense (n_centers x n) fp64 matrix, exactly as in the reference; that bounds the usable size
// (n_centers * n * 8 bytes), which is why 1M-cell runs pass `initial_archetypes` instead.
#include "common.cuh"
#include "spgemm_rows.cuh"

namespace {

constexpr int GR_THREADS = 256;
constexpr int GR_GRID = 592;

__global__ void k_csr_as_lists(int n, const int* __restrict__ indptr, long long* __restrict__ ptr, int* __restrict__ len) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i > n) return;
  ptr[i] = indptr[i];
  if (i < n) len[i] = indptr[i + 1] - indptr[i];
}

__global__ void k_init_fg(int n, RowLists K, double* __restrict__ f, double* __restrict__ g) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= n) return;
  const long long p = K.ptr[warp];
  const int L = K.len[warp];
  double acc = 0.0, diag = 0.0;
  for (int x = lane; x < L; x += 32) {
    const double v = K.val[p + x];
    acc = fma(v, v, acc);
    if (K.key[p + x] == warp) diag = v;
  }
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) {
    acc += __shfl_xor_sync(0xffffffffu, acc, d);
    diag += __shfl_xor_sync(0xffffffffu, diag, d);
  }
  if (lane == 0) {
    f[warp] = acc;
    g[warp] = diag;
  }
}

// two-stage argmax of f/g, first index on ties
__global__ void k_argmax_part(int n, const double* __restrict__ f, const double* __restrict__ g,
                              double* __restrict__ pv, int* __restrict__ pi) {
  __shared__ double sv[GR_THREADS];
  __shared__ int si[GR_THREADS];
  double bv = -INFINITY;
  int bi = 0x7fffffff;
  for (int i = blockIdx.x * GR_THREADS + threadIdx.x; i < n; i += GR_GRID * GR_THREADS) {
    const double s = f[i] / g[i];
    if (s > bv || (s == bv && i < bi)) {  // NaN never wins, as in np.argmax unless it comes first
      bv = s;
      bi = i;
    }
  }
  sv[threadIdx.x] = bv;
  si[threadIdx.x] = bi;
  __syncthreads();
  for (int st = GR_THREADS / 2; st > 0; st >>= 1) {
    if (threadIdx.x < st) {
      const double ov = sv[threadIdx.x + st];
      const int oi = si[threadIdx.x + st];
      if (ov > sv[threadIdx.x] || (ov == sv[threadIdx.x] && oi < si[threadIdx.x])) {
        sv[threadIdx.x] = ov;
        si[threadIdx.x] = oi;
      }
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    pv[blockIdx.x] = sv[0];
    pi[blockIdx.x] = si[0];
  }
}
__global__ void k_argmax_final(const double* __restrict__ pv, const int* __restrict__ pi, int* __restrict__ p_out,
                               long long* __restrict__ centers, int j) {
  __shared__ double sv[1024];
  __shared__ int si[1024];
  double bv = -INFINITY;
  int bi = 0x7fffffff;
  for (int c = threadIdx.x; c < GR_GRID; c += 1024)
    if (pv[c] > bv || (pv[c] == bv && pi[c] < bi)) {
      bv = pv[c];
      bi = pi[c];
    }
  sv[threadIdx.x] = bv;
  si[threadIdx.x] = bi;
  __syncthreads();
  for (int st = 512; st > 0; st >>= 1) {
    if (threadIdx.x < st) {
      const double ov = sv[threadIdx.x + st];
      const int oi = si[threadIdx.x + st];
      if (ov > sv[threadIdx.x] || (ov == sv[threadIdx.x] && oi < si[threadIdx.x])) {
        sv[threadIdx.x] = ov;
        si[threadIdx.x] = oi;
      }
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    const int p = si[0] == 0x7fffffff ? 0 : si[0];
    p_out[0] = p;
    centers[j] = p;
  }
}

// delta = K[:, p] - sum_r omega[r, p] omega[r, :]   (K[:, p] scattered from row p of the symmetric K)
__global__ void k_scatter_col(RowLists K, const int* __restrict__ p_in, double* __restrict__ delta) {
  const int p = p_in[0];
  const long long kp = K.ptr[p];
  const int L = K.len[p];
  for (int x = blockIdx.x * blockDim.x + threadIdx.x; x < L; x += gridDim.x * blockDim.x) delta[K.key[kp + x]] = K.val[kp + x];
}
__global__ void k_delta(int n, int j, const double* __restrict__ omega, const int* __restrict__ p_in,
                        double* __restrict__ delta) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int p = p_in[0];
  double acc = 0.0;
  for (int r = 0; r < j; ++r) acc = fma(omega[(size_t)r * n + p], omega[(size_t)r * n + i], acc);
  delta[i] = delta[i] - acc;
}
__global__ void k_make_o(int n, const int* __restrict__ p_in, const double* __restrict__ delta, double* __restrict__ o) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int p = p_in[0];
  const double dp = fmax(0.0, delta[p]);
  const double den = fmax(sqrt(dp), 1e-6);
  const double v = (i == p) ? dp : delta[i];
  o[i] = v / den;
}
// c[r] = omega_r . o for r < j ; c[j] = o . o     (one CTA per row, fixed-order tree)
__global__ void k_dots(int n, int j, const double* __restrict__ omega, const double* __restrict__ o,
                       double* __restrict__ c) {
  __shared__ double sh[GR_THREADS];
  const int r = blockIdx.x;
  const double* __restrict__ row = (r < j) ? omega + (size_t)r * n : o;
  double acc = 0.0;
  for (int i = threadIdx.x; i < n; i += GR_THREADS) acc = fma(row[i], o[i], acc);
  sh[threadIdx.x] = acc;
  __syncthreads();
  for (int st = GR_THREADS / 2; st > 0; st >>= 1) {
    if (threadIdx.x < st) sh[threadIdx.x] += sh[threadIdx.x + st];
    __syncthreads();
  }
  if (threadIdx.x == 0) c[r] = sh[0];
}
__global__ void k_update_fg(int n, int j, RowLists K, const double* __restrict__ omega, const double* __restrict__ o,
                            const double* __restrict__ c, double* __restrict__ f, double* __restrict__ g,
                            double* __restrict__ omega_out) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= n) return;
  const int i = warp;
  const long long kp = K.ptr[i];
  const int L = K.len[i];
  double ko = 0.0;
  for (int x = lane; x < L; x += 32) ko = fma(K.val[kp + x], o[K.key[kp + x]], ko);
  double pl = 0.0;
  for (int r = lane; r < j; r += 32) pl = fma(c[r], omega[(size_t)r * n + i], pl);
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) {
    ko += __shfl_xor_sync(0xffffffffu, ko, d);
    pl += __shfl_xor_sync(0xffffffffu, pl, d);
  }
  if (lane == 0) {
    const double oi = o[i];
    f[i] = f[i] + (-2.0 * (oi * (ko - pl)) + c[j] * (oi * oi));
    g[i] = g[i] + oi * oi;
    omega_out[i] = oi;
  }
}

}  // namespace

int sc_greedy_centers_impl(sc_ctx* ctx, int n, const int* indptr, const int* indices, const double* data, long long nnz,
                           int n_centers, long long* centers_host) {
  if (n <= 0 || n_centers <= 0 || n_centers > n) SC_FAIL(ctx, SC_ERR_ARG, "greedy: need 0 < n_centers <= n");
  size_t freeb = 0, totalb = 0;
  SC_CUDA(ctx, cudaMemGetInfo(&freeb, &totalb));
  const double need = (double)n_centers * n * 8.0;
  if (need > 0.8 * (double)freeb)
    SC_FAIL(ctx, SC_ERR_CAPACITY,
            "greedy initialisation keeps a dense (n_centers x n) fp64 matrix, as the reference does (cpu.py:372-373); "
            "it does not fit on this GPU - pass initial_archetypes instead");
  DBuf<int> m_ptr, m_col, lens, pi, p_dev;
  DBuf<double> m_val, f, g, delta, o, c, pv, omega;
  DBuf<long long> lptr, centers;
  SC_TRY(m_ptr.ensure(ctx, (size_t)n + 1));
  SC_TRY(m_col.ensure(ctx, (size_t)nnz));
  SC_TRY(m_val.ensure(ctx, (size_t)nnz));
  SC_CUDA(ctx, sc_copy(m_ptr.p, indptr, sizeof(int) * ((size_t)n + 1), ctx->stream));
  SC_CUDA(ctx, sc_copy(m_col.p, indices, sizeof(int) * (size_t)nnz, ctx->stream));
  SC_CUDA(ctx, sc_copy(m_val.p, data, sizeof(double) * (size_t)nnz, ctx->stream));
  SC_TRY(lptr.ensure(ctx, (size_t)n + 1));
  SC_TRY(lens.ensure(ctx, (size_t)n + 1));
  SC_LAUNCH(ctx, k_csr_as_lists, (n + 256) / 256, 256, 0, n, m_ptr.p, lptr.p, lens.p);
  CsrView M{n, m_ptr.p, m_col.p, m_val.p};
  RowLists Ml{lptr.p, lens.p, m_col.p, m_val.p};
  RowListsBuf K;
  SpgemmScratch sp;
  SC_TRY(sc_spgemm_rows(ctx, M, 0, n, Ml, n, K, sp, true));  // K = M M^T (M symmetric), rows sorted by cell
  SC_TRY(f.ensure(ctx, (size_t)n));
  SC_TRY(g.ensure(ctx, (size_t)n));
  SC_TRY(delta.ensure(ctx, (size_t)n));
  SC_TRY(o.ensure(ctx, (size_t)n));
  SC_TRY(c.ensure(ctx, (size_t)n_centers + 1));
  SC_TRY(pv.ensure(ctx, GR_GRID));
  SC_TRY(pi.ensure(ctx, GR_GRID));
  SC_TRY(p_dev.ensure(ctx, 4));
  SC_TRY(centers.ensure(ctx, (size_t)n_centers));
  SC_TRY(omega.ensure(ctx, (size_t)n_centers * n));
  const unsigned wgrid = (unsigned)(((size_t)n * 32 + 255) / 256);
  SC_LAUNCH(ctx, k_init_fg, wgrid, 256, 0, n, K.view(), f.p, g.p);
  for (int j = 0; j < n_centers; ++j) {
    SC_LAUNCH(ctx, k_argmax_part, GR_GRID, GR_THREADS, 0, n, f.p, g.p, pv.p, pi.p);
    SC_LAUNCH(ctx, k_argmax_final, 1, 1024, 0, pv.p, pi.p, p_dev.p, centers.p, j);
    SC_CUDA(ctx, cudaMemsetAsync(delta.p, 0, sizeof(double) * (size_t)n, ctx->stream));
    SC_LAUNCH(ctx, k_scatter_col, 8, 256, 0, K.view(), p_dev.p, delta.p);
    SC_LAUNCH(ctx, k_delta, (n + 255) / 256, 256, 0, n, j, omega.p, p_dev.p, delta.p);
    SC_LAUNCH(ctx, k_make_o, (n + 255) / 256, 256, 0, n, p_dev.p, delta.p, o.p);
    SC_LAUNCH(ctx, k_dots, j + 1, GR_THREADS, 0, n, j, omega.p, o.p, c.p);
    SC_LAUNCH(ctx, k_update_fg, wgrid, 256, 0, n, j, K.view(), omega.p, o.p, c.p, f.p, g.p, omega.p + (size_t)j * n);
  }
  SC_CUDA(ctx, cudaMemcpyAsync(centers_host, centers.p, sizeof(long long) * (size_t)n_centers, cudaMemcpyDeviceToHost,
                               ctx->stream));
  SC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return SC_OK;
}
