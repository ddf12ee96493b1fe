// Metacell aggregation: core.summarize_by_SEACell (core.py:190-242) and summarize_by_soft_SEACell (core.py:121-187).
//
//   out[g, :] = ( sum over the members (cell, w) of group g, cells ascending, of  w * X[cell, :] )  [ / divisor[g] ]
//
// X is the cells x genes count matrix in CSR (int64 row pointers, int32 columns, float32 or float64 values), the
// members are (group, cell, weight) triplets: one per cell with weight 1 for the hard assignment, the thresholded and
// row-normalised entries of A^T for the soft one.  The reference loops over the metacells in Python and sums a dense
// copy of every slice; here the triplets are sorted by (group, cell) and one CTA owns one output row, adding its
// members' rows in ascending cell order (columns of one CSR row are distinct, rows are separated by a barrier), so the
// summation order - and with it every fp64 result - is fixed and equals numpy's row-by-row axis-0 reduction.
// HBM-bound: reads every member row once (12 or 8 B per stored count), writes n_groups x n_genes x 8 B.
#include "common.cuh"

namespace {

__global__ void k_member_keys(long long m, const int* __restrict__ group, const int* __restrict__ cell,
                              unsigned long long* __restrict__ keys, unsigned* __restrict__ order) {
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= m) return;
  keys[e] = ((unsigned long long)(unsigned)group[e] << 32) | (unsigned)cell[e];
  order[e] = (unsigned)e;
}

__global__ void k_group_bounds(int n_groups, long long m, const unsigned long long* __restrict__ keys,
                               long long* __restrict__ gptr) {
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g > n_groups) return;
  const unsigned long long target = (unsigned long long)(unsigned)g << 32;
  long long lo = 0, hi = m;
  while (lo < hi) {
    const long long mid = (lo + hi) >> 1;
    if (keys[mid] < target) lo = mid + 1;
    else hi = mid;
  }
  gptr[g] = lo;
}

template <typename TV>
__global__ void __launch_bounds__(256)
k_summarize(int n_genes, const long long* __restrict__ x_ptr, const int* __restrict__ x_col,
            const TV* __restrict__ x_val, const long long* __restrict__ gptr,
            const unsigned long long* __restrict__ keys, const unsigned* __restrict__ order,
            const double* __restrict__ weight, const double* __restrict__ divisor, double* __restrict__ out) {
  const int g = blockIdx.x;
  double* __restrict__ orow = out + (size_t)g * n_genes;
  for (int c = threadIdx.x; c < n_genes; c += blockDim.x) orow[c] = 0.0;
  __syncthreads();
  for (long long e = gptr[g]; e < gptr[g + 1]; ++e) {
    const int cell = (int)(keys[e] & 0xffffffffu);
    const double w = weight[order[e]];
    const long long p0 = x_ptr[cell], p1 = x_ptr[cell + 1];
    for (long long p = p0 + threadIdx.x; p < p1; p += blockDim.x) {
      const int c = x_col[p];
      orow[c] = __dadd_rn(orow[c], __dmul_rn(w, (double)x_val[p]));
    }
    __syncthreads();  // the next member's row may hit the same genes
  }
  if (divisor != nullptr) {
    const double dv = divisor[g];
    for (int c = threadIdx.x; c < n_genes; c += blockDim.x) orow[c] = __ddiv_rn(orow[c], dv);
  }
}

}  // namespace

int sc_summarize_dev(sc_ctx* ctx, int n, int n_genes, const long long* x_ptr, const int* x_col, const void* x_val,
                     int val_is_f64, long long m, const int* group, const int* cell, const double* weight, int n_groups,
                     const double* divisor, double* out) {
  (void)n;
  if (n_groups <= 0 || n_genes <= 0) return SC_OK;
  DBuf<unsigned long long> k_in, k_out;
  DBuf<unsigned> o_in, o_out;
  DBuf<long long> gptr;
  DBuf<char> tmp;
  SC_TRY(k_in.ensure(ctx, (size_t)m + 1));
  SC_TRY(k_out.ensure(ctx, (size_t)m + 1));
  SC_TRY(o_in.ensure(ctx, (size_t)m + 1));
  SC_TRY(o_out.ensure(ctx, (size_t)m + 1));
  SC_TRY(gptr.ensure(ctx, (size_t)n_groups + 2));
  if (m > 0) {
    SC_LAUNCH(ctx, k_member_keys, (unsigned)((m + 255) / 256), 256, 0, m, group, cell, k_in.p, o_in.p);
    SC_TRY(sc_sort_pairs_u64_u32(ctx, k_in.p, k_out.p, o_in.p, o_out.p, (size_t)m, 64, tmp));
  }
  SC_LAUNCH(ctx, k_group_bounds, (n_groups + 256) / 256, 256, 0, n_groups, m, k_out.p, gptr.p);
  if (val_is_f64) {
    SC_LAUNCH(ctx, k_summarize<double>, n_groups, 256, 0, n_genes, x_ptr, x_col, (const double*)x_val, gptr.p, k_out.p,
              o_out.p, weight, divisor, out);
  } else {
    SC_LAUNCH(ctx, k_summarize<float>, n_groups, 256, 0, n_genes, x_ptr, x_col, (const float*)x_val, gptr.p, k_out.p,
              o_out.p, weight, divisor, out);
  }
  SC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return SC_OK;
}
