#pragma once
#include "common.cuh"

// exact brute-force kNN on device buffers; idx/d2 are [(row_end-row_begin) x (n_neighbors-1)]
int sc_knn_exact_dev(sc_ctx* ctx, const void* X, int x_is_f64, int n, int d, int n_neighbors, int row_begin,
                     int row_end, int* idx, double* d2);

// tensor-core shortlist filter (knn_tc.cu): 32 candidates per query of [row_begin,row_end) + the 32nd approximate
// squared distance; Xc / norm2 are the mean-centred fp32 embedding and its squared row norms
bool sc_knn_tc_supported(int n, int d);
double sc_knn_tc_gamma(int d);  // error bound factor of the 3xTF32 pipeline, see knn_tc.cu
int sc_knn_filter_tc(sc_ctx* ctx, const float* Xc, const float* norm2, const float* maxnorm2, int n, int d, int row_begin,
                     int row_end, int* cand_idx, float* cand_thr, float* dbg_s, int dbg_cols);
int sc_knn_tc_probe_dev(sc_ctx* ctx, const void* X, int x_is_f64, int n, int d, int cols, float* s_out, float* xc_out,
                        float* norm2_out);

// device-resident kernel matrix M (CSR, sorted columns) and the bandwidths
struct sc_kernel_graph {
  int n = 0;
  long long nnz = 0;
  DBuf<int> indptr, indices;
  DBuf<double> data, sigma;
  DBuf<unsigned long long> ekeys;
  DBuf<int> knn_idx;   // [n x kk] when built through sc_kernel_build
  DBuf<double> knn_d2;
  int kk = 0;
  int build(sc_ctx* ctx, const void* X_dev, int x_is_f64, int n, int d, int n_neighbors, const int* idx_dev,
            const double* d2_dev, int intersection);
};
