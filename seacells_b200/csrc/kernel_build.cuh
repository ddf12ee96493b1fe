#pragma once
#include "common.cuh"

// exact brute-force kNN on device buffers; idx/d2 are [(row_end-row_begin) x (n_neighbors-1)]
int sc_knn_exact_dev(sc_ctx* ctx, const void* X, int x_is_f64, int n, int d, int n_neighbors, int row_begin,
                     int row_end, int* idx, double* d2);

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
