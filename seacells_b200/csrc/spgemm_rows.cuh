#pragma once
#include "common.cuh"

struct SpgemmScratch {
  DBuf<long long> ub;
  DBuf<int> flags, big_list, dense_list;
  DBuf<double> big_acc;
  DBuf<char> tmp;
  int last_big_rows = 0;
  long long total_big_rows = 0;  // cumulative over all products of this scratch
};

// Z[i,:] = sum_q S[i,q] * X[q,:] for rows i in [row_begin,row_end); other rows of Z are empty.  Rows sorted by key.
// key_space: keys of X are in [0, key_space).  sorted = false: rows come out in first-occurrence order (fixed, cheaper).
// wide = true (first-occurrence variant only): 1024-entry tables per warp, for rows with up to ~900 distinct keys
// (K = M M^T has 300-600 per row).
int sc_spgemm_rows(sc_ctx* ctx, const CsrView& S, int row_begin, int row_end, const RowLists& X, int key_space,
                   RowListsBuf& Z, SpgemmScratch& sc, bool sorted = true, bool wide = false);
