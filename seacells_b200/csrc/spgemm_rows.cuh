#pragma once
#include "common.cuh"

struct SpgemmScratch {
  DBuf<long long> ub;
  DBuf<int> flags;
  DBuf<char> tmp;
};

// Z[i,:] = sum_q S[i,q] * X[q,:] for rows i in [row_begin,row_end); other rows of Z are empty.  Rows sorted by key.
int sc_spgemm_rows(sc_ctx* ctx, const CsrView& S, int row_begin, int row_end, const RowLists& X, RowListsBuf& Z,
                   SpgemmScratch& sc, int hash_cap = 512);
