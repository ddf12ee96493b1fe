// Frank-Wolfe fit engine: _updateA / _updateB / compute_RSS of cpu.py:425-536 on compressed A and B.
//
// Layout in HBM (n cells, k archetypes, S = max_FW_iter slots per column):
//   M            CSR int32/int32/fp64                       ~ 24 nnz per row
//   A slots      a_idx[n*S] int32, a_val[n*S] fp64, a_cnt[n]  (column j of A = slots of cell j)
//   B slots      b_idx[k*S] int32, b_val[k*S] fp64, b_cnt[k]  (column a of B = slots of archetype a)
//   K            K = M M^T (cpu.py:128,157) as row lists, first-occurrence order, ~770 entries per row at 1M cells; a rank
//                of a sharded job keeps the columns of its own cells only
//   KB rows      per-cell lists (archetype, (K B)[cell, archetype]) sorted by archetype                  ~ 55 per cell
//   t1A          dense k x k fp64, t1A[i'][i] = (B^T K B)[i, i']      (gradient term of _updateA, cpu.py:442)
//   t1B          dense k x k fp64, A A^T                                (gradient term of _updateB, cpu.py:480)
//   C_a          per-archetype candidate lists (cell, (K A^T)[cell, a]) (cpu.py:481), the support of t2, for the
//                ordinary archetypes; T2H / HH: dense [cells x 64] t2 and running product of the hub archetypes
// Nothing of size n x k or n x n is ever formed.  Multi-GPU: cells are sharded by contiguous slabs, everything a rank
// needs from the others arrives through sc_fit::exchange (an in-place all-gather of equal slabs, comm.cuh).
//
// Exactness of the candidate restriction (SURVEY.md §8a notes): K, A, B >= 0, so a gradient entry can be negative only
// on the support of t2.  If the minimum over that support is < 0 it is the column argmin; otherwise the column is
// scanned densely from index 0 with early exit at the first exact zero, which is what np.argmin over the dense column
// (cpu_dense.py:363, 398) and scipy's implicit-zero rule (scipy/sparse/_data.py, reached from cpu.py:450, 487) return.
#include "fit.cuh"

#include <algorithm>
#include <cmath>

#include "comm.cuh"
#include "fit_kernels.cuh"

constexpr long long SYNC_FREE_CAP = 4ll << 20;  // entries; see push_product

// ------------------------------------------------------------------------------------------------ small kernels
__global__ void k_fill_stride_ptr(long long* p, int rows, int stride) {
  int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r <= rows) p[r] = (long long)r * stride;
}

__global__ void k_sum_len(int rows, const int* __restrict__ len, unsigned long long* out) {
  int r = blockIdx.x * blockDim.x + threadIdx.x;
  unsigned long long v = (r < rows) ? (unsigned long long)len[r] : 0ull;
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
  if ((threadIdx.x & 31) == 0 && v) atomicAdd(out, v);
}

__global__ void k_square(const double* __restrict__ in, double* __restrict__ out, size_t n) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = in[i] * in[i];
}

// Input validation.  K = M M^T (cpu.py:128,157) is applied as M (M .) and the RSS uses the same products, which is only
// the reference's arithmetic when M equals its transpose; the candidate restriction needs M, A0 >= 0.  flag bit 0: a
// negative or NaN value, bit 1: M[i,j] has no equal M[j,i].
__global__ void k_check_kernel(int n, const int* __restrict__ ptr, const int* __restrict__ col,
                               const double* __restrict__ val, int* __restrict__ flag) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int bad = 0;
  for (int e = ptr[i]; e < ptr[i + 1]; ++e) {
    const int j = col[e];
    const double v = val[e];
    if (!(v >= 0.0)) bad |= 1;
    if (j < 0 || j >= n) {
      bad |= 2;
      continue;
    }
    int lo = ptr[j], hi = ptr[j + 1];
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      if (col[mid] < i) lo = mid + 1;
      else hi = mid;
    }
    if (!(lo < ptr[j + 1] && col[lo] == i && val[lo] == v)) bad |= 2;
  }
  if (bad) atomicOr(flag, bad);
}
__global__ void k_check_nonneg(size_t count, const double* __restrict__ val, int* __restrict__ flag) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < count && !(val[i] >= 0.0)) atomicOr(flag, 1);
}

// B by cell: (cell, archetype) pairs -> radix sort -> row bounds.  Rows come out sorted by archetype.
__global__ void k_emit_b_pairs(int k, int S, const int* __restrict__ b_idx, const double* __restrict__ b_val,
                               const int* __restrict__ b_cnt, const long long* __restrict__ off,
                               unsigned long long* __restrict__ keys, double* __restrict__ vals) {
  int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= k * S) return;
  int a = e / S, s = e - a * S;
  if (s >= b_cnt[a]) return;
  keys[off[a] + s] = ((unsigned long long)(unsigned)b_idx[e] << 32) | (unsigned)a;
  vals[off[a] + s] = b_val[e];
}
__global__ void k_lists_from_sorted(int rows, long long total, const unsigned long long* __restrict__ keys,
                                    long long* __restrict__ ptr, int* __restrict__ len, int* __restrict__ key_out) {
  // one thread per row: ptr[r] = lower_bound(r << 32); plus unpack of the low word for the first `total` entries
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t < total) key_out[t] = (int)(keys[t] & 0xffffffffu);
  if (t > rows) return;
  const unsigned long long target = (unsigned long long)(unsigned)t << 32;
  long long lo = 0, hi = total;
  while (lo < hi) {
    long long mid = (lo + hi) >> 1;
    if (keys[mid] < target) lo = mid + 1;
    else hi = mid;
  }
  ptr[t] = lo;
  if (t < rows) {
    const unsigned long long t2 = (unsigned long long)(unsigned)(t + 1) << 32;
    long long lo2 = lo, hi2 = total;
    while (lo2 < hi2) {
      long long mid = (lo2 + hi2) >> 1;
      if (keys[mid] < t2) lo2 = mid + 1;
      else hi2 = mid;
    }
    len[t] = (int)(lo2 - lo);
  }
}

// ------------------------------------------------------------------------------------------------ products with K
// K = M M^T is formed once per kernel matrix (cpu.py:128,157 `self.K = M @ M.T`), rows in first-occurrence order with
// K[i,p] = sum_q M[i,q] M[q,p] accumulated over q ascending.  K is bitwise symmetric (M is; the products commute; the
// order is the same), so column p of K is row p.  A product K X with a column-sparse X (B: ~14 entries per column; the
// one-hot E_t of one Frank-Wolfe step: 1 per column) is evaluated in "push" form: every source entry X[c, a] = w sends
// w K[c, i] to (cell i, archetype a) for the stored entries of row c of K, the triples are grouped by a stable radix
// sort on (cell, archetype) and runs of equal (cell, archetype) are summed in emission order (ascending archetype, then
// slot) - a fixed order, so the result is bitwise reproducible and independent of how the cells are sharded.
// The work is proportional to the number of sources x ~770, not to n x (row work): 10 M triples for one step's K E_t at
// 1M cells, where the row-wise two-stage product M (M E_t) launched over all n rows cost 1.5 ms per step.
// A rank of a sharded job forms only the COLUMNS of K that belong to its own slab of cells, K[:, slab] = M M[slab, :]^T:
// a 1/world share of the memory and of the build, and every entry of a row it stores is one of its own target cells, so
// the products need no filtering and no pass over other ranks' entries.
struct PushSrc {
  int m;                  // sources
  const int* cell;        // [m] source cell c
  const int* key;         // [m] archetype a, or null: a = source index
  const double* w;        // [m] weight, or null: 1.0
};
// Work is split by ENTRY, not by source: rows of K range from a few dozen to tens of thousands of entries (hub cells of the
// kNN graph), and a warp per source left the whole launch waiting for the longest row.  off = exclusive scan of the
// sources' row lengths; thread e finds its source by binary search in `off` (m + 1 values, cache resident).
__global__ void k_src_len(PushSrc src, RowLists K, int* __restrict__ len_out) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s > src.m) return;
  len_out[s] = (s < src.m) ? K.len[src.cell[s]] : 0;
}
__device__ __forceinline__ int push_find_source(const long long* __restrict__ off, int m, long long e) {
  int lo = 0, hi = m;  // largest s with off[s] <= e
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (off[mid] <= e) lo = mid;
    else hi = mid;
  }
  return lo;
}
// K B (several sources may share (cell, archetype)): 64-bit keys (cell << kbits | archetype), values w K[c, i]
__global__ void __launch_bounds__(256)
k_push_emit(PushSrc src, RowLists K, const long long* __restrict__ off, int kbits, unsigned long long* __restrict__ keys,
            double* __restrict__ vals) {
  const long long total = off[src.m];
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const int s = push_find_source(off, src.m, e);
    const long long kp = K.ptr[src.cell[s]] + (e - off[s]);
    const unsigned long long a = (unsigned long long)(unsigned)(src.key ? src.key[s] : s);
    keys[e] = ((unsigned long long)(unsigned)K.key[kp] << kbits) | a;
    vals[e] = src.w ? __dmul_rn(src.w[s], K.val[kp]) : K.val[kp];
  }
}
// One source per archetype (the one-hot E_t of a Frank-Wolfe step): nothing to sum, so the triples only have to be grouped
// by target cell: 32-bit sort keys (the cell; the sort is stable and the flattened entries are emitted source by source,
// so inside a cell they stay in ascending-archetype order) with the entry index as payload.  Two cheaper-looking
// variants were measured at 1M cells and dropped: a 64-bit (cell, archetype) radix sort of the ~10 M triples (0.6 ms in
// CUB alone) and a histogram + atomic-cursor scatter + per-row rank sort (1.1 ms: hub cells of the kNN graph receive
// hundreds of entries per step, and returning atomics on one address serialise).
__global__ void __launch_bounds__(256)
k_push_keys(PushSrc src, RowLists K, const long long* __restrict__ off, unsigned* __restrict__ keys,
            unsigned* __restrict__ idx, int* __restrict__ flat_a, double* __restrict__ flat_v) {
  const long long total = off[src.m];
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const int s = push_find_source(off, src.m, e);
    const long long kp = K.ptr[src.cell[s]] + (e - off[s]);
    keys[e] = (unsigned)K.key[kp];
    idx[e] = (unsigned)e;
    flat_a[e] = src.key ? src.key[s] : s;  // read here, coalesced along the row of K; the gather after the sort then
    flat_v[e] = K.val[kp];                 // reads two compact, L2-resident arrays instead of the 9 GB of K
  }
}
// sorted position j -> (archetype, value) of the emitted entry idx[j]
__global__ void __launch_bounds__(256)
k_push_gather(long long kept, const unsigned* __restrict__ keys, const unsigned* __restrict__ idx,
              const int* __restrict__ flat_a, const double* __restrict__ flat_v, int* __restrict__ z_key,
              double* __restrict__ z_val) {
  for (long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x; j < kept; j += (long long)gridDim.x * blockDim.x) {
    if (keys[j] == 0xffffffffu) continue;  // sentinel of the sync-free variant (unused sort slot)
    const unsigned o = idx[j];
    z_key[j] = flat_a[o];
    z_val[j] = flat_v[o];
  }
}
__global__ void k_rows_from_keys32(int r0, int rows, long long total, const unsigned* __restrict__ keys,
                                   long long* __restrict__ ptr, int* __restrict__ len) {
  const int t = r0 + blockIdx.x * blockDim.x + threadIdx.x;  // rows [r0, rows): the cells this rank's columns of K can name
  if (t > rows) return;
  auto lower = [&](unsigned target) {
    long long lo = 0, hi = total;
    while (lo < hi) {
      const long long mid = (lo + hi) >> 1;
      if (keys[mid] < target) lo = mid + 1;
      else hi = mid;
    }
    return lo;
  };
  const long long lo = lower((unsigned)t);
  ptr[t] = lo;
  if (t < rows) len[t] = (int)(lower((unsigned)t + 1u) - lo);
}
// sorted triples -> row lists.  heads[e] = 1 where a new (cell, archetype) run starts.
__global__ void k_run_heads(long long total, const unsigned long long* __restrict__ keys, long long* __restrict__ heads) {
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e > total) return;
  heads[e] = (e < total && (e == 0 || keys[e] != keys[e - 1])) ? 1 : 0;  // heads[total] = 0: the scan ends with the count
}
// one thread per run head: sum the run in order, write the compacted (cell-keyed) entry
__global__ void k_run_sum(long long total, const unsigned long long* __restrict__ keys, const double* __restrict__ vals,
                          const long long* __restrict__ pos, int kbits, unsigned long long* __restrict__ ckeys,
                          int* __restrict__ out_key, double* __restrict__ out_val) {
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= total) return;
  const unsigned long long key = keys[e];
  if (e > 0 && keys[e - 1] == key) return;
  double acc = vals[e];
  for (long long x = e + 1; x < total && keys[x] == key; ++x) acc = __dadd_rn(acc, vals[x]);
  const long long o = pos[e];
  ckeys[o] = key;
  out_key[o] = (int)(key & ((1ull << kbits) - 1ull));
  out_val[o] = acc;
}
// row bounds of the compacted, (cell << kbits | archetype)-sorted entries; `total` is read from device memory
__global__ void k_rows_from_ckeys(int r0, int rows, const long long* __restrict__ total_dev,
                                  const unsigned long long* __restrict__ ckeys, int kbits, long long* __restrict__ ptr,
                                  int* __restrict__ len) {
  const int t = r0 + blockIdx.x * blockDim.x + threadIdx.x;
  if (t > rows) return;
  const long long total = *total_dev;
  auto lower = [&](unsigned long long target) {
    long long lo = 0, hi = total;
    while (lo < hi) {
      const long long mid = (lo + hi) >> 1;
      if (ckeys[mid] < target) lo = mid + 1;
      else hi = mid;
    }
    return lo;
  };
  const long long lo = lower((unsigned long long)(unsigned)t << kbits);
  ptr[t] = lo;
  if (t < rows) len[t] = (int)(lower((unsigned long long)(unsigned)(t + 1) << kbits) - lo);
}
// B slots -> push sources (archetype-major, slot order)
__global__ void k_b_sources(int k, int S, const int* __restrict__ b_idx, const double* __restrict__ b_val,
                            const int* __restrict__ b_cnt, const long long* __restrict__ off, int* __restrict__ cell,
                            int* __restrict__ key, double* __restrict__ w) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= k * S) return;
  const int a = e / S, sl = e - a * S;
  if (sl >= b_cnt[a]) return;
  const long long o = off[a] + sl;
  cell[o] = b_idx[e];
  key[o] = a;
  w[o] = b_val[e];
}
__global__ void k_b_support_bits(int k, int S, const int* __restrict__ b_idx, const int* __restrict__ b_cnt,
                                 unsigned* __restrict__ bits) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= k * S) return;
  const int a = e / S, sl = e - a * S;
  if (sl >= b_cnt[a]) return;
  const int c = b_idx[e];
  atomicOr(&bits[c >> 5], 1u << (c & 31));
}
// rows [r0, r1) of L -> this rank's segment of the job-wide layout: entries packed from `base` on, ptr/len at the row's
// global index (rows r1 .. r0 + slab of a short last slab get empty lists)
__global__ void k_compact_segment(int r0, int r1, int slab, RowLists L, const unsigned* __restrict__ keep,
                                  const long long* __restrict__ loc_off, long long base, long long* __restrict__ g_ptr,
                                  int* __restrict__ g_len, int* __restrict__ g_key, double* __restrict__ g_val) {
  const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (w >= slab) return;
  const int i = r0 + w;
  const long long o = base + loc_off[w];
  int len = 0;
  if (i < r1 && (!keep || ((keep[i >> 5] >> (i & 31)) & 1u))) {
    len = L.len[i];
    const long long p = L.ptr[i];
    for (int x = lane; x < len; x += 32) {
      g_key[o + x] = L.key[p + x];
      g_val[o + x] = L.val[p + x];
    }
  }
  if (lane == 0) {
    g_ptr[i] = o;
    g_len[i] = len;
  }
}
// M restricted to the columns [c0, c1): lengths, then the compacted (column, value) lists (warp per row)
__global__ void k_filter_cols_count(int n, int c0, int c1, const int* __restrict__ indptr, const int* __restrict__ col,
                                    int* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i > n) return;
  int c = 0;
  if (i < n)
    for (int e = indptr[i]; e < indptr[i + 1]; ++e) c += (col[e] >= c0 && col[e] < c1) ? 1 : 0;
  out[i] = c;
}
__global__ void k_filter_cols_fill(int n, int c0, int c1, const int* __restrict__ indptr, const int* __restrict__ col,
                                   const double* __restrict__ val, const long long* __restrict__ off,
                                   int* __restrict__ len, int* __restrict__ key, double* __restrict__ v) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  long long o = off[i];
  for (int e = indptr[i]; e < indptr[i + 1]; ++e)
    if (col[e] >= c0 && col[e] < c1) {
      key[o] = col[e];
      v[o] = val[e];
      ++o;
    }
  len[i] = (int)(o - off[i]);
}
__global__ void k_masked_len(int r0, int r1, int slab, const int* __restrict__ len, const unsigned* __restrict__ keep,
                             int* __restrict__ out) {
  const int w = blockIdx.x * blockDim.x + threadIdx.x;
  if (w > slab) return;
  const int i = r0 + w;
  out[w] = (w < slab && i < r1 && (!keep || ((keep[i >> 5] >> (i & 31)) & 1u))) ? len[i] : 0;
}
__global__ void k_max_int(int n, const int* __restrict__ v, int* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  int m = (i < n) ? v[i] : 0;
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, d));
  if ((threadIdx.x & 31) == 0 && m > 0) atomicMax(out, m);
}
__global__ void k_csr_as_lists(int n, const int* __restrict__ indptr, long long* __restrict__ ptr, int* __restrict__ len) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i > n) return;
  ptr[i] = indptr[i];
  if (i < n) len[i] = indptr[i + 1] - indptr[i];
}

// t1A[i'][i] = sum over slots (cell, b) of B[:, i'] of b * KB[cell, i]      (cpu.py:442  t1 = t2 @ B)
__global__ void k_t1A(int k, int S, const int* __restrict__ b_idx, const double* __restrict__ b_val,
                      const int* __restrict__ b_cnt, RowLists KB, double* __restrict__ t1T) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= k) return;
  double* row = t1T + (size_t)warp * k;
  const int cnt = b_cnt[warp];
  for (int s = 0; s < cnt; ++s) {
    const int cell = b_idx[warp * S + s];
    const double b = b_val[warp * S + s];
    const long long p = KB.ptr[cell];
    const int L = KB.len[cell];
    for (int x = lane; x < L; x += 32) {
      const int key = KB.key[p + x];
      row[key] = __dadd_rn(row[key], __dmul_rn(KB.val[p + x], b));
    }
    __syncwarp();
  }
}

// ------------------------------------------------------------------------------------------------ _updateA
struct UpdAArgs {
  int cell_begin, cell_end, n, k, S, T;
  double l2;
  RowLists KB;        // candidates of cell j: (archetype, t2[archetype, j]) sorted by archetype
  const double* t1T;  // t1T[i' * k + i] = t1[i, i']
  RowLists Aprev;
  int* a_idx;
  double* a_val;
  int* a_cnt;
  int* trace;  // [T * n] or null
  int* slow_list;
  int* slow_cnt;
  unsigned long long* diag;  // [0] Frank-Wolfe steps whose zero scan had to go past archetype 31, [1] 32-wide batches they scanned
};

// binary search for `key` in a sorted list; returns value or 0
__device__ __forceinline__ double list_lookup(const int* __restrict__ keys, const double* __restrict__ vals, int L,
                                              int key) {
  int lo = 0, hi = L;
  while (lo < hi) {
    int mid = (lo + hi) >> 1;
    if (keys[mid] < key) lo = mid + 1;
    else hi = mid;
  }
  return (lo < L && keys[lo] == key) ? vals[lo] : 0.0;
}

// per warp: candidate t2 values, running (t1 A)[candidates], running (t1 A)[archetypes 0..31], slot weights, candidate
// ids, slot ids
constexpr size_t FWA_PER_WARP = (size_t)(2 * FWA_NC + FW_SMAX) * sizeof(double) + (size_t)(FWA_NC + FW_SMAX) * sizeof(int);
constexpr int FWA_NZ = 4;  // the running products (t1 A)[i] are kept for archetypes i < 32 FWA_NZ (registers: lane l owns l + 32 j)

// Dense-column rule when no candidate is negative: scan archetypes i0, i0+1, ... and stop at the first exact zero of
// G[i] = 2 (sum_s t1[i, act_s] a_s - t2[i]); if the column has no zero, return the first minimum (sv, si carry the
// minimum of the part of the column that was already scanned).
__device__ __forceinline__ int fwA_zero_scan(int k, const double* __restrict__ t1T, const int* act_id,
                                             const double* act_val, int nact, const int* ck, const double* ct2,
                                             int nc, int lane, int i_begin = 0, double sv = FW_INF,
                                             int si = 0x7fffffff) {
  for (int i0 = i_begin; i0 < k; i0 += 32) {
    const int i = i0 + lane;
    double G = FW_INF;
    if (i < k) {
      double acc = 0.0;
      for (int e = 0; e < nact; ++e) acc = fma(t1T[(size_t)act_id[e] * k + i], act_val[e], acc);
      G = 2.0 * (acc - list_lookup(ck, ct2, nc, i));
    }
    const unsigned z = __ballot_sync(0xffffffffu, G == 0.0);
    if (z) return i0 + __ffs(z) - 1;
    if (G < sv) {  // ascending scan: strict '<' keeps the first index
      sv = G;
      si = i;
    }
  }
  warp_argmin(sv, si);
  return si;
}

// Shared-memory path of _updateA: one warp per cell.  Handles every cell with <= FWA_NC candidates (support of
// t2[:, j]) when l2_penalty == 0, including cells whose gradient has no negative entry (zero scan; the chosen archetype
// then lies outside the candidate set and simply becomes one more active slot).
// The step A <- A + gamma (e - A) moves t1 A by gamma (t1[:, chosen] - t1 A), so the products the gradient needs - on
// the candidates, and on archetypes 0..31 where the zero scan almost always ends - are advanced in place instead of
// being re-summed over the active slots at every step (that re-summation was ~25 t1 gathers per archetype and step
// for the zero scan alone: 320 KB of L2 traffic per cell and call).  Exact zeros stay exact: a sum of non-negative
// terms is zero iff every term is.
// RESUM = true (SC_FIT_RESUM_A=1) re-sums (t1 A) over the active slots at every step, in slot order, instead of advancing
// the running products: the parity tests run both variants against the oracle's argmin traces.
template <bool RESUM>
__global__ void __launch_bounds__(FWA_WARPS * 32) k_updateA_fast(UpdAArgs g) {
  constexpr int NZ = FWA_NZ;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  unsigned char* base = smem_raw + (size_t)warp * FWA_PER_WARP;
  double* ct2 = reinterpret_cast<double*>(base);
  double* Hc = ct2 + FWA_NC;   // (t1 A)[ck[c]] for the current A
  double* sa = Hc + FWA_NC;
  int* ck = reinterpret_cast<int*>(sa + FW_SMAX);
  int* sid = ck + FWA_NC;
  const int k = g.k, S = g.S, T = g.T;
  const double* __restrict__ t1T = g.t1T;

  for (int j = g.cell_begin + blockIdx.x * FWA_WARPS + warp; j < g.cell_end; j += gridDim.x * FWA_WARPS) {
    const int nc = g.KB.len[j];
    if (nc > FWA_NC) {
      if (lane == 0) g.slow_list[atomicAdd(g.slow_cnt, 1)] = j;
      continue;
    }
    const long long kp = g.KB.ptr[j];
    for (int c = lane; c < nc; c += 32) {
      ck[c] = g.KB.key[kp + c];
      ct2[c] = g.KB.val[kp + c];
    }
    __syncwarp();
    // ---- t = 0: gradient against the previous A (cpu.py:447 with A = A_prev)
    const int m0 = g.Aprev.len[j];
    const int* pkey = g.Aprev.key + g.Aprev.ptr[j];
    const double* pval = g.Aprev.val + g.Aprev.ptr[j];
    double bv = FW_INF;
    int bi = 0x7fffffff;
    for (int c = lane; c < nc; c += 32) {
      double acc = 0.0;
      const int id = ck[c];
      for (int e = 0; e < m0; ++e) acc = fma(t1T[(size_t)pkey[e] * k + id], pval[e], acc);
      const double G = 2.0 * (acc - ct2[c]);
      if (G < bv) {
        bv = G;
        bi = c;
      }
    }
    warp_argmin(bv, bi);
    int chosen = (bv < 0.0) ? ck[bi] : fwA_zero_scan(k, t1T, pkey, pval, m0, ck, ct2, nc, lane);
    double aprev = 0.0;  // A_prev[chosen, j]
    for (int e0 = 0; e0 < m0; e0 += 32) {
      const int e = e0 + lane;
      const bool hit = e < m0 && pkey[e] == chosen;
      const unsigned ball = __ballot_sync(0xffffffffu, hit);
      if (ball) {
        aprev = __shfl_sync(0xffffffffu, hit ? pval[e] : 0.0, __ffs(ball) - 1);
        break;
      }
    }
    int nslots = 1;
    const double a0 = fw_step(aprev, 1.0, fw_gamma(0));  // every other entry becomes a + (0 - a) = 0 and is dropped
    if (lane == 0) {
      sid[0] = chosen;
      sa[0] = a0;
      if (g.trace) g.trace[j] = chosen;
    }
    // zero-rule window: archetypes 0 .. 32 NZ - 1.  hz = running (t1 A)[i], zt2 = t2[i, j] (0 outside the support), both
    // in registers of the owning lane.  With a 32-wide window 16 % of all Frank-Wolfe steps at 1M cells had to continue
    // with the re-summing scan below (3.9 batches of 32 archetypes x ~15 gathers per lane each).
    double hz[NZ], zt2[NZ];
    {
      const double* __restrict__ trow = t1T + (size_t)chosen * k;
      for (int c = lane; c < nc; c += 32) Hc[c] = __dmul_rn(trow[ck[c]], a0);
#pragma unroll
      for (int jz = 0; jz < NZ; ++jz) {
        const int zi = lane + 32 * jz;
        hz[jz] = (zi < k) ? __dmul_rn(trow[zi], a0) : 0.0;
        zt2[jz] = (zi < k) ? list_lookup(ck, ct2, nc, zi) : 0.0;
      }
    }
    __syncwarp();
    // ---- t >= 1
    for (int t = 1; t < T; ++t) {
      if (RESUM) {
        for (int c = lane; c < nc; c += 32) {
          double acc = 0.0;
          const int id = ck[c];
          for (int e = 0; e < nslots; ++e) acc = fma(t1T[(size_t)sid[e] * k + id], sa[e], acc);
          Hc[c] = acc;
        }
#pragma unroll
        for (int jz = 0; jz < NZ; ++jz) {
          const int zi = lane + 32 * jz;
          double acc = 0.0;
          if (zi < k)
            for (int e = 0; e < nslots; ++e) acc = fma(t1T[(size_t)sid[e] * k + zi], sa[e], acc);
          hz[jz] = acc;
        }
        __syncwarp();
      }
      bv = FW_INF;
      bi = 0x7fffffff;
      for (int c = lane; c < nc; c += 32) {
        const double G = 2.0 * (Hc[c] - ct2[c]);
        if (G < bv) {
          bv = G;
          bi = c;
        }
      }
      warp_argmin(bv, bi);
      if (bv < 0.0) {
        chosen = ck[bi];
      } else {
        // dense-column rule: first exact zero, archetypes ascending; the window comes from the running products
        double sv = FW_INF;
        int si = 0x7fffffff;
        bool found = false;
#pragma unroll
        for (int jz = 0; jz < NZ; ++jz) {
          if (found) break;
          const int zi = lane + 32 * jz;
          const double Gz = (zi < k) ? 2.0 * (hz[jz] - zt2[jz]) : FW_INF;
          const unsigned z = __ballot_sync(0xffffffffu, Gz == 0.0);
          if (z) {
            chosen = 32 * jz + __ffs(z) - 1;
            found = true;
          } else if (Gz < sv) {  // ascending per lane: strict '<' keeps the first index
            sv = Gz;
            si = zi;
          }
        }
        if (!found) {
          chosen = fwA_zero_scan(k, t1T, sid, sa, nslots, ck, ct2, nc, lane, 32 * NZ, sv, si);
          if (lane == 0 && g.diag) {
            atomicAdd(g.diag, 1ull);
            atomicAdd(g.diag + 1, (unsigned long long)((chosen >= 32 * NZ ? chosen - 32 * NZ : k) / 32 + 1));
          }
        }
      }
      const double gamma = fw_gamma(t);
      int pos = -1;
      for (int s0 = 0; s0 < nslots; s0 += 32) {
        const int s = s0 + lane;
        const unsigned ball = __ballot_sync(0xffffffffu, s < nslots && sid[s] == chosen);
        if (ball) {
          pos = s0 + __ffs(ball) - 1;
          break;
        }
      }
      __syncwarp();
      for (int s = lane; s < nslots; s += 32) sa[s] = fw_step(sa[s], s == pos ? 1.0 : 0.0, gamma);
      if (pos < 0) {
        if (lane == 0) {
          sid[nslots] = chosen;
          sa[nslots] = fw_step(0.0, 1.0, gamma);
        }
        nslots++;
      }
      if (!RESUM) {
        const double* __restrict__ trow = t1T + (size_t)chosen * k;
        for (int c = lane; c < nc; c += 32) Hc[c] = fw_step(Hc[c], trow[ck[c]], gamma);
#pragma unroll
        for (int jz = 0; jz < NZ; ++jz) {
          const int zi = lane + 32 * jz;
          if (zi < k) hz[jz] = fw_step(hz[jz], trow[zi], gamma);
        }
      }
      if (lane == 0 && g.trace) g.trace[(size_t)t * g.n + j] = chosen;
      __syncwarp();
    }
    for (int s = lane; s < nslots; s += 32) {
      g.a_idx[(size_t)j * S + s] = sid[s];
      g.a_val[(size_t)j * S + s] = sa[s];
    }
    if (lane == 0) g.a_cnt[j] = nslots;
    __syncwarp();
  }
}

// Generic path: exact dense-column semantics for any cell (no candidates, all-non-negative gradient, l2 > 0, ...).
__global__ void __launch_bounds__(128) k_updateA_slow(UpdAArgs g, int nslow) {
  __shared__ int s_id[4][FW_SMAX + 1];
  __shared__ double s_val[4][FW_SMAX + 1];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int k = g.k, S = g.S, T = g.T;
  const double l2 = g.l2;
  const double* __restrict__ t1T = g.t1T;
  for (int w = blockIdx.x * 4 + warp; w < nslow; w += gridDim.x * 4) {
    const int j = g.slow_list[w];
    const int nc = g.KB.len[j];
    const long long kp = g.KB.ptr[j];
    const int* ckeys = g.KB.key + kp;
    const double* cvals = g.KB.val + kp;
    const int* act_id = g.Aprev.key + g.Aprev.ptr[j];
    const double* act_val = g.Aprev.val + g.Aprev.ptr[j];
    int nact = g.Aprev.len[j];
    for (int t = 0; t < T; ++t) {
      auto evalG = [&](int id, double t2v) -> double {
        double acc = 0.0, aval = 0.0;
        for (int e = 0; e < nact; ++e) {
          const int ai = act_id[e];
          const double av = act_val[e];
          acc = fma(t1T[(size_t)ai * k + id], av, acc);
          if (ai == id) aval = av;
        }
        double G = 2.0 * (acc - t2v);
        if (l2 != 0.0) G = __dsub_rn(G, __dmul_rn(l2, aval));
        return G;
      };
      double bv = FW_INF;
      int bi = 0x7fffffff;
      for (int c0 = 0; c0 < nc; c0 += 32) {
        const int c = c0 + lane;
        if (c < nc) {
          const int id = ckeys[c];
          const double G = evalG(id, cvals[c]);
          if (G < bv || (G == bv && id < bi)) {
            bv = G;
            bi = id;
          }
        }
      }
      if (l2 != 0.0) {  // active vertices outside the support of t2 can be negative through the -l2*A term
        for (int a0 = 0; a0 < nact; a0 += 32) {
          const int e = a0 + lane;
          if (e < nact) {
            const int id = act_id[e];
            const double G = evalG(id, list_lookup(ckeys, cvals, nc, id));
            if (G < bv || (G == bv && id < bi)) {
              bv = G;
              bi = id;
            }
          }
        }
      }
      warp_argmin(bv, bi);
      int amin;
      if (bv < 0.0) {
        amin = bi;
      } else {
        bool found = false;
        double sv = FW_INF;
        int si = 0x7fffffff;
        amin = 0;
        for (int i0 = 0; i0 < k && !found; i0 += 32) {
          const int i = i0 + lane;
          double G = FW_INF;
          if (i < k) G = evalG(i, list_lookup(ckeys, cvals, nc, i));
          const unsigned z = __ballot_sync(0xffffffffu, G == 0.0);
          if (z) {
            amin = i0 + __ffs(z) - 1;
            found = true;
          } else if (G < sv) {  // ascending scan: strict '<' keeps the first index
            sv = G;
            si = i;
          }
        }
        if (!found) {
          warp_argmin(sv, si);
          amin = si;
        }
      }
      // simplex step
      const double gamma = fw_gamma(t);
      double aprev = 0.0;
      int pos = -1;
      for (int e0 = 0; e0 < nact; e0 += 32) {
        const int e = e0 + lane;
        const bool hit = e < nact && act_id[e] == amin;
        const unsigned ball = __ballot_sync(0xffffffffu, hit);
        if (ball) {
          const int src = __ffs(ball) - 1;
          aprev = __shfl_sync(0xffffffffu, hit ? act_val[e] : 0.0, src);
          pos = e0 + src;
          break;
        }
      }
      __syncwarp();
      if (t == 0) {
        if (lane == 0) {
          s_id[warp][0] = amin;
          s_val[warp][0] = fw_step(aprev, 1.0, gamma);  // every other entry becomes a + (0 - a) = 0 and is dropped
        }
        nact = 1;
        act_id = s_id[warp];
        act_val = s_val[warp];
      } else {
        for (int e = lane; e < nact; e += 32) s_val[warp][e] = fw_step(s_val[warp][e], e == pos ? 1.0 : 0.0, gamma);
        if (pos < 0) {
          if (lane == 0) {
            s_id[warp][nact] = amin;
            s_val[warp][nact] = fw_step(0.0, 1.0, gamma);
          }
          nact++;
        }
      }
      if (lane == 0 && g.trace) g.trace[(size_t)t * g.n + j] = amin;
      __syncwarp();
    }
    for (int s = lane; s < nact; s += 32) {
      g.a_idx[(size_t)j * S + s] = act_id[s];
      g.a_val[(size_t)j * S + s] = act_val[s];
    }
    if (lane == 0) g.a_cnt[j] = nact;
    __syncwarp();
  }
}

// ------------------------------------------------------------------------------------------------ Gram A A^T
__global__ void k_emit_a_pairs(int n, int S, const int* __restrict__ a_idx, const double* __restrict__ a_val,
                               const int* __restrict__ a_cnt, const long long* __restrict__ off,
                               unsigned long long* keys, double* vals) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= n) return;
  const int cnt = a_cnt[warp];
  const long long o = off[warp];
  for (int s = lane; s < cnt; s += 32) {
    keys[o + s] = ((unsigned long long)(unsigned)a_idx[(size_t)warp * S + s] << 32) | (unsigned)warp;
    vals[o + s] = a_val[(size_t)warp * S + s];
  }
}
// Row-sharded Gram: a rank accumulates the rows t1B[a, :] of its own slab of archetypes [a0, a1) (every entry is still
// summed whole, in cell order, by exactly one rank), so it transposes only the slots of A that name those archetypes.
__global__ void k_arch_hist(int r0, int r1, int S, const int* __restrict__ a_idx, const int* __restrict__ a_cnt,
                            int* __restrict__ cells_per_arch) {
  const int warp = r0 + ((blockIdx.x * blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
  if (warp >= r1) return;
  const int cnt = a_cnt[warp];
  for (int s = lane; s < cnt; s += 32) atomicAdd(&cells_per_arch[a_idx[(size_t)warp * S + s]], 1);
}
__global__ void k_count_a_range(int n, int S, int a0, int a1, const int* __restrict__ a_idx,
                                const int* __restrict__ a_cnt, int* __restrict__ out) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp > n) return;
  int c = 0;
  if (warp < n) {
    const int cnt = a_cnt[warp];
    for (int s = lane; s < cnt; s += 32) {
      const int a = a_idx[(size_t)warp * S + s];
      c += (a >= a0 && a < a1) ? 1 : 0;
    }
  }
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) c += __shfl_xor_sync(0xffffffffu, c, d);
  if (lane == 0) out[warp] = c;  // out[n] = 0: the scan ends with the total
}
__global__ void k_emit_a_pairs_range(int n, int S, int a0, int a1, const int* __restrict__ a_idx,
                                     const double* __restrict__ a_val, const int* __restrict__ a_cnt,
                                     const long long* __restrict__ off, unsigned long long* keys, double* vals) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= n) return;
  const int cnt = a_cnt[warp];
  long long o = off[warp];
  for (int s0 = 0; s0 < cnt; s0 += 32) {
    const int s = s0 + lane;
    int a = -1;
    if (s < cnt) a = a_idx[(size_t)warp * S + s];
    const bool keep = a >= a0 && a < a1;
    const unsigned ball = __ballot_sync(0xffffffffu, keep);
    if (keep) {
      const long long pos = o + __popc(ball & ((1u << lane) - 1u));
      keys[pos] = ((unsigned long long)(unsigned)a << 32) | (unsigned)warp;
      vals[pos] = a_val[(size_t)warp * S + s];
    }
    o += __popc(ball);
  }
}
__global__ void k_row_bounds(int k, const unsigned long long* __restrict__ keys, long long total, long long* rptr) {
  int a = blockIdx.x * blockDim.x + threadIdx.x;
  if (a > k) return;
  const unsigned long long target = (unsigned long long)(unsigned)a << 32;
  long long lo = 0, hi = total;
  while (lo < hi) {
    long long mid = (lo + hi) >> 1;
    if (keys[mid] < target) lo = mid + 1;
    else hi = mid;
  }
  rptr[a] = lo;
}
// t1B[l][a] = sum_j A[l,j] A[a,j], cells j ascending (the order of scipy's csr_matmat for A @ A.T, cpu.py:480).
// Rows longer than `hub_thr` cells ("hub" archetypes: low-index archetypes that collect a little weight from almost
// every cell through the argmin-of-zeros rule) are skipped here and filled by k_gram_mirror / k_gram_hub_tiles.
__global__ void k_gram(int k, int S, int a0, int a1, const unsigned char* __restrict__ is_hub, const long long* __restrict__ rptr,
                       const unsigned long long* __restrict__ keys, const double* __restrict__ vals,
                       const int* __restrict__ a_idx, const double* __restrict__ a_val,
                       const int* __restrict__ a_cnt, double* __restrict__ t1) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= k || warp < a0 || warp >= a1) return;
  if (is_hub[warp]) return;
  double* row = t1 + (size_t)warp * k;
  for (long long e = rptr[warp]; e < rptr[warp + 1]; ++e) {
    const int j = (int)(keys[e] & 0xffffffffu);
    const double v = vals[e];
    const int cnt = a_cnt[j];
    for (int s = lane; s < cnt; s += 32) {
      const int a = a_idx[(size_t)j * S + s];
      row[a] = __dadd_rn(row[a], __dmul_rn(v, a_val[(size_t)j * S + s]));
    }
    __syncwarp();
  }
}
// hub row h, ordinary column a: the product terms and their order are those of t1B[a][h], so copy it
__global__ void k_gram_mirror(int k, int nhub, const int* __restrict__ hubs, const unsigned char* __restrict__ is_hub,
                              double* __restrict__ t1) {
  const int a = blockIdx.x * blockDim.x + threadIdx.x;
  const int h = hubs[blockIdx.y];
  if (a >= k || is_hub[a]) return;
  t1[(size_t)h * k + a] = t1[(size_t)a * k + h];
}
// hub x hub entries, cell-major: a CTA walks a contiguous block of cells in index order; for each cell the weights it
// puts on hub archetypes are staged in shared memory and all pairs are added into the CTA's private H x H tile (targets
// of one cell are distinct, cells are sequential => fixed order).  The per-CTA tiles are then summed in CTA order.
constexpr int GRAM_HUB_MAX = 160;
constexpr int GRAM_HUB_CELLS = 1024;  // cells per hub x hub tile (a CTA walks them one after the other)
__global__ void __launch_bounds__(256)
k_gram_hub_tiles(int n, int S, int H, int tile0, const int* __restrict__ hub_of /* [k] -> hub index or -1 */,
                 const int* __restrict__ a_idx, const double* __restrict__ a_val, const int* __restrict__ a_cnt,
                 double* __restrict__ tiles) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* tile_acc = reinterpret_cast<double*>(smem_raw);        // [H*H]
  double* hv = tile_acc + (size_t)H * H;                          // [2][FW_SMAX] weights of the staged cell
  int* hi = reinterpret_cast<int*>(hv + 2 * FW_SMAX);             // [2][FW_SMAX] hub indices
  __shared__ int hcnt[2];
  for (int t = threadIdx.x; t < H * H; t += 256) tile_acc[t] = 0.0;
  const int tile = tile0 + blockIdx.x;  // a rank of a sharded job fills its own range of tiles
  const int c0 = tile * GRAM_HUB_CELLS;
  const int c1 = min(n, c0 + GRAM_HUB_CELLS);
  int buf = 0;
  for (int j = c0; j < c1; ++j, buf ^= 1) {
    // warp 0 compacts the hub slots of cell j (slot order) into buffer `buf`
    if (threadIdx.x < 32) {
      const int cnt = a_cnt[j];
      int w = 0;
      for (int s0 = 0; s0 < cnt; s0 += 32) {
        const int s = s0 + threadIdx.x;
        int h = -1;
        double v = 0.0;
        if (s < cnt) {
          h = hub_of[a_idx[(size_t)j * S + s]];
          v = a_val[(size_t)j * S + s];
        }
        const unsigned ball = __ballot_sync(0xffffffffu, h >= 0);
        if (h >= 0) {
          const int pos = w + __popc(ball & ((1u << threadIdx.x) - 1u));
          hi[buf * FW_SMAX + pos] = h;
          hv[buf * FW_SMAX + pos] = v;
        }
        w += __popc(ball);
      }
      if (threadIdx.x == 0) hcnt[buf] = w;
    }
    __syncthreads();
    const int m = hcnt[buf];
    for (int e = threadIdx.x; e < m * m; e += 256) {
      const int p = e / m, q = e - p * m;
      double* dst = tile_acc + (size_t)hi[buf * FW_SMAX + p] * H + hi[buf * FW_SMAX + q];
      *dst = __dadd_rn(*dst, __dmul_rn(hv[buf * FW_SMAX + p], hv[buf * FW_SMAX + q]));
    }
    // the next iteration stages into the other buffer; its __syncthreads orders these adds before the adds of j+1
  }
  __syncthreads();
  double* out = tiles + (size_t)tile * H * H;
  for (int t = threadIdx.x; t < H * H; t += 256) out[t] = tile_acc[t];
}
__global__ void k_gram_hub_reduce(int k, int H, int ntiles, const int* __restrict__ hubs,
                                  const double* __restrict__ tiles, double* __restrict__ t1) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= H * H) return;
  double acc = 0.0;
  for (int c = 0; c < ntiles; ++c) acc = __dadd_rn(acc, tiles[(size_t)c * H * H + e]);
  const int p = e / H, q = e - p * H;
  t1[(size_t)hubs[p] * k + hubs[q]] = acc;
}

// ------------------------------------------------------------------------------------------------ candidates by archetype
// Per-cell rows of t2 = K A^T  ->  per-archetype candidate lists ordered by decreasing t2 (32-bit quantised).
// key = archetype << 32 | ~(upper word of t2's bit pattern): t2 > 0, so the upper word is monotone in t2.
__global__ void k_emit_cand(int rows, RowLists L, const long long* __restrict__ off,
                            unsigned long long* __restrict__ keys, unsigned* __restrict__ pos_val,
                            int* __restrict__ cellv, double* __restrict__ t2v) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= rows) return;
  const long long p = L.ptr[warp], o = off[warp];
  const int len = L.len[warp];
  for (int x = lane; x < len; x += 32) {
    const double v = L.val[p + x];
    const unsigned hi = (unsigned)((unsigned long long)__double_as_longlong(v) >> 32);
    keys[o + x] = ((unsigned long long)(unsigned)L.key[p + x] << 32) | (unsigned)(~hi);
    pos_val[o + x] = (unsigned)(o + x);
    cellv[o + x] = warp;
    t2v[o + x] = v;
  }
}
__global__ void k_gather_cand(long long total, const unsigned* __restrict__ order, const int* __restrict__ cellv,
                              const double* __restrict__ t2v, int* __restrict__ c_cell, double* __restrict__ c_t2) {
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= total) return;
  const unsigned src = order[e];
  c_cell[e] = cellv[src];
  c_t2[e] = t2v[src];
}

// ------------------------------------------------------------------------------------------------ _updateB, one FW step
struct UpdBArgs {
  int n, k, S, t;
  const double* t1;  // Gram, row a contiguous (symmetric)
  const long long* c_ptr;
  const int* c_cell;
  const double* c_t2;
  RowLists KB;   // (K B)[cell, :] for the current B
  RowLists T2c;  // (K A^T)[cell, :], first-occurrence order (dense-scan lookups)
  int* b_idx;
  double* b_val;
  int* b_cnt;
  int* trace;  // [T * k] or null
  int* amin;   // [k] or null: the cell chosen for every archetype at this step (feeds the incremental K B update)
  unsigned long long* evaluated;  // diagnostics: candidates whose exact gradient was computed
  // round 0 seeds the pruning bound from the r0n largest-t2 candidates and - from the second Frank-Wolfe step on - from
  // the cell the archetype chose at the previous step (prev, null at step 0), whose gradient is usually the minimum again
  int r0n = 64;
  const int* prev = nullptr;
  int row_begin = 0, row_end = 0;
};

constexpr int UPB_CHUNK = 1024;  // candidates per CTA
__device__ __forceinline__ void block_argmin_256(double& v, int& i, double* sg, int* si);

__global__ void k_chunk_count(int k, const long long* __restrict__ c_ptr, const int* __restrict__ hub_slot,
                              int* __restrict__ nch) {
  int a = blockIdx.x * blockDim.x + threadIdx.x;
  if (a > k) return;
  nch[a] = (a < k && hub_slot[a] < 0) ? (int)((c_ptr[a + 1] - c_ptr[a] + UPB_CHUNK - 1) / UPB_CHUNK) : 0;
}

// ---- hub block ------------------------------------------------------------------------------------------------------
// Archetypes for which (almost) every cell is a candidate are evaluated cell-major instead: one pass over the K B rows,
// each entry (l, v) adding v * t1[l, hubs] to a register tile of all hub columns at once (coalesced 8-byte lanes over
// the hub index), i.e. a CSR x dense product with the n_hub-wide slice of t1.  Same summation order per (cell, hub) as
// the candidate path (entries of the K B row in key order), so the gradients are bitwise the same numbers.
constexpr int HUB_MAX = 128;
constexpr int HUB_CELLS_PER_CTA = 256;

__global__ void k_hub_t1(int k, int nh, int nh_pad, const int* __restrict__ hub_ids, const double* __restrict__ t1,
                         double* __restrict__ T1H) {
  const int l = blockIdx.x * blockDim.x + threadIdx.x;
  const int s = blockIdx.y;
  if (l >= k) return;
  T1H[(size_t)l * nh_pad + s] = (s < nh) ? t1[(size_t)hub_ids[s] * k + l] : 0.0;
}
// The hub archetypes' share of t2 = K A^T is dense (every cell is a candidate), so it is not pushed through the sparse
// row products and the candidate sort at all: the hub slots of A are split off into a dense [cells x padded hubs] matrix
// AH and t2[:, hubs] = M (M AH) is two CSR x dense products with 8 x nh_pad-byte rows.  Same products, same order of
// the sums as the sparse path (which adds the same non-zero terms in ascending column order; the dense rows add exact
// zeros in between), multiply and add rounded separately: the values are bitwise those of the sparse rows.
__global__ void k_split_slots(int n, int S, int nh_pad, const int* __restrict__ hub_slot, const int* __restrict__ a_idx,
                              const double* __restrict__ a_val, const int* __restrict__ a_cnt,
                              int* __restrict__ o_key, double* __restrict__ o_val, int* __restrict__ o_len,
                              double* __restrict__ AH) {
  const int j = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (j >= n) return;
  const int cnt = a_cnt[j];
  int w = 0;
  for (int s0 = 0; s0 < cnt; s0 += 32) {
    const int s = s0 + lane;
    int a = -1, hs = -1;
    double v = 0.0;
    if (s < cnt) {
      a = a_idx[(size_t)j * S + s];
      v = a_val[(size_t)j * S + s];
      hs = hub_slot[a];
      if (hs >= 0) AH[(size_t)j * nh_pad + hs] = v;
    }
    const bool keep = a >= 0 && hs < 0;
    const unsigned ball = __ballot_sync(0xffffffffu, keep);
    if (keep) {
      const int pos = w + __popc(ball & ((1u << lane) - 1u));
      o_key[(size_t)j * S + pos] = a;
      o_val[(size_t)j * S + pos] = v;
    }
    w += __popc(ball);
  }
  if (lane == 0) o_len[j] = w;
}
// Y[i - out0, :] = sum_q M[i, q] X[q, :] for rows [r0, r1), nh_pad = 32 R columns; lane l owns columns l, l + 32, ...
template <int R>
__global__ void __launch_bounds__(256)
k_spmm_hubcols(int r0, int r1, int out0, int nh_pad, const int* __restrict__ indptr, const int* __restrict__ indices,
               const double* __restrict__ data, const double* __restrict__ X, double* __restrict__ Y) {
  const int i = r0 + ((blockIdx.x * blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
  if (i >= r1) return;
  double acc[R];
#pragma unroll
  for (int r = 0; r < R; ++r) acc[r] = 0.0;
  const int s0 = indptr[i], s1 = indptr[i + 1];
  for (int base = s0; base < s1; base += 32) {
    int q = 0;
    double m = 0.0;
    if (base + lane < s1) {
      q = indices[base + lane];
      m = data[base + lane];
    }
    const int cnt = min(32, s1 - base);
#pragma unroll 4
    for (int e = 0; e < cnt; ++e) {
      const int qq = __shfl_sync(0xffffffffu, q, e);
      const double mm = __shfl_sync(0xffffffffu, m, e);
      const double* __restrict__ xr = X + (size_t)qq * nh_pad + lane;
#pragma unroll
      for (int r = 0; r < R; ++r) acc[r] = __dadd_rn(acc[r], __dmul_rn(mm, __ldg(xr + r * 32)));
    }
  }
  double* yr = Y + (size_t)(i - out0) * nh_pad + lane;
#pragma unroll
  for (int r = 0; r < R; ++r) yr[r * 32] = acc[r];
}

// out[cta][slot][4] = {candidate min G, cell, non-candidate min G, cell} over the CTA's cells
// incremental = 0: HH[cell, :] = (rows of KB) x T1H, evaluated in full.
// incremental = 1: KB holds only the step's delta rows K E (see k_kb_merge) and HH the product for the previous B:
//                  HH <- c_old HH + c_new (delta rows x T1H)  - the same identity as for K B itself, 8x fewer entries.
// Four cells per warp iteration: their HH / t2 rows and list headers are requested before any of them is consumed, so a
// warp keeps ~20 independent loads in flight instead of walking one cell through four dependent round trips (the first
// version ran at 37 % warps active and 1.4 TB/s; it was latency bound, not bandwidth bound).  Lane l owns hubs l, l + 32,
// ...; R = nh_pad / 32.  Per (cell, hub) the arithmetic and its order are unchanged.
// (A version that staged the list headers and entries of a CTA's 256 cells in shared memory and kept HH / t2 in the
// locality order was measured and dropped: 1.5 ms against 0.94 ms for this one at 1M cells - its warps walk separate
// 32-cell blocks, the L1 hit rate of the t1 gathers fell from 70 % to 29 %, and the stages serialise behind barriers.)
template <int R, int HUB_UNROLL>
__global__ void __launch_bounds__(256, (HUB_UNROLL <= 2 ? 3 : 2)) k_updateB_hub(int row_begin, int row_end, int nh_pad,
                                                        const int* __restrict__ perm, RowLists KB,
                                                        const double* __restrict__ T1H, const double* __restrict__ T2H,
                                                        double* __restrict__ HH, int incremental, double c_old,
                                                        double c_new, double* __restrict__ out) {
  __shared__ double s_v[8][2][HUB_MAX];
  __shared__ int s_i[8][2][HUB_MAX];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  double cv[R], nv[R];
  int ci[R], ni[R];
#pragma unroll
  for (int r = 0; r < R; ++r) {
    cv[r] = nv[r] = FW_INF;
    ci[r] = ni[r] = 0x7fffffff;
  }
  const int c0 = row_begin + blockIdx.x * HUB_CELLS_PER_CTA;
  const int c1 = min(row_end, c0 + HUB_CELLS_PER_CTA);
  // Software pipeline over the warp's groups of HUB_UNROLL cells: the cell ids of the group after next and the list
  // headers of the next group are requested while the current group is processed, so the dependent chain
  // perm -> (ptr, len) -> entries -> t1 rows no longer starts from scratch in every iteration.
  int n_cell[HUB_UNROLL], n_L[HUB_UNROLL], nn_cell[HUB_UNROLL];
  long long n_p[HUB_UNROLL];
  {
    const int i0 = c0 + warp * HUB_UNROLL;
#pragma unroll
    for (int u = 0; u < HUB_UNROLL; ++u) {
      n_cell[u] = (i0 + u < c1) ? perm[i0 + u - row_begin] : -1;
      nn_cell[u] = (i0 + 8 * HUB_UNROLL + u < c1) ? perm[i0 + 8 * HUB_UNROLL + u - row_begin] : -1;
    }
#pragma unroll
    for (int u = 0; u < HUB_UNROLL; ++u) {
      n_p[u] = n_cell[u] >= 0 ? KB.ptr[n_cell[u]] : 0;
      n_L[u] = n_cell[u] >= 0 ? KB.len[n_cell[u]] : 0;
    }
  }
  for (int ii = c0 + warp * HUB_UNROLL; ii < c1; ii += 8 * HUB_UNROLL) {
    int cell[HUB_UNROLL], L[HUB_UNROLL];
    long long p[HUB_UNROLL];
    double hh[HUB_UNROLL][R], t2[HUB_UNROLL][R], acc[HUB_UNROLL][R];
#pragma unroll
    for (int u = 0; u < HUB_UNROLL; ++u) {
      cell[u] = n_cell[u];  // locality order; ties are resolved on the cell index
      p[u] = n_p[u];
      L[u] = n_L[u];
      if (cell[u] >= 0) {
        const size_t ro = (size_t)(cell[u] - row_begin) * nh_pad + lane;
#pragma unroll
        for (int r = 0; r < R; ++r) {
          hh[u][r] = incremental ? HH[ro + r * 32] : 0.0;
          t2[u][r] = T2H[ro + r * 32];
        }
      }
#pragma unroll
      for (int r = 0; r < R; ++r) acc[u][r] = 0.0;
    }
#pragma unroll
    for (int u = 0; u < HUB_UNROLL; ++u) {  // prefetch: headers of the next group, cell ids of the one after
      n_cell[u] = nn_cell[u];
      n_p[u] = n_cell[u] >= 0 ? KB.ptr[n_cell[u]] : 0;
      n_L[u] = n_cell[u] >= 0 ? KB.len[n_cell[u]] : 0;
      const int j = ii + 16 * HUB_UNROLL + u;
      nn_cell[u] = (j < c1) ? perm[j - row_begin] : -1;
    }
#pragma unroll
    for (int u = 0; u < HUB_UNROLL; ++u) {
      for (int x0 = 0; x0 < L[u]; x0 += 32) {
        int key = 0;
        double val = 0.0;
        if (x0 + lane < L[u]) {
          key = KB.key[p[u] + x0 + lane];
          val = KB.val[p[u] + x0 + lane];
        }
        const int cnt = min(32, L[u] - x0);
#pragma unroll 4
        for (int e = 0; e < cnt; ++e) {
          const int l = __shfl_sync(0xffffffffu, key, e);
          const double v = __shfl_sync(0xffffffffu, val, e);
          const double* __restrict__ trow = T1H + (size_t)l * nh_pad + lane;
#pragma unroll
          for (int r = 0; r < R; ++r) acc[u][r] = fma(v, __ldg(trow + r * 32), acc[u][r]);
        }
      }
    }
#pragma unroll
    for (int u = 0; u < HUB_UNROLL; ++u) {
      if (cell[u] < 0) continue;
      const int i = cell[u];
      double* __restrict__ hrow = HH + (size_t)(i - row_begin) * nh_pad + lane;
#pragma unroll
      for (int r = 0; r < R; ++r) {
        double a = acc[u][r];
        if (incremental) a = __dadd_rn(__dmul_rn(c_old, hh[u][r]), __dmul_rn(c_new, a));
        hrow[r * 32] = a;
        const double t2v = t2[u][r];
        const double G = 2.0 * (a - t2v);
        if (t2v > 0.0) {
          if (G < cv[r] || (G == cv[r] && i < ci[r])) {
            cv[r] = G;
            ci[r] = i;
          }
        } else if (G < nv[r] || (G == nv[r] && i < ni[r])) {
          nv[r] = G;
          ni[r] = i;
        }
      }
    }
  }
#pragma unroll
  for (int r = 0; r < R; ++r) {
    s_v[warp][0][r * 32 + lane] = cv[r];
    s_i[warp][0][r * 32 + lane] = ci[r];
    s_v[warp][1][r * 32 + lane] = nv[r];
    s_i[warp][1][r * 32 + lane] = ni[r];
  }
  __syncthreads();
  for (int t = threadIdx.x; t < 2 * nh_pad; t += 256) {
    const int kind = t / nh_pad, slot = t - kind * nh_pad;
    double v = s_v[0][kind][slot];
    int i = s_i[0][kind][slot];
    for (int w = 1; w < 8; ++w) {
      const double ov = s_v[w][kind][slot];
      const int oi = s_i[w][kind][slot];
      if (ov < v || (ov == v && oi < i)) {
        v = ov;
        i = oi;
      }
    }
    double* o = out + ((size_t)blockIdx.x * nh_pad + slot) * 4 + kind * 2;
    o[0] = v;
    o[1] = (double)i;
  }
}

__global__ void __launch_bounds__(256) k_updateB_hub_reduce(int nctas, int nh_pad, const int* __restrict__ hub_ids,
                                                            const double* __restrict__ in, double* __restrict__ partial) {
  __shared__ double sg[8];
  __shared__ int si[8];
  const int slot = blockIdx.x;
  double res[4];
  for (int kind = 0; kind < 2; ++kind) {
    double v = FW_INF;
    int i = 0x7fffffff;
    for (int c = threadIdx.x; c < nctas; c += 256) {
      const double* e = in + ((size_t)c * nh_pad + slot) * 4 + kind * 2;
      const double ov = e[0];
      const int oi = (int)e[1];
      if (ov < v || (ov == v && oi < i)) {
        v = ov;
        i = oi;
      }
    }
    block_argmin_256(v, i, sg, si);
    res[kind * 2] = v;
    res[kind * 2 + 1] = (double)i;
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    double* o = partial + (size_t)hub_ids[slot] * 4;
    o[0] = res[0];
    o[1] = res[1];
    o[2] = res[2];
    o[3] = res[3];
  }
}
// job-wide pruning bound: minimum over the ranks of the round-0 minima (all-gathered, [world][k])
__global__ void k_min_over_ranks(int k, int world, const double* __restrict__ all, double* __restrict__ out) {
  const int a = blockIdx.x * blockDim.x + threadIdx.x;
  if (a >= k) return;
  double m = all[a];
  for (int r = 1; r < world; ++r) m = fmin(m, all[(size_t)r * k + a]);
  out[a] = m;
}
__global__ void k_chunk_fill(int k, const long long* __restrict__ chunk_ptr, int* __restrict__ chunk_arch) {
  const int a = blockIdx.x;
  for (long long c = chunk_ptr[a] + threadIdx.x; c < chunk_ptr[a + 1]; c += blockDim.x) chunk_arch[c] = a;
}

// block-wide lexicographic (value, index) minimum; result valid on every thread
__device__ __forceinline__ void block_argmin_256(double& v, int& i, double* sg, int* si) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  warp_argmin(v, i);
  __syncthreads();
  if (lane == 0) {
    sg[warp] = v;
    si[warp] = i;
  }
  __syncthreads();
  v = sg[0];
  i = si[0];
  for (int w = 1; w < 8; ++w)
    if (sg[w] < v || (sg[w] == v && si[w] < i)) {
      v = sg[w];
      i = si[w];
    }
}

// Locality order: cells sorted by the archetype of their largest K B entry.  Kernels that gather rows of a k x k matrix
// (or of its hub slice) per cell walk the cells in this order so that neighbouring warps hit the same rows in L1/L2.
__global__ void k_perm_keys(int row_begin, int row_end, const int* __restrict__ lmax, unsigned long long* __restrict__ keys) {
  const int i = row_begin + blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= row_end) return;
  keys[i - row_begin] = ((unsigned long long)(unsigned)lmax[i] << 32) | (unsigned)i;
}
__global__ void k_perm_unpack(int count, const unsigned long long* __restrict__ keys, int* __restrict__ perm) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t < count) perm[t] = (int)(keys[t] & 0xffffffffu);
}

// largest entry of every K B row: (value, archetype); feeds the second pruning bound of k_updateB_eval
__global__ void k_row_max(int r0, int r1, RowLists L, double* __restrict__ mv, int* __restrict__ ml) {
  const int warp = r0 + ((blockIdx.x * blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
  if (warp >= r1) return;
  const long long p = L.ptr[warp];
  const int len = L.len[warp];
  double bv = -1.0;
  int bk = 0;
  for (int x = lane; x < len; x += 32) {
    const double v = L.val[p + x];
    if (v > bv) {
      bv = v;
      bk = L.key[p + x];
    }
  }
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) {
    const double ov = __shfl_xor_sync(0xffffffffu, bv, d);
    const int ok = __shfl_xor_sync(0xffffffffu, bk, d);
    if (ov > bv || (ov == bv && ok < bk)) {
      bv = ov;
      bk = ok;
    }
  }
  if (lane == 0) {
    mv[warp] = bv > 0.0 ? bv : 0.0;
    ml[warp] = bk;
  }
}

// Gradient of _updateB on the support of t2 (cpu.py:486): G[i,a] = 2 ((K B t1)[i,a] - t2[i,a]).
// One chunk of one archetype's candidate list per CTA.  Candidates are ordered by decreasing t2.  K, B, t1 >= 0 and monotone rounding
// give two exact lower bounds on the computed G[i,a]:
//     G >= -2 t2[i,a]                                   (the product term is >= 0)
//     G >= 2 (fl(KBmax_i * t1[lmax_i, a]) - t2[i,a])    (a sum of non-negative terms is >= any one of them)
// so once a first pass over the UPB_R0 largest-t2 candidates has produced a minimum `bound`, any candidate whose lower
// bound exceeds it can neither be the argmin nor tie with it and is skipped without reading its K B row.  This is what
// keeps long candidate lists affordable.  round 0 = first UPB_R0 candidates of every archetype,
// round 1 = everything else.
constexpr int UPB_R0 = 64;  // candidates evaluated unconditionally to seed the bound at step 0 (UpdBArgs::r0n later)
__global__ void __launch_bounds__(256) k_updateB_eval(UpdBArgs g, int round, const long long* __restrict__ chunk_ptr,
                                                      const int* __restrict__ chunk_arch,
                                                      const double* __restrict__ cmax_v, const int* __restrict__ cmax_l,
                                                      double* __restrict__ part_v, int* __restrict__ part_i,
                                                      double* __restrict__ bound_v,
                                                      const double* __restrict__ prune_bound) {
  __shared__ double sg[8];
  __shared__ int si[8];
  int a, chunk;
  if (round == 0) {
    a = blockIdx.x;
    chunk = 0;
    if (chunk_ptr[a + 1] == chunk_ptr[a]) {  // no candidates of this archetype on this rank (or a hub archetype)
      if (threadIdx.x == 0) bound_v[a] = FW_INF;  // the bounds are all-gathered: never leave a stale one behind
      return;
    }
  } else {
    a = chunk_arch[blockIdx.x];
    chunk = (int)((long long)blockIdx.x - chunk_ptr[a]);
  }
  const int lane = threadIdx.x & 31;
  const double* __restrict__ row = g.t1 + (size_t)a * g.k;
  const long long c0 = g.c_ptr[a] + (long long)chunk * UPB_CHUNK;
  const int nc = (int)min((long long)UPB_CHUNK, g.c_ptr[a + 1] - c0);
  const double bound = (round == 0) ? FW_INF : prune_bound[a];  // job-wide: the smallest round-0 minimum of any rank
  const int e_begin = (round == 1 && chunk == 0) ? g.r0n : 0;
  const int e_end = (round == 0) ? min(nc, g.r0n) : nc;
  if (round == 1 && e_begin < e_end) {
    // The list is sorted by decreasing upper word of t2, so every candidate from here on has t2 <= t2_ub (the first
    // one's upper word with all lower bits set).  If even that cannot beat the bound, the whole chunk - and every later
    // chunk of this archetype - is out by the first bound, without being read.
    const long long fb = __double_as_longlong(g.c_t2[c0 + e_begin]) | 0xffffffffll;
    if (-2.0 * __longlong_as_double(fb) > bound) {
      if (threadIdx.x == 0) {
        part_v[chunk_ptr[a] + chunk] = FW_INF;
        part_i[chunk_ptr[a] + chunk] = 0x7fffffff;
        if (g.evaluated) atomicAdd(g.evaluated + 2, 1ull);
      }
      return;
    }
  }
  // phase 1: cheap bound tests, survivors compacted into shared memory
  __shared__ int q_cell[UPB_CHUNK];
  __shared__ double q_t2[UPB_CHUNK];
  __shared__ int q_n;
  if (threadIdx.x == 0) q_n = 0;
  __syncthreads();
  for (int e0 = e_begin; e0 < e_end; e0 += 256) {
    const int e = e0 + threadIdx.x;
    int cell = 0;
    double t2v = 0.0;
    bool live = false;
    if (e < e_end) {
      t2v = g.c_t2[c0 + e];
      live = !(-2.0 * t2v > bound);
      if (live) {
        cell = g.c_cell[c0 + e];
        if (round == 1) {
          const double hb = __dmul_rn(cmax_v[cell], row[cmax_l[cell]]);
          live = !(2.0 * (hb - t2v) > bound);
        }
      }
    }
    const unsigned ball = __ballot_sync(0xffffffffu, live);
    int wbase = 0;
    if (lane == 0 && ball) wbase = atomicAdd(&q_n, __popc(ball));
    wbase = __shfl_sync(0xffffffffu, wbase, 0);
    if (live) {
      const int pos = wbase + __popc(ball & ((1u << lane) - 1u));
      q_cell[pos] = cell;
      q_t2[pos] = t2v;
    }
  }
  __syncthreads();
  if (round == 0 && g.prev) {  // the previous winner, if it is one of this rank's candidates of archetype a
    const int pc = g.prev[a];
    if (pc >= g.row_begin && pc < g.row_end) {
      const long long tp = g.T2c.ptr[pc];
      const int tl = g.T2c.len[pc];
      for (int x = threadIdx.x; x < tl; x += 256)
        if (g.T2c.key[tp + x] == a) {  // at most one entry matches
          const int pos = atomicAdd(&q_n, 1);
          q_cell[pos] = pc;
          q_t2[pos] = g.T2c.val[tp + x];
        }
    }
    __syncthreads();
  }
  const int nq = q_n;
  if (threadIdx.x == 0 && g.evaluated) atomicAdd(g.evaluated, (unsigned long long)nq);
  // phase 2: exact gradient of the survivors.  A chunk keeps a few dozen of its 1024 candidates, so one survivor per
  // thread would leave most of the CTA idle behind ~50 dependent gathers: 8 lanes share a survivor instead (lane s
  // takes entries s, s+8, ... of the K B row, then a fixed xor tree), 32 survivors per pass.  The summation order is a
  // function of the row alone, so the value of G[i,a] does not depend on chunking or sharding.
  double bv = FW_INF;
  int bi = 0x7fffffff;
  const int sub = threadIdx.x & 7;
  for (int e0 = 0; e0 < nq; e0 += 32) {
    const int e = e0 + (threadIdx.x >> 3);
    int cell = 0, L = 0;
    long long p = 0;
    double t2v = 0.0;
    bool live = false;
    if (e < nq) {
      cell = q_cell[e];
      t2v = q_t2[e];
      live = !(-2.0 * t2v > bv);  // the group's own running minimum tightens the first bound further
      if (live) {
        p = g.KB.ptr[cell];
        L = g.KB.len[cell];
      }
    }
    double acc = 0.0;
    if (live) {
      const int* __restrict__ kk = g.KB.key + p;
      const double* __restrict__ vv = g.KB.val + p;
#pragma unroll 2
      for (int x = sub; x < L; x += 8) acc = fma(__ldg(vv + x), __ldg(row + __ldg(kk + x)), acc);
    }
    acc += __shfl_xor_sync(0xffffffffu, acc, 4);
    acc += __shfl_xor_sync(0xffffffffu, acc, 2);
    acc += __shfl_xor_sync(0xffffffffu, acc, 1);
    if (live) {  // all 8 lanes of the group hold the same sum and keep the same running minimum
      const double G = 2.0 * (acc - t2v);
      if (G < bv || (G == bv && cell < bi)) {
        bv = G;
        bi = cell;
      }
    }
  }
  block_argmin_256(bv, bi, sg, si);
  if (threadIdx.x == 0) {
    if (round == 0) {
      bound_v[a] = bv;
      bound_v[g.k + a] = __longlong_as_double((long long)bi);  // carried to the step kernel as the round-0 winner
    } else {
      part_v[chunk_ptr[a] + chunk] = bv;
      part_i[chunk_ptr[a] + chunk] = bi;
    }
  }
}

// Per archetype, on the rank's own cells: reduce the chunk minima; if no local candidate is negative, scan the local
// cells outside the support of t2 (their gradient is 2 (K B t1)[i,a] >= 0) in ascending order for the first exact zero.
// out[a] = { candidate min G, its cell, non-candidate result G (0 for a zero, else the local minimum), its cell }.
__global__ void __launch_bounds__(256) k_updateB_local(UpdBArgs g, int row_begin, int row_end,
                                                       const long long* __restrict__ chunk_ptr,
                                                       const double* __restrict__ part_v,
                                                       const int* __restrict__ part_i,
                                                       const double* __restrict__ bound_v,
                                                       const double* __restrict__ prune_bound,
                                                       const int* __restrict__ hub_slot, double* __restrict__ out) {
  __shared__ double sg[8];
  __shared__ int si[8];
  __shared__ int s_zero;
  const int a = blockIdx.x;
  const int k = g.k;
  if (hub_slot[a] >= 0) return;  // evaluated cell-major by k_updateB_hub
  const double* __restrict__ row = g.t1 + (size_t)a * k;
  double bv = FW_INF;
  int bi = 0x7fffffff;
  if (threadIdx.x == 0 && chunk_ptr[a + 1] > chunk_ptr[a]) {  // winner of the first 256 candidates (round 0)
    bv = bound_v[a];
    bi = (int)__double_as_longlong(bound_v[k + a]);
  }
  for (long long c = chunk_ptr[a] + threadIdx.x; c < chunk_ptr[a + 1]; c += 256) {
    const double v = part_v[c];
    const int i = part_i[c];
    if (v < bv || (v == bv && i < bi)) {
      bv = v;
      bi = i;
    }
  }
  if (threadIdx.x == 0) s_zero = 0x7fffffff;
  block_argmin_256(bv, bi, sg, si);
  double nv = FW_INF;
  int ni = 0x7fffffff;
  const long long ncand = g.c_ptr[a + 1] - g.c_ptr[a];
  // (a negative job-wide bound means some rank holds a negative candidate: the column argmin is a candidate, and the
  // scan of this rank's cells outside the support would be thrown away by the merge)
  if (!(bv < 0.0) && !(prune_bound[a] < 0.0) && ncand < (long long)(row_end - row_begin)) {
    double tv = FW_INF;
    int ti = 0x7fffffff;
    bool found = false;
    for (int i0 = row_begin; i0 < row_end; i0 += 256) {
      const int i = i0 + threadIdx.x;
      if (i < row_end) {
        const long long tp = g.T2c.ptr[i];
        const int tl = g.T2c.len[i];
        bool is_cand = false;  // rows of t2 are in first-occurrence order: linear scan (this path is rare)
        for (int x = 0; x < tl && !is_cand; ++x) is_cand = g.T2c.key[tp + x] == a;
        if (!is_cand) {
          const long long p = g.KB.ptr[i];
          const int L = g.KB.len[i];
          double acc = 0.0;
          for (int x = 0; x < L; ++x) acc = fma(g.KB.val[p + x], row[g.KB.key[p + x]], acc);
          const double G = 2.0 * acc;
          if (G == 0.0) atomicMin(&s_zero, i);
          else if (G < tv || (G == tv && i < ti)) {
            tv = G;
            ti = i;
          }
        }
      }
      __syncthreads();
      const int z = s_zero;
      __syncthreads();  // everyone has read s_zero before the next chunk may lower it
      if (z != 0x7fffffff) {
        found = true;
        nv = 0.0;
        ni = z;
        break;
      }
    }
    if (!found) {
      block_argmin_256(tv, ti, sg, si);
      nv = tv;
      ni = ti;
    }
  }
  if (threadIdx.x == 0) {
    out[(size_t)a * 4 + 0] = bv;
    out[(size_t)a * 4 + 1] = (double)bi;
    out[(size_t)a * 4 + 2] = nv;
    out[(size_t)a * 4 + 3] = (double)ni;
  }
}

// Merge the per-rank partials (gathered[r][a][4]) into the column argmin - negative candidate minimum if there is one,
// else the first exact zero / first minimum over the whole column (np.argmin on the dense column) - and take the
// simplex step on column a of B (cpu.py:487-494).  Every rank runs this on identical inputs, so B stays replicated.
__global__ void __launch_bounds__(32) k_updateB_step(UpdBArgs g, const double* __restrict__ gathered, int n_ranks) {
  const int a = blockIdx.x;
  const int lane = threadIdx.x & 31;
  const int k = g.k, S = g.S;
  double cv = FW_INF, nv = FW_INF;
  int ci = 0x7fffffff, ni = 0x7fffffff;
  for (int r = 0; r < n_ranks; ++r) {
    const double* e = gathered + ((size_t)r * k + a) * 4;
    const double v0 = e[0], v2 = e[2];
    const int i0 = (int)e[1], i2 = (int)e[3];
    if (v0 < cv || (v0 == cv && i0 < ci)) {
      cv = v0;
      ci = i0;
    }
    if (v2 < nv || (v2 == nv && i2 < ni)) {
      nv = v2;
      ni = i2;
    }
  }
  int amin = ci;
  if (!(cv < 0.0)) {
    if (nv < cv || (nv == cv && ni < ci)) amin = ni;
    if (amin == 0x7fffffff) amin = 0;
  }
  {
    const int cnt = g.b_cnt[a];
    const double gamma = fw_gamma(g.t);
    int* bidx = g.b_idx + (size_t)a * S;
    double* bval = g.b_val + (size_t)a * S;
    int pos = -1;
    double bprev = 0.0;
    for (int e0 = 0; e0 < cnt; e0 += 32) {
      const int e = e0 + lane;
      const bool hit = e < cnt && bidx[e] == amin;
      const unsigned ball = __ballot_sync(0xffffffffu, hit);
      if (ball) {
        const int src = __ffs(ball) - 1;
        bprev = __shfl_sync(0xffffffffu, hit ? bval[e] : 0.0, src);
        pos = e0 + src;
        break;
      }
    }
    __syncwarp();
    if (g.t == 0) {
      if (lane == 0) {
        bidx[0] = amin;
        bval[0] = fw_step(bprev, 1.0, gamma);
        g.b_cnt[a] = 1;
      }
    } else {
      for (int e = lane; e < cnt; e += 32) bval[e] = fw_step(bval[e], e == pos ? 1.0 : 0.0, gamma);
      if (pos < 0 && lane == 0) {
        bidx[cnt] = amin;
        bval[cnt] = fw_step(0.0, 1.0, gamma);
        g.b_cnt[a] = cnt + 1;
      }
    }
    if (lane == 0 && g.trace) g.trace[(size_t)g.t * k + a] = amin;
    if (lane == 0 && g.amin) g.amin[a] = amin;
  }
}

// ------------------------------------------------------------------------------------------------ incremental K B
// Inside one _updateB call the simplex step is B <- (1 - gamma) B + gamma E with the same gamma for every column and one
// unit entry per column of E (the chosen cells), so  K B_{t+1} = (1 - gamma) K B_t + gamma K E_t  and K E_t = M (M E_t)
// is a product with k non-zeros on the right instead of nnz(B) (60x less work at 1M cells).  The reference re-evaluates
// K @ B from scratch (cpu.py:486); the identity is exact, only the rounding of the fp64 values differs (~1e-15
// relative), which matters solely through the argmin - SC_FIT_FRESH_KB=1 restores the re-evaluation for comparison.
__global__ void k_emit_e_pairs(int k, const int* __restrict__ amin, unsigned long long* __restrict__ keys,
                               double* __restrict__ ones) {
  const int a = blockIdx.x * blockDim.x + threadIdx.x;
  if (a >= k) return;
  keys[a] = ((unsigned long long)(unsigned)amin[a] << 32) | (unsigned)a;
  ones[a] = 1.0;
}

__global__ void k_merge_len(int n, const int* __restrict__ la, const int* __restrict__ lb, long long* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i > n) return;
  out[i] = (i < n) ? (long long)la[i] + (long long)lb[i] : 0ll;
}

// New[i,:] = c_old * Old[i,:] (+) c_new * D[i,:] for rows [r0, r1), one warp per row.  Entries of Old keep their order,
// keys that are new to the row are appended in D's order (both orders are fixed => bitwise reproducible).
__device__ __forceinline__ void kb_merge_row(int i, const RowLists& Old, const RowLists& D,
                                             const long long* __restrict__ new_ptr, int* __restrict__ new_len,
                                             int* __restrict__ new_key, double* __restrict__ new_val, double c_old,
                                             double c_new, double* __restrict__ max_v, int* __restrict__ max_l) {
  const int lane = threadIdx.x & 31;
  const int lo = Old.len[i], ld = D.len[i];
  const long long po = Old.ptr[i], pd = D.ptr[i], pn = new_ptr[i];
  // the row's largest entry (value, then smaller key) rides along: it feeds the second pruning bound of k_updateB_eval
  double bv = -1.0;
  int bk = 0x7fffffff;
  auto track = [&](double v, int key) {
    if (v > bv || (v == bv && key < bk)) {
      bv = v;
      bk = key;
    }
  };
  auto finish_max = [&]() {
#pragma unroll
    for (int dd = 16; dd > 0; dd >>= 1) {
      const double ov = __shfl_xor_sync(0xffffffffu, bv, dd);
      const int ok2 = __shfl_xor_sync(0xffffffffu, bk, dd);
      if (ov > bv || (ov == bv && ok2 < bk)) {
        bv = ov;
        bk = ok2;
      }
    }
    if (lane == 0) {
      max_v[i] = bv > 0.0 ? bv : 0.0;
      max_l[i] = bv > 0.0 ? bk : 0;
    }
  };
  if (ld <= 32) {
    // common case: the delta row fits one batch - one pass over the old row
    const bool dvd = lane < ld;
    const int dk = dvd ? D.key[pd + lane] : -1;
    const double dv = dvd ? D.val[pd + lane] : 0.0;
    bool matched = false;
    for (int ob = 0; ob < lo; ob += 32) {
      const bool ov = ob + lane < lo;
      const int ok = ov ? Old.key[po + ob + lane] : -2;
      double val = ov ? __dmul_rn(c_old, Old.val[po + ob + lane]) : 0.0;
      for (int j = 0; j < ld; ++j) {
        const int kj = __shfl_sync(0xffffffffu, dk, j);
        const double vj = __shfl_sync(0xffffffffu, dv, j);
        const bool hit = ok == kj;
        if (hit) val = __dadd_rn(val, __dmul_rn(c_new, vj));
        if (__any_sync(0xffffffffu, hit) && lane == j) matched = true;
      }
      if (ov) {
        new_key[pn + ob + lane] = ok;
        new_val[pn + ob + lane] = val;
        track(val, ok);
      }
    }
    const bool app = dvd && !matched;
    const unsigned ball = __ballot_sync(0xffffffffu, app);
    if (app) {
      const int pos = lo + __popc(ball & ((1u << lane) - 1u));
      const double nv = __dmul_rn(c_new, dv);
      new_key[pn + pos] = dk;
      new_val[pn + pos] = nv;
      track(nv, dk);
    }
    if (lane == 0) new_len[i] = lo + __popc(ball);
    finish_max();
    return;
  }
  // old entries, scaled, plus the matching delta entry if there is one
  for (int ob = 0; ob < lo; ob += 32) {
    const bool ov = ob + lane < lo;
    const int ok = ov ? Old.key[po + ob + lane] : -2;
    double val = ov ? __dmul_rn(c_old, Old.val[po + ob + lane]) : 0.0;
    for (int db = 0; db < ld; db += 32) {
      const bool dvd = db + lane < ld;
      const int dk = dvd ? D.key[pd + db + lane] : -1;
      const double dv = dvd ? D.val[pd + db + lane] : 0.0;
      const int cnt = min(32, ld - db);
      for (int j = 0; j < cnt; ++j) {
        const int kj = __shfl_sync(0xffffffffu, dk, j);
        const double vj = __shfl_sync(0xffffffffu, dv, j);
        if (ok == kj) val = __dadd_rn(val, __dmul_rn(c_new, vj));
      }
    }
    if (ov) {
      new_key[pn + ob + lane] = ok;
      new_val[pn + ob + lane] = val;
      track(val, ok);
    }
  }
  // delta entries whose key is new to the row
  int napp = 0;
  for (int db = 0; db < ld; db += 32) {
    const bool dvd = db + lane < ld;
    const int dk = dvd ? D.key[pd + db + lane] : -1;
    const double dv = dvd ? D.val[pd + db + lane] : 0.0;
    bool matched = false;
    for (int ob = 0; ob < lo; ob += 32) {
      const int ok = (ob + lane < lo) ? Old.key[po + ob + lane] : -2;
      const int cnt = min(32, lo - ob);
      for (int j = 0; j < cnt; ++j) matched |= (__shfl_sync(0xffffffffu, ok, j) == dk);
    }
    const bool app = dvd && !matched;
    const unsigned ball = __ballot_sync(0xffffffffu, app);
    if (app) {
      const int pos = lo + napp + __popc(ball & ((1u << lane) - 1u));
      const double nv = __dmul_rn(c_new, dv);
      new_key[pn + pos] = dk;
      new_val[pn + pos] = nv;
      track(nv, dk);
    }
    napp += __popc(ball);
  }
  if (lane == 0) new_len[i] = lo + napp;
  finish_max();
}
// generic kernel: the listed rows (handed over by the staged kernel), one warp per row
__global__ void __launch_bounds__(256)
k_kb_merge(RowLists Old, RowLists D, const long long* __restrict__ new_ptr, int* __restrict__ new_len,
           int* __restrict__ new_key, double* __restrict__ new_val, double c_old, double c_new,
           double* __restrict__ max_v, int* __restrict__ max_l, const int* __restrict__ list,
           const int* __restrict__ list_count) {
  const int warps = (gridDim.x * blockDim.x) >> 5;
  const int count = *list_count;
  for (int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; w < count; w += warps)
    kb_merge_row(list[w], Old, D, new_ptr, new_len, new_key, new_val, c_old, c_new, max_v, max_l);
}

// Staged version of the same merge (same arithmetic, same output order).  The warp-per-row kernel above spent ~670 warp
// instructions per row broadcasting every delta entry to every batch of old entries with shuffles and votes (ncu: 60 % SM
// throughput, 1.6 GB of traffic in 0.98 ms).  Here a CTA stages the delta rows of its 64 cells in shared memory; a lane
// that holds an old entry compares it against the staged keys with broadcast shared-memory reads, marks matches in a
// shared flag array, and the unmatched delta entries are appended in their (canonical) order.
constexpr int KBM_ROWS = 64;     // rows per CTA (8 per warp)
constexpr int KBM_ECAP = 2048;   // staged delta entries per CTA
__global__ void __launch_bounds__(256)
k_kb_merge_staged(int r0, int r1, RowLists Old, RowLists D, const long long* __restrict__ new_ptr,
                  int* __restrict__ new_len, int* __restrict__ new_key, double* __restrict__ new_val, double c_old,
                  double c_new, double* __restrict__ max_v, int* __restrict__ max_l, int* __restrict__ slow_list,
                  int* __restrict__ slow_count) {
  __shared__ double s_dv[KBM_ECAP];
  __shared__ __align__(16) int s_dk[KBM_ECAP];  // every row's keys start at a multiple of 4 and are padded to one (key -3)
  __shared__ unsigned char s_match[KBM_ECAP];
  __shared__ long long s_pd[KBM_ROWS];
  __shared__ int s_ld[KBM_ROWS], s_doff[KBM_ROWS + 1];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row0 = r0 + blockIdx.x * KBM_ROWS;
  const int nrow = min(KBM_ROWS, r1 - row0);
  if (threadIdx.x < KBM_ROWS) {
    int ld = 0;
    if ((int)threadIdx.x < nrow) {
      ld = D.len[row0 + threadIdx.x];
      s_pd[threadIdx.x] = D.ptr[row0 + threadIdx.x];
    }
    s_ld[threadIdx.x] = ld;
  }
  __syncthreads();
  if (warp == 0) {  // exclusive scan of the 64 lengths, each rounded up to a multiple of 4, by one warp
    const int a = (s_ld[lane] + 3) & ~3, b = (s_ld[lane + 32] + 3) & ~3;
    int ia = a, ib = b;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const int va = __shfl_up_sync(0xffffffffu, ia, d), vb = __shfl_up_sync(0xffffffffu, ib, d);
      if (lane >= d) {
        ia += va;
        ib += vb;
      }
    }
    const int tota = __shfl_sync(0xffffffffu, ia, 31);
    s_doff[lane] = ia - a;
    s_doff[lane + 32] = tota + ib - b;
    if (lane == 31) s_doff[64] = tota + ib;
  }
  for (int t = threadIdx.x; t < KBM_ECAP; t += 256) s_match[t] = 0;
  __syncthreads();
  for (int j = 0; j < KBM_ROWS / 8; ++j) {  // stage the delta rows: warp w takes rows w, w + 8, ...
    const int rr = warp + 8 * j;
    const int ld = s_ld[rr], doff = s_doff[rr], ldp = (ld + 3) & ~3;
    if (doff + ldp > KBM_ECAP) continue;
    const long long pd = s_pd[rr];
    for (int x = lane; x < ldp; x += 32) {
      s_dk[doff + x] = x < ld ? D.key[pd + x] : -3;
      s_dv[doff + x] = x < ld ? D.val[pd + x] : 0.0;
    }
  }
  __syncthreads();
  for (int j = 0; j < KBM_ROWS / 8; ++j) {
    const int rr = warp + 8 * j;
    if (rr >= nrow) break;
    const int i = row0 + rr;
    const int ld = s_ld[rr], doff = s_doff[rr], ldp = (ld + 3) & ~3;
    if (doff + ldp > KBM_ECAP) {  // does not fit the staging area: the generic kernel takes this row
      if (lane == 0) slow_list[atomicAdd(slow_count, 1)] = i;
      continue;
    }
    const int lo = Old.len[i];
    const long long po = Old.ptr[i], pn = new_ptr[i];
    double bv = -1.0;
    int bk = 0x7fffffff;
    for (int ob = 0; ob < lo; ob += 32) {
      const bool ov = ob + lane < lo;
      int ok = -2;
      double val = 0.0;
      if (ov) {
        ok = Old.key[po + ob + lane];
        val = __dmul_rn(c_old, Old.val[po + ob + lane]);
      }
      // four staged keys per shared-memory load (an old key equals at most one delta key; pad keys never match)
      for (int d = 0; d < ldp; d += 4) {
        const int4 kk = *reinterpret_cast<const int4*>(&s_dk[doff + d]);
        int hit = -1;
        if (ok == kk.x) hit = d;
        if (ok == kk.y) hit = d + 1;
        if (ok == kk.z) hit = d + 2;
        if (ok == kk.w) hit = d + 3;
        if (hit >= 0) {
          val = __dadd_rn(val, __dmul_rn(c_new, s_dv[doff + hit]));
          s_match[doff + hit] = 1;
        }
      }
      if (ov) {
        new_key[pn + ob + lane] = ok;
        new_val[pn + ob + lane] = val;
        if (val > bv || (val == bv && ok < bk)) {
          bv = val;
          bk = ok;
        }
      }
    }
    __syncwarp();
    int napp = 0;
    for (int db = 0; db < ld; db += 32) {
      const bool app = db + lane < ld && !s_match[doff + db + lane];
      const unsigned ball = __ballot_sync(0xffffffffu, app);
      if (app) {
        const int pos = lo + napp + __popc(ball & ((1u << lane) - 1u));
        const int dk = s_dk[doff + db + lane];
        const double nvv = __dmul_rn(c_new, s_dv[doff + db + lane]);
        new_key[pn + pos] = dk;
        new_val[pn + pos] = nvv;
        if (nvv > bv || (nvv == bv && dk < bk)) {
          bv = nvv;
          bk = dk;
        }
      }
      napp += __popc(ball);
    }
    if (lane == 0) new_len[i] = lo + napp;
#pragma unroll
    for (int dd = 16; dd > 0; dd >>= 1) {
      const double ov2 = __shfl_xor_sync(0xffffffffu, bv, dd);
      const int ok2 = __shfl_xor_sync(0xffffffffu, bk, dd);
      if (ov2 > bv || (ov2 == bv && ok2 < bk)) {
        bv = ov2;
        bk = ok2;
      }
    }
    if (lane == 0) {
      max_v[i] = bv > 0.0 ? bv : 0.0;
      max_l[i] = bv > 0.0 ? bk : 0;
    }
  }
}

// ------------------------------------------------------------------------------------------------ RSS, labels
// out[j] = a_j^T t1A a_j - 2 <KB[j,:], a_j>     (||M - M B A||_F^2 = ||M||_F^2 + sum_j out[j])
__global__ void k_rss_cell(int r0, int r1, int k, int S, const int* __restrict__ a_idx, const double* __restrict__ a_val,
                           const int* __restrict__ a_cnt, RowLists KB, const double* __restrict__ t1T,
                           double* __restrict__ out) {
  const int warp = r0 + ((blockIdx.x * blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
  if (warp >= r1) return;
  const int j = warp;
  const int cnt = a_cnt[j];
  const int* idx = a_idx + (size_t)j * S;
  const double* val = a_val + (size_t)j * S;
  const long long kp = KB.ptr[j];
  const int L = KB.len[j];
  double cross = 0.0;
  for (int s = lane; s < cnt; s += 32) cross = fma(val[s], list_lookup(KB.key + kp, KB.val + kp, L, idx[s]), cross);
  cross = warp_sum(cross);
  double quad = 0.0;
  for (int s = 0; s < cnt; ++s) {
    double inner = 0.0;
    const double* trow = t1T + (size_t)idx[s] * k;
    for (int s2 = lane; s2 < cnt; s2 += 32) inner = fma(trow[idx[s2]], val[s2], inner);
    inner = warp_sum(inner);
    quad = fma(val[s], inner, quad);
  }
  if (lane == 0) out[j] = quad - 2.0 * cross;
}

// column argmax of A, first index on ties (cpu.py:717)
__global__ void k_hard_labels(int rows, int S, const int* __restrict__ idx, const double* __restrict__ val,
                              const int* __restrict__ cnt, int* __restrict__ label) {
  int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= rows) return;
  double bv = 0.0;  // implicit zeros: an empty column gives index 0, as argmax over zeros does
  int bi = 0x7fffffff;
  const int c = cnt[j];
  for (int s = 0; s < c; ++s) {
    const double v = val[(size_t)j * S + s];
    const int i = idx[(size_t)j * S + s];
    if (v > bv || (v == bv && i < bi && v > 0.0)) {
      bv = v;
      bi = i;
    }
  }
  label[j] = (bi == 0x7fffffff) ? 0 : bi;
}

// top-5 (archetype, weight) per cell by repeated argmax with masking (cpu_dense.py:586-605)
__global__ void k_soft_top(int n, int k, int S, int top, const int* __restrict__ idx, const double* __restrict__ val,
                           const int* __restrict__ cnt, int* __restrict__ out_idx, double* __restrict__ out_w) {
  int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  const int c = cnt[j];
  unsigned long long used = 0ull;  // S <= 64
  int chosen[8];
  for (int r = 0; r < top; ++r) {
    double bv = 0.0;
    int bi = -1, bs = -1;
    for (int s = 0; s < c; ++s) {
      if (used >> s & 1ull) continue;
      const double v = val[(size_t)j * S + s];
      const int i = idx[(size_t)j * S + s];
      if (v > bv || (v == bv && v > 0.0 && i < bi)) {
        bv = v;
        bi = i;
        bs = s;
      }
    }
    if (bs >= 0) {
      used |= 1ull << bs;
    } else {
      // all remaining entries are zero: argmax returns the smallest archetype not masked to -1 yet
      int cand = 0;
      bool again = true;
      while (again) {
        again = false;
        for (int q = 0; q < r; ++q)
          if (chosen[q] == cand) {
            cand++;
            again = true;
          }
      }
      bi = cand < k ? cand : 0;
      bv = 0.0;
    }
    chosen[r] = bi;
    out_idx[(size_t)j * top + r] = bi;
    out_w[(size_t)j * top + r] = bv;
  }
}

// ================================================================================================ host side
#include <chrono>
#include <cstdlib>
static double now_s() {
  return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}
enum { P_BROW = 0, P_SPGEMM_Y, P_SPGEMM_KB, P_T1A, P_UPDA_FAST, P_UPDA_SLOW, P_GRAM, P_T2ROWS, P_T2TRANS, P_GEVAL, P_RSS };
void sc_fit::pstart() {
  if (!profile) return;
  cudaStreamSynchronize(ctx->stream);
  prof_t0 = now_s();
}
void sc_fit::pstop(int stage) {
  if (!profile) return;
  cudaStreamSynchronize(ctx->stream);
  prof[stage] += now_s() - prof_t0;
}
static int set_smem_attr_once(sc_ctx* ctx) {  // per device, so once per engine
  SC_CUDA(ctx, cudaFuncSetAttribute(k_updateA_fast<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    (int)(FWA_PER_WARP * FWA_WARPS)));
  SC_CUDA(ctx, cudaFuncSetAttribute(k_updateA_fast<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    (int)(FWA_PER_WARP * FWA_WARPS)));
  return SC_OK;
}

sc_fit::~sc_fit() {
  for (auto e : ev_a) cudaEventDestroy(e);
  for (auto e : ev_b) cudaEventDestroy(e);
  for (auto e : ev_h0) cudaEventDestroy(e);
  for (auto e : ev_h1) cudaEventDestroy(e);
  if (ev_fork) cudaEventDestroy(ev_fork);
  if (ev_join) cudaEventDestroy(ev_join);
  if (stream2) {
    cudaStreamSynchronize(stream2);
    cudaStreamDestroy(stream2);
  }
}

int sc_fit::init(sc_ctx* c, int n_, int k_, int T_) {
  ctx = c;
  profile = std::getenv("SC_PROFILE") != nullptr;
  fresh_kb = std::getenv("SC_FIT_FRESH_KB") != nullptr;
  overlap_hub = std::getenv("SC_NO_OVERLAP") == nullptr;
  if (const char* e = std::getenv("SC_HUB_U")) hub_u = atoi(e) == 4 ? 4 : 2;
  SC_CUDA(ctx, cudaStreamCreateWithFlags(&stream2, cudaStreamNonBlocking));
  SC_CUDA(ctx, cudaEventCreateWithFlags(&ev_fork, cudaEventDisableTiming));
  SC_CUDA(ctx, cudaEventCreateWithFlags(&ev_join, cudaEventDisableTiming));
  resum_A = std::getenv("SC_FIT_RESUM_A") != nullptr;
  if (const char* e = std::getenv("SC_HUB_MAX")) hub_cap = std::max(0, std::min(HUB_MAX, atoi(e)));
  if (const char* e = std::getenv("SC_R0_SEED")) r0_seed = std::max(0, std::min(UPB_R0, atoi(e)));
  n = n_;
  row_begin = 0;
  row_end = n_;
  slab = n_;
  if (ctx->world() > 1) {  // contiguous, equal slabs of cells; the last ranks may own a short or empty slab
    const int w = ctx->world();
    if (w > 64) SC_FAIL(ctx, SC_ERR_ARG, "fit: at most 64 ranks");
    slab = (n_ + w - 1) / w;
    row_begin = std::min(n_, ctx->rank() * slab);
    row_end = std::min(n_, row_begin + slab);
  }
  k = k_;
  T = T_;
  S = std::max(T_, 1);
  if (n <= 0 || k <= 0 || k > n) SC_FAIL(ctx, SC_ERR_ARG, "fit: need 0 < k <= n");
  if (T < 1 || T > FW_SMAX) SC_FAIL(ctx, SC_ERR_ARG, "fit: max_FW_iter must be in [1, 64]");
  SC_TRY(slot_ptr_n.ensure(ctx, (size_t)n + 1));
  SC_LAUNCH(ctx, k_fill_stride_ptr, (n + 256) / 256, 256, 0, slot_ptr_n.p, n, S);
  for (int b = 0; b < 2; ++b) {
    SC_TRY(a_idx[b].ensure(ctx, ((size_t)n + 64) * S));  // + padding rows so equal slabs can be all-gathered in place
    SC_TRY(a_val[b].ensure(ctx, ((size_t)n + 64) * S));
    SC_TRY(a_cnt[b].ensure(ctx, (size_t)n + 64));
  }
  SC_TRY(b_idx.ensure(ctx, (size_t)k * S));
  SC_TRY(b_val.ensure(ctx, (size_t)k * S));
  SC_TRY(b_cnt.ensure(ctx, (size_t)k));
  SC_TRY(t1A.ensure(ctx, (size_t)k * k));
  {
    const int w = ctx->world();
    SC_TRY(t1B.ensure(ctx, (size_t)((k + w - 1) / w) * w * k));  // padded to equal archetype slabs (rows are all-gathered)
  }
  SC_TRY(percell.ensure(ctx, (size_t)std::max(n, 1024) + 64));
  SC_TRY(scalars.ensure(ctx, 8));
  SC_TRY(slow_list.ensure(ctx, (size_t)n));
  SC_TRY(slow_cnt.ensure(ctx, 4));
  SC_TRY(cnt_n.ensure(ctx, (size_t)n + 1));
  SC_TRY(cnt_k.ensure(ctx, (size_t)k + 1));
  SC_TRY(amin_cur.ensure(ctx, (size_t)k + 1));
  SC_TRY(set_smem_attr_once(ctx));
  return SC_OK;
}

int sc_fit::set_kernel(const int* indptr, const int* indices, const double* data, long long nnz) {
  SC_TRY(m_ptr.ensure(ctx, (size_t)n + 1));
  SC_TRY(m_col.ensure(ctx, (size_t)nnz));
  SC_TRY(m_val.ensure(ctx, (size_t)nnz));
  SC_CUDA(ctx, sc_copy(m_ptr.p, indptr, sizeof(int) * ((size_t)n + 1), ctx->stream));
  SC_CUDA(ctx, sc_copy(m_col.p, indices, sizeof(int) * (size_t)nnz, ctx->stream));
  SC_CUDA(ctx, sc_copy(m_val.p, data, sizeof(double) * (size_t)nnz, ctx->stream));
  m_nnz = nnz;
  {
    SC_CUDA(ctx, cudaMemsetAsync(slow_cnt.p, 0, sizeof(int), ctx->stream));
    SC_LAUNCH(ctx, k_check_kernel, (n + 127) / 128, 128, 0, n, m_ptr.p, m_col.p, m_val.p, slow_cnt.p);
    int flag = 0;
    SC_CUDA(ctx, cudaMemcpyAsync(&flag, slow_cnt.p, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    SC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (flag & 1) SC_FAIL(ctx, SC_ERR_ARG, "kernel matrix has negative or NaN entries (the engine needs M >= 0)");
    if (flag & 2)
      SC_FAIL(ctx, SC_ERR_ARG,
              "kernel matrix is not symmetric: the engine applies K = M M^T as M (M .), which equals the reference's "
              "K only for M == M^T (every matrix SEACellGraph.rbf builds is)");
  }
  // ||M||_F^2 (fixed-order reduction)
  DBuf<double> sq;
  SC_TRY(sq.ensure(ctx, (size_t)nnz));
  SC_LAUNCH(ctx, k_square, (unsigned)((nnz + 255) / 256), 256, 0, m_val.p, sq.p, (size_t)nnz);
  SC_TRY(sc_sum_f64(ctx, sq.p, scalars.p, (size_t)nnz, tmp));
  SC_CUDA(ctx, cudaMemcpyAsync(&m_fro2, scalars.p, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  SC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  // K = M M^T (cpu.py:128,157), rows in first-occurrence order; a rank of a sharded job keeps the columns of its own
  // slab of cells only (see "products with K" above): K[:, slab] = M (M restricted to the columns of the slab)
  {
    CsrView M{n, m_ptr.p, m_col.p, m_val.p};
    SC_TRY(m_ptr64.ensure(ctx, (size_t)n + 1));
    SC_TRY(m_len.ensure(ctx, (size_t)n + 1));
    if (row_begin == 0 && row_end == n) {
      SC_LAUNCH(ctx, k_csr_as_lists, (n + 256) / 256, 256, 0, n, m_ptr.p, m_ptr64.p, m_len.p);
      RowLists Ml{m_ptr64.p, m_len.p, m_col.p, m_val.p};
      SC_TRY(sc_spgemm_rows(ctx, M, 0, n, Ml, n, Kl, sp, false, true));
    } else {
      DBuf<int> fkey;
      DBuf<double> fval;
      SC_LAUNCH(ctx, k_filter_cols_count, (n + 256) / 256, 256, 0, n, row_begin, row_end, m_ptr.p, m_col.p, cnt_n.p);
      SC_TRY(sc_exclusive_scan_i32_to_i64(ctx, cnt_n.p, m_ptr64.p, (size_t)n + 1, tmp));
      long long fnnz = 0;
      SC_CUDA(ctx, cudaMemcpyAsync(&fnnz, m_ptr64.p + n, sizeof(long long), cudaMemcpyDeviceToHost, ctx->stream));
      SC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
      SC_TRY(fkey.ensure(ctx, (size_t)fnnz + 1));
      SC_TRY(fval.ensure(ctx, (size_t)fnnz + 1));
      SC_LAUNCH(ctx, k_filter_cols_fill, (n + 255) / 256, 256, 0, n, row_begin, row_end, m_ptr.p, m_col.p, m_val.p,
                m_ptr64.p, m_len.p, fkey.p, fval.p);
      RowLists Ml{m_ptr64.p, m_len.p, fkey.p, fval.p};
      SC_TRY(sc_spgemm_rows(ctx, M, 0, n, Ml, n, Kl, sp, false, true));
      SC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));  // fkey / fval go out of scope
    }
  }
  {
    SC_CUDA(ctx, cudaMemsetAsync(slow_cnt.p, 0, sizeof(int), ctx->stream));
    SC_LAUNCH(ctx, k_max_int, (n + 255) / 256, 256, 0, n, Kl.len.p, slow_cnt.p);
    SC_CUDA(ctx, cudaMemcpyAsync(&k_maxrow, slow_cnt.p, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    SC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  }
  have_M = true;
  pre_valid = false;
  return SC_OK;
}

int sc_fit::set_B_onehot(const long long* archetypes_host) {
  std::vector<int> idx((size_t)k * S, 0), cnt(k, 1);
  std::vector<double> val((size_t)k * S, 0.0);
  for (int a = 0; a < k; ++a) {
    long long c = archetypes_host[a];
    if (c < 0 || c >= n) SC_FAIL(ctx, SC_ERR_ARG, "fit: archetype position out of range");
    idx[(size_t)a * S] = (int)c;
    val[(size_t)a * S] = 1.0;
  }
  return set_B_slots(idx.data(), val.data(), cnt.data());
}

int sc_fit::set_B_slots(const int* idx, const double* val, const int* cnt) {
  SC_CUDA(ctx, sc_copy(b_idx.p, idx, sizeof(int) * (size_t)k * S, ctx->stream));
  SC_CUDA(ctx, sc_copy(b_val.p, val, sizeof(double) * (size_t)k * S, ctx->stream));
  SC_CUDA(ctx, sc_copy(b_cnt.p, cnt, sizeof(int) * (size_t)k, ctx->stream));
  SC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  have_B = true;
  pre_valid = false;
  return SC_OK;
}

int sc_fit::set_A_lists(const long long* colptr, const int* idx, const double* val) {
  // A as CSC (column j = cell j): becomes the "previous A" of the next _updateA
  long long nnz = 0;
  std::vector<long long> hp((size_t)n + 1);
  SC_CUDA(ctx, sc_copy(hp.data(), colptr, sizeof(long long) * ((size_t)n + 1), ctx->stream));
  SC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  nnz = hp[n];
  std::vector<int> hl(n);
  for (int j = 0; j < n; ++j) hl[j] = (int)(hp[j + 1] - hp[j]);
  SC_TRY(a0.ptr.ensure(ctx, (size_t)n + 1));
  SC_TRY(a0.len.ensure(ctx, (size_t)n));
  SC_TRY(a0.key.ensure(ctx, (size_t)nnz + 1));
  SC_TRY(a0.val.ensure(ctx, (size_t)nnz + 1));
  SC_CUDA(ctx, sc_copy(a0.ptr.p, hp.data(), sizeof(long long) * ((size_t)n + 1), ctx->stream));
  SC_CUDA(ctx, sc_copy(a0.len.p, hl.data(), sizeof(int) * (size_t)n, ctx->stream));
  SC_CUDA(ctx, sc_copy(a0.key.p, idx, sizeof(int) * (size_t)nnz, ctx->stream));
  SC_CUDA(ctx, sc_copy(a0.val.p, val, sizeof(double) * (size_t)nnz, ctx->stream));
  int flag = 0;
  if (nnz > 0) {
    SC_CUDA(ctx, cudaMemsetAsync(slow_cnt.p, 0, sizeof(int), ctx->stream));
    SC_LAUNCH(ctx, k_check_nonneg, (unsigned)((nnz + 255) / 256), 256, 0, (size_t)nnz, a0.val.p, slow_cnt.p);
    SC_CUDA(ctx, cudaMemcpyAsync(&flag, slow_cnt.p, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  }
  SC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  if (flag) SC_FAIL(ctx, SC_ERR_ARG, "initial assignments must be non-negative");
  a0.rows = n;
  a_cur = -1;
  have_A = true;
  a_is_slots = false;
  return SC_OK;
}

int sc_fit::set_A_slots(const int* idx, const double* val, const int* cnt) {
  a_cur = 0;
  SC_CUDA(ctx, sc_copy(a_idx[0].p, idx, sizeof(int) * (size_t)n * S, ctx->stream));
  SC_CUDA(ctx, sc_copy(a_val[0].p, val, sizeof(double) * (size_t)n * S, ctx->stream));
  SC_CUDA(ctx, sc_copy(a_cnt[0].p, cnt, sizeof(int) * (size_t)n, ctx->stream));
  SC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  have_A = true;
  a_is_slots = true;
  return SC_OK;
}

RowLists sc_fit::A_view() const {
  if (!a_is_slots) return a0.view();
  return RowLists{slot_ptr_n.p, a_cnt[a_cur].p, a_idx[a_cur].p, a_val[a_cur].p};
}

// Z[i, :] = sum over the sources (c, a, w) of w K[c, i] for the kept cells i; rows come out sorted by archetype.
// has_dups: several sources may share (i, a) (K B: the members of archetype a within two hops of i) and are summed in
// source order; the one-hot E_t of a Frank-Wolfe step has one source per archetype, so nothing needs summing.
int sc_fit::push_product(const PushSrc& src, bool has_dups, RowListsBuf& Z) {
  const int m = src.m;
  int kbits = 1, nbits = 1;
  while ((1ll << kbits) < k) kbits++;
  while ((1ll << nbits) < n) nbits++;
  RowLists Kv = Kl.view();
  SC_TRY(Z.ptr.ensure(ctx, (size_t)n + 1));
  SC_TRY(Z.len.ensure(ctx, (size_t)n + 1));
  Z.rows = n;
  // flattened (source, entry) pairs: off = exclusive scan of the sources' row lengths in this rank's columns of K
  SC_TRY(ps_cnt.ensure(ctx, (size_t)m + 1));
  SC_TRY(ps_off.ensure(ctx, (size_t)m + 1));
  SC_LAUNCH(ctx, k_src_len, (m + 256) / 256, 256, 0, src, Kv, ps_cnt.p);
  SC_TRY(sc_exclusive_scan_i32_to_i64(ctx, ps_cnt.p, ps_off.p, (size_t)m + 1, tmp));
  // Small problems are launch- and sync-bound (a 10k-cell fit is 500 Frank-Wolfe steps of ~16 tiny kernels): when the
  // bound m x (longest row of K) is small, the per-step product is sized by that bound and runs without reading the
  // exact count back - unused sort slots hold a sentinel key that sorts behind every cell.
  const long long cap = (long long)m * std::max(k_maxrow, 1);
  const bool sync_free = !has_dups && cap <= SYNC_FREE_CAP;
  long long total = cap;
  if (!sync_free) {
    SC_CUDA(ctx, cudaMemcpyAsync(&total, ps_off.p + m, sizeof(long long), cudaMemcpyDeviceToHost, ctx->stream));
    SC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  }
  last_push_total = total;
  last_push_sync_free = sync_free;
  // rows outside this rank's slab receive nothing: their (empty) lists must still be well formed
  SC_CUDA(ctx, cudaMemsetAsync(Z.ptr.p, 0, sizeof(long long) * ((size_t)n + 1), ctx->stream));
  SC_CUDA(ctx, cudaMemsetAsync(Z.len.p, 0, sizeof(int) * ((size_t)n + 1), ctx->stream));
  SC_TRY(Z.key.ensure(ctx, (size_t)total + 1));
  SC_TRY(Z.val.ensure(ctx, (size_t)total + 1));
  if (total == 0) return SC_OK;
  const unsigned fgrid = 148 * 16;  // grid-stride over the flattened entries
  const int nrows = row_end - row_begin;
  if (!has_dups) {
    if (total >= (1ll << 32)) SC_FAIL(ctx, SC_ERR_CAPACITY, "push product: more than 2^32 entries in one step");
    SC_TRY(pk32_in.ensure(ctx, (size_t)total + 1));
    SC_TRY(pk32_out.ensure(ctx, (size_t)total + 1));
    SC_TRY(pi32_in.ensure(ctx, (size_t)total + 1));
    SC_TRY(pi32_out.ensure(ctx, (size_t)total + 1));
    SC_TRY(flat_a.ensure(ctx, (size_t)total + 1));  // flat (archetype, value) of the emitted entries
    SC_TRY(pv_in.ensure(ctx, (size_t)total + 1));
    if (sync_free) SC_CUDA(ctx, cudaMemsetAsync(pk32_in.p, 0xff, sizeof(unsigned) * (size_t)total, ctx->stream));
    SC_LAUNCH(ctx, k_push_keys, fgrid, 256, 0, src, Kv, ps_off.p, pk32_in.p, pi32_in.p, flat_a.p, pv_in.p);
    SC_TRY(sc_sort_pairs_u32_u32(ctx, pk32_in.p, pk32_out.p, pi32_in.p, pi32_out.p, (size_t)total, sync_free ? 32 : nbits,
                                 tmp));
    SC_LAUNCH(ctx, k_push_gather, fgrid, 256, 0, total, pk32_out.p, pi32_out.p, flat_a.p, pv_in.p, Z.key.p, Z.val.p);
    SC_LAUNCH(ctx, k_rows_from_keys32, (nrows + 256) / 256, 256, 0, row_begin, row_end, total, pk32_out.p, Z.ptr.p,
              Z.len.p);
    return SC_OK;
  }
  SC_TRY(pk_keys_in.ensure(ctx, (size_t)total + 1));
  SC_TRY(pk_keys_out.ensure(ctx, (size_t)total + 1));
  SC_TRY(pv_in.ensure(ctx, (size_t)total + 1));
  SC_TRY(pv_out.ensure(ctx, (size_t)total + 1));
  SC_TRY(run_heads.ensure(ctx, (size_t)total + 1));
  SC_TRY(run_pos.ensure(ctx, (size_t)total + 1));
  SC_TRY(ckeys.ensure(ctx, (size_t)total + 1));
  SC_LAUNCH(ctx, k_push_emit, fgrid, 256, 0, src, Kv, ps_off.p, kbits, pk_keys_in.p, pv_in.p);
  const unsigned egrid = (unsigned)((total + 256) / 256);
  SC_TRY(sc_sort_pairs_u64_f64(ctx, pk_keys_in.p, pk_keys_out.p, pv_in.p, pv_out.p, (size_t)total, kbits + nbits, tmp, 0));
  SC_LAUNCH(ctx, k_run_heads, egrid, 256, 0, total, pk_keys_out.p, run_heads.p);
  SC_TRY(sc_exclusive_scan_i64(ctx, run_heads.p, run_pos.p, (size_t)total + 1, tmp));
  SC_LAUNCH(ctx, k_run_sum, egrid, 256, 0, total, pk_keys_out.p, pv_out.p, run_pos.p, kbits, ckeys.p, Z.key.p, Z.val.p);
  SC_LAUNCH(ctx, k_rows_from_ckeys, (nrows + 256) / 256, 256, 0, row_begin, row_end, run_pos.p + total, ckeys.p, kbits,
            Z.ptr.p, Z.len.p);
  return SC_OK;
}

// Rows [row_begin, row_end) of L (those whose bit is set in `keep` when it is given) -> job-wide row lists G: every
// rank packs its rows into its segment of a shared layout (segments padded to the longest one) and the segments are
// all-gathered, so G holds the selected rows of every rank.  One GPU: G is not used.
int sc_fit::allgather_rows(const RowLists& L, const unsigned* keep, RowListsBuf& G) {
  const int w = ctx->world();
  SC_TRY(t_off.ensure(ctx, (size_t)n + 1));
  SC_LAUNCH(ctx, k_masked_len, (slab + 256) / 256, 256, 0, row_begin, row_end, slab, L.len, keep, cnt_n.p);
  SC_TRY(sc_exclusive_scan_i32_to_i64(ctx, cnt_n.p, t_off.p, (size_t)slab + 1, tmp));
  SC_TRY(cand_cnt.ensure(ctx, (size_t)w * k + 64));
  SC_CUDA(ctx, cudaMemcpyAsync(cand_cnt.p + ctx->rank(), t_off.p + slab, sizeof(long long), cudaMemcpyDeviceToDevice,
                               ctx->stream));
  SC_TRY(exchange(cand_cnt.p, sizeof(long long)));
  std::vector<long long> totals((size_t)w);
  SC_CUDA(ctx, cudaMemcpyAsync(totals.data(), cand_cnt.p, sizeof(long long) * (size_t)w, cudaMemcpyDeviceToHost, ctx->stream));
  SC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  long long seg = 1;
  for (long long tt : totals) seg = std::max(seg, tt);
  SC_TRY(G.ptr.ensure(ctx, (size_t)slab * w + 1));
  SC_TRY(G.len.ensure(ctx, (size_t)slab * w + 1));
  SC_TRY(G.key.ensure(ctx, (size_t)seg * w));
  SC_TRY(G.val.ensure(ctx, (size_t)seg * w));
  G.rows = n;
  SC_LAUNCH(ctx, k_compact_segment, (unsigned)(((size_t)slab * 32 + 255) / 256), 256, 0, row_begin, row_end, slab, L, keep,
            t_off.p, (long long)ctx->rank() * seg, G.ptr.p, G.len.p, G.key.p, G.val.p);
  SC_TRY(exchange(G.ptr.p, sizeof(long long) * (size_t)slab));
  SC_TRY(exchange(G.len.p, sizeof(int) * (size_t)slab));
  SC_TRY(exchange(G.key.p, sizeof(int) * (size_t)seg));
  SC_TRY(exchange(G.val.p, sizeof(double) * (size_t)seg));
  return SC_OK;
}

// K B rows of this rank's cells for the current B (cpu.py:441 `K @ B`), sorted by archetype.  with_support: the rows of the
// cells of supp(B), which t1 = (K B)^T B needs wherever they live (cpu.py:442), are all-gathered into KBs.
int sc_fit::compute_KB(bool with_support) {
  pstart();
  const int ks = k * S;
  SC_TRY(b_off.ensure(ctx, (size_t)k + 1));
  SC_TRY(src_cell.ensure(ctx, (size_t)ks + 1));
  SC_TRY(src_key.ensure(ctx, (size_t)ks + 1));
  SC_TRY(src_w.ensure(ctx, (size_t)ks + 1));
  SC_CUDA(ctx, cudaMemcpyAsync(cnt_k.p, b_cnt.p, sizeof(int) * (size_t)k, cudaMemcpyDeviceToDevice, ctx->stream));
  SC_CUDA(ctx, cudaMemsetAsync(cnt_k.p + k, 0, sizeof(int), ctx->stream));
  SC_TRY(sc_exclusive_scan_i32_to_i64(ctx, cnt_k.p, b_off.p, (size_t)k + 1, tmp));
  long long bnnz = 0;
  SC_CUDA(ctx, cudaMemcpyAsync(&bnnz, b_off.p + k, sizeof(long long), cudaMemcpyDeviceToHost, ctx->stream));
  SC_LAUNCH(ctx, k_b_sources, (ks + 255) / 256, 256, 0, k, S, b_idx.p, b_val.p, b_cnt.p, b_off.p, src_cell.p, src_key.p,
            src_w.p);
  SC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  pstop(P_BROW);
  pstart();
  PushSrc src{(int)bnnz, src_cell.p, src_key.p, src_w.p};
  SC_TRY(push_product(src, true, KB));
  if (with_support && ctx->world() > 1) {
    const size_t words = ((size_t)n + 31) / 32;
    SC_TRY(supp_bits.ensure(ctx, words));
    SC_CUDA(ctx, cudaMemsetAsync(supp_bits.p, 0, sizeof(unsigned) * words, ctx->stream));
    SC_LAUNCH(ctx, k_b_support_bits, (ks + 255) / 256, 256, 0, k, S, b_idx.p, b_cnt.p, supp_bits.p);
    SC_TRY(allgather_rows(KB.view(), supp_bits.p, KBs));
  }
  pstop(P_SPGEMM_KB);
  return SC_OK;
}

static void swap_lists(RowListsBuf& a, RowListsBuf& b) {
  std::swap(a.ptr.p, b.ptr.p);
  std::swap(a.ptr.cap, b.ptr.cap);
  std::swap(a.len.p, b.len.p);
  std::swap(a.len.cap, b.len.cap);
  std::swap(a.key.p, b.key.p);
  std::swap(a.key.cap, b.key.cap);
  std::swap(a.val.p, b.val.p);
  std::swap(a.val.cap, b.val.cap);
  std::swap(a.rows, b.rows);
}

// K B_t from K B_{t-1} and the cells chosen at step t-1 (amin_cur, written by k_updateB_step):
//   K B_t = (1 - gamma_{t-1}) K B_{t-1} + gamma_{t-1} K E_{t-1},   gamma_0 = 1 (B was reset to E_0)
int sc_fit::advance_KB_delta() {
  pstart();
  PushSrc src{k, amin_cur.p, nullptr, nullptr};
  SC_TRY(push_product(src, false, DKB));
  pstop(P_SPGEMM_Y);
  return SC_OK;
}
int sc_fit::advance_KB_merge(int t) {
  pstart();
  if (t == 1) {
    swap_lists(KB, DKB);  // gamma_0 = 1
  } else {
    const double gamma = 2.0 / ((double)(t - 1) + 2.0);  // fw_gamma(t - 1)
    SC_TRY(sp.ub.ensure(ctx, (size_t)n + 1));
    SC_TRY(KB2.ptr.ensure(ctx, (size_t)n + 1));
    SC_TRY(KB2.len.ensure(ctx, (size_t)n));
    SC_LAUNCH(ctx, k_merge_len, (n + 256) / 256, 256, 0, n, KB.len.p, DKB.len.p, sp.ub.p);
    SC_TRY(sc_exclusive_scan_i64(ctx, sp.ub.p, KB2.ptr.p, (size_t)n + 1, sp.tmp));
    SC_CUDA(ctx, cudaMemsetAsync(KB2.len.p, 0, sizeof(int) * (size_t)n, ctx->stream));
    long long total = 0;
    if (last_push_sync_free) {
      total = (long long)T * last_push_total;  // nnz(K B_t) <= t x (bound of one step's delta): sized once, no read-back
    } else {
      SC_CUDA(ctx, cudaMemcpyAsync(&total, KB2.ptr.p + n, sizeof(long long), cudaMemcpyDeviceToHost, ctx->stream));
      SC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    }
    SC_TRY(KB2.key.ensure(ctx, (size_t)total + 1));
    SC_TRY(KB2.val.ensure(ctx, (size_t)total + 1));
    KB2.rows = n;
    const int rows = row_end - row_begin;
    if (rows > 0) {
      SC_CUDA(ctx, cudaMemsetAsync(slow_cnt.p, 0, sizeof(int), ctx->stream));
      SC_LAUNCH(ctx, k_kb_merge_staged, (rows + KBM_ROWS - 1) / KBM_ROWS, 256, 0, row_begin, row_end, KB.view(), DKB.view(),
                KB2.ptr.p, KB2.len.p, KB2.key.p, KB2.val.p, 1.0 - gamma, gamma, cmax_v.p, cmax_l.p, slow_list.p,
                slow_cnt.p);
      // rows whose CTA overflowed the staging area (cells next to very many chosen cells): generic kernel, a few warps
      SC_LAUNCH(ctx, k_kb_merge, 148, 256, 0, KB.view(), DKB.view(), KB2.ptr.p, KB2.len.p, KB2.key.p, KB2.val.p,
                1.0 - gamma, gamma, cmax_v.p, cmax_l.p, slow_list.p, slow_cnt.p);
    }
    swap_lists(KB, KB2);
  }
  pstop(P_SPGEMM_KB);
  return SC_OK;
}

// perm = this rank's cells sorted by (archetype of the largest entry of the current K B row, cell)
int sc_fit::build_perm() {
  const int nloc = row_end - row_begin;
  SC_TRY(cmax_v.ensure(ctx, (size_t)n));
  SC_TRY(cmax_l.ensure(ctx, (size_t)n));
  SC_TRY(perm.ensure(ctx, (size_t)nloc + 1));
  SC_TRY(pk_in.ensure(ctx, (size_t)nloc + 1));
  SC_TRY(pk_out.ensure(ctx, (size_t)nloc + 1));
  if (nloc == 0) return SC_OK;
  SC_LAUNCH(ctx, k_row_max, (unsigned)(((size_t)nloc * 32 + 255) / 256), 256, 0, row_begin, row_end, KB.view(), cmax_v.p,
            cmax_l.p);
  SC_LAUNCH(ctx, k_perm_keys, (nloc + 255) / 256, 256, 0, row_begin, row_end, cmax_l.p, pk_in.p);
  int kb = 1;
  while ((1ll << kb) < k) kb++;
  SC_TRY(sc_sort_keys_u64(ctx, pk_in.p, pk_out.p, (size_t)nloc, 32 + kb, tmp));
  SC_LAUNCH(ctx, k_perm_unpack, (nloc + 255) / 256, 256, 0, nloc, pk_out.p, perm.p);
  return SC_OK;
}

int sc_fit::precompute_A() {
  if (pre_valid) return SC_OK;
  if (!have_M || !have_B) SC_FAIL(ctx, SC_ERR_STATE, "fit: kernel matrix and B must be set first");
  SC_TRY(compute_KB(true));  // this rank's cells (candidates of _updateA, RSS); the rows of supp(B) are gathered for t1
  pstart();
  SC_CUDA(ctx, cudaMemsetAsync(t1A.p, 0, sizeof(double) * (size_t)k * k, ctx->stream));
  SC_LAUNCH(ctx, k_t1A, (k * 32 + 127) / 128, 128, 0, k, S, b_idx.p, b_val.p, b_cnt.p,
            ctx->world() > 1 ? KBs.view() : KB.view(), t1A.p);
  pstop(P_T1A);
  pre_valid = true;
  return SC_OK;
}

int sc_fit::update_A(double l2, int* trace_dev) {
  if (!have_A) SC_FAIL(ctx, SC_ERR_STATE, "fit: A has not been initialised");
  SC_TRY(precompute_A());
  const int nxt = (a_cur < 0) ? 0 : 1 - a_cur;
  UpdAArgs g;
  g.cell_begin = row_begin;
  g.cell_end = row_end;
  g.n = n;
  g.k = k;
  g.S = S;
  g.T = T;
  g.l2 = l2;
  g.KB = KB.view();
  g.t1T = t1A.p;
  // (walking the cells in the locality order of build_perm() was measured slower here: hot t1 rows collide in L2)
  g.Aprev = A_view();
  g.a_idx = a_idx[nxt].p;
  g.a_val = a_val[nxt].p;
  g.a_cnt = a_cnt[nxt].p;
  g.trace = trace_dev;
  g.slow_list = slow_list.p;
  g.slow_cnt = slow_cnt.p;
  SC_TRY(updA_diag.ensure(ctx, 2));
  SC_CUDA(ctx, cudaMemsetAsync(updA_diag.p, 0, sizeof(unsigned long long) * 2, ctx->stream));
  g.diag = updA_diag.p;
  SC_CUDA(ctx, cudaMemsetAsync(slow_cnt.p, 0, sizeof(int), ctx->stream));
  int nslow = 0;
  pstart();
  if (l2 == 0.0) {
    int grid = std::min((row_end - row_begin + FWA_WARPS - 1) / FWA_WARPS, 148 * 16);
    if (resum_A) SC_LAUNCH(ctx, k_updateA_fast<true>, grid, FWA_WARPS * 32, FWA_PER_WARP * FWA_WARPS, g);
    else SC_LAUNCH(ctx, k_updateA_fast<false>, grid, FWA_WARPS * 32, FWA_PER_WARP * FWA_WARPS, g);
    SC_CUDA(ctx, cudaMemcpyAsync(&nslow, slow_cnt.p, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    SC_CUDA(ctx, cudaMemcpyAsync(last_updA_diag, updA_diag.p, sizeof(unsigned long long) * 2, cudaMemcpyDeviceToHost,
                                 ctx->stream));
    SC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  } else {
    // the -l2*A term can make entries outside the support of t2 negative: every cell takes the generic path
    const int nloc = row_end - row_begin;
    std::vector<int> all(nloc);
    for (int j = 0; j < nloc; ++j) all[j] = row_begin + j;
    SC_CUDA(ctx, sc_copy(slow_list.p, all.data(), sizeof(int) * (size_t)nloc, ctx->stream));
    SC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    nslow = nloc;
  }
  last_slow_cells = nslow;
  pstop(P_UPDA_FAST);
  pstart();
  if (nslow > 0) {
    int grid = std::min((nslow + 3) / 4, 148 * 16);
    SC_LAUNCH(ctx, k_updateA_slow, grid, 128, 0, g, nslow);
  }
  pstop(P_UPDA_SLOW);
  a_cur = nxt;
  a_is_slots = true;
  // every rank needs all of A for A A^T and K A^T: all-gather the slot arrays (equal slabs, in place)
  SC_TRY(exchange(a_idx[nxt].p, sizeof(int) * (size_t)slab * S));
  SC_TRY(exchange(a_val[nxt].p, sizeof(double) * (size_t)slab * S));
  SC_TRY(exchange(a_cnt[nxt].p, sizeof(int) * (size_t)slab));
  return SC_OK;
}

// in-place all-gather of equal slabs over the ranks of the job (no-op on one GPU)
int sc_fit::exchange(void* buf, size_t bytes_per_rank) {
  if (!ctx->comm || ctx->comm->world == 1) return SC_OK;
  const int rc = ctx->comm->allgather(buf, bytes_per_rank, ctx->stream);
  if (rc != SC_OK) ctx->err = ctx->comm->err;
  return rc;
}

int sc_fit::update_B_begin() {
  if (!have_A || !a_is_slots) SC_FAIL(ctx, SC_ERR_STATE, "fit: run update_A before update_B");
  if (!have_B) SC_FAIL(ctx, SC_ERR_STATE, "fit: B has not been initialised");
  CsrView M{n, m_ptr.p, m_col.p, m_val.p};
  const int cur = a_cur;
  // ---- t1B = A A^T: transpose A by a stable sort on (archetype, cell), then accumulate rows in cell order.
  // Rows are sharded over the ranks by archetype slab [a0, a1); the slabs are all-gathered (rows of t1B, per-tile hub x hub
  // partial sums), the mirror and the tile reduction then run on every rank - each entry is produced whole by one rank.
  pstart();
  const int w = ctx->world();
  const int ka = (k + w - 1) / w;  // archetypes per rank
  const int a0 = std::min(k, ctx->rank() * ka), a1 = std::min(k, a0 + ka);
  SC_TRY(a_off.ensure(ctx, (size_t)n + 1));
  SC_TRY(arow_ptr.ensure(ctx, (size_t)k + 1));
  // cells per archetype (hub selection): counted on this rank's cells, the counts are all-gathered and summed on the host
  SC_TRY(arch_cnt_all.ensure(ctx, (size_t)w * k));
  {
    int* mine = arch_cnt_all.p + (size_t)ctx->rank() * k;
    const int nloc = row_end - row_begin;
    SC_CUDA(ctx, cudaMemsetAsync(mine, 0, sizeof(int) * (size_t)k, ctx->stream));
    if (nloc > 0)
      SC_LAUNCH(ctx, k_arch_hist, (unsigned)(((size_t)nloc * 32 + 255) / 256), 256, 0, row_begin, row_end, S, a_idx[cur].p,
                a_cnt[cur].p, mine);
    SC_TRY(exchange(arch_cnt_all.p, sizeof(int) * (size_t)k));
  }
  std::vector<int> cnt_all((size_t)w * k), cells_per_arch((size_t)k, 0);
  SC_CUDA(ctx, cudaMemcpyAsync(cnt_all.data(), arch_cnt_all.p, sizeof(int) * cnt_all.size(), cudaMemcpyDeviceToHost, ctx->stream));
  SC_LAUNCH(ctx, k_count_a_range, (unsigned)((((size_t)n + 1) * 32 + 255) / 256), 256, 0, n, S, a0, a1, a_idx[cur].p,
            a_cnt[cur].p, cnt_n.p);
  SC_TRY(sc_exclusive_scan_i32_to_i64(ctx, cnt_n.p, a_off.p, (size_t)n + 1, tmp));
  long long annz = 0;
  SC_CUDA(ctx, cudaMemcpyAsync(&annz, a_off.p + n, sizeof(long long), cudaMemcpyDeviceToHost, ctx->stream));
  SC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  for (int r = 0; r < w; ++r)
    for (int a = 0; a < k; ++a) cells_per_arch[a] += cnt_all[(size_t)r * k + a];
  SC_TRY(gk_in.ensure(ctx, (size_t)annz + 1));
  SC_TRY(gk_out.ensure(ctx, (size_t)annz + 1));
  SC_TRY(gv_in.ensure(ctx, (size_t)annz + 1));
  SC_TRY(gv_out.ensure(ctx, (size_t)annz + 1));
  SC_LAUNCH(ctx, k_emit_a_pairs_range, (unsigned)(((size_t)n * 32 + 255) / 256), 256, 0, n, S, a0, a1, a_idx[cur].p,
            a_val[cur].p, a_cnt[cur].p, a_off.p, gk_in.p, gv_in.p);
  int kbits = 1;
  while ((1ll << kbits) < k) kbits++;
  SC_TRY(sc_sort_pairs_u64_f64(ctx, gk_in.p, gk_out.p, gv_in.p, gv_out.p, (size_t)annz, 32 + kbits, tmp));
  SC_LAUNCH(ctx, k_row_bounds, (k + 256) / 256, 256, 0, k, gk_out.p, annz, arow_ptr.p);
  if (a1 > a0) SC_CUDA(ctx, cudaMemsetAsync(t1B.p + (size_t)a0 * k, 0, sizeof(double) * (size_t)(a1 - a0) * k, ctx->stream));
  {
    // hub rows: chosen on the host from the number of cells per archetype (the same on every rank)
    const long long hub_thr = 8192;
    std::vector<std::pair<long long, int>> cand_h;
    for (int a = 0; a < k; ++a)
      if (cells_per_arch[a] > hub_thr) cand_h.push_back({-(long long)cells_per_arch[a], a});
    std::sort(cand_h.begin(), cand_h.end());
    if ((int)cand_h.size() > GRAM_HUB_MAX) cand_h.resize(GRAM_HUB_MAX);  // the rest stay ordinary (sequential) rows
    std::vector<int> hubs;
    for (auto& c : cand_h) hubs.push_back(c.second);
    std::sort(hubs.begin(), hubs.end());
    std::vector<unsigned char> ishub(k, 0);
    std::vector<int> hubof(k, -1);
    for (size_t q = 0; q < hubs.size(); ++q) {
      ishub[hubs[q]] = 1;
      hubof[hubs[q]] = (int)q;
    }
    SC_TRY(hub_flag.ensure(ctx, (size_t)k));
    SC_CUDA(ctx, cudaMemcpyAsync(hub_flag.p, ishub.data(), (size_t)k, cudaMemcpyHostToDevice, ctx->stream));
    SC_LAUNCH(ctx, k_gram, (k * 32 + 127) / 128, 128, 0, k, S, a0, a1, hub_flag.p, arow_ptr.p, gk_out.p, gv_out.p,
              a_idx[cur].p, a_val[cur].p, a_cnt[cur].p, t1B.p);
    const int H = (int)hubs.size();
    const int ntiles = (n + GRAM_HUB_CELLS - 1) / GRAM_HUB_CELLS;
    const int tper = (ntiles + w - 1) / w;  // tiles per rank
    if (H > 0) {
      const int tl0 = std::min(ntiles, ctx->rank() * tper), tl1 = std::min(ntiles, tl0 + tper);
      SC_TRY(hub_list.ensure(ctx, (size_t)H));
      SC_TRY(gram_hubof.ensure(ctx, (size_t)k));
      SC_TRY(gram_tiles.ensure(ctx, (size_t)tper * w * H * H));
      SC_CUDA(ctx, cudaMemcpyAsync(hub_list.p, hubs.data(), sizeof(int) * H, cudaMemcpyHostToDevice, ctx->stream));
      SC_CUDA(ctx, cudaMemcpyAsync(gram_hubof.p, hubof.data(), sizeof(int) * (size_t)k, cudaMemcpyHostToDevice, ctx->stream));
      const size_t smem = (size_t)H * H * sizeof(double) + 2 * FW_SMAX * (sizeof(double) + sizeof(int));
      SC_CUDA(ctx, cudaFuncSetAttribute(k_gram_hub_tiles, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      if (tl1 > tl0)
        SC_LAUNCH(ctx, k_gram_hub_tiles, tl1 - tl0, 256, smem, n, S, H, tl0, gram_hubof.p, a_idx[cur].p, a_val[cur].p,
                  a_cnt[cur].p, gram_tiles.p);
      SC_TRY(exchange(gram_tiles.p, sizeof(double) * (size_t)tper * H * H));
    }
    SC_TRY(exchange(t1B.p, sizeof(double) * (size_t)ka * k));
    if (H > 0) {
      SC_LAUNCH(ctx, k_gram_mirror, dim3((k + 255) / 256, H), 256, 0, k, H, hub_list.p, hub_flag.p, t1B.p);
      SC_LAUNCH(ctx, k_gram_hub_reduce, (H * H + 255) / 256, 256, 0, k, H, ntiles, hub_list.p, gram_tiles.p, t1B.p);
    }
    SC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));  // the host vectors go out of scope
  }
  pstop(P_GRAM);
  // ---- hub block of this call: the archetypes that carry weight from more than 1/64 of all cells (at most hub_cap of
  // them, the most populated first).  Their gradient is evaluated cell-major by k_updateB_hub and their share of t2 is
  // dense; the choice depends on A alone, which every rank holds in full.
  {
    std::vector<std::pair<long long, int>> big;
    const long long thr = std::max<long long>(32, n / 64);
    for (int a = 0; a < k; ++a)
      if (cells_per_arch[a] > thr) big.push_back({-(long long)cells_per_arch[a], a});
    std::sort(big.begin(), big.end());
    if ((int)big.size() > hub_cap) big.resize(hub_cap);
    std::vector<int> ids, slot(k, -1);
    for (auto& b : big) ids.push_back(b.second);
    std::sort(ids.begin(), ids.end());
    for (size_t q = 0; q < ids.size(); ++q) slot[ids[q]] = (int)q;
    n_hub = (int)ids.size();
    n_hub_pad = ((n_hub + 31) / 32) * 32;
    SC_TRY(hub_ids.ensure(ctx, (size_t)HUB_MAX));
    SC_TRY(hub_slot.ensure(ctx, (size_t)k));
    if (n_hub) SC_CUDA(ctx, cudaMemcpyAsync(hub_ids.p, ids.data(), sizeof(int) * n_hub, cudaMemcpyHostToDevice, ctx->stream));
    SC_CUDA(ctx, cudaMemcpyAsync(hub_slot.p, slot.data(), sizeof(int) * (size_t)k, cudaMemcpyHostToDevice, ctx->stream));
    SC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    last_hub_block = n_hub;
  }
  // ---- t2 = K A^T as per-cell rows, then as per-archetype candidate lists
  pstart();
  RowLists Al{slot_ptr_n.p, a_cnt[cur].p, a_idx[cur].p, a_val[cur].p};
  if (n_hub > 0) {
    const int nloc = row_end - row_begin;
    SC_TRY(ah_len.ensure(ctx, (size_t)n + 1));
    SC_TRY(ah_key.ensure(ctx, (size_t)n * S));
    SC_TRY(ah_val.ensure(ctx, (size_t)n * S));
    SC_TRY(AH.ensure(ctx, (size_t)n * n_hub_pad));
    SC_TRY(YH.ensure(ctx, (size_t)n * n_hub_pad));
    SC_TRY(T2H.ensure(ctx, (size_t)std::max(nloc, 1) * n_hub_pad));
    SC_CUDA(ctx, cudaMemsetAsync(AH.p, 0, sizeof(double) * (size_t)n * n_hub_pad, ctx->stream));
    SC_LAUNCH(ctx, k_split_slots, (unsigned)(((size_t)n * 32 + 255) / 256), 256, 0, n, S, n_hub_pad, hub_slot.p, a_idx[cur].p,
              a_val[cur].p, a_cnt[cur].p, ah_key.p, ah_val.p, ah_len.p, AH.p);
    Al = RowLists{slot_ptr_n.p, ah_len.p, ah_key.p, ah_val.p};  // A without its hub rows
#define SC_HUBCOLS(RR, A0, A1, OUT0, XX, YY)                                                                            \
  SC_LAUNCH(ctx, k_spmm_hubcols<RR>, (unsigned)(((size_t)((A1) - (A0)) * 32 + 255) / 256), 256, 0, A0, A1, OUT0, n_hub_pad, \
            m_ptr.p, m_col.p, m_val.p, XX, YY)
    switch (n_hub_pad / 32) {
      case 1: SC_HUBCOLS(1, 0, n, 0, AH.p, YH.p); if (nloc > 0) SC_HUBCOLS(1, row_begin, row_end, row_begin, YH.p, T2H.p); break;
      case 2: SC_HUBCOLS(2, 0, n, 0, AH.p, YH.p); if (nloc > 0) SC_HUBCOLS(2, row_begin, row_end, row_begin, YH.p, T2H.p); break;
      case 3: SC_HUBCOLS(3, 0, n, 0, AH.p, YH.p); if (nloc > 0) SC_HUBCOLS(3, row_begin, row_end, row_begin, YH.p, T2H.p); break;
      default: SC_HUBCOLS(4, 0, n, 0, AH.p, YH.p); if (nloc > 0) SC_HUBCOLS(4, row_begin, row_end, row_begin, YH.p, T2H.p); break;
    }
#undef SC_HUBCOLS
  }
  // rows in first-occurrence order (no compaction, no sort): the values do not depend on the order inside a row of the
  // right-hand side (a key occurs once per row), and the candidate lists are ordered by the radix sort below
  if (ctx->world() == 1) {
    SC_TRY(sc_spgemm_rows(ctx, M, 0, n, Al, k, MA, sp, false));
    SC_TRY(sc_spgemm_rows(ctx, M, row_begin, row_end, MA.view(), k, T2c, sp, false));
  } else {
    // M A^T for this rank's rows only, all-gathered (allgather_rows), so the second product sees every row of M A^T
    SC_TRY(sc_spgemm_rows(ctx, M, row_begin, row_end, Al, k, MA, sp, false));
    SC_TRY(allgather_rows(MA.view(), nullptr, MAg));
    SC_TRY(sc_spgemm_rows(ctx, M, row_begin, row_end, MAg.view(), k, T2c, sp, false));  // candidates: this rank's cells
  }
  pstop(P_T2ROWS);
  pstart();
  {
    SC_TRY(t_off.ensure(ctx, (size_t)n + 1));
    SC_CUDA(ctx, cudaMemcpyAsync(cnt_n.p, T2c.len.p, sizeof(int) * (size_t)n, cudaMemcpyDeviceToDevice, ctx->stream));
    SC_CUDA(ctx, cudaMemsetAsync(cnt_n.p + n, 0, sizeof(int), ctx->stream));
    SC_TRY(sc_exclusive_scan_i32_to_i64(ctx, cnt_n.p, t_off.p, (size_t)n + 1, tmp));
    long long cnnz = 0;
    SC_CUDA(ctx, cudaMemcpyAsync(&cnnz, t_off.p + n, sizeof(long long), cudaMemcpyDeviceToHost, ctx->stream));
    SC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (cnnz >= (1ll << 32)) SC_FAIL(ctx, SC_ERR_CAPACITY, "update_B: more than 2^32 candidate pairs");
    last_t2_nnz = cnnz + (long long)n_hub * (row_end - row_begin);  // sparse candidates + the dense hub columns
    SC_TRY(ck_in.ensure(ctx, (size_t)cnnz + 1));
    SC_TRY(ck_out.ensure(ctx, (size_t)cnnz + 1));
    SC_TRY(co_in.ensure(ctx, (size_t)cnnz + 1));
    SC_TRY(co_out.ensure(ctx, (size_t)cnnz + 1));
    SC_TRY(c_cellv.ensure(ctx, (size_t)cnnz + 1));
    SC_TRY(c_t2v.ensure(ctx, (size_t)cnnz + 1));
    SC_TRY(c_cell.ensure(ctx, (size_t)cnnz + 1));
    SC_TRY(c_t2.ensure(ctx, (size_t)cnnz + 1));
    SC_TRY(c_ptr.ensure(ctx, (size_t)k + 1));
    SC_LAUNCH(ctx, k_emit_cand, (unsigned)(((size_t)n * 32 + 255) / 256), 256, 0, n, T2c.view(), t_off.p, ck_in.p,
              co_in.p, c_cellv.p, c_t2v.p);
    SC_TRY(sc_sort_pairs_u64_u32(ctx, ck_in.p, ck_out.p, co_in.p, co_out.p, (size_t)cnnz, 32 + kbits, tmp));
    SC_LAUNCH(ctx, k_row_bounds, (k + 256) / 256, 256, 0, k, ck_out.p, cnnz, c_ptr.p);
    if (cnnz > 0)  // (every archetype can be a hub archetype on tiny problems: no sparse candidates at all)
      SC_LAUNCH(ctx, k_gather_cand, (unsigned)((cnnz + 255) / 256), 256, 0, cnnz, co_out.p, c_cellv.p, c_t2v.p, c_cell.p,
                c_t2.p);
  }
  if (n_hub > 0) {  // hub block (chosen above): slices of t1 by hub column, running product, per-CTA partial results
    const int nloc = row_end - row_begin;
    const int nctas = (nloc + HUB_CELLS_PER_CTA - 1) / HUB_CELLS_PER_CTA;
    SC_TRY(T1H.ensure(ctx, (size_t)k * n_hub_pad));
    SC_TRY(HH.ensure(ctx, (size_t)std::max(nloc, 1) * n_hub_pad));
    SC_TRY(hub_part.ensure(ctx, (size_t)std::max(nctas, 1) * n_hub_pad * 4));
    SC_LAUNCH(ctx, k_hub_t1, dim3((k + 255) / 256, n_hub_pad), 256, 0, k, n_hub, n_hub_pad, hub_ids.p, t1B.p, T1H.p);
  }
  // chunk table: archetypes with long candidate lists are spread over many CTAs
  SC_TRY(chunk_ptr.ensure(ctx, (size_t)k + 1));
  SC_LAUNCH(ctx, k_chunk_count, (k + 256) / 256, 256, 0, k, c_ptr.p, hub_slot.p, cnt_k.p);
  SC_TRY(sc_exclusive_scan_i32_to_i64(ctx, cnt_k.p, chunk_ptr.p, (size_t)k + 1, tmp));
  long long nchunks = 0;
  SC_CUDA(ctx, cudaMemcpyAsync(&nchunks, chunk_ptr.p + k, sizeof(long long), cudaMemcpyDeviceToHost, ctx->stream));
  SC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  SC_TRY(chunk_arch.ensure(ctx, (size_t)nchunks + 1));
  SC_TRY(cmax_v.ensure(ctx, (size_t)n));
  SC_TRY(cmax_l.ensure(ctx, (size_t)n));
  SC_TRY(bound_v.ensure(ctx, (size_t)2 * k));
  SC_TRY(eval_cnt.ensure(ctx, 4));
  SC_CUDA(ctx, cudaMemsetAsync(eval_cnt.p, 0, sizeof(unsigned long long) * 4, ctx->stream));
  SC_TRY(part_v.ensure(ctx, (size_t)nchunks + 1));
  SC_TRY(part_i.ensure(ctx, (size_t)nchunks + 1));
  SC_LAUNCH(ctx, k_chunk_fill, k, 128, 0, k, chunk_ptr.p, chunk_arch.p);
  pstop(P_T2TRANS);
  nchunks_cur = nchunks;
  SC_TRY(part_out.ensure(ctx, (size_t)k * 4 * ctx->world()));
  in_update_B = true;
  return SC_OK;
}

int sc_fit::update_B_eval(int t, double* partial_dev) {
  if (!in_update_B) SC_FAIL(ctx, SC_ERR_STATE, "update_B_eval outside update_B_begin/end");
  if (t < 0 || t >= T) SC_FAIL(ctx, SC_ERR_ARG, "update_B_eval: bad step");
  bool hub_async = false;
  // K B for the current B (cpu.py:486), this rank's rows, first-occurrence order (the rows are only walked here):
  // evaluated from B at the first step, then advanced by the step's rank-k change (see k_kb_merge)
  if (t == 0) {
    if (!pre_valid) SC_TRY(compute_KB(false));  // else: the rows _updateA / the RSS just used
  } else if (fresh_kb) {
    SC_TRY(compute_KB(false));
  } else {
    SC_TRY(advance_KB_delta());
    if (overlap_hub && t >= 2 && n_hub > 0) {  // fork: the hub gradient needs the delta rows only
      SC_CUDA(ctx, cudaEventRecord(ev_fork, ctx->stream));
      SC_CUDA(ctx, cudaStreamWaitEvent(stream2, ev_fork, 0));
      SC_TRY(launch_hub(t, stream2, partial_dev));
      SC_CUDA(ctx, cudaEventRecord(ev_join, stream2));
      hub_async = true;
    }
    SC_TRY(advance_KB_merge(t));
  }
  UpdBArgs g;
  g.n = n;
  g.k = k;
  g.S = S;
  g.t = t;
  g.t1 = t1B.p;
  g.c_ptr = c_ptr.p;
  g.c_cell = c_cell.p;
  g.c_t2 = c_t2.p;
  g.KB = KB.view();
  g.T2c = T2c.view();
  g.b_idx = b_idx.p;
  g.b_val = b_val.p;
  g.b_cnt = b_cnt.p;
  g.trace = nullptr;
  g.amin = nullptr;
  g.evaluated = eval_cnt.p;
  g.row_begin = row_begin;
  g.row_end = row_end;
  g.prev = (t > 0 && r0_seed > 0) ? amin_cur.p : nullptr;
  g.r0n = (t > 0 && r0_seed > 0) ? r0_seed : std::max(8, (UPB_R0 + ctx->world() - 1) / ctx->world());
  const long long nchunks = nchunks_cur;
  const double* prune = bound_v.p;
  pstart();
  if ((int)ev_a.size() <= t) {
    cudaEvent_t ea, eb;
    SC_CUDA(ctx, cudaEventCreate(&ea));
    SC_CUDA(ctx, cudaEventCreate(&eb));
    ev_a.push_back(ea);
    ev_b.push_back(eb);
  }
  SC_CUDA(ctx, cudaEventRecord(ev_a[t], ctx->stream));
  if (t == 0 && n_hub > 0) SC_TRY(build_perm());  // the order only affects locality; reused for the remaining steps
  {
    // largest entry of every K B row (second pruning bound): produced by k_kb_merge when it ran, else computed here
    if (nchunks > 0 && row_end > row_begin && (t < 2 || fresh_kb))
      SC_LAUNCH(ctx, k_row_max, (unsigned)(((size_t)(row_end - row_begin) * 32 + 255) / 256), 256, 0, row_begin, row_end,
                KB.view(), cmax_v.p, cmax_l.p);
    // round 0 runs for every archetype on every rank (an archetype without local candidates reports +inf): the
    // all-gather of the bounds below is a collective, so it cannot depend on what this rank happens to own
    SC_LAUNCH(ctx, k_updateB_eval, k, 256, 0, g, 0, chunk_ptr.p, chunk_arch.p, cmax_v.p, cmax_l.p, part_v.p, part_i.p,
              bound_v.p, bound_v.p);  // round 0 prunes against +inf: the last argument is never read
    if (ctx->world() > 1) {
      // every rank evaluated the 64 / world largest-t2 candidates of each archetype among ITS cells; the smallest of the
      // round-0 minima is the bound all of them prune with (one more k x 8-byte all-gather per step, instead of every
      // rank evaluating 64 candidates per archetype on its own)
      const int w = ctx->world();
      SC_TRY(bound_all.ensure(ctx, (size_t)w * k + k));
      SC_CUDA(ctx, cudaMemcpyAsync(bound_all.p + (size_t)ctx->rank() * k, bound_v.p, sizeof(double) * (size_t)k,
                                   cudaMemcpyDeviceToDevice, ctx->stream));
      SC_TRY(exchange(bound_all.p, sizeof(double) * (size_t)k));
      prune = bound_all.p + (size_t)w * k;
      SC_LAUNCH(ctx, k_min_over_ranks, (k + 255) / 256, 256, 0, k, w, bound_all.p, bound_all.p + (size_t)w * k);
    }
    if (nchunks > 0)
      SC_LAUNCH(ctx, k_updateB_eval, (unsigned)nchunks, 256, 0, g, 1, chunk_ptr.p, chunk_arch.p, cmax_v.p, cmax_l.p,
                part_v.p, part_i.p, bound_v.p, prune);
  }
  if (n_hub > 0 && !hub_async) SC_TRY(launch_hub(t, ctx->stream, partial_dev));
  SC_CUDA(ctx, cudaEventRecord(ev_b[t], ctx->stream));
  SC_LAUNCH(ctx, k_updateB_local, k, 256, 0, g, row_begin, row_end, chunk_ptr.p, part_v.p, part_i.p, bound_v.p, prune,
            hub_slot.p, partial_dev);
  if (hub_async) SC_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ev_join, 0));  // the hub archetypes' partials are in
  pstop(P_GEVAL);
  return SC_OK;
}

// The cell-major gradient of the hub archetypes for step t on `stream` (see k_updateB_hub).  From step 2 on it depends on
// the step's delta rows only - not on the merged K B - so update_B_eval runs it on a second stream next to the merge and
// the two rounds of k_updateB_eval: both chains are latency bound and share the SMs well.
int sc_fit::launch_hub(int t, cudaStream_t stream, double* partial_dev) {
  const int nloc = row_end - row_begin;
  const int nctas = (nloc + HUB_CELLS_PER_CTA - 1) / HUB_CELLS_PER_CTA;
  // steps 0 and 1 evaluate the product in full (K B from the previous B, then K E_0); later steps advance it
  const bool inc = !fresh_kb && t >= 2;
  const double gamma = inc ? 2.0 / ((double)(t - 1) + 2.0) : 0.0;
  const RowLists hub_rows = inc ? DKB.view() : KB.view();
  if ((int)ev_h0.size() <= t) {
    cudaEvent_t ea, eb;
    SC_CUDA(ctx, cudaEventCreate(&ea));
    SC_CUDA(ctx, cudaEventCreate(&eb));
    ev_h0.push_back(ea);
    ev_h1.push_back(eb);
  }
  if (inc) {  // algorithmic bytes of this launch: HH read + write and t2, 8 B each per (cell, hub), + the delta rows
    hub_bytes_total += 24.0 * (double)nloc * n_hub + 12.0 * (double)last_push_total;
    hub_launches_counted++;
  }
  SC_CUDA(ctx, cudaEventRecord(ev_h0[t], stream));
#define SC_HUB_LAUNCH_U(RR, UU)                                                                                        \
  do {                                                                                                                 \
    k_updateB_hub<RR, UU><<<nctas, 256, 0, stream>>>(row_begin, row_end, n_hub_pad, perm.p, hub_rows, T1H.p, T2H.p, HH.p, \
                                                     inc ? 1 : 0, 1.0 - gamma, gamma, hub_part.p);                      \
    ctx->launches++;                                                                                                   \
    SC_CUDA(ctx, cudaGetLastError());                                                                                  \
  } while (0)
#define SC_HUB_LAUNCH(RR)          \
  do {                             \
    if (hub_u == 4 && (RR) <= 2) { \
      SC_HUB_LAUNCH_U(RR, 4);      \
    } else {                       \
      SC_HUB_LAUNCH_U(RR, 2);      \
    }                              \
  } while (0)
  if (nctas > 0) switch (n_hub_pad / 32) {
      case 1: SC_HUB_LAUNCH(1); break;
      case 2: SC_HUB_LAUNCH(2); break;
      case 3: SC_HUB_LAUNCH(3); break;
      default: SC_HUB_LAUNCH(4); break;
    }
#undef SC_HUB_LAUNCH_U
#undef SC_HUB_LAUNCH
  SC_CUDA(ctx, cudaEventRecord(ev_h1[t], stream));
  k_updateB_hub_reduce<<<n_hub, 256, 0, stream>>>(nctas, n_hub_pad, hub_ids.p, hub_part.p, partial_dev);
  ctx->launches++;
  SC_CUDA(ctx, cudaGetLastError());
  return SC_OK;
}

int sc_fit::update_B_step(int t, const double* gathered_dev, int n_ranks, int* trace_dev) {
  if (!in_update_B) SC_FAIL(ctx, SC_ERR_STATE, "update_B_step outside update_B_begin/end");
  UpdBArgs g;
  g.n = n;
  g.k = k;
  g.S = S;
  g.t = t;
  g.t1 = t1B.p;
  g.c_ptr = c_ptr.p;
  g.c_cell = c_cell.p;
  g.c_t2 = c_t2.p;
  g.KB = KB.view();
  g.T2c = T2c.view();
  g.b_idx = b_idx.p;
  g.b_val = b_val.p;
  g.b_cnt = b_cnt.p;
  g.trace = trace_dev;
  g.amin = amin_cur.p;
  g.evaluated = nullptr;
  SC_LAUNCH(ctx, k_updateB_step, k, 32, 0, g, gathered_dev, n_ranks);
  return SC_OK;
}

int sc_fit::update_B_end() {
  if (!in_update_B) SC_FAIL(ctx, SC_ERR_STATE, "update_B_end without update_B_begin");
  in_update_B = false;
  {
    unsigned long long ev = 0;
    unsigned long long pruned = 0;
    SC_CUDA(ctx, cudaMemcpyAsync(&pruned, eval_cnt.p + 2, sizeof(pruned), cudaMemcpyDeviceToHost, ctx->stream));
    SC_CUDA(ctx, cudaMemcpyAsync(&ev, eval_cnt.p, sizeof(ev), cudaMemcpyDeviceToHost, ctx->stream));
    SC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    last_evaluated = (long long)ev;
    last_pruned_chunks = (long long)pruned;
    last_chunk_launches = nchunks_cur * T;
    for (int t = 0; t < T; ++t) {
      float ms = 0.f;
      SC_CUDA(ctx, cudaEventElapsedTime(&ms, ev_a[t], ev_b[t]));
      eval_ms_total += ms;
      eval_steps++;
      if (n_hub > 0 && !fresh_kb && t >= 2 && (int)ev_h0.size() > t) {  // incremental launches only (steps 0, 1 evaluate in full)
        SC_CUDA(ctx, cudaEventElapsedTime(&ms, ev_h0[t], ev_h1[t]));
        hub_ms_total += ms;
        hub_steps++;
      }
    }
    long long kbn = 0;  // nnz(K B) of the last Frank-Wolfe step (this rank's rows)
    SC_CUDA(ctx, cudaMemsetAsync(eval_cnt.p + 1, 0, sizeof(unsigned long long), ctx->stream));
    SC_LAUNCH(ctx, k_sum_len, (n + 255) / 256, 256, 0, n, KB.len.p, eval_cnt.p + 1);
    SC_CUDA(ctx, cudaMemcpyAsync(&kbn, eval_cnt.p + 1, sizeof(long long), cudaMemcpyDeviceToHost, ctx->stream));
    SC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    last_kb_nnz = kbn;
  }
  pre_valid = false;
  return SC_OK;
}

// _updateB on one GPU: the split-phase calls back to back (the multi-GPU host inserts one all-gather per step)
int sc_fit::update_B(int* trace_dev) {
  SC_TRY(update_B_begin());
  const int w = ctx->world();
  const size_t pbytes = sizeof(double) * (size_t)k * 4;
  for (int t = 0; t < T; ++t) {
    // this rank's per-archetype {candidate min G, cell, non-candidate G, cell} goes straight into its slab of the gathered
    // buffer; the all-gather is the north_star's exchange of (value, index) argmin pairs
    SC_TRY(update_B_eval(t, part_out.p + (size_t)ctx->rank() * k * 4));
    SC_TRY(exchange(part_out.p, pbytes));
    SC_TRY(update_B_step(t, part_out.p, w, trace_dev));
  }
  return update_B_end();
}

int sc_fit::rss(double* out_host) {
  if (!have_A || !a_is_slots) SC_FAIL(ctx, SC_ERR_STATE, "fit: A has not been computed");
  SC_TRY(precompute_A());
  const int cur = a_cur;
  pstart();
  // per-cell terms on this rank's cells; every rank then sums all n of them in the same fixed order (no floating-point
  // reduction across ranks: the RSS is bitwise the single-GPU value)
  const int nloc = row_end - row_begin;
  if (nloc > 0)
    SC_LAUNCH(ctx, k_rss_cell, (unsigned)(((size_t)nloc * 32 + 255) / 256), 256, 0, row_begin, row_end, k, S, a_idx[cur].p,
              a_val[cur].p, a_cnt[cur].p, KB.view(), t1A.p, percell.p);
  SC_TRY(exchange(percell.p, sizeof(double) * (size_t)slab));
  SC_TRY(sc_sum_f64(ctx, percell.p, scalars.p, (size_t)n, tmp));
  double s = 0.0;
  SC_CUDA(ctx, cudaMemcpyAsync(&s, scalars.p, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  SC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  pstop(P_RSS);
  const double sq = m_fro2 + s;
  *out_host = std::sqrt(sq > 0.0 ? sq : 0.0);
  return SC_OK;
}

int sc_fit::run(int max_iter, int min_iter, double eps, double l2, double* threshold_inout, int* n_iter_out,
                int* converged_out, double* rss_out, int rss_cap, int* n_rss_out) {
  // initialize (cpu.py:204-285) + _fit loop (cpu.py:595-651); B and the previous A must already be set
  int nr = 0;
  SC_TRY(update_A(l2, nullptr));
  double r = 0.0;
  SC_TRY(rss(&r));
  if (nr < rss_cap) rss_out[nr] = r;
  nr++;
  double prev = r;
  if (!(*threshold_inout > 0.0)) *threshold_inout = eps * r;  // sticky once set (cpu.py:280)
  const double thr = *threshold_inout;
  bool converged = false;
  int it = 0;
  while ((!converged && it < max_iter) || it < min_iter) {
    it++;
    SC_TRY(update_A(l2, nullptr));
    SC_TRY(update_B(nullptr));
    SC_TRY(rss(&r));
    if (nr < rss_cap) rss_out[nr] = r;
    nr++;
    if (std::fabs(prev - r) < thr) converged = true;
    prev = r;
  }
  *n_iter_out = it;
  *converged_out = converged ? 1 : 0;
  *n_rss_out = nr;
  return SC_OK;
}

__global__ void k_lists_to_csr(int rows, RowLists L, const long long* __restrict__ off, int* __restrict__ indices,
                               double* __restrict__ data) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= rows) return;
  const long long p = L.ptr[warp], o = off[warp];
  const int len = L.len[warp];
  for (int x = lane; x < len; x += 32) {
    indices[o + x] = L.key[p + x];
    data[o + x] = L.val[p + x];
  }
}

// K B as CSR (n x k): the transpose of Z_ = B_.T @ K (cpu.py:641).  Pass null pointers to get only nnz.
int sc_fit::get_KB_csr(long long* nnz_out, long long* indptr, int* indices, double* data) {
  if (!have_B) SC_FAIL(ctx, SC_ERR_STATE, "fit: B has not been initialised");
  if (row_begin == 0 && row_end == n) {
    SC_TRY(precompute_A());
  } else {  // a rank of a sharded job holds K B for its own cells only: gather every row for this accessor
    SC_TRY(precompute_A());
    SC_TRY(allgather_rows(KB.view(), nullptr, KBs));
    pre_valid = false;  // KBs no longer holds the support rows of t1
  }
  const RowLists KBall = (row_begin == 0 && row_end == n) ? KB.view() : KBs.view();
  SC_TRY(t_off.ensure(ctx, (size_t)n + 1));
  SC_CUDA(ctx, cudaMemcpyAsync(cnt_n.p, KBall.len, sizeof(int) * (size_t)n, cudaMemcpyDeviceToDevice, ctx->stream));
  SC_CUDA(ctx, cudaMemsetAsync(cnt_n.p + n, 0, sizeof(int), ctx->stream));
  SC_TRY(sc_exclusive_scan_i32_to_i64(ctx, cnt_n.p, t_off.p, (size_t)n + 1, tmp));
  long long nnz = 0;
  SC_CUDA(ctx, cudaMemcpyAsync(&nnz, t_off.p + n, sizeof(long long), cudaMemcpyDeviceToHost, ctx->stream));
  SC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  if (nnz_out) *nnz_out = nnz;
  if (!indptr || !indices || !data) return SC_OK;
  DBuf<int> di;
  DBuf<double> dv;
  SC_TRY(di.ensure(ctx, (size_t)nnz + 1));
  SC_TRY(dv.ensure(ctx, (size_t)nnz + 1));
  SC_LAUNCH(ctx, k_lists_to_csr, (unsigned)(((size_t)n * 32 + 255) / 256), 256, 0, n, KBall, t_off.p, di.p, dv.p);
  SC_CUDA(ctx, sc_copy(indptr, t_off.p, sizeof(long long) * ((size_t)n + 1), ctx->stream));
  SC_CUDA(ctx, sc_copy(indices, di.p, sizeof(int) * (size_t)nnz, ctx->stream));
  SC_CUDA(ctx, sc_copy(data, dv.p, sizeof(double) * (size_t)nnz, ctx->stream));
  SC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return SC_OK;
}

int sc_fit::get_A(int* idx, double* val, int* cnt) {
  if (!have_A || !a_is_slots) SC_FAIL(ctx, SC_ERR_STATE, "fit: A has not been computed");
  SC_CUDA(ctx, sc_copy(idx, a_idx[a_cur].p, sizeof(int) * (size_t)n * S, ctx->stream));
  SC_CUDA(ctx, sc_copy(val, a_val[a_cur].p, sizeof(double) * (size_t)n * S, ctx->stream));
  SC_CUDA(ctx, sc_copy(cnt, a_cnt[a_cur].p, sizeof(int) * (size_t)n, ctx->stream));
  SC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return SC_OK;
}

int sc_fit::get_B(int* idx, double* val, int* cnt) {
  if (!have_B) SC_FAIL(ctx, SC_ERR_STATE, "fit: B has not been initialised");
  SC_CUDA(ctx, sc_copy(idx, b_idx.p, sizeof(int) * (size_t)k * S, ctx->stream));
  SC_CUDA(ctx, sc_copy(val, b_val.p, sizeof(double) * (size_t)k * S, ctx->stream));
  SC_CUDA(ctx, sc_copy(cnt, b_cnt.p, sizeof(int) * (size_t)k, ctx->stream));
  SC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return SC_OK;
}

int sc_fit::hard_assignments(int* labels) {
  if (!have_A || !a_is_slots) SC_FAIL(ctx, SC_ERR_STATE, "fit: A has not been computed");
  DBuf<int> lab;
  SC_TRY(lab.ensure(ctx, (size_t)n));
  SC_LAUNCH(ctx, k_hard_labels, (n + 255) / 256, 256, 0, n, S, a_idx[a_cur].p, a_val[a_cur].p, a_cnt[a_cur].p, lab.p);
  SC_CUDA(ctx, sc_copy(labels, lab.p, sizeof(int) * (size_t)n, ctx->stream));
  SC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return SC_OK;
}

int sc_fit::hard_archetypes(int* cells) {
  if (!have_B) SC_FAIL(ctx, SC_ERR_STATE, "fit: B has not been initialised");
  DBuf<int> lab;
  SC_TRY(lab.ensure(ctx, (size_t)k));
  SC_LAUNCH(ctx, k_hard_labels, (k + 255) / 256, 256, 0, k, S, b_idx.p, b_val.p, b_cnt.p, lab.p);
  SC_CUDA(ctx, sc_copy(cells, lab.p, sizeof(int) * (size_t)k, ctx->stream));
  SC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return SC_OK;
}

int sc_fit::soft_assignments(int top, int* idx, double* w) {
  if (!have_A || !a_is_slots) SC_FAIL(ctx, SC_ERR_STATE, "fit: A has not been computed");
  if (top < 1 || top > 8) SC_FAIL(ctx, SC_ERR_ARG, "soft_assignments: top must be in [1, 8]");
  DBuf<int> di;
  DBuf<double> dw;
  SC_TRY(di.ensure(ctx, (size_t)n * top));
  SC_TRY(dw.ensure(ctx, (size_t)n * top));
  SC_LAUNCH(ctx, k_soft_top, (n + 255) / 256, 256, 0, n, k, S, top, a_idx[a_cur].p, a_val[a_cur].p, a_cnt[a_cur].p,
            di.p, dw.p);
  SC_CUDA(ctx, sc_copy(idx, di.p, sizeof(int) * (size_t)n * top, ctx->stream));
  SC_CUDA(ctx, sc_copy(w, dw.p, sizeof(double) * (size_t)n * top, ctx->stream));
  SC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return SC_OK;
}
