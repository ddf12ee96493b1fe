// Frank-Wolfe fit engine: _updateA / _updateB / compute_RSS of cpu.py:425-536 on compressed A and B.
//
// Layout in HBM (n cells, k archetypes, S = max_FW_iter slots per column):
//   M            CSR int32/int32/fp64                       ~ 24 nnz per row
//   A slots      a_idx[n*S] int32, a_val[n*S] fp64, a_cnt[n]  (column j of A = slots of cell j)
//   B slots      b_idx[k*S] int32, b_val[k*S] fp64, b_cnt[k]  (column a of B = slots of archetype a)
//   KB rows      per-cell lists (archetype, (K B)[cell, archetype]), K = M M^T applied as M (M .)   ~ 40 per cell
//   t1A          dense k x k fp64, t1A[i'][i] = (B^T K B)[i, i']      (gradient term of _updateA, cpu.py:442)
//   t1B          dense k x k fp64, A A^T                                (gradient term of _updateB, cpu.py:480)
//   C_a          per-archetype candidate lists (cell, (K A^T)[cell, a]) (cpu.py:481), the support of t2
// Nothing of size n x k or n x n is ever formed.
//
// Exactness of the candidate restriction (SURVEY.md §8a notes): K, A, B >= 0, so a gradient entry can be negative only
// on the support of t2.  If the minimum over that support is < 0 it is the column argmin; otherwise the column is
// scanned densely from index 0 with early exit at the first exact zero, which is what np.argmin over the dense column
// (cpu_dense.py:363, 398) and scipy's implicit-zero rule (scipy/sparse/_data.py, reached from cpu.py:450, 487) return.
#include "fit.cuh"

#include <algorithm>
#include <cmath>

#include "fit_kernels.cuh"

// ------------------------------------------------------------------------------------------------ small kernels
__global__ void k_fill_stride_ptr(long long* p, int rows, int stride) {
  int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r <= rows) p[r] = (long long)r * stride;
}

__global__ void k_square(const double* __restrict__ in, double* __restrict__ out, size_t n) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = in[i] * in[i];
}

// B by cell: count, fill, sort
__global__ void k_brow_count(int k, int S, const int* __restrict__ b_idx, const int* __restrict__ b_cnt, int* cnt) {
  int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= k * S) return;
  int a = e / S, s = e - a * S;
  if (s < b_cnt[a]) atomicAdd(&cnt[b_idx[e]], 1);
}
__global__ void k_brow_fill(int k, int S, const int* __restrict__ b_idx, const double* __restrict__ b_val,
                            const int* __restrict__ b_cnt, const long long* __restrict__ ptr, int* cursor, int* key,
                            double* val) {
  int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= k * S) return;
  int a = e / S, s = e - a * S;
  if (s >= b_cnt[a]) return;
  int cell = b_idx[e];
  int pos = atomicAdd(&cursor[cell], 1);
  key[ptr[cell] + pos] = a;
  val[ptr[cell] + pos] = b_val[e];
}
__global__ void k_rows_sort_small(int rows, const long long* __restrict__ ptr, const int* __restrict__ len, int* key,
                                  double* val) {
  int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= rows) return;
  int L = len[r];
  if (L < 2) return;
  long long p = ptr[r];
  for (int i = 1; i < L; ++i) {
    int kk = key[p + i];
    double vv = val[p + i];
    int j = i - 1;
    while (j >= 0 && key[p + j] > kk) {
      key[p + j + 1] = key[p + j];
      val[p + j + 1] = val[p + j];
      --j;
    }
    key[p + j + 1] = kk;
    val[p + j + 1] = vv;
  }
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
};

__global__ void __launch_bounds__(FWA_WARPS * 32) k_updateA_fast(UpdAArgs g) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr size_t PER_WARP = (size_t)(2 * FWA_NC + FWA_TCAP) * sizeof(double) + (size_t)(FWA_NC + FW_SMAX) * sizeof(int);
  unsigned char* base = smem_raw + (size_t)warp * PER_WARP;
  double* ct2 = reinterpret_cast<double*>(base);
  double* ca = ct2 + FWA_NC;
  double* Tc = ca + FWA_NC;
  int* ck = reinterpret_cast<int*>(Tc + FWA_TCAP);
  int* slv = ck + FWA_NC;
  const int k = g.k, S = g.S, T = g.T;
  const double* __restrict__ t1T = g.t1T;

  for (int j = g.cell_begin + blockIdx.x * FWA_WARPS + warp; j < g.cell_end; j += gridDim.x * FWA_WARPS) {
    const int nc = g.KB.len[j];
    bool slow = (nc == 0 || nc > FWA_NC);
    if (!slow) {
      const long long kp = g.KB.ptr[j];
      for (int c = lane; c < nc; c += 32) {
        ck[c] = g.KB.key[kp + c];
        ct2[c] = g.KB.val[kp + c];
        ca[c] = 0.0;
      }
      __syncwarp();
      const int scap = min(FWA_TCAP / nc, FW_SMAX);
      // ---- t = 0: gradient against the previous A (cpu.py:447 with A = A_prev)
      const int m0 = g.Aprev.len[j];
      const long long pp = g.Aprev.ptr[j];
      double bv = FW_INF;
      int bi = 0x7fffffff;
      for (int c = lane; c < nc; c += 32) {
        double acc = 0.0;
        const int id = ck[c];
        for (int e = 0; e < m0; ++e) acc = fma(t1T[(size_t)g.Aprev.key[pp + e] * k + id], g.Aprev.val[pp + e], acc);
        const double G = 2.0 * (acc - ct2[c]);
        if (G < bv) {
          bv = G;
          bi = c;
        }
      }
      warp_argmin(bv, bi);
      if (!(bv < 0.0)) {
        slow = true;
      } else {
        int cstar = bi;
        // previous weight of the chosen archetype (A_prev[amin, j])
        double aprev = 0.0;
        for (int e0 = 0; e0 < m0; e0 += 32) {
          const int e = e0 + lane;
          const bool hit = e < m0 && g.Aprev.key[pp + e] == ck[cstar];
          const unsigned ball = __ballot_sync(0xffffffffu, hit);
          if (ball) {
            const int src = __ffs(ball) - 1;
            aprev = __shfl_sync(0xffffffffu, hit ? g.Aprev.val[pp + e] : 0.0, src);
            break;
          }
        }
        int nslots = 1;
        if (lane == 0) {
          slv[0] = cstar;
          ca[cstar] = fw_step(aprev, 1.0, fw_gamma(0));
          if (g.trace) g.trace[j] = ck[cstar];
        }
        if (scap > 0)
          for (int c = lane; c < nc; c += 32) Tc[c] = t1T[(size_t)ck[cstar] * k + ck[c]];
        __syncwarp();
        // ---- t >= 1
        for (int t = 1; t < T; ++t) {
          bv = FW_INF;
          bi = 0x7fffffff;
          for (int c = lane; c < nc; c += 32) {
            double acc = 0.0;
            const int id = ck[c];
            for (int s = 0; s < nslots; ++s) {
              const int v = slv[s];
              const double tv = (s < scap) ? Tc[s * nc + c] : t1T[(size_t)ck[v] * k + id];
              acc = fma(tv, ca[v], acc);
            }
            const double G = 2.0 * (acc - ct2[c]);
            if (G < bv) {
              bv = G;
              bi = c;
            }
          }
          warp_argmin(bv, bi);
          if (!(bv < 0.0)) {
            slow = true;
            break;
          }
          cstar = bi;
          const double gamma = fw_gamma(t);
          const bool was_active = ca[cstar] != 0.0;
          __syncwarp();
          for (int s = lane; s < nslots; s += 32) {
            const int v = slv[s];
            ca[v] = fw_step(ca[v], v == cstar ? 1.0 : 0.0, gamma);
          }
          if (!was_active) {
            if (lane == 0) {
              ca[cstar] = fw_step(0.0, 1.0, gamma);
              slv[nslots] = cstar;
            }
            if (nslots < scap)
              for (int c = lane; c < nc; c += 32) Tc[nslots * nc + c] = t1T[(size_t)ck[cstar] * k + ck[c]];
            nslots++;
          }
          if (lane == 0 && g.trace) g.trace[(size_t)t * g.n + j] = ck[cstar];
          __syncwarp();
        }
        if (!slow) {
          for (int s = lane; s < nslots; s += 32) {
            g.a_idx[(size_t)j * S + s] = ck[slv[s]];
            g.a_val[(size_t)j * S + s] = ca[slv[s]];
          }
          if (lane == 0) g.a_cnt[j] = nslots;
        }
      }
    }
    if (slow && lane == 0) g.slow_list[atomicAdd(g.slow_cnt, 1)] = j;
    __syncwarp();
  }
}

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
// t1B[l][a] = sum_j A[l,j] A[a,j], cells j ascending (the order of scipy's csr_matmat for A @ A.T, cpu.py:480)
__global__ void k_gram(int k, int S, const long long* __restrict__ rptr, const unsigned long long* __restrict__ keys,
                       const double* __restrict__ vals, const int* __restrict__ a_idx,
                       const double* __restrict__ a_val, const int* __restrict__ a_cnt, double* __restrict__ t1) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= k) return;
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

// ------------------------------------------------------------------------------------------------ candidates by archetype
__global__ void k_count_keys(int rows, RowLists L, int* cnt) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= rows) return;
  const long long p = L.ptr[warp];
  const int len = L.len[warp];
  for (int x = lane; x < len; x += 32) atomicAdd(&cnt[L.key[p + x]], 1);
}
__global__ void k_scatter_by_key(int rows, RowLists L, const long long* __restrict__ cptr, int* cursor, int* c_cell,
                                 double* c_val) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= rows) return;
  const long long p = L.ptr[warp];
  const int len = L.len[warp];
  for (int x = lane; x < len; x += 32) {
    const int a = L.key[p + x];
    const long long pos = cptr[a] + atomicAdd(&cursor[a], 1);
    c_cell[pos] = warp;
    c_val[pos] = L.val[p + x];
  }
}

// ------------------------------------------------------------------------------------------------ _updateB, one FW step
struct UpdBArgs {
  int n, k, S, t;
  const double* t1;  // Gram, row a contiguous (symmetric)
  const long long* c_ptr;
  const int* c_cell;
  const double* c_t2;
  RowLists KB;   // (K B)[cell, :] for the current B
  RowLists T2c;  // (K A^T)[cell, :] sorted by archetype (dense-scan lookups)
  int* b_idx;
  double* b_val;
  int* b_cnt;
  int* trace;  // [T * k] or null
};

__global__ void __launch_bounds__(256) k_updateB_iter(UpdBArgs g) {
  __shared__ double sg[8];
  __shared__ int si[8];
  __shared__ double s_bestv;
  __shared__ int s_besti;
  __shared__ int s_zero;
  const int a = blockIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int k = g.k, S = g.S;
  const double* __restrict__ row = g.t1 + (size_t)a * k;
  const long long c0 = g.c_ptr[a];
  const int nc = (int)(g.c_ptr[a + 1] - c0);

  double bv = FW_INF;
  int bi = 0x7fffffff;
  for (int e = warp; e < nc; e += 8) {
    const int cell = g.c_cell[c0 + e];
    const double t2v = g.c_t2[c0 + e];
    const long long p = g.KB.ptr[cell];
    const int L = g.KB.len[cell];
    double acc = 0.0;
    for (int x = lane; x < L; x += 32) acc = fma(g.KB.val[p + x], row[g.KB.key[p + x]], acc);
    const double H = warp_sum(acc);
    const double G = 2.0 * (H - t2v);
    if (G < bv || (G == bv && cell < bi)) {
      bv = G;
      bi = cell;
    }
  }
  if (lane == 0) {
    sg[warp] = bv;
    si[warp] = bi;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double v = sg[0];
    int i = si[0];
    for (int w = 1; w < 8; ++w)
      if (sg[w] < v || (sg[w] == v && si[w] < i)) {
        v = sg[w];
        i = si[w];
      }
    s_bestv = v;
    s_besti = i;
    s_zero = 0x7fffffff;
  }
  __syncthreads();
  int amin = s_besti;
  if (!(s_bestv < 0.0)) {
    // dense column scan from cell 0 with early exit at the first exact zero
    double tv = FW_INF;
    int ti = 0x7fffffff;
    bool found = false;
    for (int i0 = 0; i0 < g.n; i0 += 256) {
      const int i = i0 + threadIdx.x;
      double G = FW_INF;
      if (i < g.n) {
        const long long p = g.KB.ptr[i];
        const int L = g.KB.len[i];
        double acc = 0.0;
        for (int x = 0; x < L; ++x) acc = fma(g.KB.val[p + x], row[g.KB.key[p + x]], acc);
        const double t2v = list_lookup(g.T2c.key + g.T2c.ptr[i], g.T2c.val + g.T2c.ptr[i], g.T2c.len[i], a);
        G = 2.0 * (acc - t2v);
        if (G == 0.0) atomicMin(&s_zero, i);
        else if (G < tv) {
          tv = G;
          ti = i;
        }
      }
      __syncthreads();
      const int z = s_zero;
      __syncthreads();  // everyone has read s_zero before the next chunk may lower it
      if (z != 0x7fffffff) {
        found = true;
        break;
      }
    }
    if (found) {
      amin = s_zero;
    } else {
      // block-wide lexicographic minimum
      warp_argmin(tv, ti);
      if (lane == 0) {
        sg[warp] = tv;
        si[warp] = ti;
      }
      __syncthreads();
      if (threadIdx.x == 0) {
        double v = sg[0];
        int i = si[0];
        for (int w = 1; w < 8; ++w)
          if (sg[w] < v || (sg[w] == v && si[w] < i)) {
            v = sg[w];
            i = si[w];
          }
        s_besti = i;
      }
      __syncthreads();
      amin = s_besti;
    }
  }
  // simplex step on column a of B (cpu.py:494)
  if (warp == 0) {
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
  }
}

// ------------------------------------------------------------------------------------------------ RSS, labels
// out[j] = a_j^T t1A a_j - 2 <KB[j,:], a_j>     (||M - M B A||_F^2 = ||M||_F^2 + sum_j out[j])
__global__ void k_rss_cell(int n, int k, int S, const int* __restrict__ a_idx, const double* __restrict__ a_val,
                           const int* __restrict__ a_cnt, RowLists KB, const double* __restrict__ t1T,
                           double* __restrict__ out) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= n) return;
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
static int set_smem_attr_once(sc_ctx* ctx) {
  static bool done = false;
  if (done) return SC_OK;
  constexpr size_t PER_WARP = (size_t)(2 * FWA_NC + FWA_TCAP) * sizeof(double) + (size_t)(FWA_NC + FW_SMAX) * sizeof(int);
  SC_CUDA(ctx, cudaFuncSetAttribute(k_updateA_fast, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    (int)(PER_WARP * FWA_WARPS)));
  done = true;
  return SC_OK;
}

int sc_fit::init(sc_ctx* c, int n_, int k_, int T_) {
  ctx = c;
  n = n_;
  k = k_;
  T = T_;
  S = std::max(T_, 1);
  if (n <= 0 || k <= 0 || k > n) SC_FAIL(ctx, SC_ERR_ARG, "fit: need 0 < k <= n");
  if (T < 1 || T > FW_SMAX) SC_FAIL(ctx, SC_ERR_ARG, "fit: max_FW_iter must be in [1, 64]");
  SC_TRY(slot_ptr_n.ensure(ctx, (size_t)n + 1));
  SC_LAUNCH(ctx, k_fill_stride_ptr, (n + 256) / 256, 256, 0, slot_ptr_n.p, n, S);
  for (int b = 0; b < 2; ++b) {
    SC_TRY(a_idx[b].ensure(ctx, (size_t)n * S));
    SC_TRY(a_val[b].ensure(ctx, (size_t)n * S));
    SC_TRY(a_cnt[b].ensure(ctx, (size_t)n));
  }
  SC_TRY(b_idx.ensure(ctx, (size_t)k * S));
  SC_TRY(b_val.ensure(ctx, (size_t)k * S));
  SC_TRY(b_cnt.ensure(ctx, (size_t)k));
  SC_TRY(t1A.ensure(ctx, (size_t)k * k));
  SC_TRY(t1B.ensure(ctx, (size_t)k * k));
  SC_TRY(percell.ensure(ctx, (size_t)std::max(n, 1024)));
  SC_TRY(scalars.ensure(ctx, 8));
  SC_TRY(slow_list.ensure(ctx, (size_t)n));
  SC_TRY(slow_cnt.ensure(ctx, 4));
  SC_TRY(cnt_n.ensure(ctx, (size_t)n + 1));
  SC_TRY(cnt_k.ensure(ctx, (size_t)k + 1));
  SC_TRY(cursor_n.ensure(ctx, (size_t)n + 1));
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
  // ||M||_F^2 (fixed-order reduction)
  DBuf<double> sq;
  SC_TRY(sq.ensure(ctx, (size_t)nnz));
  SC_LAUNCH(ctx, k_square, (unsigned)((nnz + 255) / 256), 256, 0, m_val.p, sq.p, (size_t)nnz);
  SC_TRY(sc_sum_f64(ctx, sq.p, scalars.p, (size_t)nnz, tmp));
  SC_CUDA(ctx, cudaMemcpyAsync(&m_fro2, scalars.p, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  SC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
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
  SC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
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

// KB rows and t1A for the current B (shared by _updateA and compute_RSS)
int sc_fit::compute_KB() {
  CsrView M{n, m_ptr.p, m_col.p, m_val.p};
  // B by cell
  SC_CUDA(ctx, cudaMemsetAsync(cnt_n.p, 0, sizeof(int) * ((size_t)n + 1), ctx->stream));
  SC_CUDA(ctx, cudaMemsetAsync(cursor_n.p, 0, sizeof(int) * ((size_t)n + 1), ctx->stream));
  const int ks = k * S;
  SC_LAUNCH(ctx, k_brow_count, (ks + 255) / 256, 256, 0, k, S, b_idx.p, b_cnt.p, cnt_n.p);
  SC_TRY(brow.ptr.ensure(ctx, (size_t)n + 1));
  SC_TRY(brow.key.ensure(ctx, (size_t)ks));
  SC_TRY(brow.val.ensure(ctx, (size_t)ks));
  SC_TRY(sc_exclusive_scan_i32_to_i64(ctx, cnt_n.p, brow.ptr.p, (size_t)n + 1, tmp));
  SC_LAUNCH(ctx, k_brow_fill, (ks + 255) / 256, 256, 0, k, S, b_idx.p, b_val.p, b_cnt.p, brow.ptr.p, cursor_n.p,
            brow.key.p, brow.val.p);
  SC_LAUNCH(ctx, k_rows_sort_small, (n + 255) / 256, 256, 0, n, brow.ptr.p, cnt_n.p, brow.key.p, brow.val.p);
  RowLists Bl{brow.ptr.p, cnt_n.p, brow.key.p, brow.val.p};
  SC_TRY(sc_spgemm_rows(ctx, M, 0, n, Bl, Y, sp, hash_cap));
  SC_TRY(sc_spgemm_rows(ctx, M, 0, n, Y.view(), KB, sp, hash_cap));
  return SC_OK;
}

int sc_fit::precompute_A() {
  if (pre_valid) return SC_OK;
  if (!have_M || !have_B) SC_FAIL(ctx, SC_ERR_STATE, "fit: kernel matrix and B must be set first");
  SC_TRY(compute_KB());
  SC_CUDA(ctx, cudaMemsetAsync(t1A.p, 0, sizeof(double) * (size_t)k * k, ctx->stream));
  SC_LAUNCH(ctx, k_t1A, (k * 32 + 127) / 128, 128, 0, k, S, b_idx.p, b_val.p, b_cnt.p, KB.view(), t1A.p);
  pre_valid = true;
  return SC_OK;
}

int sc_fit::update_A(double l2, int* trace_dev) {
  if (!have_A) SC_FAIL(ctx, SC_ERR_STATE, "fit: A has not been initialised");
  SC_TRY(precompute_A());
  const int nxt = (a_cur < 0) ? 0 : 1 - a_cur;
  UpdAArgs g;
  g.cell_begin = 0;
  g.cell_end = n;
  g.n = n;
  g.k = k;
  g.S = S;
  g.T = T;
  g.l2 = l2;
  g.KB = KB.view();
  g.t1T = t1A.p;
  g.Aprev = A_view();
  g.a_idx = a_idx[nxt].p;
  g.a_val = a_val[nxt].p;
  g.a_cnt = a_cnt[nxt].p;
  g.trace = trace_dev;
  g.slow_list = slow_list.p;
  g.slow_cnt = slow_cnt.p;
  SC_CUDA(ctx, cudaMemsetAsync(slow_cnt.p, 0, sizeof(int), ctx->stream));
  int nslow = 0;
  if (l2 == 0.0) {
    constexpr size_t PER_WARP =
        (size_t)(2 * FWA_NC + FWA_TCAP) * sizeof(double) + (size_t)(FWA_NC + FW_SMAX) * sizeof(int);
    int grid = std::min((n + FWA_WARPS - 1) / FWA_WARPS, 148 * 16);
    SC_LAUNCH(ctx, k_updateA_fast, grid, FWA_WARPS * 32, PER_WARP * FWA_WARPS, g);
    SC_CUDA(ctx, cudaMemcpyAsync(&nslow, slow_cnt.p, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    SC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  } else {
    // the -l2*A term can make entries outside the support of t2 negative: every cell takes the generic path
    std::vector<int> all(n);
    for (int j = 0; j < n; ++j) all[j] = j;
    SC_CUDA(ctx, sc_copy(slow_list.p, all.data(), sizeof(int) * (size_t)n, ctx->stream));
    SC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    nslow = n;
  }
  last_slow_cells = nslow;
  if (nslow > 0) {
    int grid = std::min((nslow + 3) / 4, 148 * 16);
    SC_LAUNCH(ctx, k_updateA_slow, grid, 128, 0, g, nslow);
  }
  a_cur = nxt;
  a_is_slots = true;
  return SC_OK;
}

int sc_fit::update_B(int* trace_dev) {
  if (!have_A || !a_is_slots) SC_FAIL(ctx, SC_ERR_STATE, "fit: run update_A before update_B");
  if (!have_B) SC_FAIL(ctx, SC_ERR_STATE, "fit: B has not been initialised");
  CsrView M{n, m_ptr.p, m_col.p, m_val.p};
  const int cur = a_cur;
  // ---- t1B = A A^T: transpose A by a stable sort on (archetype, cell), then accumulate rows in cell order
  SC_TRY(a_off.ensure(ctx, (size_t)n + 1));
  SC_CUDA(ctx, cudaMemcpyAsync(cnt_n.p, a_cnt[cur].p, sizeof(int) * (size_t)n, cudaMemcpyDeviceToDevice, ctx->stream));
  SC_CUDA(ctx, cudaMemsetAsync(cnt_n.p + n, 0, sizeof(int), ctx->stream));
  SC_TRY(sc_exclusive_scan_i32_to_i64(ctx, cnt_n.p, a_off.p, (size_t)n + 1, tmp));
  long long annz = 0;
  SC_CUDA(ctx, cudaMemcpyAsync(&annz, a_off.p + n, sizeof(long long), cudaMemcpyDeviceToHost, ctx->stream));
  SC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  SC_TRY(gk_in.ensure(ctx, (size_t)annz + 1));
  SC_TRY(gk_out.ensure(ctx, (size_t)annz + 1));
  SC_TRY(gv_in.ensure(ctx, (size_t)annz + 1));
  SC_TRY(gv_out.ensure(ctx, (size_t)annz + 1));
  SC_TRY(arow_ptr.ensure(ctx, (size_t)k + 1));
  SC_LAUNCH(ctx, k_emit_a_pairs, (unsigned)(((size_t)n * 32 + 255) / 256), 256, 0, n, S, a_idx[cur].p, a_val[cur].p,
            a_cnt[cur].p, a_off.p, gk_in.p, gv_in.p);
  int kbits = 1;
  while ((1ll << kbits) < k) kbits++;
  SC_TRY(sc_sort_pairs_u64_f64(ctx, gk_in.p, gk_out.p, gv_in.p, gv_out.p, (size_t)annz, 32 + kbits, tmp));
  SC_LAUNCH(ctx, k_row_bounds, (k + 256) / 256, 256, 0, k, gk_out.p, annz, arow_ptr.p);
  SC_CUDA(ctx, cudaMemsetAsync(t1B.p, 0, sizeof(double) * (size_t)k * k, ctx->stream));
  SC_LAUNCH(ctx, k_gram, (k * 32 + 127) / 128, 128, 0, k, S, arow_ptr.p, gk_out.p, gv_out.p, a_idx[cur].p,
            a_val[cur].p, a_cnt[cur].p, t1B.p);
  // ---- t2 = K A^T as per-cell rows, then as per-archetype candidate lists
  RowLists Al{slot_ptr_n.p, a_cnt[cur].p, a_idx[cur].p, a_val[cur].p};
  SC_TRY(sc_spgemm_rows(ctx, M, 0, n, Al, MA, sp, hash_cap));
  SC_TRY(sc_spgemm_rows(ctx, M, 0, n, MA.view(), T2c, sp, hash_cap));
  SC_CUDA(ctx, cudaMemsetAsync(cnt_k.p, 0, sizeof(int) * ((size_t)k + 1), ctx->stream));
  SC_LAUNCH(ctx, k_count_keys, (unsigned)(((size_t)n * 32 + 255) / 256), 256, 0, n, T2c.view(), cnt_k.p);
  SC_TRY(c_ptr.ensure(ctx, (size_t)k + 1));
  SC_TRY(sc_exclusive_scan_i32_to_i64(ctx, cnt_k.p, c_ptr.p, (size_t)k + 1, tmp));
  long long cnnz = 0;
  SC_CUDA(ctx, cudaMemcpyAsync(&cnnz, c_ptr.p + k, sizeof(long long), cudaMemcpyDeviceToHost, ctx->stream));
  SC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  last_t2_nnz = cnnz;
  SC_TRY(c_cell.ensure(ctx, (size_t)cnnz + 1));
  SC_TRY(c_t2.ensure(ctx, (size_t)cnnz + 1));
  SC_CUDA(ctx, cudaMemsetAsync(cnt_k.p, 0, sizeof(int) * ((size_t)k + 1), ctx->stream));
  SC_LAUNCH(ctx, k_scatter_by_key, (unsigned)(((size_t)n * 32 + 255) / 256), 256, 0, n, T2c.view(), c_ptr.p, cnt_k.p,
            c_cell.p, c_t2.p);
  // ---- T Frank-Wolfe steps; K B is re-evaluated from the current B at every step (cpu.py:486)
  for (int t = 0; t < T; ++t) {
    SC_TRY(compute_KB());
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
    SC_LAUNCH(ctx, k_updateB_iter, k, 256, 0, g);
  }
  pre_valid = false;
  return SC_OK;
}

int sc_fit::rss(double* out_host) {
  if (!have_A || !a_is_slots) SC_FAIL(ctx, SC_ERR_STATE, "fit: A has not been computed");
  SC_TRY(precompute_A());
  const int cur = a_cur;
  SC_LAUNCH(ctx, k_rss_cell, (unsigned)(((size_t)n * 32 + 255) / 256), 256, 0, n, k, S, a_idx[cur].p, a_val[cur].p,
            a_cnt[cur].p, KB.view(), t1A.p, percell.p);
  SC_TRY(sc_sum_f64(ctx, percell.p, scalars.p, (size_t)n, tmp));
  double s = 0.0;
  SC_CUDA(ctx, cudaMemcpyAsync(&s, scalars.p, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  SC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
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
