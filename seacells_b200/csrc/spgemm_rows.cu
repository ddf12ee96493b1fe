// Z = S * X  for a CSR matrix S (n x m, the kernel matrix M) and per-row sparse lists X (m rows keyed by archetype).
//
// This is the engine's "SpMM for K*B" (cpu.py:441, 481, 484 compute K@B, K@A.T with K = M M^T): two passes of this
// kernel apply M twice, and the right-hand side stays in its compressed form, so the output never becomes the dense
// n x k matrix the reference's gpu.py forms (107 GB at n = 1M, k = 13333).
//
// One warp per output row.  The row's (key, value) pairs are accumulated in a shared-memory open-addressing table.
// Accumulation order is fixed (ascending position in the row of S, then position in the row of X): colliding
// contributions inside one 32-wide batch are applied rank by rank, so results are bitwise reproducible.
// Output rows are sorted by key.
#include "common.cuh"
#include "spgemm_rows.cuh"

#include <algorithm>

__global__ void k_spgemm_ub(int row_begin, int row_end, const int* __restrict__ s_ptr, const int* __restrict__ s_col,
                            const int* __restrict__ x_len, long long* __restrict__ ub, int nrows_total) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i > nrows_total) return;
  long long u = 0;
  if (i >= row_begin && i < row_end) {
    for (int p = s_ptr[i]; p < s_ptr[i + 1]; ++p) u += x_len[s_col[p]];
  }
  ub[i] = u;  // ub[nrows_total] = 0 so that the exclusive scan ends with the total
}

__device__ __forceinline__ unsigned hash_key(int k) { return (unsigned)k * 2654435761u; }

template <int CAP, int WARPS>
__global__ void __launch_bounds__(WARPS * 32)
k_spgemm_fill(int row_begin, int row_end, const int* __restrict__ s_ptr, const int* __restrict__ s_col,
              const double* __restrict__ s_val, RowLists X, const long long* __restrict__ z_ptr,
              int* __restrict__ z_len, int* __restrict__ z_key, double* __restrict__ z_val, int* __restrict__ flags,
              int* __restrict__ big_list) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  double* tvals = reinterpret_cast<double*>(smem_raw) + (size_t)warp * CAP;
  int* tkeys = reinterpret_cast<int*>(smem_raw + (size_t)WARPS * CAP * sizeof(double)) + (size_t)warp * CAP;
  const unsigned FULL = 0xffffffffu;
  const unsigned lt_mask = (1u << lane) - 1u;

  for (int i = row_begin + blockIdx.x * WARPS + warp; i < row_end; i += gridDim.x * WARPS) {
    const long long zp = z_ptr[i];
    const long long ub = z_ptr[i + 1] - zp;
    if (ub == 0) {
      if (lane == 0) z_len[i] = 0;
      continue;
    }
    // effective table size: power of two >= 2*ub, capped at CAP
    int cap = 64;
    while (cap < CAP && (long long)cap < 2 * ub) cap <<= 1;
    const int capm = cap - 1;
    const int limit = cap - (cap >> 3);  // refuse to fill beyond 7/8
    for (int t = lane; t < cap; t += 32) {
      tkeys[t] = -1;
      tvals[t] = 0.0;
    }
    __syncwarp();
    int uniq = 0;
    bool overflow = false;
    const int s0 = s_ptr[i], s1 = s_ptr[i + 1];
    for (int base = s0; base < s1 && !overflow; base += 32) {
      int q = -1, xl = 0;
      double m = 0.0;
      long long xp = 0;
      if (base + lane < s1) {
        q = s_col[base + lane];
        m = s_val[base + lane];
        xl = X.len[q];
        xp = X.ptr[q];
      }
      int incl = xl;  // inclusive warp scan of list lengths
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        int v = __shfl_up_sync(FULL, incl, d);
        if (lane >= d) incl += v;
      }
      const int tot = __shfl_sync(FULL, incl, 31);
      for (int f0 = 0; f0 < tot; f0 += 32) {
        const int f = f0 + lane;
        const bool valid = f < tot;
        // owner lane: smallest o with incl[o] > f
        int lo = 0;
#pragma unroll
        for (int step = 16; step > 0; step >>= 1) {
          int pv = __shfl_sync(FULL, incl, lo + step - 1);
          if (pv <= f) lo += step;
        }
        lo = min(lo, 31);
        const int o_incl = __shfl_sync(FULL, incl, lo);
        const int o_len = __shfl_sync(FULL, xl, lo);
        const long long o_ptr = __shfl_sync(FULL, xp, lo);
        const double o_m = __shfl_sync(FULL, m, lo);
        int key = 0, slot = -1 - lane;
        double c = 0.0;
        bool is_new = false;
        if (valid) {
          const long long e = o_ptr + (f - (o_incl - o_len));
          key = X.key[e];
          c = __dmul_rn(o_m, X.val[e]);
          int s = hash_key(key) & capm;
          while (true) {
            int prev = atomicCAS(&tkeys[s], -1, key);
            if (prev == -1) {
              is_new = true;
              break;
            }
            if (prev == key) break;
            s = (s + 1) & capm;
          }
          slot = s;
        }
        uniq += __popc(__ballot_sync(FULL, is_new));
        if (uniq > limit) {
          overflow = true;
          break;
        }
        __syncwarp();
        const unsigned grp = __match_any_sync(FULL, slot);
        const int rank = __popc(grp & lt_mask);
        const int maxcnt = __reduce_max_sync(FULL, __popc(grp));
        for (int r = 0; r < maxcnt; ++r) {
          if (valid && rank == r) tvals[slot] = __dadd_rn(tvals[slot], c);
          __syncwarp();
        }
      }
    }
    if (overflow) {  // too many distinct keys for the shared-memory table: hand the row to k_spgemm_big
      if (lane == 0) {
        big_list[atomicAdd(&flags[0], 1)] = i;
        z_len[i] = 0;
      }
      __syncwarp();
      continue;
    }
    // in-place compaction of occupied slots to [0, uniq)
    int w = 0;
    for (int base = 0; base < cap; base += 32) {
      const int kk = tkeys[base + lane];
      const double vv = tvals[base + lane];
      const bool occ = kk != -1;
      const unsigned ball = __ballot_sync(FULL, occ);
      __syncwarp();
      if (occ) {
        const int pos = w + __popc(ball & lt_mask);
        tkeys[pos] = kk;
        tvals[pos] = vv;
      }
      w += __popc(ball);
      __syncwarp();
    }
    // bitonic sort by key over the next power of two
    int P = 1;
    while (P < uniq) P <<= 1;
    for (int t = uniq + lane; t < P; t += 32) tkeys[t] = 0x7fffffff;
    __syncwarp();
    for (int k2 = 2; k2 <= P; k2 <<= 1) {
      for (int j = k2 >> 1; j > 0; j >>= 1) {
        for (int t = lane; t < P; t += 32) {
          const int x = t ^ j;
          if (x > t) {
            const int ka = tkeys[t], kb = tkeys[x];
            const bool asc = (t & k2) == 0;
            if ((ka > kb) == asc) {
              tkeys[t] = kb;
              tkeys[x] = ka;
              const double va = tvals[t];
              tvals[t] = tvals[x];
              tvals[x] = va;
            }
          }
        }
        __syncwarp();
      }
    }
    if (lane == 0) z_len[i] = uniq;
    for (int t = lane; t < uniq; t += 32) {
      z_key[zp + t] = tkeys[t];
      z_val[zp + t] = tvals[t];
    }
    __syncwarp();
  }
}

// First-occurrence variant used inside the Frank-Wolfe loop of _updateB, where the consumers only walk the rows: keys
// are appended in the order they first appear in the (row of S, row of X) sequence - a fixed order - so neither the
// compaction nor the sort is needed.  The table maps key -> list index; one leader lane per distinct key of a batch does
// the find-or-insert, new indices are handed out in lane order, duplicates are accumulated rank by rank as above.
// Rows with at most 32 products and at most 32 stored entries of S: everything fits one 32-wide batch, so the row is
// finished in registers - no shared-memory table.  Same arithmetic and the same first-occurrence order as
// k_spgemm_fill_fo (duplicates of a key are added in ascending batch position, starting from the first one).  This is
// the common row of the per-step delta products K E_t of the incremental K B (fit.cu), which the table kernel handled
// at ~1 us of instructions per row.
constexpr int TINY_MAX = 32;
__global__ void __launch_bounds__(256)
k_spgemm_tiny_fo(int row_begin, int row_end, const int* __restrict__ s_ptr, const int* __restrict__ s_col,
                 const double* __restrict__ s_val, RowLists X, const long long* __restrict__ z_ptr,
                 int* __restrict__ z_len, int* __restrict__ z_key, double* __restrict__ z_val) {
  const int lane = threadIdx.x & 31;
  const unsigned FULL = 0xffffffffu;
  const unsigned lt_mask = (1u << lane) - 1u;
  const int warps = (gridDim.x * blockDim.x) >> 5;
  for (int i = row_begin + ((blockIdx.x * blockDim.x + threadIdx.x) >> 5); i < row_end; i += warps) {
    const long long zp = z_ptr[i];
    const long long ub = z_ptr[i + 1] - zp;
    if (ub == 0) {
      if (lane == 0) z_len[i] = 0;
      continue;
    }
    const int s0 = s_ptr[i], s1 = s_ptr[i + 1];
    if (ub > TINY_MAX || s1 - s0 > TINY_MAX) continue;  // k_spgemm_fill_fo takes this row
    int q = -1, xl = 0;
    double m = 0.0;
    long long xp = 0;
    if (s0 + lane < s1) {
      q = s_col[s0 + lane];
      m = s_val[s0 + lane];
      xl = X.len[q];
      xp = X.ptr[q];
    }
    int incl = xl;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const int v = __shfl_up_sync(FULL, incl, d);
      if (lane >= d) incl += v;
    }
    const int tot = (int)ub;
    const bool valid = lane < tot;
    int lo = 0;
#pragma unroll
    for (int step = 16; step > 0; step >>= 1) {
      const int pv = __shfl_sync(FULL, incl, lo + step - 1);
      if (pv <= lane) lo += step;
    }
    lo = min(lo, 31);
    const int o_incl = __shfl_sync(FULL, incl, lo);
    const int o_len = __shfl_sync(FULL, xl, lo);
    const long long o_ptr = __shfl_sync(FULL, xp, lo);
    const double o_m = __shfl_sync(FULL, m, lo);
    int key = -1 - lane;
    double c = 0.0;
    if (valid) {
      const long long e = o_ptr + (lane - (o_incl - o_len));
      key = X.key[e];
      c = __dmul_rn(o_m, X.val[e]);
    }
    const unsigned kgrp = __match_any_sync(FULL, key);
    const bool leader = valid && (__ffs(kgrp) - 1) == lane;
    const int maxcnt = __reduce_max_sync(FULL, valid ? __popc(kgrp) : 0);
    unsigned rem = kgrp & ~(1u << lane);  // only meaningful on leaders: the other members, ascending
    double acc = c;
    for (int r = 1; r < maxcnt; ++r) {
      const int src = rem ? (__ffs(rem) - 1) : lane;
      const double v = __shfl_sync(FULL, c, src);
      if (leader && rem) acc = __dadd_rn(acc, v);
      rem &= rem - 1;
    }
    const unsigned lead = __ballot_sync(FULL, leader);
    if (leader) {
      const int pos = __popc(lead & lt_mask);
      z_key[zp + pos] = key;
      z_val[zp + pos] = acc;
    }
    if (lane == 0) z_len[i] = __popc(lead);
  }
}

template <int CAP, int WARPS>
__global__ void __launch_bounds__(WARPS * 32)
k_spgemm_fill_fo(int row_begin, int row_end, const int* __restrict__ s_ptr, const int* __restrict__ s_col,
                 const double* __restrict__ s_val, RowLists X, const long long* __restrict__ z_ptr,
                 int* __restrict__ z_len, int* __restrict__ z_key, double* __restrict__ z_val, int* __restrict__ flags,
                 int* __restrict__ big_list) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  double* lvals = reinterpret_cast<double*>(smem_raw) + (size_t)warp * CAP;
  int* tkeys = reinterpret_cast<int*>(smem_raw + (size_t)WARPS * CAP * sizeof(double)) + (size_t)warp * CAP;
  unsigned short* tidx = reinterpret_cast<unsigned short*>(smem_raw + (size_t)WARPS * CAP * (sizeof(double) + sizeof(int))) +
                         (size_t)warp * CAP;
  const unsigned FULL = 0xffffffffu;
  const unsigned lt_mask = (1u << lane) - 1u;

  for (int i = row_begin + blockIdx.x * WARPS + warp; i < row_end; i += gridDim.x * WARPS) {
    const long long zp = z_ptr[i];
    const long long ub = z_ptr[i + 1] - zp;
    if (ub == 0) {
      if (lane == 0) z_len[i] = 0;
      continue;
    }
    if (ub <= TINY_MAX && s_ptr[i + 1] - s_ptr[i] <= TINY_MAX) continue;  // done by k_spgemm_tiny_fo
    int cap = 64;
    while (cap < CAP && (long long)cap * 4 < 5 * ub) cap <<= 1;  // load factor <= 0.8 even if every key is distinct
    const int capm = cap - 1;
    const int limit = min((long long)(CAP - (CAP >> 3)), ub);
    for (int t = lane; t < cap; t += 32) tkeys[t] = -1;
    __syncwarp();
    int uniq = 0;
    bool overflow = false;
    const int s0 = s_ptr[i], s1 = s_ptr[i + 1];
    for (int base = s0; base < s1 && !overflow; base += 32) {
      int q = -1, xl = 0;
      double m = 0.0;
      long long xp = 0;
      if (base + lane < s1) {
        q = s_col[base + lane];
        m = s_val[base + lane];
        xl = X.len[q];
        xp = X.ptr[q];
      }
      int incl = xl;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        int v = __shfl_up_sync(FULL, incl, d);
        if (lane >= d) incl += v;
      }
      const int tot = __shfl_sync(FULL, incl, 31);
      for (int f0 = 0; f0 < tot; f0 += 32) {
        const int f = f0 + lane;
        const bool valid = f < tot;
        int lo = 0;
#pragma unroll
        for (int step = 16; step > 0; step >>= 1) {
          int pv = __shfl_sync(FULL, incl, lo + step - 1);
          if (pv <= f) lo += step;
        }
        lo = min(lo, 31);
        const int o_incl = __shfl_sync(FULL, incl, lo);
        const int o_len = __shfl_sync(FULL, xl, lo);
        const long long o_ptr = __shfl_sync(FULL, xp, lo);
        const double o_m = __shfl_sync(FULL, m, lo);
        int key = -1 - lane;
        double c = 0.0;
        if (valid) {
          const long long e = o_ptr + (f - (o_incl - o_len));
          key = X.key[e];
          c = __dmul_rn(o_m, X.val[e]);
        }
        const unsigned kgrp = __match_any_sync(FULL, key);
        const int leader = __ffs(kgrp) - 1;
        bool is_new = false;
        int slot = 0, idx = 0;
        if (valid && lane == leader) {
          int s = hash_key(key) & capm;
          while (true) {
            const int prev = atomicCAS(&tkeys[s], -1, key);
            if (prev == -1) {
              is_new = true;
              break;
            }
            if (prev == key) break;
            s = (s + 1) & capm;
          }
          slot = s;
        }
        const unsigned newmask = __ballot_sync(FULL, is_new);
        if (uniq + __popc(newmask) > limit) {
          overflow = true;
          break;
        }
        if (valid && lane == leader) {
          if (is_new) {
            idx = uniq + __popc(newmask & lt_mask);
            tidx[slot] = (unsigned short)idx;
            lvals[idx] = 0.0;
            z_key[zp + idx] = key;
          } else {
            idx = tidx[slot];
          }
        }
        uniq += __popc(newmask);
        __syncwarp();
        const int myidx = __shfl_sync(FULL, idx, leader);
        const int rank = __popc(kgrp & lt_mask);
        const int maxcnt = __reduce_max_sync(FULL, valid ? __popc(kgrp) : 0);
        for (int r = 0; r < maxcnt; ++r) {
          if (valid && rank == r) lvals[myidx] = __dadd_rn(lvals[myidx], c);
          __syncwarp();
        }
      }
    }
    if (overflow) {
      if (lane == 0) {
        big_list[atomicAdd(&flags[0], 1)] = i;
        z_len[i] = 0;
      }
      __syncwarp();
      continue;
    }
    if (lane == 0) z_len[i] = uniq;
    for (int t = lane; t < uniq; t += 32) z_val[zp + t] = lvals[t];
    __syncwarp();
  }
}

// Rows with more distinct keys than the shared-memory table holds (cells next to a "hub" cell that many archetypes
// put weight on - a by-product of the reference's argmin-of-zeros rule): one CTA per row, dense accumulator over the
// whole key space in global scratch.  Keys of one X row are distinct, X rows are applied one after the other, so the
// accumulation order is again fixed; the output comes out sorted by key without a sort.
constexpr int BIG_THREADS = 256;
__global__ void __launch_bounds__(BIG_THREADS)
k_spgemm_big(int nbig, const int* __restrict__ big_list, int key_space, const int* __restrict__ s_ptr,
             const int* __restrict__ s_col, const double* __restrict__ s_val, RowLists X,
             const long long* __restrict__ z_ptr, int* __restrict__ z_len, int* __restrict__ z_key,
             double* __restrict__ z_val, double* __restrict__ scratch) {
  __shared__ int wtot[BIG_THREADS / 32];
  __shared__ int s_running;
  double* acc = scratch + (size_t)blockIdx.x * key_space;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int w = blockIdx.x; w < nbig; w += gridDim.x) {
    const int i = big_list[w];
    for (int t = tid; t < key_space; t += BIG_THREADS) acc[t] = 0.0;
    __syncthreads();
    for (int p = s_ptr[i]; p < s_ptr[i + 1]; ++p) {
      const int q = s_col[p];
      const double m = s_val[p];
      const long long xp = X.ptr[q];
      const int xl = X.len[q];
      for (int x = tid; x < xl; x += BIG_THREADS) {
        const int key = X.key[xp + x];
        acc[key] = __dadd_rn(acc[key], __dmul_rn(m, X.val[xp + x]));
      }
      __syncthreads();
    }
    if (tid == 0) s_running = 0;
    __syncthreads();
    const long long zp = z_ptr[i];
    for (int base = 0; base < key_space; base += BIG_THREADS) {
      const int t = base + tid;
      const double v = (t < key_space) ? acc[t] : 0.0;
      const bool occ = v != 0.0;
      const unsigned ball = __ballot_sync(0xffffffffu, occ);
      if (lane == 0) wtot[warp] = __popc(ball);
      __syncthreads();
      int before = s_running;
      for (int ww = 0; ww < warp; ++ww) before += wtot[ww];
      if (occ) {
        const int pos = before + __popc(ball & ((1u << lane) - 1u));
        z_key[zp + pos] = t;
        z_val[zp + pos] = v;
      }
      __syncthreads();
      if (tid == 0) {
        int tot = 0;
        for (int ww = 0; ww < BIG_THREADS / 32; ++ww) tot += wtot[ww];
        s_running += tot;
      }
      __syncthreads();
    }
    if (tid == 0) z_len[i] = s_running;
    __syncthreads();
  }
}

// First-occurrence callers (order inside a row is irrelevant to them) get a cheaper big-row kernel: one CTA per row, an
// open-addressing table of HCAP slots in shared memory instead of a dense accumulator over the key space (K = M M^T has
// key space n = 10^6 and ~10 % of its rows exceed the per-warp table: the dense kernel spent 1.2 s zeroing and scanning
// 8 MB per row).  The rows of X are applied one after the other (a key occurs once per row of X, so the adds of one pass
// hit distinct slots): per key the accumulation order is again ascending position in the row of S => fixed values.
// Which slot a key lands in depends on the race between inserting threads, so the ORDER of the output row is not
// reproducible; rows that do not fit (more than 7/8 HCAP distinct keys) are handed on to the dense kernel.
constexpr int HCAP = 8192;
__global__ void __launch_bounds__(BIG_THREADS)
k_spgemm_big_hash(int nbig, const int* __restrict__ big_list, const int* __restrict__ s_ptr,
                  const int* __restrict__ s_col, const double* __restrict__ s_val, RowLists X,
                  const long long* __restrict__ z_ptr, int* __restrict__ z_len, int* __restrict__ z_key,
                  double* __restrict__ z_val, int* __restrict__ flags, int* __restrict__ dense_list) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* tvals = reinterpret_cast<double*>(smem_raw);
  int* tkeys = reinterpret_cast<int*>(smem_raw + (size_t)HCAP * sizeof(double));
  __shared__ int s_count, s_over, s_out;
  const int tid = threadIdx.x;
  constexpr int LIMIT = HCAP - (HCAP >> 3);
  for (int w = blockIdx.x; w < nbig; w += gridDim.x) {
    const int i = big_list[w];
    for (int t = tid; t < HCAP; t += BIG_THREADS) {
      tkeys[t] = -1;
      tvals[t] = 0.0;
    }
    if (tid == 0) {
      s_count = 0;
      s_over = 0;
      s_out = 0;
    }
    __syncthreads();
    for (int p = s_ptr[i]; p < s_ptr[i + 1]; ++p) {
      const int q = s_col[p];
      const double m = s_val[p];
      const long long xp = X.ptr[q];
      const int xl = X.len[q];
      for (int x = tid; x < xl; x += BIG_THREADS) {
        const int key = X.key[xp + x];
        const double c = __dmul_rn(m, X.val[xp + x]);
        int sl = hash_key(key) & (HCAP - 1);
        bool placed = false;
        for (int probe = 0; probe < HCAP; ++probe) {
          const int prev = atomicCAS(&tkeys[sl], -1, key);
          if (prev == -1) {
            if (atomicAdd(&s_count, 1) >= LIMIT) s_over = 1;
            placed = true;
            break;
          }
          if (prev == key) {
            placed = true;
            break;
          }
          sl = (sl + 1) & (HCAP - 1);
        }
        if (placed) tvals[sl] = __dadd_rn(tvals[sl], c);
        else s_over = 1;
      }
      __syncthreads();
      if (s_over) break;
    }
    __syncthreads();
    if (s_over) {
      if (tid == 0) {
        dense_list[atomicAdd(&flags[1], 1)] = i;
        z_len[i] = 0;
      }
      __syncthreads();
      continue;
    }
    const long long zp = z_ptr[i];
    for (int t = tid; t < HCAP; t += BIG_THREADS) {
      const int key = tkeys[t];
      if (key != -1) {
        const int pos = atomicAdd(&s_out, 1);
        z_key[zp + pos] = key;
        z_val[zp + pos] = tvals[t];
      }
    }
    __syncthreads();
    if (tid == 0) z_len[i] = s_out;
    __syncthreads();
  }
}

int sc_spgemm_rows(sc_ctx* ctx, const CsrView& S, int row_begin, int row_end, const RowLists& X, int key_space,
                   RowListsBuf& Z, SpgemmScratch& sc, bool sorted, bool wide) {
  constexpr int CAP = 512, WARPS = 8;
  const int n = S.n;
  SC_TRY(sc.ub.ensure(ctx, (size_t)n + 1));
  SC_TRY(Z.ptr.ensure(ctx, (size_t)n + 1));
  SC_TRY(Z.len.ensure(ctx, (size_t)n));
  SC_TRY(sc.flags.ensure(ctx, 4));
  SC_TRY(sc.big_list.ensure(ctx, (size_t)n));
  Z.rows = n;
  SC_CUDA(ctx, cudaMemsetAsync(Z.len.p, 0, sizeof(int) * (size_t)n, ctx->stream));
  SC_CUDA(ctx, cudaMemsetAsync(sc.flags.p, 0, sizeof(int) * 4, ctx->stream));
  SC_LAUNCH(ctx, k_spgemm_ub, (n + 1 + 255) / 256, 256, 0, row_begin, row_end, S.indptr, S.indices, X.len, sc.ub.p, n);
  SC_TRY(sc_exclusive_scan_i64(ctx, sc.ub.p, Z.ptr.p, (size_t)n + 1, sc.tmp));
  long long total = 0;
  SC_CUDA(ctx, cudaMemcpyAsync(&total, Z.ptr.p + n, sizeof(long long), cudaMemcpyDeviceToHost, ctx->stream));
  SC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  SC_TRY(Z.key.ensure(ctx, (size_t)total + 1));
  SC_TRY(Z.val.ensure(ctx, (size_t)total + 1));
  const int rows = row_end - row_begin;
  if (rows <= 0) return SC_OK;
  const size_t smem = (size_t)WARPS * CAP * (sizeof(double) + sizeof(int));
  // function attributes are per device: set them on every call (a process may drive several GPUs)
  SC_CUDA(ctx, cudaFuncSetAttribute(k_spgemm_fill<CAP, WARPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int grid = std::min((rows + WARPS - 1) / WARPS, 148 * 16);
  if (sorted) {
    SC_LAUNCH(ctx, (k_spgemm_fill<CAP, WARPS>), grid, WARPS * 32, smem, row_begin, row_end, S.indptr, S.indices, S.data,
              X, Z.ptr.p, Z.len.p, Z.key.p, Z.val.p, sc.flags.p, sc.big_list.p);
  } else {
    const size_t smem_fo = (size_t)WARPS * CAP * (sizeof(double) + sizeof(int) + sizeof(unsigned short));
    SC_CUDA(ctx, cudaFuncSetAttribute(k_spgemm_fill_fo<CAP, WARPS>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)smem_fo));
    SC_LAUNCH(ctx, k_spgemm_tiny_fo, std::min((rows + 7) / 8, 148 * 32), 256, 0, row_begin, row_end, S.indptr, S.indices,
              S.data, X, Z.ptr.p, Z.len.p, Z.key.p, Z.val.p);
    if (wide) {
      constexpr int WCAP = 1024;
      const size_t smem_w = (size_t)WARPS * WCAP * (sizeof(double) + sizeof(int) + sizeof(unsigned short));
      SC_CUDA(ctx, cudaFuncSetAttribute(k_spgemm_fill_fo<WCAP, WARPS>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        (int)smem_w));
      SC_LAUNCH(ctx, (k_spgemm_fill_fo<WCAP, WARPS>), grid, WARPS * 32, smem_w, row_begin, row_end, S.indptr, S.indices,
                S.data, X, Z.ptr.p, Z.len.p, Z.key.p, Z.val.p, sc.flags.p, sc.big_list.p);
    } else {
      SC_LAUNCH(ctx, (k_spgemm_fill_fo<CAP, WARPS>), grid, WARPS * 32, smem_fo, row_begin, row_end, S.indptr, S.indices,
                S.data, X, Z.ptr.p, Z.len.p, Z.key.p, Z.val.p, sc.flags.p, sc.big_list.p);
    }
  }
  int nbig = 0;
  SC_CUDA(ctx, cudaMemcpyAsync(&nbig, sc.flags.p, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  SC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  sc.last_big_rows = nbig;
  sc.total_big_rows += nbig;
  const int* dense_rows = sc.big_list.p;
  if (nbig > 0 && !sorted) {
    SC_TRY(sc.dense_list.ensure(ctx, (size_t)n));
    const size_t hsmem = (size_t)HCAP * (sizeof(double) + sizeof(int));
    SC_CUDA(ctx, cudaFuncSetAttribute(k_spgemm_big_hash, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)hsmem));
    SC_LAUNCH(ctx, k_spgemm_big_hash, std::min(nbig, 148 * 2), BIG_THREADS, hsmem, nbig, sc.big_list.p, S.indptr, S.indices,
              S.data, X, Z.ptr.p, Z.len.p, Z.key.p, Z.val.p, sc.flags.p, sc.dense_list.p);
    SC_CUDA(ctx, cudaMemcpyAsync(&nbig, sc.flags.p + 1, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    SC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    dense_rows = sc.dense_list.p;
  }
  if (nbig > 0) {
    const int bgrid = std::min(nbig, 148 * 2);
    SC_TRY(sc.big_acc.ensure(ctx, (size_t)bgrid * key_space));
    SC_LAUNCH(ctx, k_spgemm_big, bgrid, BIG_THREADS, 0, nbig, dense_rows, key_space, S.indptr, S.indices, S.data, X,
              Z.ptr.p, Z.len.p, Z.key.p, Z.val.p, sc.big_acc.p);
  }
  return SC_OK;
}
