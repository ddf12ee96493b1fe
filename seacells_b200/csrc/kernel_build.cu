// Kernel construction: SEACellGraph.rbf (build_graph.py:113-189) on the GPU.
//
//   knn      exact brute-force kNN replacing scanpy.pp.neighbors (build_graph.py:125): d2(i,j) = sum_c (x_ic-x_jc)^2 with
//            the inputs widened to fp64, accumulated sequentially over c with separately rounded multiply and add,
//            neighbours ordered by (d2, j), self excluded by index.  Bit-identical to oracle knn_exact by construction.
//   sigma    adaptive bandwidth: the (kn//2)-th smallest non-zero neighbour distance (build_graph.py:18-34)
//   graph    binarise + self loop + union / intersection (build_graph.py:129-164) as a sort + unique over 64-bit edges
//   values   exp(-||xi-xj||^2 / (sigma_i sigma_j)) on the edges only (build_graph.py:37-61), the squared distance
//            summed in numpy's pairwise order so the exponent's argument is bit-identical to the reference's
#include "kernel_build.cuh"

#include <math.h>
#include <stdlib.h>

constexpr int KNN_QT = 64;   // queries per CTA
constexpr int KNN_RT = 64;   // references per tile
constexpr int KNN_DC = 64;   // embedding dimensions per shared-memory chunk
constexpr int KNN_THREADS = 256;
constexpr int KNN_QPW = KNN_QT / (KNN_THREADS / 32);  // queries per warp in the selection phase

template <typename TX>
__global__ void __launch_bounds__(KNN_THREADS)
k_knn_exact(const TX* __restrict__ X, int n, int d, int kk, int row_begin, int row_end,
            const int* __restrict__ row_list, int out_base, int* __restrict__ out_idx, double* __restrict__ out_d2) {
  // queries are positions [row_begin,row_end); with a row_list they are row_list[pos] (fallback rows of the filter path)
  extern __shared__ __align__(16) unsigned char smem_raw[];
  TX* sq = reinterpret_cast<TX*>(smem_raw);                    // [KNN_DC][KNN_QT]
  TX* sr = sq + KNN_DC * KNN_QT;                                // [KNN_DC][KNN_RT]
  double* sd = reinterpret_cast<double*>(sr + KNN_DC * KNN_RT);  // [KNN_QT][KNN_RT + 1]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int tx = tid & 15, ty = tid >> 4;  // 16 x 16 threads, 4 x 4 register tile each
  const int q0 = row_begin + blockIdx.x * KNN_QT;
  const bool single_chunk = d <= KNN_DC;

  double ld[KNN_QPW];
  int li[KNN_QPW];
#pragma unroll
  for (int q = 0; q < KNN_QPW; ++q) {
    ld[q] = INFINITY;
    li[q] = 0x7fffffff;
  }
  const unsigned kmask = (kk >= 32) ? 0xffffffffu : ((1u << kk) - 1u);

  auto load_q = [&](int c0) {
    const int dc = min(KNN_DC, d - c0);
    for (int e = tid; e < KNN_QT * dc; e += KNN_THREADS) {
      const int qq = e / dc, c = e - qq * dc;
      const int pq = q0 + qq;
      const int gq = (pq < row_end) ? (row_list ? row_list[pq] : pq) : 0;
      sq[c * KNN_QT + qq] = (pq < row_end) ? X[(size_t)gq * d + c0 + c] : (TX)0;
    }
  };
  if (single_chunk) load_q(0);

  for (int r0 = 0; r0 < n; r0 += KNN_RT) {
    double acc[4][4];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int b = 0; b < 4; ++b) acc[a][b] = 0.0;
    for (int c0 = 0; c0 < d; c0 += KNN_DC) {
      const int dc = min(KNN_DC, d - c0);
      __syncthreads();
      if (!single_chunk) load_q(c0);
      for (int e = tid; e < KNN_RT * dc; e += KNN_THREADS) {
        const int rr = e / dc, c = e - rr * dc;
        const int gr = r0 + rr;
        sr[c * KNN_RT + rr] = (gr < n) ? X[(size_t)gr * d + c0 + c] : (TX)0;
      }
      __syncthreads();
      for (int c = 0; c < dc; ++c) {
        double qv[4], rv[4];
#pragma unroll
        for (int a = 0; a < 4; ++a) qv[a] = (double)sq[c * KNN_QT + ty * 4 + a];
#pragma unroll
        for (int b = 0; b < 4; ++b) rv[b] = (double)sr[c * KNN_RT + tx * 4 + b];
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
          for (int b = 0; b < 4; ++b) {
            const double df = __dsub_rn(qv[a], rv[b]);
            acc[a][b] = __dadd_rn(acc[a][b], __dmul_rn(df, df));
          }
      }
    }
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int b = 0; b < 4; ++b) sd[(ty * 4 + a) * (KNN_RT + 1) + tx * 4 + b] = acc[a][b];
    __syncthreads();
    // selection: each warp owns KNN_QPW queries; lane l < kk holds the l-th best (d2, index)
#pragma unroll
    for (int q = 0; q < KNN_QPW; ++q) {
      const int qq = warp * KNN_QPW + q;
      const int pq = q0 + qq;
      if (pq >= row_end) continue;
      const int gq = row_list ? row_list[pq] : pq;
#pragma unroll
      for (int h = 0; h < KNN_RT / 32; ++h) {
        const int rr = h * 32 + lane;
        const int gr = r0 + rr;
        const double dv = sd[qq * (KNN_RT + 1) + rr];
        const double td = __shfl_sync(0xffffffffu, ld[q], kk - 1);
        const int ti = __shfl_sync(0xffffffffu, li[q], kk - 1);
        const bool cand = gr < n && gr != gq && (dv < td || (dv == td && gr < ti));
        unsigned ball = __ballot_sync(0xffffffffu, cand);
        while (ball) {
          const int src = __ffs(ball) - 1;
          ball &= ball - 1;
          const double cd = __shfl_sync(0xffffffffu, dv, src);
          const int ci = __shfl_sync(0xffffffffu, gr, src);
          const bool less = (ld[q] < cd) || (ld[q] == cd && li[q] < ci);
          const int p = __popc(__ballot_sync(0xffffffffu, less) & kmask);
          const double ud = __shfl_up_sync(0xffffffffu, ld[q], 1);
          const int ui = __shfl_up_sync(0xffffffffu, li[q], 1);
          if (p < kk) {
            if (lane == p) {
              ld[q] = cd;
              li[q] = ci;
            } else if (lane > p) {
              ld[q] = ud;
              li[q] = ui;
            }
          }
        }
      }
    }
  }
#pragma unroll
  for (int q = 0; q < KNN_QPW; ++q) {
    const int pq = q0 + warp * KNN_QPW + q;
    if (pq < row_end && lane < kk) {
      const int gq = row_list ? row_list[pq] : pq;
      out_idx[(size_t)(gq - out_base) * kk + lane] = li[q];
      out_d2[(size_t)(gq - out_base) * kk + lane] = ld[q];
    }
  }
}

template <typename TX>
static int launch_knn(sc_ctx* ctx, const TX* X, int n, int d, int kk, int row_begin, int row_end, const int* row_list,
                      int out_base, int* idx, double* d2) {
  const size_t smem = (size_t)KNN_DC * (KNN_QT + KNN_RT) * sizeof(TX) + (size_t)KNN_QT * (KNN_RT + 1) * sizeof(double);
  SC_CUDA(ctx, cudaFuncSetAttribute(k_knn_exact<TX>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int rows = row_end - row_begin;
  if (rows <= 0) return SC_OK;
  SC_LAUNCH(ctx, k_knn_exact<TX>, (rows + KNN_QT - 1) / KNN_QT, KNN_THREADS, smem, X, n, d, kk, row_begin, row_end,
            row_list, out_base, idx, d2);
  return SC_OK;
}

// ------------------------------------------------------------------------------------------------ fp32 filter + exact re-rank
// Stage 1: approximate squared distances ||xi||^2 + ||xj||^2 - 2 xi.xj on the mean-centred embedding in fp32 (FFMA
//          tiles), fused warp-level top-32 per query.
// Stage 2: the 32 survivors are re-ranked with the exact fp64 difference-form distance (same operation sequence as
//          k_knn_exact), sorted by (d2, index), the best n_neighbors-1 kept.
// Certificate: every excluded j has approx(i,j) >= thr_i (the 32nd approximate distance) and |approx - exact| <= eps_i,
//          so exact(i,j) >= thr_i - eps_i; if the exact (n_neighbors-1)-th distance is < thr_i - eps_i no excluded point
//          can belong to (or tie with) the result.  Queries that fail the certificate are recomputed by k_knn_exact.
//          The result is therefore bit-identical to the exact kernel's.
constexpr int KF_NC = 32;

__global__ void k_col_mean(const void* __restrict__ X, int x_is_f64, int n, int d, double* __restrict__ mean) {
  __shared__ double sh[256];
  const int c = blockIdx.x;
  double acc = 0.0;
  for (int i = threadIdx.x; i < n; i += 256)
    acc += x_is_f64 ? ((const double*)X)[(size_t)i * d + c] : (double)((const float*)X)[(size_t)i * d + c];
  sh[threadIdx.x] = acc;
  __syncthreads();
  for (int st = 128; st > 0; st >>= 1) {
    if (threadIdx.x < st) sh[threadIdx.x] += sh[threadIdx.x + st];
    __syncthreads();
  }
  if (threadIdx.x == 0) mean[c] = sh[0] / (double)n;
}

__global__ void k_center(const void* __restrict__ X, int x_is_f64, int n, int d, const double* __restrict__ mean,
                         float* __restrict__ Xc, float* __restrict__ norm2, float* __restrict__ maxnorm_bits) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= n) return;
  float acc = 0.f;
  for (int c = lane; c < d; c += 32) {
    const double x = x_is_f64 ? ((const double*)X)[(size_t)warp * d + c] : (double)((const float*)X)[(size_t)warp * d + c];
    const float v = (float)(x - mean[c]);
    Xc[(size_t)warp * d + c] = v;
    acc = fmaf(v, v, acc);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (lane == 0) {
    norm2[warp] = acc;
    atomicMax(reinterpret_cast<int*>(maxnorm_bits), __float_as_int(acc));  // acc >= 0: int order == float order
  }
}

__global__ void __launch_bounds__(KNN_THREADS)
k_knn_filter(const float* __restrict__ Xc, const float* __restrict__ norm2, int n, int d, int row_begin, int row_end,
             int* __restrict__ cand_idx, float* __restrict__ cand_thr) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float* sq = reinterpret_cast<float*>(smem_raw);   // [KNN_DC][KNN_QT]
  float* sr = sq + KNN_DC * KNN_QT;                   // [KNN_DC][KNN_RT]
  float* sd = sr + KNN_DC * KNN_RT;                   // [KNN_QT][KNN_RT + 1]
  float* snr = sd + KNN_QT * (KNN_RT + 1);            // [KNN_RT] reference norms
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int tx = tid & 15, ty = tid >> 4;
  const int q0 = row_begin + blockIdx.x * KNN_QT;
  const bool single_chunk = d <= KNN_DC;
  float ld[KNN_QPW];
  int li[KNN_QPW];
#pragma unroll
  for (int q = 0; q < KNN_QPW; ++q) {
    ld[q] = INFINITY;
    li[q] = 0x7fffffff;
  }
  float qn[4];
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    const int gq = q0 + ty * 4 + a;
    qn[a] = gq < row_end ? norm2[gq] : 0.f;
  }
  auto load_q = [&](int c0) {
    const int dc = min(KNN_DC, d - c0);
    const int qq = tid & (KNN_QT - 1);
    const int gq = q0 + qq;
    const float* __restrict__ src = Xc + (size_t)min(gq, n - 1) * d + c0;
    for (int c = tid >> 6; c < dc; c += KNN_THREADS / KNN_QT) sq[c * KNN_QT + qq] = (gq < row_end) ? src[c] : 0.f;
  };
  if (single_chunk) load_q(0);
  for (int r0 = 0; r0 < n; r0 += KNN_RT) {
    float acc[4][4];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int b = 0; b < 4; ++b) acc[a][b] = 0.f;
    for (int c0 = 0; c0 < d; c0 += KNN_DC) {
      const int dc = min(KNN_DC, d - c0);
      __syncthreads();
      if (!single_chunk) load_q(c0);
      {
        const int rr = tid & (KNN_RT - 1);
        const int gr = r0 + rr;
        const float* __restrict__ src = Xc + (size_t)min(gr, n - 1) * d + c0;
        for (int c = tid >> 6; c < dc; c += KNN_THREADS / KNN_RT) sr[c * KNN_RT + rr] = (gr < n) ? src[c] : 0.f;
      }
      if (c0 == 0 && tid < KNN_RT) snr[tid] = (r0 + tid < n) ? norm2[r0 + tid] : 0.f;
      __syncthreads();
#pragma unroll 2
      for (int c = 0; c < dc; ++c) {
        const float4 qv = *reinterpret_cast<const float4*>(sq + c * KNN_QT + ty * 4);
        const float4 rv = *reinterpret_cast<const float4*>(sr + c * KNN_RT + tx * 4);
        const float qa[4] = {qv.x, qv.y, qv.z, qv.w}, ra[4] = {rv.x, rv.y, rv.z, rv.w};
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
          for (int b = 0; b < 4; ++b) acc[a][b] = fmaf(qa[a], ra[b], acc[a][b]);
      }
    }
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int b = 0; b < 4; ++b)
        sd[(ty * 4 + a) * (KNN_RT + 1) + tx * 4 + b] = fmaf(-2.f, acc[a][b], qn[a] + snr[tx * 4 + b]);
    __syncthreads();
#pragma unroll
    for (int q = 0; q < KNN_QPW; ++q) {
      const int qq = warp * KNN_QPW + q;
      const int gq = q0 + qq;
      if (gq >= row_end) continue;
#pragma unroll
      for (int h = 0; h < KNN_RT / 32; ++h) {
        const int rr = h * 32 + lane;
        const int gr = r0 + rr;
        const float dv = sd[qq * (KNN_RT + 1) + rr];
        const float td = __shfl_sync(0xffffffffu, ld[q], KF_NC - 1);
        const bool cand = gr < n && gr != gq && dv < td;
        unsigned ball = __ballot_sync(0xffffffffu, cand);
        while (ball) {
          const int src = __ffs(ball) - 1;
          ball &= ball - 1;
          const float cd = __shfl_sync(0xffffffffu, dv, src);
          const int ci = __shfl_sync(0xffffffffu, gr, src);
          const bool less = ld[q] <= cd;  // ties keep the earlier reference; any fixed rule works for a filter
          const int p = __popc(__ballot_sync(0xffffffffu, less));
          const float ud = __shfl_up_sync(0xffffffffu, ld[q], 1);
          const int ui = __shfl_up_sync(0xffffffffu, li[q], 1);
          if (p < KF_NC) {
            if (lane == p) {
              ld[q] = cd;
              li[q] = ci;
            } else if (lane > p) {
              ld[q] = ud;
              li[q] = ui;
            }
          }
        }
      }
    }
  }
#pragma unroll
  for (int q = 0; q < KNN_QPW; ++q) {
    const int gq = q0 + warp * KNN_QPW + q;
    if (gq < row_end) {
      cand_idx[(size_t)(gq - row_begin) * KF_NC + lane] = li[q];
      if (lane == KF_NC - 1) cand_thr[gq - row_begin] = ld[q];
    }
  }
}

template <typename TX>
__global__ void k_knn_rerank(const TX* __restrict__ X, const float* __restrict__ norm2, const float* __restrict__ maxnorm,
                             int n, int d, int kk, int row_begin, int row_end, const int* __restrict__ cand_idx,
                             const float* __restrict__ cand_thr, double gamma, int* __restrict__ out_idx,
                             double* __restrict__ out_d2, int* __restrict__ fail_list, int* __restrict__ fail_cnt) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  const int gq = row_begin + warp;
  if (gq >= row_end) return;
  int j = cand_idx[(size_t)warp * KF_NC + lane];
  double dd = INFINITY;
  if (j != 0x7fffffff && j != gq) {  // the tensor-core filter leaves the query itself in its shortlist
    const TX* __restrict__ xi = X + (size_t)gq * d;
    const TX* __restrict__ xj = X + (size_t)j * d;
    double acc = 0.0;
    for (int c = 0; c < d; ++c) {
      const double df = __dsub_rn((double)xi[c], (double)xj[c]);
      acc = __dadd_rn(acc, __dmul_rn(df, df));
    }
    dd = acc;
  }
  // bitonic sort of the 32 (d2, index) pairs across the warp, ascending
#pragma unroll
  for (int k2 = 2; k2 <= 32; k2 <<= 1) {
#pragma unroll
    for (int st = k2 >> 1; st > 0; st >>= 1) {
      const double od = __shfl_xor_sync(0xffffffffu, dd, st);
      const int oj = __shfl_xor_sync(0xffffffffu, j, st);
      const bool up = (lane & k2) == 0;
      const bool lower = (lane & st) == 0;
      const bool mine_less = dd < od || (dd == od && j < oj);
      const bool keep_mine = (lower == up) ? mine_less : !mine_less;
      if (!keep_mine) {
        dd = od;
        j = oj;
      }
    }
  }
  // certificate
  const double dk = __shfl_sync(0xffffffffu, dd, kk - 1);
  const double nq = (double)norm2[gq], nm = (double)maxnorm[0];
  const double eps = gamma * (nq + nm + 2.0 * sqrt(nq * nm));
  const double thr = (double)cand_thr[warp];
  const bool ok = dk < thr - eps;  // false for NaN / inf as well
  if (!ok) {
    if (lane == 0) fail_list[atomicAdd(fail_cnt, 1)] = gq;
    return;
  }
  if (lane < kk) {
    out_idx[(size_t)warp * kk + lane] = j;
    out_d2[(size_t)warp * kk + lane] = dd;
  }
}

int sc_knn_exact_dev(sc_ctx* ctx, const void* X, int x_is_f64, int n, int d, int n_neighbors, int row_begin,
                     int row_end, int* idx, double* d2) {
  const int kk = n_neighbors - 1;
  if (kk < 1 || kk > 32) SC_FAIL(ctx, SC_ERR_ARG, "knn: n_neighbors must be in [2, 33]");
  if (n <= kk) SC_FAIL(ctx, SC_ERR_ARG, "knn: need more cells than neighbours");
  if (row_begin < 0 || row_end > n || row_begin > row_end) SC_FAIL(ctx, SC_ERR_ARG, "knn: bad row range");
  const int rows = row_end - row_begin;
  if (rows == 0) return SC_OK;
  const bool use_filter = n >= 2048 && kk <= 24 && n > KF_NC + 1 && getenv("SC_KNN_EXACT") == nullptr;
  ctx->last_knn_filter_ms = 0.f;
  ctx->last_knn_filter_kind = 0;
  if (!use_filter) {
    if (x_is_f64) return launch_knn<double>(ctx, (const double*)X, n, d, kk, row_begin, row_end, nullptr, row_begin, idx, d2);
    return launch_knn<float>(ctx, (const float*)X, n, d, kk, row_begin, row_end, nullptr, row_begin, idx, d2);
  }
  DBuf<double> mean;
  DBuf<float> Xc, norm2, maxnorm, thr;
  DBuf<int> cand, fail;
  SC_TRY(mean.ensure(ctx, (size_t)d));
  SC_TRY(Xc.ensure(ctx, (size_t)n * d));
  SC_TRY(norm2.ensure(ctx, (size_t)n));
  SC_TRY(maxnorm.ensure(ctx, 4));
  SC_TRY(thr.ensure(ctx, (size_t)rows));
  SC_TRY(cand.ensure(ctx, (size_t)rows * KF_NC));
  SC_TRY(fail.ensure(ctx, (size_t)rows + 4));
  SC_CUDA(ctx, cudaMemsetAsync(maxnorm.p, 0, sizeof(float) * 4, ctx->stream));
  SC_CUDA(ctx, cudaMemsetAsync(fail.p + rows, 0, sizeof(int) * 4, ctx->stream));
  SC_LAUNCH(ctx, k_col_mean, d, 256, 0, X, x_is_f64, n, d, mean.p);
  SC_LAUNCH(ctx, k_center, (unsigned)(((size_t)n * 32 + 255) / 256), 256, 0, X, x_is_f64, n, d, mean.p, Xc.p, norm2.p,
            maxnorm.p);
  double gamma;
  if (sc_knn_tc_supported(n, d) && getenv("SC_KNN_FFMA") == nullptr) {
    // tensor-core filter (knn_tc.cu): 2xFP16 split product (22-bit operands), fp32 accumulation in tensor memory
    SC_TRY(sc_knn_filter_tc(ctx, Xc.p, norm2.p, maxnorm.p, n, d, row_begin, row_end, cand.p, thr.p, nullptr, 0));
    gamma = sc_knn_tc_gamma(d);
  } else {
    const size_t smem = (size_t)KNN_DC * (KNN_QT + KNN_RT) * sizeof(float) +
                        (size_t)KNN_QT * (KNN_RT + 1) * sizeof(float) + KNN_RT * sizeof(float);
    SC_CUDA(ctx, cudaFuncSetAttribute(k_knn_filter, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    SC_CUDA(ctx, cudaEventCreate(&e0));
    SC_CUDA(ctx, cudaEventCreate(&e1));
    SC_CUDA(ctx, cudaEventRecord(e0, ctx->stream));
    SC_LAUNCH(ctx, k_knn_filter, (rows + KNN_QT - 1) / KNN_QT, KNN_THREADS, smem, Xc.p, norm2.p, n, d, row_begin,
              row_end, cand.p, thr.p);
    SC_CUDA(ctx, cudaEventRecord(e1, ctx->stream));
    SC_CUDA(ctx, cudaEventSynchronize(e1));
    cudaEventElapsedTime(&ctx->last_knn_filter_ms, e0, e1);
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    ctx->last_knn_filter_kind = 1;
    // error bound of the fp32 pipeline (conversion, norms, d-term dot product, final combination), 4x conservative
    gamma = (double)(d + 8) * 4.0 * 5.9604644775390625e-08;
  }
  if (x_is_f64) {
    SC_LAUNCH(ctx, k_knn_rerank<double>, (unsigned)(((size_t)rows * 32 + 255) / 256), 256, 0, (const double*)X, norm2.p,
              maxnorm.p, n, d, kk, row_begin, row_end, cand.p, thr.p, gamma, idx, d2, fail.p, fail.p + rows);
  } else {
    SC_LAUNCH(ctx, k_knn_rerank<float>, (unsigned)(((size_t)rows * 32 + 255) / 256), 256, 0, (const float*)X, norm2.p,
              maxnorm.p, n, d, kk, row_begin, row_end, cand.p, thr.p, gamma, idx, d2, fail.p, fail.p + rows);
  }
  int nfail = 0;
  SC_CUDA(ctx, cudaMemcpyAsync(&nfail, fail.p + rows, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  SC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  ctx->last_knn_fallback = nfail;
  if (nfail > 0) {
    if (x_is_f64) SC_TRY(launch_knn<double>(ctx, (const double*)X, n, d, kk, 0, nfail, fail.p, row_begin, idx, d2));
    else SC_TRY(launch_knn<float>(ctx, (const float*)X, n, d, kk, 0, nfail, fail.p, row_begin, idx, d2));
    SC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  }
  return SC_OK;
}

// Test probe: centre X, run the tensor-core filter on the first 128 queries and return the raw s(i,j) tile together
// with the centred fp32 embedding and the fp32 norms it was computed from (all device buffers).
int sc_knn_tc_probe_dev(sc_ctx* ctx, const void* X, int x_is_f64, int n, int d, int cols, float* s_out, float* xc_out,
                        float* norm2_out) {
  if (!sc_knn_tc_supported(n, d)) SC_FAIL(ctx, SC_ERR_ARG, "knn tensor-core probe: unsupported shape");
  DBuf<double> mean;
  DBuf<float> maxnorm, thr;
  DBuf<int> cand;
  const int rows = n < 128 ? n : 128;
  SC_TRY(mean.ensure(ctx, (size_t)d));
  SC_TRY(maxnorm.ensure(ctx, 4));
  SC_TRY(thr.ensure(ctx, (size_t)rows));
  SC_TRY(cand.ensure(ctx, (size_t)rows * KF_NC));
  SC_CUDA(ctx, cudaMemsetAsync(maxnorm.p, 0, sizeof(float) * 4, ctx->stream));
  SC_LAUNCH(ctx, k_col_mean, d, 256, 0, X, x_is_f64, n, d, mean.p);
  SC_LAUNCH(ctx, k_center, (unsigned)(((size_t)n * 32 + 255) / 256), 256, 0, X, x_is_f64, n, d, mean.p, xc_out,
            norm2_out, maxnorm.p);
  SC_TRY(sc_knn_filter_tc(ctx, xc_out, norm2_out, maxnorm.p, n, d, 0, rows, cand.p, thr.p, s_out, cols));
  return SC_OK;
}

// ---- sigma_i: (kn//2)-th smallest non-zero Euclidean neighbour distance, 0 when there are fewer (np.linalg.norm([]))
__global__ void k_sigma(int n, int kk, int med, const double* __restrict__ d2, double* __restrict__ sigma) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int seen = 0;
  double s = 0.0;
  for (int l = 0; l < kk; ++l) {
    const double v = d2[(size_t)i * kk + l];
    if (v != 0.0) {
      seen++;
      if (seen == med) {
        s = sqrt(v);
        break;
      }
    }
  }
  sigma[i] = s;
}

// ---- edges: (i,i), (i,j), (j,i) for every stored (non-zero distance) neighbour j of i
__global__ void k_emit_edges(int n, int kk, const int* __restrict__ idx, const double* __restrict__ d2,
                             unsigned long long* __restrict__ keys) {
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long per = 1 + 2 * (long long)kk;
  if (e >= (long long)n * per) return;
  const int i = (int)(e / per);
  const int r = (int)(e - (long long)i * per);
  unsigned long long key;
  if (r == 0) {
    key = ((unsigned long long)(unsigned)i << 32) | (unsigned)i;
  } else {
    const int l = (r - 1) >> 1;
    const int j = idx[(size_t)i * kk + l];
    if (d2[(size_t)i * kk + l] == 0.0) key = ~0ull;  // scanpy's graph has explicit zeros eliminated
    else key = ((r - 1) & 1) ? (((unsigned long long)(unsigned)j << 32) | (unsigned)i)
                             : (((unsigned long long)(unsigned)i << 32) | (unsigned)j);
  }
  keys[e] = key;
}

__global__ void k_edge_flags(long long N, const unsigned long long* __restrict__ keys, int intersection,
                             int* __restrict__ flag) {
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e > N) return;
  if (e == N) {
    flag[e] = 0;
    return;
  }
  const unsigned long long kx = keys[e];
  int f = 0;
  if (kx != ~0ull) {
    const bool first = (e == 0) || keys[e - 1] != kx;
    if (!intersection) f = first;
    else {
      const bool diag = (unsigned)(kx >> 32) == (unsigned)(kx & 0xffffffffu);
      f = first && (diag || (e + 1 < N && keys[e + 1] == kx));
    }
  }
  flag[e] = f;
}

__global__ void k_edge_compact(long long N, const unsigned long long* __restrict__ keys, const int* __restrict__ flag,
                               const long long* __restrict__ pos, unsigned long long* __restrict__ out) {
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= N) return;
  if (flag[e]) out[pos[e]] = keys[e];
}

__global__ void k_edge_rowptr(int n, long long nnz, const unsigned long long* __restrict__ ekeys, int* __restrict__ indptr) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i > n) return;
  const unsigned long long target = (unsigned long long)(unsigned)i << 32;
  long long lo = 0, hi = nnz;
  while (lo < hi) {
    long long mid = (lo + hi) >> 1;
    if (ekeys[mid] < target) lo = mid + 1;
    else hi = mid;
  }
  indptr[i] = (int)lo;
}

// numpy's pairwise summation (np.sum over a contiguous axis of length len), applied to (xi - xj)^2
template <typename TX>
__device__ double pairwise_sqdist(const TX* __restrict__ xi, const TX* __restrict__ xj, int len) {
  auto term = [&](int c) {
    const double df = __dsub_rn((double)xi[c], (double)xj[c]);
    return __dmul_rn(df, df);
  };
  if (len < 8) {
    double res = 0.0;
    for (int c = 0; c < len; ++c) res = __dadd_rn(res, term(c));
    return res;
  }
  if (len <= 128) {
    double r[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) r[q] = term(q);
    int c = 8;
    for (; c < len - (len % 8); c += 8) {
#pragma unroll
      for (int q = 0; q < 8; ++q) r[q] = __dadd_rn(r[q], term(c + q));
    }
    double res = __dadd_rn(__dadd_rn(__dadd_rn(r[0], r[1]), __dadd_rn(r[2], r[3])),
                           __dadd_rn(__dadd_rn(r[4], r[5]), __dadd_rn(r[6], r[7])));
    for (; c < len; ++c) res = __dadd_rn(res, term(c));
    return res;
  }
  int n2 = len / 2;
  n2 -= n2 % 8;
  return __dadd_rn(pairwise_sqdist(xi, xj, n2), pairwise_sqdist(xi + n2, xj + n2, len - n2));
}

template <typename TX>
__global__ void k_edge_values(long long nnz, int d, const TX* __restrict__ X, const double* __restrict__ sigma,
                              const unsigned long long* __restrict__ ekeys, int* __restrict__ cols,
                              double* __restrict__ vals) {
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= nnz) return;
  const int i = (int)(ekeys[e] >> 32), j = (int)(ekeys[e] & 0xffffffffu);
  const double num = pairwise_sqdist(X + (size_t)i * d, X + (size_t)j * d, d);
  const double den = __dmul_rn(sigma[i], sigma[j]);
  cols[e] = j;
  vals[e] = exp(__ddiv_rn(-num, den));
}

__global__ void k_nonzero_flags(long long nnz, const double* __restrict__ vals, int* __restrict__ flag) {
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e > nnz) return;
  flag[e] = (e < nnz && vals[e] != 0.0) ? 1 : 0;
}
__global__ void k_nonzero_compact(long long nnz, const int* __restrict__ flag, const long long* __restrict__ pos,
                                  const unsigned long long* __restrict__ k_in, const int* __restrict__ c_in,
                                  const double* __restrict__ v_in, unsigned long long* __restrict__ k_out,
                                  int* __restrict__ c_out, double* __restrict__ v_out) {
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= nnz || !flag[e]) return;
  k_out[pos[e]] = k_in[e];
  c_out[pos[e]] = c_in[e];
  v_out[pos[e]] = v_in[e];
}

int sc_kernel_graph::build(sc_ctx* ctx, const void* X_dev, int x_is_f64, int n_, int d, int n_neighbors,
                           const int* idx, const double* d2, int intersection) {
  n = n_;
  const int kk = n_neighbors - 1;
  const int med = n_neighbors / 2;
  SC_TRY(sigma.ensure(ctx, (size_t)n));
  SC_LAUNCH(ctx, k_sigma, (n + 255) / 256, 256, 0, n, kk, med, d2, sigma.p);
  const long long N = (long long)n * (1 + 2 * kk);
  DBuf<unsigned long long> k_in, k_out;
  DBuf<int> flag;
  DBuf<long long> pos;
  DBuf<char> tmp;
  SC_TRY(k_in.ensure(ctx, (size_t)N));
  SC_TRY(k_out.ensure(ctx, (size_t)N));
  SC_TRY(flag.ensure(ctx, (size_t)N + 1));
  SC_TRY(pos.ensure(ctx, (size_t)N + 1));
  SC_LAUNCH(ctx, k_emit_edges, (unsigned)((N + 255) / 256), 256, 0, n, kk, idx, d2, k_in.p);
  SC_TRY(sc_sort_keys_u64(ctx, k_in.p, k_out.p, (size_t)N, 64, tmp));
  SC_LAUNCH(ctx, k_edge_flags, (unsigned)((N + 256) / 256), 256, 0, N, k_out.p, intersection, flag.p);
  SC_TRY(sc_exclusive_scan_i32_to_i64(ctx, flag.p, pos.p, (size_t)N + 1, tmp));
  long long total = 0;
  SC_CUDA(ctx, cudaMemcpyAsync(&total, pos.p + N, sizeof(long long), cudaMemcpyDeviceToHost, ctx->stream));
  SC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  nnz = total;
  if (nnz >= (1ll << 31)) SC_FAIL(ctx, SC_ERR_CAPACITY, "kernel matrix has more than 2^31 entries");
  SC_TRY(ekeys.ensure(ctx, (size_t)nnz + 1));
  SC_TRY(indptr.ensure(ctx, (size_t)n + 1));
  SC_TRY(indices.ensure(ctx, (size_t)nnz + 1));
  SC_TRY(data.ensure(ctx, (size_t)nnz + 1));
  SC_LAUNCH(ctx, k_edge_compact, (unsigned)((N + 255) / 256), 256, 0, N, k_out.p, flag.p, pos.p, ekeys.p);
  SC_LAUNCH(ctx, k_edge_rowptr, (n + 256) / 256, 256, 0, n, nnz, ekeys.p, indptr.p);
  if (x_is_f64) {
    SC_LAUNCH(ctx, k_edge_values<double>, (unsigned)((nnz + 127) / 128), 128, 0, nnz, d, (const double*)X_dev, sigma.p,
              ekeys.p, indices.p, data.p);
  } else {
    SC_LAUNCH(ctx, k_edge_values<float>, (unsigned)((nnz + 127) / 128), 128, 0, nnz, d, (const float*)X_dev, sigma.p,
              ekeys.p, indices.p, data.p);
  }
  // entries that evaluate to exactly 0 are not stored by the reference (lil_matrix(masked_row), build_graph.py:61)
  SC_LAUNCH(ctx, k_nonzero_flags, (unsigned)((nnz + 256) / 256), 256, 0, nnz, data.p, flag.p);
  SC_TRY(sc_exclusive_scan_i32_to_i64(ctx, flag.p, pos.p, (size_t)nnz + 1, tmp));
  long long kept = 0;
  SC_CUDA(ctx, cudaMemcpyAsync(&kept, pos.p + nnz, sizeof(long long), cudaMemcpyDeviceToHost, ctx->stream));
  SC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  if (kept != nnz) {
    DBuf<unsigned long long> k2;
    DBuf<int> c2;
    DBuf<double> v2;
    SC_TRY(k2.ensure(ctx, (size_t)kept + 1));
    SC_TRY(c2.ensure(ctx, (size_t)kept + 1));
    SC_TRY(v2.ensure(ctx, (size_t)kept + 1));
    SC_LAUNCH(ctx, k_nonzero_compact, (unsigned)((nnz + 255) / 256), 256, 0, nnz, flag.p, pos.p, ekeys.p, indices.p,
              data.p, k2.p, c2.p, v2.p);
    SC_CUDA(ctx, cudaMemcpyAsync(ekeys.p, k2.p, sizeof(unsigned long long) * kept, cudaMemcpyDeviceToDevice, ctx->stream));
    SC_CUDA(ctx, cudaMemcpyAsync(indices.p, c2.p, sizeof(int) * kept, cudaMemcpyDeviceToDevice, ctx->stream));
    SC_CUDA(ctx, cudaMemcpyAsync(data.p, v2.p, sizeof(double) * kept, cudaMemcpyDeviceToDevice, ctx->stream));
    nnz = kept;
    SC_LAUNCH(ctx, k_edge_rowptr, (n + 256) / 256, 256, 0, n, nnz, ekeys.p, indptr.p);
    SC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  }
  return SC_OK;
}
