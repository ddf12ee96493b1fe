// kNN shortlist filter on the 5th-generation tensor cores (tcgen05, sm_100a).
//
// Replaces the distance contraction of scanpy.pp.neighbors (build_graph.py:125) for the filter stage of the exact kNN
// (kernel_build.cu): for every query i and reference j it evaluates
//        s(i,j) = xi.xj - ||xj||^2 / 2          (so that  ||xi-xj||^2 = ||xi||^2 - 2 s(i,j))
// on the mean-centred fp32 embedding as a split-precision product on the tensor cores
//        A_hi B_hi + A_hi B_lo + A_lo B_hi,   x = hi + lo,  hi = fp16(x),  lo = fp16(x - hi)
// with fp32 accumulation in tensor memory, and keeps the 32 references with the largest s per query together with the
// 32nd value (the threshold of the certificate checked by k_knn_rerank).  Two fp16 halves carry 22 significand bits -
// the same operand precision as the classical 3xTF32 split (two 11-bit tf32 halves) - but kind::f16 MMAs run at twice
// the tf32 rate for the same power, and the kernel is power-bound (the 3xTF32 version of this kernel held the 1 kW cap
// at 558 ms for 1M x 50; see profiles/r1_knn_filter_tc.summary.txt).  The embedding is pre-scaled by a power of two so
// that every |x| < 256 (exact; fp16 then never overflows and its subnormal floor 2^-25 is < 2^-32 of the largest norm,
// which the error bound absorbs).  The -||xj||^2/2 term rides along as one extra K column (A holds 128 there, B holds
// -||xj||^2 scale^2 / 256), so the epilogue is a pure "is anything in this row chunk above my threshold" scan.
//
// Layout
//   references pre-tiled in global memory (k_knn_pack): one block of KP*512 bytes per 128 cells = hi part + lo part,
//              each already in the shared-memory image tcgen05.mma wants (K-major, no swizzle: 8x16-byte core
//              matrices = 8 rows x 8 halves, row groups 128 B apart (SBO), K chunks 2048 B apart (LBO));
//              KP = round_up(d+1, 16).
//   CTA        persistent, one per SM, 384 threads: warp 0 = bulk-copy producer (cp.async.bulk -> UBLKCP),
//              warp 1 = MMA issuer (one thread; tcgen05.mma kind::f16, M=128 N=128 K=16), warps 4-11 = epilogue
//              (tcgen05.ld 32x32b: a query row is shared by two threads, one per 64-column half of the tile).  The query tile (128 rows, hi and lo) lives in TENSOR
//              MEMORY (written once per query tile by the epilogue threads with tcgen05.st, A operand of the MMAs):
//              with both operands in shared memory the MMAs were bound by shared-memory bandwidth (8 KB of operand
//              reads per 64-cycle MMA next to the bulk copies: 95 cycles per MMA measured).  Reference tiles stream
//              through a ring of up to 6 stages (as many as fit: in-flight bytes hide the L2 latency); accumulators are triple buffered in TMEM (3 x 128 columns) so the MMAs of the
//              next tiles run under the scan of this one.  TMEM map: [0,384) accumulators, [384,448) A hi, [448,512) A lo.
//   selection  per row: threshold register + append buffer (128 entries, L2-resident scratch); a row whose buffer
//              passes 96 entries is compacted to its best 32 by the whole warp (bitwise radix select over votes).
//              The query itself is not excluded here (it is simply the best candidate): k_knn_rerank drops it.
//
//   order      cells are grouped by their nearest pivot (n/512 pivot cells, fp32 distances).  Reference tiles and query
//              tiles are formed in that order and a query tile visits the reference tiles outwards from its own
//              position (c, c+1, c-1, c+2, ...), so the near neighbours arrive first, the per-row thresholds are
//              almost final after a few tiles and the rest of the stream takes the one-vote fast path.
//              (In input order the first ~10 % of every stream was selection-bound: ~40 % of the kernel's time.)
//              The order only affects speed; any visiting order yields a valid shortlist and threshold.
//
// Every mbarrier wait is bounded (2 s): a broken pipeline sets an error flag and drains instead of hanging the GPU.
#include "kernel_build.cuh"

#include <cuda_fp16.h>
#include <math.h>
#include <stdlib.h>

namespace {

constexpr int TC_M = 128;          // queries per CTA tile
constexpr int TC_N = 128;          // references per streamed tile
constexpr int TC_THREADS = 384;    // producer warp, MMA warp, 2 idle, 8 epilogue warps
constexpr int TC_CAP = 128;        // append-buffer entries per row
constexpr int TC_TRIG = 96;        // compact when a row holds more than this after a chunk
constexpr int TC_KEEP = 32;        // shortlist size (== KF_NC in kernel_build.cu)
constexpr int TC_TMEM_COLS = 512;  // 3 accumulator buffers x 128 fp32 columns + the query tile (hi, lo: 64 columns each)
constexpr int TC_ACC = 3;          // accumulator buffers in tensor memory
constexpr int TC_MAX_STAGES = 6;   // reference-tile stages in shared memory: as many as fit (<= 6), chosen on the host
constexpr uint32_t TC_COL_AHI = 384, TC_COL_ALO = 448;
constexpr unsigned FULLMASK = 0xffffffffu;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t"
      "}\n"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ bool mbar_test(uint32_t bar, uint32_t parity) {  // non-blocking probe
  uint32_t ok;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t"
      "}\n"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
// bounded wait: false (and the shared abort flag set) when the barrier did not flip within ~2 s
__device__ __forceinline__ bool mbar_wait(uint32_t bar, uint32_t parity, volatile int* abort_flag) {
  if (mbar_try_wait(bar, parity)) return true;
  const long long t0 = clock64();
  while (true) {
#pragma unroll 1
    for (int spin = 0; spin < 64; ++spin)
      if (mbar_try_wait(bar, parity)) return true;
    if (*abort_flag) return false;
    if (clock64() - t0 > 4000000000ll) {
      *abort_flag = 1;
      return false;
    }
  }
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// A operand from tensor memory (lane = row, column = k), B from shared memory
__device__ __forceinline__ void tc_mma_f16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc,
                                               uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t"
      "}\n" ::"r"(d_tmem),
      "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tc_st8(uint32_t taddr, const uint32_t (&v)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(v[0]),
               "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
               : "memory");
}
__device__ __forceinline__ void tc_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tc_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// shared-memory matrix descriptor: K-major, SWIZZLE_NONE, LBO = 2048 B (next K chunk), SBO = 128 B (next 8 rows)
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
  return (uint64_t)((saddr >> 4) & 0x3fffu) | ((uint64_t)(2048u >> 4) << 16) | ((uint64_t)(128u >> 4) << 32) |
         (1ull << 46);
}
// instruction descriptor: D fp32, A/B fp16 (format 0), both K-major, N=128, M=128
constexpr uint32_t TC_IDESC = (1u << 4) | (0u << 7) | (0u << 10) | ((uint32_t)(TC_N >> 3) << 17) |
                              ((uint32_t)(TC_M >> 4) << 24);

// x -> (hi, lo) fp16 halves, bit patterns; x must already be scaled into fp16's range
__device__ __forceinline__ void split_f16(float x, unsigned short& hi, unsigned short& lo) {
  const __half h = __float2half_rn(x);
  const __half l = __float2half_rn(x - __half2float(h));
  hi = __half_as_ushort(h);
  lo = __half_as_ushort(l);
}
// power-of-two scale that brings every centred |x| (<= the largest row norm) below 256; exact, same on every thread
__device__ __forceinline__ float knn_scale(const float* __restrict__ maxnorm2) {
  const float mn = sqrtf(maxnorm2[0]);
  if (!(mn > 0.f) || !isfinite(mn)) return 1.f;
  int e;
  frexpf(mn, &e);  // mn = f * 2^e, 0.5 <= f < 1
  return ldexpf(1.f, 8 - e);
}
// i-th reference tile of the outward walk from tile c over T tiles: c, c+1, c-1, c+2, c-2, ... then the longer side
__device__ __forceinline__ int tile_seq(int i, int c, int T) {
  const int lo = c, hi = T - 1 - c, m = min(lo, hi);
  if (i == 0) return c;
  if (i <= 2 * m) {
    const int j = (i + 1) >> 1;
    return (i & 1) ? c + j : c - j;
  }
  return (hi > lo) ? c + (i - m) : c - (i - m);
}

constexpr float TC_ONE_COL = 128.f;  // value of the query tile's extra column; the references carry -||x||^2 scale^2 / 256

__device__ __forceinline__ bool elect_one() {
  uint32_t p;
  asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}\n" : "=r"(p));
  return p != 0;
}

// order-preserving map of fp32 bit patterns onto unsigned integers, and its inverse
__device__ __forceinline__ uint32_t ordered_bits(uint32_t f) { return (f & 0x80000000u) ? ~f : (f | 0x80000000u); }
__device__ __forceinline__ uint32_t unordered_bits(uint32_t u) { return (u & 0x80000000u) ? (u & 0x7fffffffu) : ~u; }

}  // namespace

// ---- locality order -----------------------------------------------------------------------------------------------
// Cells are grouped by their nearest pivot (P cells taken at a regular stride, fp32 distances on the centred
// embedding).  Reference tiles and query tiles are formed in that order, so the few tiles around a query's own position
// hold most of its near neighbours.
constexpr int PIV_CHUNK = 64;
template <int DREG>
__global__ void __launch_bounds__(256)
k_knn_nearest_pivot(const float* __restrict__ Xc, const float* __restrict__ norm2, int n, int d, int P,
                    int cell_begin, int cell_end, unsigned long long* __restrict__ keys) {
  __shared__ __align__(16) float piv[PIV_CHUNK][DREG];
  __shared__ float pn[PIV_CHUNK];
  const int i = cell_begin + blockIdx.x * blockDim.x + threadIdx.x;
  const bool live = i < cell_end;
  float x[DREG];
#pragma unroll
  for (int c = 0; c < DREG; ++c) x[c] = (live && c < d) ? Xc[(size_t)i * d + c] : 0.f;
  float best = INFINITY;
  int bp = 0;
  for (int p0 = 0; p0 < P; p0 += PIV_CHUNK) {
    __syncthreads();
    for (int t = threadIdx.x; t < PIV_CHUNK * DREG; t += blockDim.x) {
      const int pp = t / DREG, c = t - pp * DREG;
      const int p = p0 + pp;
      const long long cell = (p < P) ? ((long long)p * n) / P : 0;
      piv[pp][c] = (p < P && c < d) ? Xc[(size_t)cell * d + c] : 0.f;
      if (c == 0) pn[pp] = (p < P) ? norm2[cell] : INFINITY;
    }
    __syncthreads();
    const int cnt = min(PIV_CHUNK, P - p0);
    for (int pp = 0; pp < cnt; ++pp) {
      float dot = 0.f;
#pragma unroll
      for (int c = 0; c < DREG; c += 4) {
        const float4 pv = *reinterpret_cast<const float4*>(&piv[pp][c]);
        dot = fmaf(x[c], pv.x, dot);
        dot = fmaf(x[c + 1], pv.y, dot);
        dot = fmaf(x[c + 2], pv.z, dot);
        dot = fmaf(x[c + 3], pv.w, dot);
      }
      const float dist = fmaf(-2.f, dot, pn[pp]);  // + ||x||^2, constant per cell
      if (dist < best) {
        best = dist;
        bp = p0 + pp;
      }
    }
  }
  if (live) keys[i - cell_begin] = ((unsigned long long)(unsigned)bp << 32) | (unsigned)i;
}

static int launch_nearest_pivot(sc_ctx* ctx, const float* Xc, const float* norm2, int n, int d, int P, int cell_begin,
                                int cell_end, unsigned long long* keys) {
  const int cells = cell_end - cell_begin;
  if (cells <= 0) return SC_OK;
  const int grid = (cells + 255) / 256;
  if (d <= 32) {
    SC_LAUNCH(ctx, k_knn_nearest_pivot<32>, grid, 256, 0, Xc, norm2, n, d, P, cell_begin, cell_end, keys);
  } else if (d <= 64) {
    SC_LAUNCH(ctx, k_knn_nearest_pivot<64>, grid, 256, 0, Xc, norm2, n, d, P, cell_begin, cell_end, keys);
  } else {
    SC_LAUNCH(ctx, k_knn_nearest_pivot<128>, grid, 256, 0, Xc, norm2, n, d, P, cell_begin, cell_end, keys);
  }
  return SC_OK;
}

__global__ void k_knn_unpack_order(int count, const unsigned long long* __restrict__ keys, int* __restrict__ order,
                                   int* __restrict__ inv) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= count) return;
  const int cell = (int)(keys[p] & 0xffffffffu);
  order[p] = cell;
  if (inv) inv[cell] = p;
}

// ---- operand packing --------------------------------------------------------------------------------------------
// Reference tiles: one 16-byte unit (8 consecutive K columns of one cell, fp16) per thread, for the hi and the lo part.
// Column d holds -||x||^2 scale^2 / 256 (times the query tile's 128 = -||x||^2 scale^2 / 2); rows past n get -60000
// there (never shortlisted) and zeros elsewhere.
__global__ void k_knn_pack(const float* __restrict__ Xc, const float* __restrict__ norm2,
                           const float* __restrict__ maxnorm2, const int* __restrict__ order, int n, int d, int KP,
                           uint4* __restrict__ out) {
  const int KC = KP >> 3;
  const long long u = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long n_tiles = (n + TC_N - 1) / TC_N;
  if (u >= n_tiles * KC * 128) return;
  const float scale = knn_scale(maxnorm2);
  const long long tile = u / (KC * 128);
  const int w = (int)(u - tile * (KC * 128));
  const int kc = w >> 7, r = w & 127;
  const long long pos = tile * 128 + r;  // position in the Morton order
  const long long cell = pos < n ? order[pos] : n;
  unsigned short hi[8], lo[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    const int c = kc * 8 + e;
    float v = 0.f;
    if (cell < n) {
      if (c < d) v = Xc[cell * d + c] * scale;
      else if (c == d) v = -norm2[cell] * scale * scale * (1.f / 256.f);
    } else if (c == d) {
      v = -60000.f;
    }
    split_f16(v, hi[e], lo[e]);
  }
  uint4* tb = out + tile * (size_t)(2 * KC * 128);
  tb[w] = make_uint4(hi[0] | ((uint32_t)hi[1] << 16), hi[2] | ((uint32_t)hi[3] << 16), hi[4] | ((uint32_t)hi[5] << 16),
                     hi[6] | ((uint32_t)hi[7] << 16));
  tb[KC * 128 + w] = make_uint4(lo[0] | ((uint32_t)lo[1] << 16), lo[2] | ((uint32_t)lo[3] << 16),
                                lo[4] | ((uint32_t)lo[5] << 16), lo[6] | ((uint32_t)lo[7] << 16));
}

// ---- the filter -------------------------------------------------------------------------------------------------
#define TC_TRACE(ev) do { if (a.trace && blockIdx.x == 0 && qcount == 0 && rt >= a.trace_off && rt < a.trace_off + 256) a.trace[(rt - a.trace_off) * 8 + (ev)] = clock64(); } while (0)

struct TcArgs {
  const float* Xc;           // centred embedding [n x d] (the query tile is split and written to TMEM in-kernel)
  int d;
  const unsigned char* TB;   // packed reference tiles of all n cells
  const float* norm2;
  const float* maxnorm2;     // largest squared row norm (fixes the power-of-two pre-scale)
  const int* order;          // [n] Morton position -> cell (reference tiles are packed in this order)
  const int* inv;            // [n] cell -> Morton position
  const int* qorder;         // [rows] query position -> cell (the queries of this call in Morton order)
  int n, KP, row_begin, row_end, n_q_tiles, n_r_tiles, stages;
  int* cand_idx;             // [rows x 32]
  float* cand_thr;           // [rows]
  unsigned long long* scratch;  // [gridDim.x x 128 rows x TC_CAP]
  int* err;
  float* dbg_s;              // optional [128 x dbg_cols]: raw s of query tile 0
  int dbg_cols;
  int trace_off;
  long long* trace;          // development: clock64 stamps of CTA 0, [256 tiles x 8 events]
};

// warp-cooperative compaction of one row buffer (c > TC_KEEP entries) to its best TC_KEEP entries, written to
// rowbuf[0..TC_KEEP) in no particular order; returns the TC_KEEP-th largest s.  Exact selection by a bitwise radix
// descent on the order-preserving 32-bit image of s: one vote per key register and bit, no sorting.
__device__ __forceinline__ float compact_row(unsigned long long* __restrict__ rowbuf, int c, int lane) {
  constexpr int R = TC_CAP / 32;
  uint32_t u[R], x[R];
  bool valid[R];
#pragma unroll
  for (int i = 0; i < R; ++i) {
    const int e = i * 32 + lane;
    valid[i] = e < c;
    const uint2 k = valid[i] ? reinterpret_cast<const uint2*>(rowbuf)[e] : make_uint2(0u, 0u);
    u[i] = valid[i] ? ordered_bits(k.y) : 0u;  // order-preserving image of s
    x[i] = k.x;                                // reference index
  }
  __syncwarp();
  // bits above the highest bit in which two valid keys differ are common to all of them
  uint32_t v_or = 0u, v_and = 0xffffffffu;
#pragma unroll
  for (int i = 0; i < R; ++i)
    if (valid[i]) {
      v_or |= u[i];
      v_and &= u[i];
    }
  v_or = __reduce_or_sync(FULLMASK, v_or);
  v_and = __reduce_and_sync(FULLMASK, v_and);
  const uint32_t diff = v_or ^ v_and;
  uint32_t prefix = v_and;
  if (diff != 0u) {
    const int hb = 31 - __clz(diff);
    prefix = (hb == 31) ? 0u : (v_and & ~((2u << hb) - 1u));
    for (int bit = hb; bit >= 0; --bit) {
      const uint32_t cand = prefix | (1u << bit);
      int n_ge = 0;
#pragma unroll
      for (int i = 0; i < R; ++i) n_ge += __popc(__ballot_sync(FULLMASK, valid[i] && u[i] >= cand));
      if (n_ge >= TC_KEEP) prefix = cand;
    }
  }
  // prefix == TC_KEEP-th largest value: keep everything above it, and as many ties as are needed to reach TC_KEEP
  const uint32_t lt_mask = (1u << lane) - 1u;
  int n_gt = 0;
#pragma unroll
  for (int i = 0; i < R; ++i) n_gt += __popc(__ballot_sync(FULLMASK, valid[i] && u[i] > prefix));
  int base_gt = 0, base_eq = n_gt;
#pragma unroll
  for (int i = 0; i < R; ++i) {
    const bool gt = valid[i] && u[i] > prefix, eq = valid[i] && u[i] == prefix;
    const unsigned bg = __ballot_sync(FULLMASK, gt), be = __ballot_sync(FULLMASK, eq);
    const uint2 k = make_uint2(x[i], unordered_bits(u[i]));
    if (gt) reinterpret_cast<uint2*>(rowbuf)[base_gt + __popc(bg & lt_mask)] = k;
    if (eq) {
      const int pos = base_eq + __popc(be & lt_mask);
      if (pos < TC_KEEP) reinterpret_cast<uint2*>(rowbuf)[pos] = k;
    }
    base_gt += __popc(bg);
    base_eq += __popc(be);
  }
  __syncwarp();
  return __uint_as_float(unordered_bits(prefix));
}

__global__ void __launch_bounds__(TC_THREADS, 1) k_knn_filter_tc(TcArgs a) {
  extern __shared__ __align__(1024) unsigned char smem[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t part_bytes = (uint32_t)a.KP * 256u;  // hi (or lo) part of one 128-row tile (fp16)
  const uint32_t tile_bytes = 2u * part_bytes;
  unsigned char* sB = smem;  // a.stages reference-tile stages
  const uint32_t NS = (uint32_t)a.stages;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + NS * tile_bytes);
  // bars[0..5] full, [6..11] empty, [12..14] tmem full, [15..17] tmem empty, [18] query tile written to TMEM
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 19);
  volatile int* abort_flag = reinterpret_cast<volatile int*>(tmem_slot + 1);
  // the two epilogue threads of a row (column halves) publish their thresholds and final counts here
  volatile float* sthr = reinterpret_cast<volatile float*>(smem + NS * tile_bytes + 256);  // [2][TC_M]
  volatile int* scnt = reinterpret_cast<volatile int*>(sthr + 2 * TC_M);                          // [2][TC_M]
  const uint32_t bar0 = smem_u32(bars);
  auto BAR = [&](int i) { return bar0 + 8u * (uint32_t)i; };
  constexpr int B_FULL = 0, B_EMPTY = TC_MAX_STAGES, B_TFULL = 2 * TC_MAX_STAGES, B_TEMPTY = 2 * TC_MAX_STAGES + TC_ACC,
                B_A = 2 * TC_MAX_STAGES + 2 * TC_ACC;

  if (tid == 0) {
    for (int b = 0; b < TC_MAX_STAGES; ++b) {
      mbar_init(BAR(B_FULL + b), 1);
      mbar_init(BAR(B_EMPTY + b), 1);
    }
    for (int b = 0; b < TC_ACC; ++b) {
      mbar_init(BAR(B_TFULL + b), 1);
      mbar_init(BAR(B_TEMPTY + b), 8);
    }
    mbar_init(BAR(B_A), 4);
    *abort_flag = 0;
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (tid < 2 * TC_M) sthr[tid] = -INFINITY;
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"((uint32_t)TC_TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  uint32_t it = 0;      // reference-tile iterations done by this CTA (ring / accumulator phase bookkeeping)
  uint32_t qcount = 0;  // query tiles done by this CTA
  for (int qt = blockIdx.x; qt < a.n_q_tiles; qt += gridDim.x, ++qcount) {
    // reference tiles are visited outwards from the tile that holds this query tile's first cell
    const int ctile = a.inv[a.qorder[(size_t)qt * TC_M]] / TC_N;
    // ring position of this query tile's first reference tile (stage ring and accumulator ring move together)
    uint32_t s = it % NS, ph = (it / NS) & 1u;            // shared-memory stage ring
    uint32_t b = it % TC_ACC, bph = (it / TC_ACC) & 1u;   // accumulator ring
    auto advance = [&]() {
      if (++s == NS) {
        s = 0;
        ph ^= 1u;
      }
      if (++b == TC_ACC) {
        b = 0;
        bph ^= 1u;
      }
    };
    // Producer and MMA warps run their loops warp-uniformly and elect one lane only around the asynchronous
    // instructions: addresses and descriptors then live in uniform registers.  (Issued from inside an `if (lane == 0)`
    // region every tcgen05.mma needed five R2UR moves and cost ~95 cycles instead of its 64-cycle floor;
    // scripts/microbench/mma_rate.cu.)
    if (warp == 0) {
      // ===== producer: streams the packed reference tiles with bulk copies =====
      const bool leader = elect_one();
      for (int rt = 0; rt < a.n_r_tiles; ++rt, advance()) {
        bool ok = mbar_wait(BAR(B_EMPTY + s), ph ^ 1u, abort_flag);
        ok = __all_sync(FULLMASK, ok);
        if (!ok) break;
        if (leader) {
          TC_TRACE(0);
          mbar_expect_tx(BAR(B_FULL + s), tile_bytes);
          const unsigned char* srcb = a.TB + (size_t)tile_seq(rt, ctile, a.n_r_tiles) * tile_bytes;
          const uint32_t dst = smem_u32(sB) + s * tile_bytes;
#pragma unroll
          for (int p = 0; p < 4; ++p)
            bulk_g2s(dst + p * (tile_bytes / 4), srcb + p * (tile_bytes / 4), tile_bytes / 4, BAR(B_FULL + s));
        }
        __syncwarp();
      }
    } else if (warp == 1) {
      // ===== MMA issuer: 3 x KP/8 tcgen05.mma per reference tile, issued by one elected lane =====
      const bool leader = elect_one();
      bool ok = mbar_wait(BAR(B_A), qcount & 1u, abort_flag);  // query tile is in tensor memory
      ok = __all_sync(FULLMASK, ok);
      const uint32_t tA_hi = tmem_base + TC_COL_AHI, tA_lo = tmem_base + TC_COL_ALO;
      const int ksteps = a.KP >> 4;
      bool pre_ok = false;  // barriers of this tile were already seen complete while the previous tile was issued
      for (int rt = 0; ok && rt < a.n_r_tiles; ++rt, advance()) {
        if (!pre_ok) {
          ok = mbar_wait(BAR(B_TEMPTY + b), bph ^ 1u, abort_flag);          // accumulator drained by the epilogue
          if (leader) TC_TRACE(1);
          ok = ok && mbar_wait(BAR(B_FULL + s), ph, abort_flag);            // reference tile landed
          if (leader) TC_TRACE(2);
          ok = __all_sync(FULLMASK, ok);
          if (!ok) break;
        }
        tc_fence_after();
        const uint32_t sb = smem_u32(sB) + s * tile_bytes;
        const uint64_t dB_hi = make_desc(sb), dB_lo = make_desc(sb + part_bytes);
        const uint32_t dtm = tmem_base + b * (uint32_t)TC_N;
        // per K step (16 halves): A advances 8 TMEM columns, B by 2 chunks = 2 * LBO = 4096 B = 256 descriptor units
        if (leader) {
          for (int k = 0; k < ksteps; ++k) tc_mma_f16_ts(dtm, tA_lo + (uint32_t)(k * 8), dB_hi + (uint64_t)(k * 256), TC_IDESC, k > 0);
          for (int k = 0; k < ksteps; ++k) tc_mma_f16_ts(dtm, tA_hi + (uint32_t)(k * 8), dB_lo + (uint64_t)(k * 256), TC_IDESC, 1u);
        }
        __syncwarp();
        // the issue above blocks on the tensor pipe's queue; look at the next tile's barriers now, while the pipe
        // still has work queued, so that the hand-over between tiles does not drain it
        pre_ok = false;
        if (rt + 1 < a.n_r_tiles) {
          const uint32_t s1 = (s + 1 == NS) ? 0u : s + 1u, ph1 = (s + 1 == NS) ? (ph ^ 1u) : ph;
          const uint32_t b1 = (b + 1 == TC_ACC) ? 0u : b + 1u, bph1 = (b + 1 == TC_ACC) ? (bph ^ 1u) : bph;
          pre_ok = mbar_test(BAR(B_TEMPTY + b1), bph1 ^ 1u) && mbar_test(BAR(B_FULL + s1), ph1);
          pre_ok = __all_sync(FULLMASK, pre_ok);
        }
        if (leader) {
          for (int k = 0; k < ksteps; ++k) tc_mma_f16_ts(dtm, tA_hi + (uint32_t)(k * 8), dB_hi + (uint64_t)(k * 256), TC_IDESC, 1u);
          tc_commit(BAR(B_EMPTY + s));  // shared-memory stage free once these MMAs have read it
          tc_commit(BAR(B_TFULL + b));  // accumulator ready for the epilogue
          TC_TRACE(3);
        }
        __syncwarp();
      }
    } else if (warp >= 4) {
      // ===== epilogue: 8 warps; a query row is shared by two threads (lane quadrant q = warp % 4 as tcgen05.ld demands,
      // column half h): each keeps the best 32 of its half of every tile, the halves are merged at the end.  The larger
      // of the two thresholds bounds both (the union's 32nd best is at least either half's). =====
      const int q = warp & 3, h = (warp - 4) >> 2;
      const int row_in_tile = q * 32 + lane;
      unsigned long long* rowbuf = a.scratch + (((size_t)blockIdx.x * 2 + h) * TC_M + row_in_tile) * TC_CAP;
      unsigned long long* warpbuf = a.scratch + (((size_t)blockIdx.x * 2 + h) * TC_M + q * 32) * TC_CAP;
      unsigned long long* otherbuf = a.scratch + (((size_t)blockIdx.x * 2 + (1 - h)) * TC_M + q * 32) * TC_CAP;
      float thr = -INFINITY;
      int cnt = 0;
      bool ok = true;
      const float inv_s2 = 1.f / (knn_scale(a.maxnorm2) * knn_scale(a.maxnorm2));  // power of two: exact
      if (h == 0) {
        // this thread's query row -> tensor memory: x = hi + lo (fp16 each, two elements per 32-bit column, the lower
        // index in the low half), extra column d holds 128 (it multiplies the references' -||xj||^2 scale^2/256
        // column), padding columns and rows past row_end are zero
        const long long qpos = (long long)qt * TC_M + row_in_tile;
        const bool live = qpos < (long long)(a.row_end - a.row_begin);
        const float* __restrict__ xr = a.Xc + (size_t)(live ? a.qorder[qpos] : 0) * a.d;
        const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16);
        const float scale = knn_scale(a.maxnorm2);
        for (int k16 = 0; k16 < (a.KP >> 4); ++k16) {
          uint32_t hi[8], lo[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            unsigned short h2[2], l2[2];
#pragma unroll
            for (int t = 0; t < 2; ++t) {
              const int c = k16 * 16 + e * 2 + t;
              float x = 0.f;
              if (live) x = (c < a.d) ? xr[c] * scale : (c == a.d ? TC_ONE_COL : 0.f);
              split_f16(x, h2[t], l2[t]);
            }
            hi[e] = h2[0] | ((uint32_t)h2[1] << 16);
            lo[e] = l2[0] | ((uint32_t)l2[1] << 16);
          }
          tc_st8(trow + TC_COL_AHI + (uint32_t)(k16 * 8), hi);
          tc_st8(trow + TC_COL_ALO + (uint32_t)(k16 * 8), lo);
        }
        tc_wait_st();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(BAR(B_A));
      }
      // one 32-column chunk of this thread's row: the common case (nothing above the threshold in any of the warp's
      // 32 rows) costs a max tree and one vote
      auto scan_chunk = [&](const uint32_t (&v)[32], int col0) {
        if (a.dbg_s != nullptr && qt == 0 && row_in_tile < a.row_end - a.row_begin) {
          // test probe: raw values back in input order (rows = this call's first 128 cells, columns = cells)
          const int rcell = a.qorder[row_in_tile] - a.row_begin;
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const int ccell = (col0 + j < a.n) ? a.order[col0 + j] : a.dbg_cols;
            if (ccell < a.dbg_cols) a.dbg_s[(size_t)rcell * a.dbg_cols + ccell] = __uint_as_float(v[j]) * inv_s2;
          }
        }
        float g[4];
#pragma unroll
        for (int q4 = 0; q4 < 4; ++q4) {
          const float m01 = fmaxf(__uint_as_float(v[q4 * 8 + 0]), __uint_as_float(v[q4 * 8 + 1]));
          const float m23 = fmaxf(__uint_as_float(v[q4 * 8 + 2]), __uint_as_float(v[q4 * 8 + 3]));
          const float m45 = fmaxf(__uint_as_float(v[q4 * 8 + 4]), __uint_as_float(v[q4 * 8 + 5]));
          const float m67 = fmaxf(__uint_as_float(v[q4 * 8 + 6]), __uint_as_float(v[q4 * 8 + 7]));
          g[q4] = fmaxf(fmaxf(m01, m23), fmaxf(m45, m67));
        }
        const bool hit = fmaxf(fmaxf(g[0], g[1]), fmaxf(g[2], g[3])) > thr;
        if (__any_sync(FULLMASK, hit)) {
          // hits are rare: one vote per group of 8 columns, then predicated (branch-free) appends inside a group
#pragma unroll
          for (int q4 = 0; q4 < 4; ++q4) {
            if (__any_sync(FULLMASK, g[q4] > thr)) {
#pragma unroll
              for (int jj = 0; jj < 8; ++jj) {
                const uint32_t sb = v[q4 * 8 + jj];
                if (__uint_as_float(sb) > thr) {
                  reinterpret_cast<uint2*>(rowbuf)[cnt] = make_uint2((uint32_t)(col0 + q4 * 8 + jj), sb);
                  ++cnt;
                }
              }
            }
          }
          __syncwarp();
          unsigned need = __ballot_sync(FULLMASK, cnt > TC_TRIG);
          while (need) {
            const int L = __ffs(need) - 1;
            need &= need - 1;
            const int c = __shfl_sync(FULLMASK, cnt, L);
            const float t = compact_row(warpbuf + (size_t)L * TC_CAP, c, lane);
            if (lane == L) {
              thr = fmaxf(thr, t);
              cnt = TC_KEEP;
              sthr[h * TC_M + row_in_tile] = thr;
            }
          }
        }
      };
      for (int rt = 0; rt < a.n_r_tiles; ++rt, advance()) {
        ok = mbar_wait(BAR(B_TFULL + b), bph, abort_flag);
        ok = __all_sync(FULLMASK, ok);
        if (!ok) break;
        if (warp == 4 && lane == 0) TC_TRACE(4);
        tc_fence_after();
        thr = fmaxf(thr, sthr[(1 - h) * TC_M + row_in_tile]);  // the other half's threshold is valid for this half too
        const uint32_t tbase = tmem_base + ((uint32_t)(q * 32) << 16) + b * (uint32_t)TC_N + (uint32_t)(h * 64);
        const int colt = tile_seq(rt, ctile, a.n_r_tiles) * TC_N + h * 64;
        {
          // the tensor-memory load of the second chunk is in flight while the first one is scanned
          uint32_t va[32], vb[32];
          tc_ld32(tbase, va);
          tc_wait_ld();
          tc_ld32(tbase + 32, vb);
          scan_chunk(va, colt);
          tc_wait_ld();
          // everything this warp needs of the accumulator buffer is in registers: hand it back before the last scan
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(BAR(B_TEMPTY + b));
          scan_chunk(vb, colt + 32);
        }
        if (warp == 4 && lane == 0) TC_TRACE(5);
      }
      // final shortlist of this half of the warp's 32 rows
      __syncwarp();
      {
        unsigned need = __ballot_sync(FULLMASK, cnt > TC_KEEP);
        while (need) {
          const int L = __ffs(need) - 1;
          need &= need - 1;
          const int c = __shfl_sync(FULLMASK, cnt, L);
          const float t = compact_row(warpbuf + (size_t)L * TC_CAP, c, lane);
          if (lane == L) {
            thr = fmaxf(thr, t);
            cnt = TC_KEEP;
          }
        }
      }
      scnt[h * TC_M + row_in_tile] = cnt;
      sthr[h * TC_M + row_in_tile] = -INFINITY;  // reset for the next query tile (the other half may still read it: harmless)
      __threadfence_block();
      asm volatile("bar.sync 1, 256;" ::: "memory");  // the 8 epilogue warps: both halves of every row are final
      if (h == 0) {
        // merge the two halves: append the other half's entries to this buffer, keep the best 32 of the union
        for (int L = 0; L < 32; ++L) {
          const int c0 = __shfl_sync(FULLMASK, cnt, L);
          const int c1 = scnt[TC_M + q * 32 + L];
          if (lane < c1)
            warpbuf[(size_t)L * TC_CAP + c0 + lane] = otherbuf[(size_t)L * TC_CAP + lane];
          __syncwarp();
          const int c = c0 + c1;
          float t = -INFINITY;
          if (c > TC_KEEP) t = compact_row(warpbuf + (size_t)L * TC_CAP, c, lane);
          const int cf = min(c, TC_KEEP);
          const long long qpos = (long long)qt * TC_M + q * 32 + L;
          if (ok && qpos < (long long)(a.row_end - a.row_begin)) {
            const int grow = a.qorder[qpos];
            const size_t orow = (size_t)(grow - a.row_begin);
            int idx = 0x7fffffff;
            if (lane < cf) {
              const int pos = (int)reinterpret_cast<const uint2*>(warpbuf)[(size_t)L * TC_CAP + lane].x;
              idx = pos < a.n ? a.order[pos] : 0x7fffffff;
            }
            a.cand_idx[orow * TC_KEEP + lane] = idx;
            if (lane == 0) a.cand_thr[orow] = (c > TC_KEEP) ? fmaf(-2.f, t * inv_s2, a.norm2[grow]) : INFINITY;
          }
        }
      }
    }
    it += (uint32_t)a.n_r_tiles;
    tc_fence_before();
    __syncthreads();  // every MMA of this query tile has completed (the epilogue consumed the last accumulator)
    tc_fence_after();
  }
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)TC_TMEM_COLS) : "memory");
  }
  if (tid == 0 && *abort_flag) atomicExch(a.err, 1);
}

// ---- host side ----------------------------------------------------------------------------------------------------
static inline int knn_tc_kp(int d) { return ((d + 1 + 15) / 16) * 16; }

bool sc_knn_tc_supported(int n, int d) {
  return knn_tc_kp(d) <= 128 && n >= 2048;  // >= 3 stages of KP/2 KB + barriers fit the 227 KB of shared memory, A fits TMEM
}

// Error bound of the filter, as a multiple of (||xi|| + ||xj||)^2, for |(||xi||^2 - 2 s_computed) - ||xi - xj||^2|:
//   * operand split: x = hi + lo + r with |r| <= 2^-22 |x| (two fp16 roundings, 11 significant bits each; the
//     pre-scale keeps every |x| < 256, so the subnormal floor 2^-25 is below 2^-32 of the largest norm), and the dropped
//     lo*lo products <= 2^-22 |xi||xj| per term                               -> 3 * 2^-22 per unit of sum |xi_c||xj_c|
//   * fp16 x fp16 products are exact in fp32; each of the 3*KP accumulations loses at most one fp32 ulp (2^-23,
//     truncation allowed) of a partial sum bounded by ||xi|| ||xj|| + ||xj||^2 / 2
//   => |error(s)| <= (3*2^-22 + 3*KP*2^-23) (||xi|| ||xj|| + ||xj||^2/2) <= half that times (||xi|| + ||xj||)^2,
//      and the distance carries 2 * error(s):  (12 + 6 KP) u  with u = 2^-24,
//   * plus the fp32 centring / norm / final-combination terms of the FFMA filter's bound, (d + 8) u, and 4 u for the
//     subnormal floor.
// The whole bound is doubled.  tests/test_gpu_parity.py measures the realised error against it (sc_knn_tc_probe).
double sc_knn_tc_gamma(int d) {
  const int KP = knn_tc_kp(d);
  return 2.0 * (double)(6 * KP + 12 + d + 8 + 4) * 5.9604644775390625e-08;
}

int sc_knn_filter_tc(sc_ctx* ctx, const float* Xc, const float* norm2, const float* maxnorm2, int n, int d, int row_begin,
                     int row_end,
                     int* cand_idx, float* cand_thr, float* dbg_s, int dbg_cols) {
  const int KP = knn_tc_kp(d);
  if (!sc_knn_tc_supported(n, d)) SC_FAIL(ctx, SC_ERR_ARG, "knn tensor-core filter: unsupported shape");
  const int rows = row_end - row_begin;
  if (rows <= 0) return SC_OK;
  const int n_q_tiles = (rows + TC_M - 1) / TC_M, n_r_tiles = (n + TC_N - 1) / TC_N;
  const size_t tile_bytes = (size_t)KP * 512;
  int dev = 0, sms = 0;
  SC_CUDA(ctx, cudaGetDevice(&dev));
  SC_CUDA(ctx, cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const int grid = n_q_tiles < sms ? n_q_tiles : sms;
  DBuf<unsigned char> TB;
  DBuf<unsigned long long> scratch, mk_in, mk_out;
  DBuf<int> err, order, inv, qorder;
  DBuf<char> stmp;
  SC_TRY(TB.ensure(ctx, (size_t)n_r_tiles * tile_bytes));
  SC_TRY(scratch.ensure(ctx, (size_t)grid * 2 * TC_M * TC_CAP));
  SC_TRY(err.ensure(ctx, 4));
  SC_CUDA(ctx, cudaMemsetAsync(err.p, 0, sizeof(int) * 4, ctx->stream));
  {
    // locality order of all cells (references) and of this call's queries: sort by nearest pivot
    int P = n / 512;
    P = P < 16 ? 16 : (P > 4096 ? 4096 : P);
    SC_TRY(mk_in.ensure(ctx, (size_t)n + 1));
    SC_TRY(mk_out.ensure(ctx, (size_t)n + 1));
    SC_TRY(order.ensure(ctx, (size_t)n + 1));
    SC_TRY(inv.ensure(ctx, (size_t)n + 1));
    SC_TRY(qorder.ensure(ctx, (size_t)rows + 1));
    SC_TRY(launch_nearest_pivot(ctx, Xc, norm2, n, d, P, 0, n, mk_in.p));
    SC_TRY(sc_sort_keys_u64(ctx, mk_in.p, mk_out.p, (size_t)n, 48, stmp));
    SC_LAUNCH(ctx, k_knn_unpack_order, (n + 255) / 256, 256, 0, n, mk_out.p, order.p, inv.p);
    if (rows == n) {
      SC_CUDA(ctx, cudaMemcpyAsync(qorder.p, order.p, sizeof(int) * (size_t)n, cudaMemcpyDeviceToDevice, ctx->stream));
    } else {
      SC_TRY(launch_nearest_pivot(ctx, Xc, norm2, n, d, P, row_begin, row_end, mk_in.p));
      SC_TRY(sc_sort_keys_u64(ctx, mk_in.p, mk_out.p, (size_t)rows, 48, stmp));
      SC_LAUNCH(ctx, k_knn_unpack_order, (rows + 255) / 256, 256, 0, rows, mk_out.p, qorder.p, (int*)nullptr);
    }
    const long long units = (long long)n_r_tiles * (KP / 8) * 128;
    SC_LAUNCH(ctx, k_knn_pack, (unsigned)((units + 255) / 256), 256, 0, Xc, norm2, maxnorm2, order.p, n, d, KP,
              reinterpret_cast<uint4*>(TB.p));
  }
  TcArgs a;
  a.Xc = Xc;
  a.d = d;
  a.TB = TB.p;
  a.norm2 = norm2;
  a.maxnorm2 = maxnorm2;
  a.order = order.p;
  a.inv = inv.p;
  a.qorder = qorder.p;
  a.n = n;
  a.KP = KP;
  a.row_begin = row_begin;
  a.row_end = row_end;
  a.n_q_tiles = n_q_tiles;
  a.n_r_tiles = n_r_tiles;
  a.cand_idx = cand_idx;
  a.cand_thr = cand_thr;
  a.scratch = scratch.p;
  a.err = err.p;
  a.dbg_s = dbg_s;
  a.dbg_cols = dbg_cols;
  DBuf<long long> trace;
  a.trace = nullptr;
  a.trace_off = 0;
  if (getenv("SC_KNN_TC_TRACE")) {
    SC_TRY(trace.ensure(ctx, 256 * 8));
    SC_CUDA(ctx, cudaMemsetAsync(trace.p, 0, sizeof(long long) * 256 * 8, ctx->stream));
    a.trace = trace.p;
    a.trace_off = getenv("SC_KNN_TC_TRACE_OFF") ? atoi(getenv("SC_KNN_TC_TRACE_OFF")) : 0;
  }
  int stages = (int)((size_t)(200 * 1024) / tile_bytes);
  stages = stages < 2 ? 2 : (stages > TC_MAX_STAGES ? TC_MAX_STAGES : stages);
  a.stages = stages;
  const size_t smem = (size_t)stages * tile_bytes + 256 + 4 * TC_M * sizeof(float);
  SC_CUDA(ctx, cudaFuncSetAttribute(k_knn_filter_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  SC_CUDA(ctx, cudaEventCreate(&e0));
  SC_CUDA(ctx, cudaEventCreate(&e1));
  SC_CUDA(ctx, cudaEventRecord(e0, ctx->stream));
  SC_LAUNCH(ctx, k_knn_filter_tc, grid, TC_THREADS, smem, a);
  SC_CUDA(ctx, cudaEventRecord(e1, ctx->stream));
  SC_CUDA(ctx, cudaEventSynchronize(e1));
  float ms = 0.f;
  cudaEventElapsedTime(&ms, e0, e1);
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  ctx->last_knn_filter_ms = ms;
  ctx->last_knn_filter_kind = 2;
  ctx->last_knn_kp = KP;
  if (getenv("SC_KNN_TC_TIME"))
    fprintf(stderr, "[seacells_b200] k_knn_filter_tc: %.3f ms (rows %d, n %d, KP %d, grid %d)\n", ms, rows, n, KP, grid);
  int herr = 0;
  SC_CUDA(ctx, cudaMemcpyAsync(&herr, err.p, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  SC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  if (a.trace && rows > 4096) {
    std::vector<long long> h(256 * 8);
    cudaMemcpy(h.data(), trace.p, sizeof(long long) * 256 * 8, cudaMemcpyDeviceToHost);
    FILE* f = fopen(getenv("SC_KNN_TC_TRACE"), "w");
    if (f) {
      for (int t = 0; t < 256; ++t) {
        for (int e = 0; e < 8; ++e) fprintf(f, "%lld ", h[t * 8 + e] ? h[t * 8 + e] - h[0] : -1ll);
        fprintf(f, "\n");
      }
      fclose(f);
    }
  }
  if (herr) SC_FAIL(ctx, SC_ERR_CUDA, "knn tensor-core filter: pipeline barrier timed out");
  return SC_OK;
}
