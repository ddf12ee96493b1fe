// kNN shortlist filter on the 5th-generation tensor cores (tcgen05, sm_100a).
//
// Replaces the distance contraction of scanpy.pp.neighbors (build_graph.py:125) for the filter stage of the exact kNN
// (kernel_build.cu): for every query i and reference j it evaluates
//        s(i,j) = xi.xj - ||xj||^2 / 2          (so that  ||xi-xj||^2 = ||xi||^2 - 2 s(i,j))
// on the mean-centred fp32 embedding as a 3xTF32 split product
//        A_hi B_hi + A_hi B_lo + A_lo B_hi,   x = hi + lo,  hi = tf32(x),  lo = tf32(x - hi)
// with fp32 accumulation in tensor memory, and keeps the 32 references with the largest s per query together with the
// 32nd value (the threshold of the certificate checked by k_knn_rerank).  The -||xj||^2/2 term rides along as one
// extra K column (A holds 1 there), so the epilogue is a pure "is anything in this row chunk above my threshold" scan.
//
// Layout
//   operands   pre-tiled in global memory (k_knn_pack): one block of KP*1024 bytes per 128 cells = hi part + lo part,
//              each already in the shared-memory image tcgen05.mma wants (K-major, no swizzle: 8x16-byte core
//              matrices, row groups 128 B apart (SBO), K chunks 2048 B apart (LBO)); KP = round_up(d+1, 8).
//   CTA        persistent, one per SM, 256 threads: warp 0 = bulk-copy producer (cp.async.bulk -> UBLKCP),
//              warp 1 = MMA issuer (one thread; tcgen05.mma kind::tf32, M=128 N=128 K=8), warps 4-7 = epilogue
//              (tcgen05.ld 32x32b: one query row per thread).  A query tile (128 rows) stays in shared memory while
//              all reference tiles stream through a 2-stage ring; accumulators are double buffered in TMEM
//              (2 x 128 columns) so the MMAs of tile t+1 run under the scan of tile t.
//   selection  per row: threshold register + append buffer (128 entries, L2-resident scratch); a row whose buffer
//              passes 96 entries is compacted to its best 32 by the whole warp (rank counting over shuffles).
//
// Every mbarrier wait is bounded (2 s): a broken pipeline sets an error flag and drains instead of hanging the GPU.
#include "kernel_build.cuh"

#include <math.h>

namespace {

constexpr int TC_M = 128;          // queries per CTA tile
constexpr int TC_N = 128;          // references per streamed tile
constexpr int TC_THREADS = 256;
constexpr int TC_CAP = 128;        // append-buffer entries per row
constexpr int TC_TRIG = 96;        // compact when a row holds more than this after a chunk
constexpr int TC_KEEP = 32;        // shortlist size (== KF_NC in kernel_build.cu)
constexpr int TC_TMEM_COLS = 256;  // 2 accumulator buffers x 128 fp32 columns
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
// bounded wait: false (and the shared abort flag set) when the barrier did not flip within ~2 s
__device__ __forceinline__ bool mbar_wait(uint32_t bar, uint32_t parity, volatile int* abort_flag) {
  if (mbar_try_wait(bar, parity)) return true;
  const long long t0 = clock64();
  while (true) {
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
__device__ __forceinline__ void tc_mma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
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
// instruction descriptor: D fp32, A/B tf32, both K-major, N=128, M=128
constexpr uint32_t TC_IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(TC_N >> 3) << 17) |
                              ((uint32_t)(TC_M >> 4) << 24);

__device__ __forceinline__ float tf32_rna(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}

// (s, reference) -> 64-bit key: larger key = better candidate (larger s, then smaller reference index)
__device__ __forceinline__ unsigned long long cand_key(float s, int idx) {
  uint32_t u = __float_as_uint(s);
  u = (u & 0x80000000u) ? ~u : (u | 0x80000000u);
  return ((unsigned long long)u << 32) | (uint32_t)(~(uint32_t)idx);
}

}  // namespace

// ---- operand packing --------------------------------------------------------------------------------------------
// One 16-byte unit (4 consecutive K columns of one cell) per thread, for the hi and the lo part.
// is_query: extra column = 1, padding rows all zero.   reference: extra column = -||x||^2/2, padding rows -1e30 there.
__global__ void k_knn_pack(const float* __restrict__ Xc, const float* __restrict__ norm2, int n, int d, int KP,
                           int row_begin, int row_end, int is_query, float4* __restrict__ out) {
  const int KC = KP >> 2;
  const long long u = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int rows = row_end - row_begin;
  const long long n_tiles = (rows + TC_M - 1) / TC_M;
  if (u >= n_tiles * KC * 128) return;
  const long long tile = u / (KC * 128);
  const int w = (int)(u - tile * (KC * 128));
  const int kc = w >> 7, r = w & 127;
  const long long cell = (long long)row_begin + tile * 128 + r;
  float x[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const int c = kc * 4 + e;
    float v = 0.f;
    if (cell < row_end) {
      if (c < d) v = Xc[cell * d + c];
      else if (c == d) v = is_query ? 1.f : -0.5f * norm2[cell];
    } else if (!is_query && c == d) {
      v = -1e30f;
    }
    x[e] = v;
  }
  float hi[4], lo[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    hi[e] = tf32_rna(x[e]);
    lo[e] = tf32_rna(x[e] - hi[e]);
  }
  float4* tb = out + tile * (size_t)(2 * KC * 128);
  tb[w] = make_float4(hi[0], hi[1], hi[2], hi[3]);
  tb[KC * 128 + w] = make_float4(lo[0], lo[1], lo[2], lo[3]);
}

// ---- the filter -------------------------------------------------------------------------------------------------
struct TcArgs {
  const unsigned char* TA;   // packed query tiles of [row_begin, row_end)
  const unsigned char* TB;   // packed reference tiles of all n cells
  const float* norm2;
  int n, KP, row_begin, row_end, n_q_tiles, n_r_tiles;
  int* cand_idx;             // [rows x 32]
  float* cand_thr;           // [rows]
  unsigned long long* scratch;  // [gridDim.x x 128 rows x TC_CAP]
  int* err;
  float* dbg_s;              // optional [128 x dbg_cols]: raw s of query tile 0
  int dbg_cols;
};

// warp-cooperative compaction of one row buffer to its best TC_KEEP entries (sorted, best first); returns the
// TC_KEEP-th best s (or -inf when the row holds fewer)
__device__ __forceinline__ float compact_row(unsigned long long* __restrict__ rowbuf, int c, int lane) {
  unsigned long long key[TC_CAP / 32];
#pragma unroll
  for (int i = 0; i < TC_CAP / 32; ++i) {
    const int e = i * 32 + lane;
    key[i] = (e < c) ? rowbuf[e] : 0ull;
  }
  __syncwarp();
  int rank[TC_CAP / 32];
#pragma unroll
  for (int i = 0; i < TC_CAP / 32; ++i) rank[i] = 0;
#pragma unroll
  for (int i2 = 0; i2 < TC_CAP / 32; ++i2) {
    if (i2 * 32 < c) {
      const int lim = min(32, c - i2 * 32);
      for (int jj = 0; jj < lim; ++jj) {
        const unsigned long long kj = __shfl_sync(FULLMASK, key[i2], jj);
#pragma unroll
        for (int i = 0; i < TC_CAP / 32; ++i) rank[i] += (kj > key[i]) ? 1 : 0;
      }
    }
  }
  unsigned long long k32 = 0ull;
#pragma unroll
  for (int i = 0; i < TC_CAP / 32; ++i) {
    const int e = i * 32 + lane;
    if (e < c && rank[i] < TC_KEEP) rowbuf[rank[i]] = key[i];
    if (e < c && rank[i] == TC_KEEP - 1) k32 = key[i];
  }
  // exactly one lane (or none) holds the TC_KEEP-th key
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) k32 |= __shfl_xor_sync(FULLMASK, k32, o);
  __syncwarp();
  if (k32 == 0ull) return -INFINITY;
  uint32_t u = (uint32_t)(k32 >> 32);
  u = (u & 0x80000000u) ? (u & 0x7fffffffu) : ~u;
  return __uint_as_float(u);
}

__global__ void __launch_bounds__(TC_THREADS, 1) k_knn_filter_tc(TcArgs a) {
  extern __shared__ __align__(1024) unsigned char smem[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t part_bytes = (uint32_t)a.KP * 512u;  // hi (or lo) part of one 128-row tile
  const uint32_t tile_bytes = 2u * part_bytes;
  unsigned char* sA = smem;
  unsigned char* sB = smem + tile_bytes;  // 2 stages
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 3u * tile_bytes);
  // bars[0..1] full, [2..3] empty, [4..5] tmem full, [6..7] tmem empty, [8] A tile
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 9);
  volatile int* abort_flag = reinterpret_cast<volatile int*>(tmem_slot + 1);
  const uint32_t bar0 = smem_u32(bars);
  auto BAR = [&](int i) { return bar0 + 8u * (uint32_t)i; };

  if (tid == 0) {
    mbar_init(BAR(0), 1);
    mbar_init(BAR(1), 1);
    mbar_init(BAR(2), 1);
    mbar_init(BAR(3), 1);
    mbar_init(BAR(4), 1);
    mbar_init(BAR(5), 1);
    mbar_init(BAR(6), 4);
    mbar_init(BAR(7), 4);
    mbar_init(BAR(8), 1);
    *abort_flag = 0;
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
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
    if (warp == 0) {
      // ===== producer: one thread streams the packed tiles with bulk copies =====
      if (lane == 0) {
        mbar_expect_tx(BAR(8), tile_bytes);
        const unsigned char* src = a.TA + (size_t)qt * tile_bytes;
#pragma unroll
        for (int p = 0; p < 4; ++p) bulk_g2s(smem_u32(sA) + p * (tile_bytes / 4), src + p * (tile_bytes / 4), tile_bytes / 4, BAR(8));
        uint32_t i = it;
        for (int rt = 0; rt < a.n_r_tiles; ++rt, ++i) {
          const uint32_t s = i & 1u, ph = (i >> 1) & 1u;
          if (!mbar_wait(BAR(2 + s), ph ^ 1u, abort_flag)) break;
          mbar_expect_tx(BAR(s), tile_bytes);
          const unsigned char* srcb = a.TB + (size_t)rt * tile_bytes;
          const uint32_t dst = smem_u32(sB) + s * tile_bytes;
#pragma unroll
          for (int p = 0; p < 4; ++p) bulk_g2s(dst + p * (tile_bytes / 4), srcb + p * (tile_bytes / 4), tile_bytes / 4, BAR(s));
        }
      }
      __syncwarp();
    } else if (warp == 1) {
      // ===== MMA issuer: one thread, 3 x KP/8 tcgen05.mma per reference tile =====
      if (lane == 0) {
        bool ok = mbar_wait(BAR(8), qcount & 1u, abort_flag);
        const uint64_t dA_hi = make_desc(smem_u32(sA)), dA_lo = make_desc(smem_u32(sA) + part_bytes);
        const int ksteps = a.KP >> 3;
        uint32_t i = it;
        for (int rt = 0; ok && rt < a.n_r_tiles; ++rt, ++i) {
          const uint32_t s = i & 1u, ph = (i >> 1) & 1u;
          if (!mbar_wait(BAR(6 + s), ph ^ 1u, abort_flag)) break;  // accumulator buffer drained by the epilogue
          if (!mbar_wait(BAR(s), ph, abort_flag)) break;            // reference tile landed
          tc_fence_after();
          const uint32_t sb = smem_u32(sB) + s * tile_bytes;
          const uint64_t dB_hi = make_desc(sb), dB_lo = make_desc(sb + part_bytes);
          const uint32_t dtm = tmem_base + s * (uint32_t)TC_N;
          // each K step (8 tf32 = 2 chunks of 16 B) advances the start address by 2 * LBO = 4096 B = 256 units
          for (int k = 0; k < ksteps; ++k) tc_mma_tf32(dtm, dA_lo + (uint64_t)(k * 256), dB_hi + (uint64_t)(k * 256), TC_IDESC, k > 0);
          for (int k = 0; k < ksteps; ++k) tc_mma_tf32(dtm, dA_hi + (uint64_t)(k * 256), dB_lo + (uint64_t)(k * 256), TC_IDESC, 1u);
          for (int k = 0; k < ksteps; ++k) tc_mma_tf32(dtm, dA_hi + (uint64_t)(k * 256), dB_hi + (uint64_t)(k * 256), TC_IDESC, 1u);
          tc_commit(BAR(2 + s));  // shared-memory stage free once these MMAs have read it
          tc_commit(BAR(4 + s));  // accumulator ready for the epilogue
        }
      }
      __syncwarp();
    } else if (warp >= 4) {
      // ===== epilogue: one query row per thread =====
      const int q = warp & 3;
      const int row_in_tile = q * 32 + lane;
      const long long self_ll = (long long)a.row_begin + (long long)qt * TC_M + row_in_tile;
      const int self = (int)self_ll;
      unsigned long long* rowbuf = a.scratch + ((size_t)blockIdx.x * TC_M + row_in_tile) * TC_CAP;
      unsigned long long* warpbuf = a.scratch + ((size_t)blockIdx.x * TC_M + q * 32) * TC_CAP;
      float thr = -INFINITY;
      int cnt = 0;
      uint32_t i = it;
      bool ok = true;
      for (int rt = 0; rt < a.n_r_tiles; ++rt, ++i) {
        const uint32_t s = i & 1u, ph = (i >> 1) & 1u;
        ok = mbar_wait(BAR(4 + s), ph, abort_flag);
        ok = __all_sync(FULLMASK, ok);
        if (!ok) break;
        tc_fence_after();
        const uint32_t tbase = tmem_base + ((uint32_t)(q * 32) << 16) + s * (uint32_t)TC_N;
#pragma unroll 1
        for (int ch = 0; ch < TC_N / 32; ++ch) {
          uint32_t v[32];
          tc_ld32(tbase + ch * 32, v);
          tc_wait_ld();
          const int col0 = rt * TC_N + ch * 32;
          if (a.dbg_s != nullptr && qt == 0) {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (col0 + j < a.dbg_cols) a.dbg_s[(size_t)row_in_tile * a.dbg_cols + col0 + j] = __uint_as_float(v[j]);
          }
          float m0 = __uint_as_float(v[0]), m1 = __uint_as_float(v[1]), m2 = __uint_as_float(v[2]), m3 = __uint_as_float(v[3]);
#pragma unroll
          for (int j = 4; j < 32; j += 4) {
            m0 = fmaxf(m0, __uint_as_float(v[j]));
            m1 = fmaxf(m1, __uint_as_float(v[j + 1]));
            m2 = fmaxf(m2, __uint_as_float(v[j + 2]));
            m3 = fmaxf(m3, __uint_as_float(v[j + 3]));
          }
          const bool hit = fmaxf(fmaxf(m0, m1), fmaxf(m2, m3)) > thr;
          if (__any_sync(FULLMASK, hit)) {
            if (hit) {
#pragma unroll
              for (int j = 0; j < 32; ++j) {
                const float sv = __uint_as_float(v[j]);
                const int col = col0 + j;
                if (sv > thr && col != self) {
                  rowbuf[cnt] = cand_key(sv, col);
                  ++cnt;
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
                thr = t;
                cnt = TC_KEEP;
              }
            }
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(BAR(6 + s));
      }
      // final shortlist of the 32 rows of this warp
      __syncwarp();
      for (int L = 0; L < 32; ++L) {
        const int c = __shfl_sync(FULLMASK, cnt, L);
        const float t = compact_row(warpbuf + (size_t)L * TC_CAP, c, lane);
        const long long grow = (long long)a.row_begin + (long long)qt * TC_M + q * 32 + L;
        if (ok && grow < a.row_end) {
          const size_t orow = (size_t)(grow - a.row_begin);
          int idx = 0x7fffffff;
          if (lane < min(c, TC_KEEP)) idx = (int)(~(uint32_t)(warpbuf[(size_t)L * TC_CAP + lane] & 0xffffffffull));
          a.cand_idx[orow * TC_KEEP + lane] = idx;
          if (lane == 0) a.cand_thr[orow] = (c >= TC_KEEP) ? fmaf(-2.f, t, a.norm2[grow]) : INFINITY;
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
bool sc_knn_tc_supported(int n, int d) {
  const int KP = ((d + 1 + 7) / 8) * 8;
  return KP <= 64 && n >= 2048;  // 3 tiles of KP KB + barriers must fit the 227 KB of shared memory
}

// Error bound of the filter, as a multiple of (||xi|| + ||xj||)^2, for |(||xi||^2 - 2 s_computed) - ||xi - xj||^2|:
//   * operand split: x = hi + lo + r with |r| <= 2^-22 |x| (two tf32 roundings), and the dropped lo*lo products
//     <= 2^-22 |xi||xj| per term                                             -> 3 * 2^-22 per unit of sum |xi_c||xj_c|
//   * tf32 x tf32 products are exact in fp32; each of the 3*KP accumulations loses at most one fp32 ulp (2^-23,
//     truncation allowed) of a partial sum bounded by ||xi|| ||xj|| + ||xj||^2 / 2
//   => |error(s)| <= (3*2^-22 + 3*KP*2^-23) (||xi|| ||xj|| + ||xj||^2/2) <= half that times (||xi|| + ||xj||)^2,
//      and the distance carries 2 * error(s):  (12 + 6 KP) u  with u = 2^-24,
//   * plus the fp32 centring / norm / final-combination terms of the FFMA filter's bound, (d + 8) u.
// The whole bound is doubled.  tests/test_gpu_parity.py measures the realised error against it (sc_knn_tc_probe).
double sc_knn_tc_gamma(int d) {
  const int KP = ((d + 1 + 7) / 8) * 8;
  return 2.0 * (double)(6 * KP + 12 + d + 8) * 5.9604644775390625e-08;
}

int sc_knn_filter_tc(sc_ctx* ctx, const float* Xc, const float* norm2, int n, int d, int row_begin, int row_end,
                     int* cand_idx, float* cand_thr, float* dbg_s, int dbg_cols) {
  const int KP = ((d + 1 + 7) / 8) * 8;
  if (!sc_knn_tc_supported(n, d)) SC_FAIL(ctx, SC_ERR_ARG, "knn tensor-core filter: unsupported shape");
  const int rows = row_end - row_begin;
  if (rows <= 0) return SC_OK;
  const int n_q_tiles = (rows + TC_M - 1) / TC_M, n_r_tiles = (n + TC_N - 1) / TC_N;
  const size_t tile_bytes = (size_t)KP * 1024;
  int dev = 0, sms = 0;
  SC_CUDA(ctx, cudaGetDevice(&dev));
  SC_CUDA(ctx, cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const int grid = n_q_tiles < sms ? n_q_tiles : sms;
  DBuf<unsigned char> TA, TB;
  DBuf<unsigned long long> scratch;
  DBuf<int> err;
  SC_TRY(TA.ensure(ctx, (size_t)n_q_tiles * tile_bytes));
  SC_TRY(TB.ensure(ctx, (size_t)n_r_tiles * tile_bytes));
  SC_TRY(scratch.ensure(ctx, (size_t)grid * TC_M * TC_CAP));
  SC_TRY(err.ensure(ctx, 4));
  SC_CUDA(ctx, cudaMemsetAsync(err.p, 0, sizeof(int) * 4, ctx->stream));
  {
    const long long units = (long long)n_q_tiles * (KP / 4) * 128;
    SC_LAUNCH(ctx, k_knn_pack, (unsigned)((units + 255) / 256), 256, 0, Xc, norm2, n, d, KP, row_begin, row_end, 1,
              reinterpret_cast<float4*>(TA.p));
  }
  {
    const long long units = (long long)n_r_tiles * (KP / 4) * 128;
    SC_LAUNCH(ctx, k_knn_pack, (unsigned)((units + 255) / 256), 256, 0, Xc, norm2, n, d, KP, 0, n, 0,
              reinterpret_cast<float4*>(TB.p));
  }
  TcArgs a;
  a.TA = TA.p;
  a.TB = TB.p;
  a.norm2 = norm2;
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
  const size_t smem = 3 * tile_bytes + 128;
  SC_CUDA(ctx, cudaFuncSetAttribute(k_knn_filter_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  SC_LAUNCH(ctx, k_knn_filter_tc, grid, TC_THREADS, smem, a);
  int herr = 0;
  SC_CUDA(ctx, cudaMemcpyAsync(&herr, err.p, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  SC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  if (herr) SC_FAIL(ctx, SC_ERR_CUDA, "knn tensor-core filter: pipeline barrier timed out");
  return SC_OK;
}
