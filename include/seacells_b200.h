/* seacells_b200 - C ABI of the B200-native SEACells hot paths.
 *
 * The reference (dpeerlab/SEACells v0.3.3) has no FFI: its boundary is the Python object returned by
 * SEACells.core.SEACells (core.py:13-100).  These entry points are what that class's methods bind when the two
 * data-parallel hot paths move to the GPU; each one cites the reference method it replaces.  The Python mirror of the
 * class lives in seacells_b200/core.py and reaches this library through ctypes.
 *
 * Conventions
 *   - every function returns 0 on success or an SC_ERR_* code; sc_last_error() gives the message; nothing throws
 *   - data pointers may be HOST or DEVICE addresses (detected with cudaPointerGetAttributes); results are written to
 *     whichever kind of buffer the caller supplies; the caller owns every buffer it passes
 *   - handles own their device scratch; work is queued on the context's stream and the call returns after the
 *     results it promises are complete
 *   - indices are int32, archetype/cell positions are 0-based, values are IEEE fp64
 */
#ifndef SEACELLS_B200_H
#define SEACELLS_B200_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SC_OK 0
#define SC_ERR_CUDA 1
#define SC_ERR_ARG 2
#define SC_ERR_STATE 3
#define SC_ERR_CAPACITY 4
#define SC_ERR_NCCL 5

typedef struct sc_ctx sc_ctx;             /* one per process / GPU */
typedef struct sc_kernel_graph sc_kernel; /* device-resident kernel matrix M (build_graph.py:188 self.M) */
typedef struct sc_fit sc_fit;             /* device-resident model state (cpu.py:97-111: K, A_, B_, A0, B0) */

/* ---- context ------------------------------------------------------------------------------------------------ */
int sc_ctx_create(int device, sc_ctx** out);
void sc_ctx_destroy(sc_ctx* ctx);
const char* sc_last_error(sc_ctx* ctx);
unsigned long long sc_launch_count(sc_ctx* ctx); /* kernels launched by this library so far */
int sc_sync(sc_ctx* ctx);
void* sc_stream(sc_ctx* ctx); /* cudaStream_t the library launches on (for CUDA-event timing) */
const char* sc_version(void);

/* ---- row-sharded jobs: one context per rank, one rank per GPU ------------------------------
 * The reference has no distributed code; this is the engine's own scale-out.  A context that belongs to a job of `world`
 * ranks owns the cells [rank * slab, (rank + 1) * slab), slab = ceil(n / world): sc_kernel_build searches the neighbours
 * of its slab only, sc_fit_update_A fills its columns of A, sc_fit_update_B evaluates its rows of the gradient, and the
 * library exchanges what the other ranks need on its own stream with ONE collective, an in-place all-gather of equal
 * slabs (compressed A, per-cell RSS terms, rows of A A^T by archetype slab, rows of M A^T, the K B rows of supp(B), kNN
 * rows, per step the round-0 pruning bounds and the per-archetype {candidate min G, cell, non-candidate G, cell} argmin
 * partials).  No floating-point value is reduced across ranks, so results are bitwise identical to a single-GPU run.
 * Every rank must make the same calls in the same order with the same M, archetypes and A0, and a context must join
 * its job BEFORE any handle (sc_kernel, sc_fit) is created on it.  A rank's sc_fit forms only the columns of K = M M^T
 * that belong to its own cells.  The argmin trace of sc_fit_update_A covers the rank's own cells only.
 *   sc_nccl_unique_id + sc_ctx_init_nccl   one process per GPU: rank 0 creates the 128-byte id, the host broadcasts it
 *                                          (torch.distributed, MPI, a file ...), every rank joins; libnccl.so.2 is
 *                                          resolved with dlopen (sc_nccl_load_library gives an explicit path)
 *   sc_local_group_new + sc_ctx_attach_local   ranks are threads of one process (one context each, on one or on several
 *                                          GPUs); slabs move with peer copies between thread barriers */
int sc_nccl_load_library(const char* path);
int sc_nccl_unique_id(void* out128);
int sc_ctx_init_nccl(sc_ctx* ctx, int rank, int world, const void* id128);
int sc_local_group_new(int world, void** group_out);
void sc_local_group_free(void* group);
void sc_local_group_abort(void* group); /* a rank failed: peers blocked in an exchange return SC_ERR_NCCL */
int sc_ctx_attach_local(sc_ctx* ctx, void* group, int rank);
int sc_ctx_rank(sc_ctx* ctx);
int sc_ctx_world(sc_ctx* ctx);

/* ---- hot path 1: kernel construction  (build_graph.SEACellGraph.rbf, build_graph.py:113-189) -------------------- */

/* Replaces scanpy.pp.neighbors(ad, use_rep, n_neighbors, knn=True) at build_graph.py:125 with an exact brute-force
 * kNN over X [n x d] (float32 when x_is_f64 == 0, else float64).  Rows [row_begin,row_end) are the queries.
 * idx_out / d2_out: [(row_end-row_begin) x (n_neighbors-1)], neighbours by (squared distance, index), self excluded. */
int sc_knn(sc_ctx* ctx, const void* X, int x_is_f64, int n, int d, int n_neighbors, int row_begin, int row_end,
           int32_t* idx_out, double* d2_out);
/* For n >= 2048 sc_knn shortlists 32 candidates per query with an fp32 distance tile kernel and re-ranks them with the
 * exact fp64 distance; a per-query certificate proves that nothing was missed, otherwise the query is recomputed by the
 * all-fp64 kernel (results are bit-identical either way; SC_KNN_EXACT=1 in the environment forces the all-fp64 kernel).
 * Number of queries of the last call that needed the recomputation: */
int sc_knn_fallback_rows(sc_ctx* ctx);
/* Shortlist stage of the last sc_knn / sc_kernel_build call: out4 = { device milliseconds of the shortlist kernel (CUDA
 * events), kind (0 none, 1 fp32 FFMA tiles, 2 tcgen05 split-fp16), padded K of the tensor-core tiles, fallback rows }. */
int sc_knn_profile(sc_ctx* ctx, double* out4);

/* The shortlist stage runs on the tensor cores (tcgen05 kind::f16, split product x = hi + lo with two fp16 halves = the
 * 22-bit operand precision of 3xTF32, fp32 accumulation) when d <= 127; SC_KNN_FFMA=1 forces the fp32 FFMA tile kernel.  Test probe of that stage: s(i,j) = xi.xj - ||xj||^2/2 as the tensor cores computed it
 * for the first 128 queries against references [0, cols)  (s_out [128 x cols]), plus the mean-centred fp32 embedding
 * (xc_out [n x d]) and its squared norms (norm2_out [n]) the products were formed from; host buffers.
 * sc_knn_tc_error_factor(d) is the certificate's bound on |(||xi||^2 - 2 s) - ||xi-xj||^2| / (||xi|| + ||xj||)^2. */
int sc_knn_tc_probe(sc_ctx* ctx, const void* X, int x_is_f64, int n, int d, int cols, float* s_out, float* xc_out,
                    float* norm2_out);
double sc_knn_tc_error_factor(int d);

/* SEACellGraph.rbf(k, graph_construction) (build_graph.py:113): kNN, sigma (18-34), union/intersection (155-164),
 * Gaussian values (37-61), CSR (181-188).  intersection: 0 = "union", 1 = "intersect"/"intersection". */
int sc_kernel_build(sc_ctx* ctx, const void* X, int x_is_f64, int n, int d, int n_neighbors, int intersection,
                    sc_kernel** out, long long* nnz_out);
/* Same, from a kNN graph the caller already holds (idx/d2 as produced by sc_knn for all n rows). */
int sc_kernel_build_from_knn(sc_ctx* ctx, const void* X, int x_is_f64, int n, int d, int n_neighbors,
                             const int32_t* idx, const double* d2, int intersection, sc_kernel** out,
                             long long* nnz_out);
/* Copy M out as CSR (indptr [n+1], indices/data [nnz]) plus sigma [n] and the kNN graph; any pointer may be NULL. */
int sc_kernel_export(sc_ctx* ctx, sc_kernel* km, int32_t* indptr, int32_t* indices, double* data, double* sigma,
                     int32_t* knn_idx, double* knn_d2);
void sc_kernel_free(sc_kernel* km);

/* Y[r x c] = S[r x m] * X[m x c], S in CSR, X and Y dense row-major fp64 (the reference's K @ B with a dense B,
 * cpu_dense.py:353,389 / gpu.py:429,477).  Utility + roofline probe; the fit loop itself uses compressed operands. */
int sc_spmm_csr_f64(sc_ctx* ctx, int r, int m, int c, const int32_t* indptr, const int32_t* indices,
                    const double* data, const double* X, double* Y);

/* ---- next row: metacell aggregation  (core.py:121-242 summarize_by_SEACell / summarize_by_soft_SEACell) -------------- */
/* out[g, :] = sum over the members (cell, weight) of group g, in ascending cell order, of weight * X[cell, :], divided
 * by divisor[g] when divisor != NULL.  X: cells x genes counts in CSR (int64 indptr [n+1], int32 indices, float32 or
 * float64 data); members: m (group, cell, weight) triplets in any order; out: dense row-major [n_groups x n_genes] fp64.
 * Hard assignment (core.py:215-225): one triplet per cell, weight 1, no divisor.  Soft assignment (core.py:158-166):
 * the thresholded, row-normalised entries of A^T, divisor = their column sums.  Host or device pointers. */
int sc_summarize(sc_ctx* ctx, int n, int n_genes, const int64_t* indptr, const int32_t* indices, const void* data,
                 int data_is_f64, long long m, const int32_t* group, const int32_t* cell, const double* weight,
                 int n_groups, const double* divisor, double* out);

/* ---- next row: greedy initialisation  (cpu.py:342-423 _get_greedy_centers) ---------------------------------------- */
/* Greedy adaptive column-subset selection on K = M M^T; M as CSR (host or device pointers), centers_out: host int64
 * [n_centers].  Keeps the reference's dense (n_centers x n) fp64 work matrix, so SC_ERR_CAPACITY when that does not fit. */
int sc_greedy_centers(sc_ctx* ctx, int n, const int32_t* indptr, const int32_t* indices, const double* data,
                      long long nnz, int n_centers, int64_t* centers_out);

/* ---- hot path 2: Frank-Wolfe fit  (cpu.py:204-285, 425-536, 558-651) -------------------------------------------- */
int sc_fit_create(sc_ctx* ctx, int n_cells, int n_archetypes, int max_FW_iter, sc_fit** out);
void sc_fit_destroy(sc_fit* fit);

/* add_precomputed_kernel_matrix (cpu.py:115-128): M as CSR with sorted columns, symmetric and non-negative (checked);
 * K = M M^T is formed on the device here, as the reference does at cpu.py:127,157 */
int sc_fit_set_kernel(sc_fit* fit, const int32_t* indptr, const int32_t* indices, const double* data, long long nnz);
int sc_fit_set_kernel_from(sc_fit* fit, sc_kernel* km); /* device-to-device hand-off from sc_kernel_build */

/* initialize (cpu.py:237-244): B0 = one-hot at (archetypes[a], a) */
int sc_fit_set_archetypes(sc_fit* fit, const int64_t* cell_positions /* [k], host */);
/* initial_assignments (cpu.py:246-264): A0 as CSC (column j = cell j), already L1-column-normalised by the caller */
int sc_fit_set_A_csc(sc_fit* fit, const int64_t* colptr /* [n+1] */, const int32_t* rows, const double* vals);
/* restore compressed state (resume): slots are [cols x S] with S = max_FW_iter, cnt [cols] */
int sc_fit_set_A_slots(sc_fit* fit, const int32_t* idx, const double* val, const int32_t* cnt);
int sc_fit_set_B_slots(sc_fit* fit, const int32_t* idx, const double* val, const int32_t* cnt);

/* _updateA (cpu.py:425-459; l2_penalty as cpu_dense.py:360).  trace: NULL or host int32 [max_FW_iter x n] that
 * receives the argmin of every Frank-Wolfe step (test hook for argmin-sequence parity). */
int sc_fit_update_A(sc_fit* fit, double l2_penalty, int32_t* trace);
/* _updateB (cpu.py:461-498).  trace: NULL or host int32 [max_FW_iter x k]. */
int sc_fit_update_B(sc_fit* fit, int32_t* trace);
/* cells [row_begin, row_end) owned by this handle's rank (the whole range on one GPU) */
int sc_fit_shard(sc_fit* fit, int* row_begin, int* row_end);

/* compute_RSS (cpu.py:520-536): || M - (M B) A ||_F via the trace identity, no n x n product */
int sc_fit_rss(sc_fit* fit, double* rss_out);
/* initialize + _fit (cpu.py:204-285, 595-651) for already-set archetypes and A0.  threshold_inout: <= 0 on entry
 * means "unset" (then epsilon * RSS0 is stored, cpu.py:279-285); rss_out receives RSS_iters (RSS0 first). */
int sc_fit_run(sc_fit* fit, int max_iter, int min_iter, double convergence_epsilon, double l2_penalty,
               double* threshold_inout, int* n_iter_out, int* converged_out, double* rss_out, int rss_cap,
               int* n_rss_out);

/* K B as CSR (n x k, int64 indptr): its transpose is Z_ = B_.T @ K (cpu.py:641, get_archetype_matrix cpu.py:680).
 * Call once with null arrays to learn nnz, then with buffers. */
int sc_fit_get_KB(sc_fit* fit, long long* nnz_out, int64_t* indptr, int32_t* indices, double* data);
/* compressed results: idx/val [cols x S], cnt [cols]; column j of A_ is cell j, column a of B_ is archetype a */
int sc_fit_slots(sc_fit* fit); /* S */
int sc_fit_get_A(sc_fit* fit, int32_t* idx, double* val, int32_t* cnt);
int sc_fit_get_B(sc_fit* fit, int32_t* idx, double* val, int32_t* cnt);
/* get_hard_assignments (cpu.py:710-722): argmax over each column of A_, first index on ties */
int sc_fit_hard_assignments(sc_fit* fit, int32_t* labels /* [n] */);
/* get_hard_archetypes (cpu.py:724-729): B_.argmax(0) as cell positions */
int sc_fit_hard_archetypes(sc_fit* fit, int32_t* cells /* [k] */);
/* get_soft_assignments (cpu_dense.py:586-605): `top` rounds of argmax with masking; idx/w [n x top] */
int sc_fit_soft_assignments(sc_fit* fit, int top, int32_t* idx, double* w);
/* diagnostics: [0] cells that took the generic path in the last _updateA, [1] nnz(t2) of the last _updateB,
 * [2] archetypes in the cell-major hub block of the last _updateB, [3] rows (cumulative over the life of the handle)
 * that needed the dense-accumulator kernel of the sparse product, [4] (cell, archetype) gradients evaluated exactly in the last _updateB (all
 * Frank-Wolfe steps together; the rest of nnz(t2) x max_FW_iter was pruned by the lower bounds), [5] Frank-Wolfe
 * steps of _updateB timed so far and [6] their summed CUDA-event time in ns for the gradient launch group (row maxima when
 * not produced by the merge, both rounds of k_updateB_eval, k_updateB_hub and its reduction, the locality permutation at
 * step 0), [7] nnz of the last K*B row lists, [8] steps timed and [9] summed CUDA-event ns of k_updateB_hub alone,
 * [10] its algorithmic bytes per launch (24 B x cells x hubs + 12 B x delta entries), [11] 1024-candidate chunks of the last
 * _updateB that were discarded unread by the sorted-list bound and [12] chunk CTAs launched (all steps), [13] Frank-Wolfe
 * steps of the last _updateA whose zero scan went past archetype 31 and [14] the 32-wide batches they scanned.
 * out: 16 values. */
int sc_fit_stats(sc_fit* fit, long long* out16);
/* with SC_PROFILE=1 in the environment: accumulated wall seconds per stage (synchronising; development aid).
 * [0] B-by-cell [1] M*B [2] M*(M*B) [3] t1=B^T K B [4] _updateA shared-memory path [5] _updateA generic path
 * [6] A A^T [7] K A^T rows [8] candidates by archetype [9] _updateB gradient+argmin+step [10] RSS
 * ([0] push sources of K B, [1] K E_t of the Frank-Wolfe steps, [2] K B and the per-step merges since the products use the
 * explicit K) */
int sc_fit_profile(sc_fit* fit, double* out16);

#ifdef __cplusplus
}
#endif
#endif
