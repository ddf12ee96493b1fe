#pragma once
#include <vector>

#include "common.cuh"
#include "spgemm_rows.cuh"

// State of one SEACells model on one GPU.  See fit.cu for the layout.
struct sc_fit {
  sc_ctx* ctx = nullptr;
  int n = 0, k = 0, T = 50, S = 50;
  int row_begin = 0, row_end = 0;  // cells owned by this rank (whole range on one GPU)
  int slab = 0;                    // cells per rank (equal slabs; the buffers that are all-gathered are padded to world * slab)
  int exchange(void* buf, size_t bytes_per_rank);
  long long nchunks_cur = 0;
  bool in_update_B = false;
  DBuf<int> hub_ids, hub_slot, gram_hubof, perm;
  DBuf<unsigned long long> pk_in, pk_out;
  DBuf<double> gram_tiles;
  DBuf<double> T1H, T2H, HH, hub_part, AH, YH, ah_val;  // AH: hub slots of A, dense; YH = M AH; ah_*: A without them
  DBuf<int> ah_len, ah_key;  // HH: (K B t1)[cell, hubs], advanced step by step
  int n_hub = 0, n_hub_pad = 0, last_hub_block = 0;
  DBuf<double> part_out;  // [k][4] per-archetype partial result of one Frank-Wolfe step of _updateB on this rank
  // kernel matrix M
  DBuf<int> m_ptr, m_col;
  DBuf<double> m_val;
  long long m_nnz = 0;
  double m_fro2 = 0.0;
  bool have_M = false, have_A = false, have_B = false, a_is_slots = false, pre_valid = false;
  // compressed A (double buffered) and the initial-assignment lists
  DBuf<int> a_idx[2], a_cnt[2];
  DBuf<double> a_val[2];
  int a_cur = -1;
  DBuf<long long> slot_ptr_n;
  RowListsBuf a0;
  // K = M M^T, formed once per kernel matrix (first-occurrence row order), and the scratch of the push products
  DBuf<long long> m_ptr64;
  DBuf<int> m_len;
  RowListsBuf Kl;
  DBuf<int> ps_cnt, src_cell, src_key;
  DBuf<double> src_w, pv_in, pv_out;
  DBuf<long long> ps_off, run_heads, run_pos;
  DBuf<unsigned long long> pk_keys_in, pk_keys_out, ckeys;
  DBuf<unsigned> supp_bits, pk32_in, pk32_out, pi32_in, pi32_out;
  DBuf<int> flat_a;
  // compressed B
  DBuf<int> b_idx, b_cnt;
  DBuf<double> b_val;
  // K*B rows of this rank's cells; KBs: rows gathered from every rank (the support of B for t1, or all rows for Z_)
  RowListsBuf KB, KBs;
  // incremental K B inside one _updateB call: chosen cells of the last step, K E, second K B buffer
  DBuf<int> amin_cur;
  RowListsBuf DKB, KB2;
  // SC_R0_SEED = m > 0: after step 0, round 0 of the gradient kernel evaluates the previous winner + the m largest-t2
  // candidates instead of the 64 largest.  Measured at 1M cells: m = 16 triples the survivors of round 1 (fit 4.69 s vs
  // 3.97 s), m = 64 changes nothing - the previous winner is almost always among the 64 anyway.  Off by default.
  int r0_seed = 0;
  int hub_cap = 64;       // archetypes evaluated by the dense hub block at most (<= HUB_MAX; SC_HUB_MAX overrides)
  cudaStream_t stream2 = nullptr;  // the hub gradient of a Frank-Wolfe step runs here, next to the merge + chunk kernels
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  int hub_u = 2;   // cells in flight per warp of the hub kernel (SC_HUB_U=4: 125 registers, 2 CTAs per SM; 2: 80 registers, 3 CTAs)
  bool overlap_hub = true;         // SC_NO_OVERLAP=1: everything on one stream
  bool fresh_kb = false;  // SC_FIT_FRESH_KB=1: re-evaluate K B from B at every Frank-Wolfe step instead
  bool resum_A = false;   // SC_FIT_RESUM_A=1: _updateA re-sums (t1 A) over the active slots at every step
  DBuf<long long> b_off;
  DBuf<double> t1A, t1B;
  // _updateB precompute
  RowListsBuf MA, MAg, T2c;
  DBuf<long long> c_ptr, a_off, arow_ptr, chunk_ptr, t_off, cand_cnt;
  DBuf<unsigned long long> ck_in, ck_out;
  DBuf<unsigned> co_in, co_out;
  DBuf<int> c_cellv;
  DBuf<double> c_t2v;
  DBuf<int> chunk_arch, part_i, hub_list, arch_cnt_all;
  DBuf<double> part_v, cmax_v, bound_v, bound_all;
  DBuf<int> cmax_l;
  DBuf<unsigned char> hub_flag;
  DBuf<unsigned long long> eval_cnt, updA_diag;
  unsigned long long last_updA_diag[2] = {0, 0};
  long long last_evaluated = 0, last_pruned_chunks = 0, last_chunk_launches = 0;
  // CUDA-event timing of the dominant kernel (k_updateB_eval, both rounds of one Frank-Wolfe step), always on
  std::vector<cudaEvent_t> ev_a, ev_b, ev_h0, ev_h1;
  double hub_ms_total = 0.0, hub_bytes_total = 0.0;  // k_updateB_hub alone (incremental launches)
  long long hub_steps = 0, hub_launches_counted = 0, last_push_total = 0;
  bool last_push_sync_free = false;
  int k_maxrow = 0;  // longest row of this rank's columns of K
  double eval_ms_total = 0.0;
  long long eval_steps = 0;
  long long last_kb_nnz = 0;
  ~sc_fit();
  DBuf<int> c_cell;
  DBuf<double> c_t2;
  DBuf<unsigned long long> gk_in, gk_out;
  DBuf<double> gv_in, gv_out;
  // scratch
  DBuf<int> cnt_n, cnt_k, slow_list, slow_cnt;
  DBuf<double> percell, scalars;
  SpgemmScratch sp;
  DBuf<char> tmp;
  // diagnostics
  bool profile = false;      // SC_PROFILE=1: synchronise around every stage and accumulate wall seconds
  double prof[16] = {0};
  double prof_t0 = 0.0;
  void pstart();
  void pstop(int stage);
  int last_slow_cells = 0;
  long long last_t2_nnz = 0;

  int init(sc_ctx* c, int n_, int k_, int T_);
  int set_kernel(const int* indptr, const int* indices, const double* data, long long nnz);
  int set_B_onehot(const long long* archetypes_host);
  int set_B_slots(const int* idx, const double* val, const int* cnt);
  int set_A_lists(const long long* colptr, const int* idx, const double* val);
  int set_A_slots(const int* idx, const double* val, const int* cnt);
  RowLists A_view() const;
  int push_product(const struct PushSrc& src, bool has_dups, RowListsBuf& Z);
  int allgather_rows(const RowLists& L, const unsigned* keep, RowListsBuf& G);
  int compute_KB(bool with_support);
  int advance_KB_delta();
  int advance_KB_merge(int t);
  int launch_hub(int t, cudaStream_t stream, double* partial_dev);
  int build_perm();
  int update_B_begin();
  int update_B_eval(int t, double* partial_dev);
  int update_B_step(int t, const double* gathered_dev, int n_ranks, int* trace_dev);
  int update_B_end();
  int precompute_A();
  int update_A(double l2, int* trace_dev);
  int update_B(int* trace_dev);
  int rss(double* out_host);
  int run(int max_iter, int min_iter, double eps, double l2, double* threshold_inout, int* n_iter_out,
          int* converged_out, double* rss_out, int rss_cap, int* n_rss_out);
  int get_KB_csr(long long* nnz_out, long long* indptr, int* indices, double* data);
  int get_A(int* idx, double* val, int* cnt);
  int get_B(int* idx, double* val, int* cnt);
  int hard_assignments(int* labels);
  int hard_archetypes(int* cells);
  int soft_assignments(int top, int* idx, double* w);
};
