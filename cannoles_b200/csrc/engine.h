// engine.h -- host side of the single-system engine: owns the symbolic plan, the device
// buffers, the per-level launch lists and (on a GPU) the CUDA graphs that replay them.
#pragma once
#include <vector>

#include "b2_cuda.h"
#include "plan.h"
#include "symbolic.h"

namespace b2 {

enum LaunchKind : int {
  LK_FRONT_SMALL = 0,
  LK_ASSEMBLE_LARGE,
  LK_DIAG_WRITEBACK,
  LK_TRSM,
  LK_UPDATE,
  LK_FWD,
  LK_BWD,
  LK_FWD_BIG,
  LK_BWD_BIG,
  LK_FRONT_TINY,
  LK_FWD_TINY,
  LK_BWD_TINY,
  LK_DIAG,
  LK_DAG,
  LK_LINV,
};

struct Launch {
  int kind = 0;
  int cls = 0;        // thread-count class for per-front kernels (0:32 1:64 2:128 3:256)
  int64_t off = 0;    // offset into the device item buffer (in int32 units)
  int count = 0;      // items == CTAs
  int jb = 0;         // pivot block origin (tiled path)
  int mode = 0;       // k_update mode
  int smem = 0;       // dynamic shared memory (bytes)
  int level = 0;      // assembly-tree level (launches of one level are independent across branches)
  int branch = 0;     // kernel family: launches of one (level, branch) are ordered, branches run concurrently
  int flag = 0;       // k_trsm: diagonal block prefactored by k_diag; k_update: leave the next diagonal block alone; k_diag: first block
  int side = 0;       // launch on the side stream, concurrently with what follows on the chain
  int join_side = 0;  // wait for the side stream before this launch
};

struct Engine {
  Symbolic sym;
  int device = 0;
  cudaStream_t stream = nullptr;
  static constexpr int NBRANCH = 10;
  cudaStream_t bstream[NBRANCH] = {};   // side streams: the branches of one tree level run concurrently
  cudaEvent_t ev_fork = nullptr, ev_join[NBRANCH] = {};
  cudaStream_t sstream = nullptr;       // side stream of the tiled chain (look-ahead diagonal factorization)
  cudaEvent_t ev_sfork = nullptr, ev_sjoin = nullptr;
  // look-ahead factorization of the next diagonal block on a side stream (k_diag): measured SLOWER on
  // C4 (6.8 vs 5.5 ms: k_diag takes 40 us and the diag -> trsm chain stays serial), so off by default;
  // B2_LOOKAHEAD=1 turns it on
  bool lookahead = false;
  bool use_branches = true;
  // dataflow factorization of the tiled fronts (k_front_dag: one launch per tree level instead of a
  // k_trsm / k_update launch pair per 64-column pivot block); B2_DAG=0 selects the launch chain
  bool use_dag = true;
  int dag_level_max = 600;     // ... covering the highest run of tree levels with at most this many fronts each (B2_DAG_LEVEL_MAX)
  int dag_from_level = 0;      // first level of the dataflow launch (== nlevels: none)
  bool dag_sched = true;       // ticket order = simulated schedule (B2_DAG_SCHED=0: level by level, wave by wave)
  int64_t ndcnt = 0;           // counters of the dataflow launch (after the tile flags and the ticket)
  int64_t* d_sf_ptr = nullptr;
  int32_t* d_sf_ent = nullptr;
  int64_t flat_small_max = (int64_t)1 << 29;   // flat extend-add lists of the small fronts: at most this many elements (8 bytes each)
  int32_t *d_dfr = nullptr, *d_tl_ptr = nullptr, *d_tl_ent = nullptr, *d_fl_ptr = nullptr, *d_fl_ent = nullptr;
  int dag_excl_max = 0;        // k_front_dag launches with at most this many tasks run one CTA per SM (B2_DAG_EXCL_MAX)
  int dag_ctas = 0;            // CTAs of a k_front_dag launch (resident CTAs of the device)
  int64_t ntflag = 0;          // tile flags of all tiled fronts; the ticket counters of the launches follow them
  int ndag = 0;
  int* d_tflag = nullptr;
  int solve_fork = 1;    // fork the kernel families of a solve level (round 2: with the chained big-front solves shortened by the explicit inverses forking wins on C4 too, 1.54 -> 1.47 ms); -1: only when there are no big fronts; 0: never
  double small_max_m = 96;   // fronts up to this order take the shared-memory path (round 2 sweep on C4: 56: 3.42, 72: 3.32, 96: 3.22, 112: 3.33, 128: 3.58 ms)
  int tiny_max_m = 8;        // fronts up to this order (4 / 8 classes) take the one-thread-per-front kernels
  int tiny_solve_max_m = 8;  // ... and up to this order in the solves (round 2 sweep on C4: 8: 1.45, 16: 1.54, 32: 1.77 ms)
  // fronts of the multi-CTA solves with at least this many pivot blocks get explicit inverses of their
  // unit-lower 64 x 64 diagonal blocks (k_linv, at the end of the factorization): the in-block
  // substitution by one warp (2.4 k cycles on the chain of every block) becomes a mat-vec by the CTA
  int inv_min_blk = 2;
  bool update_tma = false;   // k_update<true>: operand tiles by TMA bulk copies (B2_UPDATE_TMA=1; measured slower on C3, see kernels.cuh)
  double solve_big_m = 96;   // fronts above this order take the multi-CTA solve kernels (measured: 96 < 192 < 384)

  // device buffers
  int32_t *d_slot_ptr = nullptr, *d_coo_sorted = nullptr;
  double *d_vals = nullptr, *d_nzval = nullptr;
  // speculative rho retry: the caller's values are uploaded to a second buffer on a copy stream and
  // compared with the previous upload while the shifted matrix is already being factorized
  double* d_vals2 = nullptr;
  int* d_mismatch = nullptr;
  int* h_mismatch = nullptr;
  cudaStream_t cstream = nullptr;
  cudaEvent_t ev_cmp = nullptr, ev_cfork = nullptr;
  int32_t *d_rho_slot = nullptr, *d_delta_slot = nullptr;
  double *d_rho_base = nullptr, *d_delta_base = nullptr;
  int32_t *d_scol = nullptr, *d_rowidx = nullptr, *d_rel = nullptr, *d_child_ptr = nullptr,
          *d_child_idx = nullptr, *d_amap_slot = nullptr, *d_amap_pos = nullptr, *d_perm = nullptr;
  int64_t *d_rptr = nullptr, *d_lptr = nullptr, *d_cbptr = nullptr, *d_uptr = nullptr,
          *d_amap_ptr = nullptr;
  double *d_Lx = nullptr, *d_CB = nullptr, *d_dvec = nullptr, *d_dstage = nullptr;
  int64_t *d_dsptr = nullptr, *d_asm_cptr = nullptr, *d_asm_off = nullptr, *d_sb_ptr = nullptr;
  int32_t *d_asm_ent = nullptr, *d_asm_rc = nullptr, *d_sb_src = nullptr, *d_sb_flag = nullptr, *d_linv_idx = nullptr;
  double* d_linv = nullptr;
  int64_t n_linv = 0;
  uint8_t* d_ug_row = nullptr;
  int32_t *d_ug_ptr = nullptr, *d_ug_src = nullptr;   // per-row gather of the children's update vectors (k_fwd, k_fwd_tiny)
  int64_t nsflag = 0;          // pivot blocks of the big fronts (0: no multi-CTA solves)
  double* d_ypub = nullptr;    // 2 N publication slots (forward y | backward x) of the multi-CTA solves
  int* d_stk = nullptr;        // one ticket counter per multi-CTA solve launch (Launch::jb): virtual block indices
  int n_stk = 0;
  int* d_flags = nullptr;
  unsigned long long* d_counts = nullptr;
  int32_t* d_items = nullptr;
  // solve
  double *d_x = nullptr, *d_upd = nullptr, *d_rhs = nullptr, *d_sol = nullptr, *d_res = nullptr,
         *d_out = nullptr, *d_part = nullptr;
  int64_t* d_Sp = nullptr;
  int32_t *d_Sj = nullptr, *d_Sslot = nullptr;
  // pinned host mirrors
  unsigned long long* h_counts = nullptr;  // [0..3] counts, [4] breakdown flag
  double* h_scalars = nullptr;

  PlanDev plan{};
  std::vector<Launch> fact_launches, fwd_launches, bwd_launches;
  int64_t n_small = 0, n_large = 0;
  double bytes_device = 0, t_plan = 0;
  bool have_vals = false, factored = false;
  double last_ms[5] = {0, 0, 0, 0, 0};
  cudaEvent_t ev[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  cudaEvent_t tev[2] = {nullptr, nullptr};  // b2_timer_start / b2_timer_stop
#ifndef B2_EMULATE
  cudaGraphExec_t g_fact = nullptr, g_fwdbwd = nullptr;
#endif
  bool use_graph = true;
  double refine_tol = 5e-13;  // > 0: stop refining as soon as ||K x - b|| / ||b|| <= refine_tol (north-star bar 1e-12); 0: always refine_steps sweeps
  int last_sweeps = 0;      // forward/backward sweeps used by the last solve
  std::vector<void*> registered;

  int init(int dev);
  void destroy();
  int build_plan();
  int run_factor_launches();
  int run_solve_launches();
  int launch_one(const Launch& L, cudaStream_t st);
  int run_list(const std::vector<Launch>& LL, bool allow_fork);
  int profile(int which, int max, int* kinds, int* cls, int* counts, double* ms, int* n);
  int assemble_and_factor(double eig_tol, int64_t* npos, int64_t* nzero, int64_t* nneg, int* breakdown,
                          bool do_assemble);
  int factorize_host(const double* vals, double eig_tol, int64_t* npos, int64_t* nzero, int64_t* nneg,
                     int* breakdown);
  int factorize_dev(const double* d_vals_in, double eig_tol, int64_t* npos, int64_t* nzero,
                    int64_t* nneg, int* breakdown);
  int refactorize_shift(double rho, double delta, double eig_tol, int64_t* npos, int64_t* nzero,
                        int64_t* nneg, int* breakdown);
  int factorize_retry(const double* vals, double rho, double eig_tol, int64_t* npos, int64_t* nzero,
                      int64_t* nneg, int* breakdown, int* speculation_held);
  int solve_core(const double* d_b, double* d_o, int negate, int refine_steps, double* relres);
  int solve_host(const double* rhs, double* out, int negate, int refine_steps, double* relres);
};

}  // namespace b2
