/*
 * cannoles_b200.h -- C ABI of libcannoles_b200.so, the B200 (sm_100a) `linsolve` backend for the
 * KKT factor/solve path of CaNNOLeS.jl.
 *
 * Each entry point replaces one step of the reference's linear-solver struct interface
 * (reference/src/solver_types.jl); INTEGRATION.md shows the Julia `ccall` glue.
 *
 *   b2_analyze            <-  LDLFactStruct(N, rows, cols, vals) ctor: `sparse`+`triu`+`ldl_analyze`
 *                             (src/solver_types.jl:61-65, called at src/CaNNOLeS.jl:327)
 *   b2_factorize          <-  try_to_factorize: set_vals! + ldl_factorize! + inertia loop
 *                             (src/solver_types.jl:79-98; set_vals! :53-59)
 *   b2_refactorize_shift  <-  the rho / delta retry of newton_system! (src/CaNNOLeS.jl:1029-1043),
 *                             which in the reference is a full try_to_factorize again
 *   b2_solve              <-  solve_ldl!: ldiv! then negate (src/solver_types.jl:69-77)
 *   b2_free               <-  Julia finalizer of the struct
 *   b2b_*                 <-  the same verbs for a batch of independent KKT systems sharing one
 *                             sparsity pattern (no counterpart in the reference: multi-start /
 *                             per-sample estimation, BASELINE.json config 5)
 *
 * Conventions: every function returns 0 on success and a negative value on a runtime error
 * (CUDA failure, malformed input); `b2_last_error()` describes it.  NUMERICAL failure (wrong
 * inertia, zero pivot) is NOT an error: it is reported through the out-parameters, as the
 * reference reports it through the Bool of try_to_factorize (src/solver_types.jl:96-97).
 * All host pointers are plain C arrays owned by the caller; indices are 1-based Int64 exactly as
 * CaNNOLeS builds them (src/CaNNOLeS.jl:276-315); values are Float64.  The library is not
 * re-entrant per handle (the reference drives it from one thread).  There is no CPU fallback:
 * without a CUDA device every call fails.
 */
#ifndef CANNOLES_B200_H
#define CANNOLES_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct b2_handle b2_handle;   /* one KKT system: symbolic plan + device factor */
typedef struct b2b_handle b2b_handle; /* a batch of systems with one shared pattern      */

/* orderings for b2_analyze / b2b_analyze */
#define B2_ORDER_ND 0      /* nested dissection (default): of the compressed x-vertex graph (H + J'J
                              pattern) when the trailing block of K is diagonal, as in a KKT matrix */
#define B2_ORDER_NATURAL 1 /* identity                                          */
#define B2_ORDER_USER 2    /* user_perm[k] = 0-based original index of pivot k  */
#define B2_ORDER_AMD 3     /* approximate minimum degree                        */
#define B2_ORDER_ND_RAW 4  /* nested dissection of the raw N-vertex graph       */

typedef struct {
  int64_t N, nnz, nnzA;           /* order, COO entries, distinct upper-triangular positions   */
  int64_t nnzL;                   /* exact off-diagonal nonzeros of L (no padding)             */
  int64_t nnzL_store, cb_store;   /* doubles held by the factor panels / contribution blocks   */
  int64_t nsuper, nlevels;        /* fronts (supernodes) and assembly-tree levels              */
  int64_t max_front, max_width;   /* largest front order and pivot-block width                 */
  int64_t n_small, n_large;       /* fronts on the shared-memory path / the tiled path         */
  int64_t launches_factor, launches_solve; /* kernel launches per factorize / per solve        */
  double flops;                   /* sum_j (c_j^2 + 3 c_j), exact symbolic       */
  double flops_store;             /* flops executed on the dense fronts (with padding)         */
  double t_order, t_symbolic, t_plan; /* seconds spent in ordering / symbolic / device set-up  */
  double bytes_device;            /* device memory held by the handle                          */
} b2_stats_t;

const char* b2_last_error(void);
int b2_version(void);
int b2_device_count(void);

/* Symbolic analysis, once per solver.  rows1/cols1: COO lower triangle (rows1[t] >= cols1[t]),
 * 1-based, duplicates allowed; a strictly-upper entry is rejected (the reference's `triu`
 * would silently drop it).  nvar/nequ/ncon give the block sizes (N = nvar+nequ+ncon) so that the
 * trailing -delta / rho diagonal segments of the COO layout can be shifted on the device. */
int b2_analyze(int64_t N, int64_t nnz, const int64_t* rows1, const int64_t* cols1, int64_t nvar,
               int64_t nequ, int64_t ncon, int ordering, const int64_t* user_perm, int device,
               b2_handle** out);

/* try_to_factorize: upload vals (nnz doubles, COO order), COO->CSC accumulate, numeric LDL^T
 * without pivoting, inertia reduction on the device.  npos = #{d > eig_tol},
 * nzero = #{|d| <= eig_tol}, nneg = #{d < -eig_tol} (NaN pivots are in none of the three);
 * breakdown = 1 if an exactly zero pivot was met.  Success in the reference's sense is
 * npos == nvar && nzero == 0. */
int b2_factorize(b2_handle* h, const double* vals, double eig_tol, int64_t* npos, int64_t* nzero,
                 int64_t* nneg, int* breakdown);

/* Retry with new regularisation without re-uploading vals: the rho segment (last nvar COO
 * entries) becomes rho, and if delta is not NaN the delta segment becomes -delta; the CSC values
 * are bit-identical to a full b2_factorize of the edited vals.  Requires the canonical COO
 * layout (b2_analyze detects it; otherwise returns an error and the caller re-uploads). */
int b2_refactorize_shift(b2_handle* h, double rho, double delta_or_nan, double eig_tol,
                         int64_t* npos, int64_t* nzero, int64_t* nneg, int* breakdown);

/* try_to_factorize as newton_system! re-enters it for a rho retry (src/CaNNOLeS.jl:1032, :1039: the
 * same vals buffer with only the trailing rho segment rewritten).  SAME RESULT as
 * b2_factorize(h, vals, ...) bit for bit, whatever the caller did to vals: the matrix shifted by
 * `rho` on the device is factorized at once while vals is uploaded on a copy stream and compared
 * on the device with the previous upload; if anything but the rho segment changed (or the segment
 * is not the constant `rho`) the factorization is redone from the fresh upload.
 * *speculation_held (may be NULL) = 1 if the shifted factorization was the answer. */
int b2_factorize_retry(b2_handle* h, const double* vals, double rho, double eig_tol, int64_t* npos,
                       int64_t* nzero, int64_t* nneg, int* breakdown, int* speculation_held);

/* solve_ldl!: d_out = (negate ? -1 : +1) * K^{-1} rhs with the last factorization.
 * refine_steps >= 0 iterative-refinement sweeps with the assembled K; if relres != NULL it
 * receives ||K x - rhs||_2 / ||rhs||_2 of the returned (un-negated) solution. */
int b2_solve(b2_handle* h, const double* rhs, double* d_out, int negate, int refine_steps,
             double* relres);
/* forward/backward sweeps the last solve used (1 + refinement sweeps actually taken); with
 * b2_set_option(h, "refine_tol", tol): refinement stops as soon as the residual is <= tol (default 5e-13;
 * 0: always refine_steps sweeps) */
int b2_last_sweeps(const b2_handle* h);

/* Device-resident variants (inputs/outputs already in HBM); used to time the kernels alone. */
int b2_factorize_dev(b2_handle* h, const double* d_vals, double eig_tol, int64_t* npos,
                     int64_t* nzero, int64_t* nneg, int* breakdown);
int b2_solve_dev(b2_handle* h, const double* d_rhs, double* d_out, int negate, int refine_steps,
                 double* relres);

/* Pin caller-owned host buffers (vals, rhs, d have stable addresses in the reference:
 * src/CaNNOLeS.jl:241-243, 276-279) so uploads run at full PCIe speed. */
int b2_register_host(b2_handle* h, void* ptr, size_t bytes);
int b2_unregister_host(b2_handle* h, void* ptr);

int b2_stats(const b2_handle* h, b2_stats_t* out);
/* device milliseconds of the phases of the last call: [0] upload, [1] COO->CSC assembly,
 * [2] numeric factorization (+ inertia), [3] solve (+ refinement), [4] download */
int b2_last_timings(const b2_handle* h, double* ms5);
/* pivot-block width, order (rows) and assembly-tree level of every front (supernode-size
 * histogram for the roofline discussion); arrays of b2_stats().nsuper entries */
int b2_front_sizes(const b2_handle* h, int64_t max, int32_t* width, int32_t* order, int32_t* level);
/* Developer aid: replay the factorization (which = 0) or one forward+backward sweep
 * (which = 1) launch by launch outside the CUDA graph, an event after every launch; returns per
 * launch the kernel kind, its class / mode, the CTA count and the warm-cache device time. */
int b2_profile(b2_handle* h, int which, int max, int* kinds, int* cls, int* counts, double* ms, int* n);
/* CUDA events on the handle's own stream (bench.py times K steps between the two; a
 * torch.cuda.Event would only see torch's current stream) */
int b2_timer_start(b2_handle* h);
int b2_timer_stop(b2_handle* h, double* ms);
/* inspection (tests): permutation, assembled CSC values / pattern, pivots in pivot order */
int b2_get_perm(const b2_handle* h, int64_t* perm0);
int b2_get_csc(const b2_handle* h, int64_t* colptr0, int64_t* rowval0);
int b2_get_nzval(b2_handle* h, double* nzval);
int b2_get_d(b2_handle* h, double* d);
int b2_set_option(b2_handle* h, const char* key, double value);
int b2_free(b2_handle* h);

/* ---- batch of independent KKT systems with one pattern (config 5) --------------------- */
int b2b_analyze(int64_t N, int64_t nnz, const int64_t* rows1, const int64_t* cols1, int64_t nvar,
                int64_t nequ, int64_t ncon, int64_t batch, int ordering, const int64_t* user_perm,
                int device, b2b_handle** out);
/* vals: batch x nnz (instance-major).  active == NULL means all instances; otherwise instances
 * with active[b] == 0 are skipped (their factor and outputs are left untouched). */
int b2b_factorize(b2b_handle* h, const double* vals, const uint8_t* active, double eig_tol,
                  int64_t* npos, int64_t* nzero, int64_t* nneg, int32_t* breakdown);
/* per-instance rho (and delta unless NULL); instances with active[b] == 0 are skipped */
int b2b_refactorize_shift(b2b_handle* h, const double* rho, const double* delta_or_null,
                          const uint8_t* active, double eig_tol, int64_t* npos, int64_t* nzero,
                          int64_t* nneg, int32_t* breakdown);
/* rhs, d_out: batch x N */
int b2b_solve(b2b_handle* h, const double* rhs, double* d_out, const uint8_t* active, int negate);
int b2b_factorize_dev(b2b_handle* h, const double* d_vals, const uint8_t* d_active, double eig_tol,
                      int64_t* d_counts4 /* batch x 4: pos, zero, neg, breakdown */);
int b2b_solve_dev(b2b_handle* h, const double* d_rhs, double* d_out, const uint8_t* d_active,
                  int negate);
/* Fused verb: factorize every active instance and, where the inertia is the expected one
 * (npos == nvar && nzero == 0), solve right away from the factor still resident in shared
 * memory (d_out = (negate ? -1 : +1) K^{-1} rhs).  Instances whose inertia is wrong leave their
 * slice of d_out untouched; the caller retries them with b2b_refactorize_shift + b2b_solve, as
 * newton_system! does one system at a time (src/CaNNOLeS.jl:1029-1049). */
int b2b_factor_solve(b2b_handle* h, const double* vals, const double* rhs, double* d_out,
                     const uint8_t* active, double eig_tol, int negate, int64_t* npos, int64_t* nzero,
                     int64_t* nneg, int32_t* breakdown);
int b2b_factor_solve_dev(b2b_handle* h, const double* d_vals, const double* d_rhs, double* d_out,
                         const uint8_t* d_active, double eig_tol, int negate, int store_factor,
                         int64_t* d_counts4);
int b2b_stats(const b2b_handle* h, b2_stats_t* out);
int b2b_last_ms(const b2b_handle* h, double* ms);    /* device time of the last kernel launch */
int b2b_timer_start(b2b_handle* h);
int b2b_timer_stop(b2b_handle* h, double* ms);
int b2b_get_perm(const b2b_handle* h, int64_t* perm0);
int b2b_get_d(b2b_handle* h, int64_t instance, double* d); /* pivots of one stored factor */

/* ---- device-resident CaNNOLeS loop for a batch of small dense instances -----------------
 * Replaces, per instance and with no host round trip, what `solve!` does around the linear
 * solver (src/CaNNOLeS.jl:418-864): prepare_newton_system! (:947-981),
 * newton_system! (:1008-1052), the trial-point products and norms (:508, 521-525, 722-732,
 * 753-755; N2), the CGLS multiplier estimate (:513, 872-897; N3), the extrapolation step, the
 * line search (:1054-1112) and get_status.  The handle is the one b2b_analyze made for the
 * Newton-mode KKT layout of the model.  Model: DenseBatchNLS,
 *   F(x) = A x + 0.1 sin(B x) - y,  c(x) = C x + 0.05 (x.x)[0:ncon] - e = 0,
 * A, B (m x n) and C (ncon x n) stored column-major per instance. */
typedef struct {
  double eig_tol, delta_min, kappa_dec, kappa_inc, kappa_largeinc, rho0, rho_max, rho_min, gamma_A; /* ParamCaNNOLeS :36-87 */
  double atol, rtol, Fatol, Frtol, delta_dec;   /* keyword arguments of solve! (:418-436) */
  double cgls_tol;                              /* Krylov.cgls atol = rtol */
  int32_t max_iter, max_eval, max_inner, always_accept_extrapolation, use_initial_multiplier, reserved;
} b2_nls_params_t;
void b2_nls_default_params(b2_nls_params_t* p);
typedef struct {
  int64_t n, m, ncon;
  int64_t shared_model;   /* != 0: one (A, B, C, y, e) for all instances, only x0 varies (multi-start) */
  const double *At, *Bt;  /* per instance n x m:    At[j * m + i] = A[i][j] */
  const double* Ct;       /* per instance n x ncon: Ct[j * ncon + k] = C[k][j] */
  const double *y, *e;    /* m, ncon */
  const double* x0;       /* n per instance (always per instance) */
  const double* y0;       /* ncon per instance or NULL (_dev verb only; used with use_initial_multiplier) */
} b2_dense_nls_t;
/* doubles per instance record: status, iter, nfact, nlinsolve, nbk, neval_residual, neval_cons,
 * objective, primal_feas, dual_feas, rho, delta, x[n], lambda[ncon].  status: 0 unknown, 1 first_order,
 * 2 small_residual, 3 stalled, 4 exception, 5 max_eval, 6 max_time, 7 max_iter; 8 / 9 / 10 = the
 * errors the reference throws (NaN at x0 :484-487, Dphi >= 0 :1085, alpha too small :1097) */
int64_t b2b_nls_record_len(const b2b_handle* h);
int b2b_nls_dense_solve_dev(b2b_handle* h, const b2_dense_nls_t* model_dev, int64_t count,
                            const b2_nls_params_t* params, double* d_records, double* d_dbg_vals);
int b2b_nls_dense_solve(b2b_handle* h, const b2_dense_nls_t* model_host, int64_t count,
                        const b2_nls_params_t* params, double* records, int64_t chunk);
/* Asynchronous form: queue one batch (upload, solve, download in stream order on one of the
 * handle's 128 lanes) and return; b2b_nls_wait blocks until every queued batch is complete.
 * Batches in flight overlap: PCIe of one with the SMs of another, and the SMs a batch leaves idle
 * while its slowest instance finishes are taken by the next batch.  where = 0: *model and records
 * are host memory (pin them); where = 1: device memory.  Arrays stay untouched until the wait. */
int b2b_nls_dense_submit(b2b_handle* h, const b2_dense_nls_t* model, int64_t count,
                         const b2_nls_params_t* params, double* records, int where);
int b2b_nls_wait(b2b_handle* h);
int b2b_free(b2b_handle* h);

/* device utilities used by bench.py / tests (plain cudaMalloc / cudaMemcpy wrappers so that
 * callers need no CUDA binding of their own) */
int b2_dev_malloc(void** dptr, size_t bytes);
int b2_dev_free(void* dptr);
int b2_dev_upload(void* dptr, const void* hptr, size_t bytes);
int b2_dev_download(void* hptr, const void* dptr, size_t bytes);
int b2_dev_sync(void);
/* pin / unpin a caller-owned host buffer without a handle (batched values, right-hand sides) */
int b2_host_register(void* ptr, size_t bytes);
int b2_host_unregister(void* ptr);
/* FP64 GEMM throughput of this library's own tile kernel and of cuBLAS DGEMM (TFLOP/s), the
 * roofline denominator for the frontal updates (MEASURED_PEAKS.json has no FP64 entry). */
int b2_measure_dgemm(int n, int reps, double* tflops_own, double* tflops_cublas);
int b2_measure_hbm(size_t bytes, int reps, double* gbs_copy);

#ifdef __cplusplus
}
#endif
#endif /* CANNOLES_B200_H */
