// nls_kernels.cuh -- the whole CaNNOLeS iteration (reference/src/CaNNOLeS.jl:418-864) of one small
// dense constrained NLS instance inside ONE CTA, for batches of independent instances that share a
// KKT sparsity pattern (BASELINE.json config 5: multi-start / per-sample estimation, n = 64, m = 128,
// 16 constraints -> N = 208).  SURVEY.md 8(e) "per-instance loop state lives on device; converged
// instances are masked out" becomes: a CTA takes an instance from a ticket counter, runs it to its
// final status and writes one fixed-size record; nothing crosses PCIe or NVLink inside the loop.
//
//   rows of SURVEY 8 that run here, per instance and per inner iteration:
//   N1  prepare_newton_system! (:947-981): the seven COO value segments are produced on the device
//       (residual Hessian B' diag(w) B in FP64 tensor-core tiles, Jacobians, -delta, rho) into the
//       CTA's private slice of `vals`, which never exists on the host
//   a3-a7  assembly, LDL', inertia, rho retries (newton_system! :1008-1052), solve: batched_instance()
//   N2  J'v products, dual / primal residuals, infinity norms (:508, 521-525, 722-732, 753-755)
//   N3  CGLS multiplier estimate (:513, 872-897; Krylov.jl defaults)
//   plus the extrapolation step, the line search (:1054-1112) and get_status.
//
// The model is DenseBatchNLS of cannoles_b200/models.py:
//   F(x) = A x + 0.1 sin(B x) - y  (m),   c(x) = C x + 0.05 (x.x)[0:ncon] - e  (ncon), lcon = 0.
// A, B, C are stored column-major (rows fastest: "At" = numpy A.T, C-contiguous) so that a warp
// reads consecutive rows of one column.
//
// All threads of the CTA execute the same control flow: every scalar the loop branches on is the
// result of a fixed-order CTA reduction that every thread computes identically.
#pragma once
#include "batched_kernels.cuh"

namespace b2 {

// ParamCaNNOLeS (:36-87) + the keyword arguments of solve! (:418-436)
struct NlsParams {
  double eig_tol, delta_min, kappa_dec, kappa_inc, kappa_largeinc, rho0, rho_max, rho_min, gamma_A;
  double atol, rtol, Fatol, Frtol, delta_dec, cgls_tol, eps2;
  int max_iter, max_eval, max_inner, always_accept_extrapolation, use_initial_multiplier;
};

struct DenseNlsModel {
  int n, m, ncon;
  long long stride_A, stride_C, stride_y, stride_e, stride_x0;   // per-instance strides (0 = shared by all instances)
  const double* At;   // (n x m) per instance: At[j * m + i] = A[i][j]
  const double* Bt;
  const double* Ct;   // (n x ncon): Ct[j * ncon + k] = C[k][j]
  const double* y;    // m
  const double* e;    // ncon
  const double* x0;   // n
  const double* y0;   // ncon or NULL (zeros)
};

// record layout (doubles): status, iter, nfact, nlinsolve, nbk, neval_residual, neval_cons,
// objective, primal_feas, dual_feas, rho, delta, x[n], lambda[ncon]
constexpr int NLS_REC_HEAD = 12;
enum NlsStatus { NLS_UNKNOWN = 0, NLS_FIRST_ORDER = 1, NLS_SMALL_RESIDUAL = 2, NLS_STALLED = 3, NLS_EXCEPTION = 4,
                 NLS_MAX_EVAL = 5, NLS_MAX_TIME = 6, NLS_MAX_ITER = 7,
                 NLS_ERR_NAN_INIT = 8,   // "Initial point gives Inf or Nan" (:484-487)
                 NLS_ERR_DPHI = 9,       // @assert Dphi < 0 (:1085)
                 NLS_ERR_ALPHA = 10 };   // "alpha too small" (:1097)

inline size_t nls_state_doubles(int n, int m, int nc, int NT) {
  return 6 * (size_t)n + 9 * (size_t)m + 9 * (size_t)nc + 2 * (size_t)(n + m + nc) + 2 * (size_t)NT + 64;
}

// ------------------------------------------------------------------------------------------
// CTA reductions with a fixed order (identical result in every thread)
template <int NT>
__device__ __forceinline__ double cta_sum(double v, double* red) {
  B2_UNROLL
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  double s = 0.0;
  B2_UNROLL
  for (int w = 0; w < NT / 32; w++) s += red[w];
  return s;
}
__device__ __forceinline__ double nan_max(double a, double b) {   // NaN-propagating max, as norm(., Inf)
  return (a != a) ? a : ((b != b) ? b : (a > b ? a : b));
}
template <int NT>
__device__ __forceinline__ double cta_max(double v, double* red) {
  B2_UNROLL
  for (int o = 16; o > 0; o >>= 1) v = nan_max(v, __shfl_xor_sync(0xffffffffu, v, o));
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  double s = red[0];
  B2_UNROLL
  for (int w = 1; w < NT / 32; w++) s = nan_max(s, red[w]);
  return s;
}
template <int NT>
__device__ __forceinline__ double cta_dot(const double* a, const double* b, int len, double* red) {
  double s = 0.0;
  for (int i = threadIdx.x; i < len; i += NT) s += a[i] * b[i];
  return cta_sum<NT>(s, red);
}
template <int NT>
__device__ __forceinline__ double cta_absmax(const double* a, int len, double* red) {
  double s = 0.0;
  for (int i = threadIdx.x; i < len; i += NT) s = nan_max(s, fabs(a[i]));
  return cta_max<NT>(s, red);
}
template <int NT>
__device__ __forceinline__ double cta_abssum(const double* a, int len, double* red) {
  double s = 0.0;
  for (int i = threadIdx.x; i < len; i += NT) s += fabs(a[i]);
  return cta_sum<NT>(s, red);
}
template <int NT>
__device__ __forceinline__ bool cta_nonfinite(const double* a, int len) {   // check_nan_inf (:902-909)
  int bad = 0;
  for (int i = threadIdx.x; i < len; i += NT) bad |= !isfinite(a[i]);
  return __syncthreads_or(bad) != 0;
}

// ------------------------------------------------------------------------------------------
// TMA (bulk asynchronous copy, global -> shared, completion on an mbarrier): the model matrices A and
// B of the instance (n columns of m doubles each, 2 x 64 KB for config 5) are staged into the shared
// memory of the packed KKT triangle while that triangle is dead -- from the end of a solve to the
// next assembly, which is exactly when the model is evaluated (trial residual, J'v products, the
// Jacobian and Hessian segments of the next system).  One warp issues one cp.async.bulk per column
// (a padded column stride keeps the FP64 tensor-core fragment loads of the Hessian conflict-free),
// nobody holds the data in registers on the way, and every evaluator reads shared memory instead of
// going back to L2 five or six times per Newton system.

// ------------------------------------------------------------------------------------------
// model evaluators (all CTA-collective; they end with a barrier)
struct DenseInst {
  int n, m, nc;
  const double *At, *Bt, *Ct, *y, *e;   // the instance's arrays in HBM
  const double *Aw, *Bw;                // A and B as the evaluators read them: staged in shared memory (or = At, Bt)
  const double* Cw;                     // C likewise (column stride nc either way)
  int ldw;                              // column stride of Aw / Bw
};

// Stage A and B of the instance into `dstA` / `dstB` (shared memory, column stride lds).  CTA-collective.
// `phase` is the mbarrier parity of this use (the caller flips it after every call).
template <int NT>
__device__ __forceinline__ bool dn_stage_model(DenseInst& I, double* dstA, double* dstB, double* dstC, int lds,
                                               unsigned long long* bar, uint32_t phase) {
  const int n = I.n, m = I.m, nc = I.nc, tid = threadIdx.x;
  __syncthreads();                      // every generic-proxy access to the destination is complete
#ifndef B2_EMULATE
  const bool tma_ok = (m % 2 == 0) && (lds % 2 == 0) && ((n * nc) % 2 == 0) &&
                      ((reinterpret_cast<uintptr_t>(I.At) | reinterpret_cast<uintptr_t>(I.Bt) | reinterpret_cast<uintptr_t>(I.Ct)) % 16 == 0);
  if (tma_ok) {
    if (tid < 32) {
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      if (tid == 0) {
        mbar_expect_tx(bar, (uint32_t)((2 * n * m + n * nc) * sizeof(double)));
        if (nc > 0) tma_bulk_g2s(dstC, I.Ct, (uint32_t)(n * nc * sizeof(double)), bar);
      }
      __syncwarp();
      for (int j = tid; j < n; j += 32) {
        tma_bulk_g2s(dstA + (size_t)j * lds, I.At + (size_t)j * m, (uint32_t)(m * sizeof(double)), bar);
        tma_bulk_g2s(dstB + (size_t)j * lds, I.Bt + (size_t)j * m, (uint32_t)(m * sizeof(double)), bar);
      }
    }
    mbar_wait(bar, phase);
    I.Aw = dstA; I.Bw = dstB; I.Cw = dstC; I.ldw = lds;
    return true;                        // the mbarrier went through one phase
  }
#endif
  {
    for (int q = tid; q < n * m; q += NT) {
      const int j = q / m, i = q - j * m;
      dstA[(size_t)j * lds + i] = I.At[q];
      dstB[(size_t)j * lds + i] = I.Bt[q];
    }
    for (int q = tid; q < n * nc; q += NT) dstC[q] = I.Ct[q];
    __syncthreads();
  }
  I.Aw = dstA; I.Bw = dstB; I.Cw = dstC; I.ldw = lds;
  return false;
}

// Fx = A x + 0.1 sin(B x) - y; sn = sin(B x), cs = cos(B x) are kept: they define J(x) and H(x)
template <int NT>
__device__ __forceinline__ void dn_residual(const DenseInst& I, const double* x, double* Fx, double* sn, double* cs,
                                            double* scr) {
  const int m = I.m, n = I.n, tid = threadIdx.x;
  const int npart = NT / m;                 // column subsets summed separately (m <= NT checked on the host)
  const int p = tid / m, i = tid - p * m;
  double sa = 0.0, sb = 0.0;
  if (p < npart) {
#pragma unroll 4
    for (int j = p; j < n; j += npart) {
      const double xj = x[j];
      sa += I.Aw[(size_t)j * I.ldw + i] * xj;
      sb += I.Bw[(size_t)j * I.ldw + i] * xj;
    }
  }
  scr[tid] = sa;
  scr[NT + tid] = sb;
  __syncthreads();
  if (tid < m) {
    double a = 0.0, b = 0.0;
    for (int q = 0; q < npart; q++) { a += scr[q * m + tid]; b += scr[NT + q * m + tid]; }
    double s, c;
    sincos(b, &s, &c);
    sn[tid] = s;
    cs[tid] = c;
    Fx[tid] = a + 0.1 * s - I.y[tid];
  }
  __syncthreads();
}

// cx = C x + 0.05 x[0:nc]^2 - e   (lcon = 0)
template <int NT>
__device__ __forceinline__ void dn_cons(const DenseInst& I, const double* x, double* cx, double* scr) {
  const int nc = I.nc, n = I.n, tid = threadIdx.x;
  if (nc == 0) return;
  const int npart = NT / nc;
  const int p = tid / nc, k = tid - p * nc;
  double s = 0.0;
  if (p < npart)
    for (int j = p; j < n; j += npart) s += I.Cw[(size_t)j * nc + k] * x[j];
  scr[tid] = s;
  __syncthreads();
  if (tid < nc) {
    double a = 0.0;
    for (int q = 0; q < npart; q++) a += scr[q * nc + tid];
    cx[tid] = a + 0.05 * (x[tid] * x[tid]) - I.e[tid];
  }
  __syncthreads();
}

// out = J(x)' v,  J = A + 0.1 diag(cos(B x)) B : a warp per column, fixed-order shuffle tree
template <int NT>
__device__ __forceinline__ void dn_jtprod_res(const DenseInst& I, const double* cs, const double* v, double* out) {
  const int m = I.m, n = I.n, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int j = warp; j < n; j += NT / 32) {
    const double* a = I.Aw + (size_t)j * I.ldw;
    const double* b = I.Bw + (size_t)j * I.ldw;
    double s = 0.0;
    for (int i = lane; i < m; i += 32) s += (a[i] + 0.1 * cs[i] * b[i]) * v[i];
    B2_UNROLL
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) out[j] = s;
  }
  __syncthreads();
}

// Jc(x)[k][j] = C[k][j] + (k == j) 0.1 x[k]
__device__ __forceinline__ double dn_jc(const DenseInst& I, const double* x, int k, int j) {
  const double c = I.Cw[(size_t)j * I.nc + k];
  return k == j ? c + 0.1 * x[k] : c;
}
// out(n) = Jc(x)' lam
template <int NT>
__device__ __forceinline__ void dn_jtprod_cons(const DenseInst& I, const double* x, const double* lam, double* out) {
  const int nc = I.nc, n = I.n;
  for (int j = threadIdx.x; j < n; j += NT) {
    double s = 0.0;
    for (int k = 0; k < nc; k++) s += dn_jc(I, x, k, j) * lam[k];
    out[j] = s;
  }
  __syncthreads();
}
// out(nc) = Jc(x) v
template <int NT>
__device__ __forceinline__ void dn_jprod_cons(const DenseInst& I, const double* x, const double* v, double* out,
                                              double* scr) {
  const int nc = I.nc, n = I.n, tid = threadIdx.x;
  if (nc == 0) return;
  const int npart = NT / nc;
  const int p = tid / nc, k = tid - p * nc;
  double s = 0.0;
  if (p < npart)
    for (int j = p; j < n; j += npart) s += dn_jc(I, x, k, j) * v[j];
  scr[tid] = s;
  __syncthreads();
  if (tid < nc) {
    double a = 0.0;
    for (int q = 0; q < npart; q++) a += scr[q * nc + tid];
    out[tid] = a;
  }
  __syncthreads();
}

// prepare_newton_system! (:947-981) for this model and the Newton (exact residual Hessian) mode:
// the value segments S1..S7 of SURVEY App. B in the CTA's private slice `vs` of the COO values.
//   S1 hess_coord_residual(x, r) = lower(B' diag(w) B), w = -0.1 sin(B x) . r      (n(n+1)/2, column-major)
//   S2 -hess_coord(x, lam; obj_weight = 0): -(0.1 lam_j) on the diagonal entries j < ncon, -0 elsewhere
//   S3 Jx column-major (n m)       S4 Jc column-major (n ncon)
//   S5 -1 (constant, written once by the caller)       S6 -delta       S7 0 (rho goes in through the override)
template <int NT>
__device__ __forceinline__ void dn_fill_vals(const DenseInst& I, const double* x, const double* lam, const double* r,
                                             const double* sn, const double* cs, double delta, double* wv,
                                             double* vs) {
  const int n = I.n, m = I.m, nc = I.nc, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int nh = n * (n + 1) / 2;
  for (int i = tid; i < m; i += NT) wv[i] = -0.1 * sn[i] * r[i];
  __syncthreads();
  {  // S1 in 8 x 8 x 4 FP64 tensor-core tiles: H(a, b) = sum_i (B(i, a) w_i) B(i, b), lower tiles only
    const int g = lane >> 2, t4 = lane & 3;
    const int nt = (n + 7) >> 3, ntiles = nt * (nt + 1) / 2;
    for (int t = warp; t < ntiles; t += NT / 32) {
      int tb = 0, rem = t;
      while (rem >= nt - tb) { rem -= nt - tb; tb++; }
      const int a0 = (tb + rem) * 8, b0 = tb * 8;
      const int ra = a0 + g, cb = b0 + g;
      const double* pa = I.Bw + (size_t)(ra < n ? ra : 0) * I.ldw;
      const double* pb = I.Bw + (size_t)(cb < n ? cb : 0) * I.ldw;
      double c0 = 0.0, c1 = 0.0;
#pragma unroll 4
      for (int k0 = 0; k0 < m; k0 += 4) {
        const int k = k0 + t4;
        const double av = (ra < n && k < m) ? pa[k] * wv[k] : 0.0;
        const double bv = (cb < n && k < m) ? pb[k] : 0.0;
        dmma_8x8x4(c0, c1, av, bv);
      }
      const int row = a0 + g, col = b0 + 2 * t4;
      if (row < n && col < n && row >= col) vs[col * n - col * (col - 1) / 2 + (row - col)] = c0;
      if (row < n && col + 1 < n && row >= col + 1) vs[(col + 1) * n - (col + 1) * col / 2 + (row - col - 1)] = c1;
    }
  }
  int o = nh;
  if (nc > 0) {   // S2
    for (int q = tid; q < nh; q += NT) vs[o + q] = -0.0;
    __syncthreads();
    for (int j = tid; j < nc; j += NT) vs[o + j * n - j * (j - 1) / 2] = -(0.1 * lam[j]);
    o += nh;
  }
  for (int q = tid; q < n * m; q += NT) {   // S3
    const int j = q / m, i = q - j * m;
    vs[o + q] = I.Aw[(size_t)j * I.ldw + i] + 0.1 * cs[i] * I.Bw[(size_t)j * I.ldw + i];
  }
  o += n * m;
  for (int q = tid; q < n * nc; q += NT) {  // S4
    const int j = q / nc, k = q - j * nc;
    vs[o + q] = dn_jc(I, x, k, j);
  }
  o += n * nc + m;                          // S5 stays -1
  for (int k = tid; k < nc; k += NT) vs[o + k] = -delta;   // S6
  o += nc;
  for (int j = tid; j < n; j += NT) vs[o + j] = 0.0;       // S7
  __syncthreads();
}

// Krylov.cgls on the operator Jc(x)' (n x ncon): lam = argmin ||Jc' lam - b||  (:513, :887; defaults
// atol = rtol = sqrt(eps) on ||Op' r||, itmax = n + ncon)
template <int NT>
__device__ __forceinline__ void dn_cgls(const DenseInst& I, const double* x, const double* b, double* lam,
                                        double* cr, double* cq, double* cs_, double* cp, double tol, double* scr,
                                        double* red) {
  const int n = I.n, nc = I.nc, tid = threadIdx.x;
  for (int k = tid; k < nc; k += NT) lam[k] = 0.0;
  __syncthreads();
  if (nc == 0) return;
  const double bnorm = sqrt(cta_dot<NT>(b, b, n, red));
  if (bnorm == 0.0) return;
  for (int j = tid; j < n; j += NT) cr[j] = b[j];
  __syncthreads();
  dn_jprod_cons<NT>(I, x, cr, cs_, scr);
  for (int k = tid; k < nc; k += NT) cp[k] = cs_[k];
  double gamma = cta_dot<NT>(cs_, cs_, nc, red);
  int it = 0;
  const int itmax = n + nc;
  double arnorm = sqrt(gamma);
  const double eps_ = tol + tol * arnorm;
  bool solved = arnorm <= eps_;
  while (!(solved || it >= itmax)) {
    dn_jtprod_cons<NT>(I, x, cp, cq);
    const double dl = cta_dot<NT>(cq, cq, n, red);
    if (dl <= 0.0) break;
    const double alpha = gamma / dl;
    for (int k = tid; k < nc; k += NT) lam[k] += alpha * cp[k];
    for (int j = tid; j < n; j += NT) cr[j] -= alpha * cq[j];
    __syncthreads();
    dn_jprod_cons<NT>(I, x, cr, cs_, scr);
    const double gnext = cta_dot<NT>(cs_, cs_, nc, red);
    const double beta = gnext / gamma;
    for (int k = tid; k < nc; k += NT) cp[k] = cs_[k] + beta * cp[k];
    __syncthreads();
    gamma = gnext;
    arnorm = sqrt(gamma);
    it++;
    solved = arnorm <= eps_;
  }
}

__device__ __forceinline__ int nls_get_status(bool optimal, bool small_residual, bool stalled, bool exception,
                                              int eval_fun, int max_eval, int iter, int max_iter) {
  if (optimal) return NLS_FIRST_ORDER;
  if (small_residual) return NLS_SMALL_RESIDUAL;
  if (stalled) return NLS_STALLED;
  if (exception) return NLS_EXCEPTION;
  if (max_eval >= 0 && eval_fun > max_eval) return NLS_MAX_EVAL;
  if (max_iter >= 0 && iter > max_iter) return NLS_MAX_ITER;
  return NLS_UNKNOWN;
}

// ------------------------------------------------------------------------------------------
// One CTA per SM (the packed KKT triangle fills the shared memory); CTAs take instances
// first .. first + count - 1 from *ticket.  vals_scratch: gridDim.x x P.nnz doubles.
template <int NT>
__global__ void __launch_bounds__(NT) k_nls_dense(BatchPlanDev P, DenseNlsModel M, NlsParams prm, int first, int count,
                                                  int* ticket, double* vals_scratch, double* rec, int rec_stride,
                                                  double* dbg_vals /* first system of every instance, or NULL */) {
  B2_DYN_SMEM(raw);
  __shared__ int cnt[4];
  __shared__ int s_inst;
  __shared__ __align__(8) unsigned long long mbar;
  const int tid = threadIdx.x;
  const int n = M.n, m = M.m, nc = M.ncon, N = n + m + nc;
#ifndef B2_EMULATE
  if (tid == 0) mbar_init(&mbar, 1);
#endif
  // A and B of the instance live in the (dead) packed-triangle area while the model is evaluated
  const int lds = (m % 2 == 0) ? m + 4 : m;             // padded: conflict-free tensor-core fragment loads
  const bool can_stage = 2 * (long long)n * lds + (long long)n * nc <= P.npacked;
  uint32_t tma_phase = 0;
  bool staged = false;
  double* st = reinterpret_cast<double*>(raw + ((batched_smem_bytes(N, P.npacked) + 15) & ~(size_t)15));
  double* x = st;            double* xt = x + n;       double* Jxtr = xt + n;   double* t1 = Jxtr + n;
  double* cr = t1 + n;       double* cq = cr + n;
  double* r = cq + n;        double* Fx = r + m;       double* rt = Fx + m;     double* Ft = rt + m;
  double* snx = Ft + m;      double* csx = snx + m;    double* snt = csx + m;   double* cst = snt + m;
  double* wv = cst + m;
  double* lam = wv + m;      double* cx = lam + nc;    double* lamt = cx + nc;  double* ct = lamt + nc;
  double* dlam = ct + nc;    double* cgs = dlam + nc;  double* cgp = cgs + nc;  double* t3 = cgp + nc;
  double* spare = t3 + nc;
  double* d = spare + nc;    double* rhs = d + N;
  double* scr = rhs + N;     // 2 NT
  double* red = scr + 2 * NT;
  double* dual = rhs;        // rhs = [dual; primal] (:631-632): aliased, the loop never needs both
  double* primal = rhs + n;
  double* dx = d;
  double* dr = d + n;
  double* vs = vals_scratch + (size_t)blockIdx.x * P.nnz;
  {  // S5 = -1, once (:306)
    const int o5 = P.nnz - n - nc - m;
    for (int i = tid; i < m; i += NT) vs[o5 + i] = -1.0;
  }
  const double smax = 100.0;

  for (;;) {
    __syncthreads();
    if (tid == 0) s_inst = atomicAdd(ticket, 1);
    __syncthreads();
    const int li = s_inst;
    if (li >= count) break;
    const int b = first + li;
    DenseInst I;
    I.n = n; I.m = m; I.nc = nc;
    I.At = M.At + (size_t)b * M.stride_A;
    I.Bt = M.Bt + (size_t)b * M.stride_A;
    I.Ct = M.Ct + (size_t)b * M.stride_C;
    I.y = M.y + (size_t)b * M.stride_y;
    I.e = M.e + (size_t)b * M.stride_e;
    I.Aw = I.At; I.Bw = I.Bt; I.Cw = I.Ct; I.ldw = m;
    staged = false;
    auto ensure_staged = [&]() {
      if (can_stage && !staged) {
        double* Pk = reinterpret_cast<double*>(raw);
        if (dn_stage_model<NT>(I, Pk, Pk + (size_t)n * lds, Pk + 2 * (size_t)n * lds, lds, &mbar, tma_phase)) tma_phase ^= 1;
        staged = true;
      }
    };
    const double* x0 = M.x0 + (size_t)b * M.stride_x0;
    for (int j = tid; j < n; j += NT) x[j] = x0[j];
    for (int k = tid; k < nc; k += NT) lam[k] = M.y0 ? M.y0[(size_t)b * nc + k] : 0.0;
    for (int k = tid; k < N; k += NT) d[k] = 0.0;
    __syncthreads();

    B2_T0(tq_all);
    double rho = 0.0, rho_old = 0.0, delta = 1.0;
    int neval_res = 0, neval_cons = 0, iter = 0, inner_iter = 0, nbk = 0, nfact = 0, nlinsolve = 0;
    int status = NLS_UNKNOWN;
    double fx = 0.0, normdual = 0.0, normprimal = 0.0, normdualhat = 0.0, normprimalhat = 0.0;
    bool first_system = true;

    ensure_staged();
    dn_residual<NT>(I, x, Fx, snx, csx, scr);
    neval_res++;
    if (cta_nonfinite<NT>(Fx, m)) {
      status = NLS_ERR_NAN_INIT;
    } else {
      fx = cta_dot<NT>(Fx, Fx, m, red) / 2;
      if (nc > 0) { dn_cons<NT>(I, x, cx, scr); neval_cons++; }
      for (int i = tid; i < m; i += NT) r[i] = Fx[i];
      __syncthreads();
      dn_jtprod_res<NT>(I, csx, r, Jxtr);
      B2_T0(tq4);
      if (!prm.use_initial_multiplier) {
        dn_cgls<NT>(I, x, Jxtr, lam, cr, cq, cgs, cgp, prm.cgls_tol, scr, red);
        if (nc > 0 && cta_dot<NT>(lam, lam, nc, red) == 0.0) {
          for (int k = tid; k < nc; k += NT) lam[k] = 1.0;
          __syncthreads();
        }
      }
      B2_ACC(7, tq4);
      dn_jtprod_cons<NT>(I, x, lam, t1);
      for (int j = tid; j < n; j += NT) dual[j] = Jxtr[j] - t1[j];
      for (int i = tid; i < m; i += NT) primal[i] = Fx[i] - r[i];
      for (int k = tid; k < nc; k += NT) primal[m + k] = cx[k];
      __syncthreads();
      normdualhat = normdual = cta_absmax<NT>(dual, n, red);
      normprimalhat = normprimal = cta_absmax<NT>(primal, m + nc, red);
      const double epsF = prm.Fatol + prm.Frtol * 2 * sqrt(fx);
      const double epstol = prm.atol + prm.rtol * normdual;
      const double epsc = sqrt(epstol);

      // optimality_check_small_residual! (:872-897): returns (||cx||inf, ||dual||inf)
      auto small_res_check = [&](double& npz, double& nd) {
        for (int i = tid; i < m; i += NT) r[i] = Fx[i];
        __syncthreads();
        dn_jtprod_res<NT>(I, csx, r, Jxtr);
        dn_cgls<NT>(I, x, Jxtr, lam, cr, cq, cgs, cgp, prm.cgls_tol, scr, red);
        dn_jtprod_cons<NT>(I, x, lam, t1);
        for (int j = tid; j < n; j += NT) dual[j] = Jxtr[j] - t1[j];
        for (int i = tid; i < m; i += NT) primal[i] = 0.0;
        for (int k = tid; k < nc; k += NT) primal[m + k] = cx[k];
        __syncthreads();
        nd = cta_absmax<NT>(dual, n, red);
        npz = nc > 0 ? cta_absmax<NT>(cx, nc, red) : 0.0;
      };
      auto dual_scaling = [&]() -> double {   // :917-920
        if (nc == 0) return 1.0;
        const double l1 = cta_abssum<NT>(lam, nc, red) / nc;
        return (smax > l1 ? smax : l1) / smax;
      };
      auto norm_cx = [&]() -> double { return nc > 0 ? sqrt(cta_dot<NT>(cx, cx, nc, red)) : 0.0; };

      bool small_residual = (2 * sqrt(fx) <= epsF) && norm_cx() <= epsc;
      double sd = dual_scaling();
      bool first_order = fmax(normdual / sd, normprimal) <= epstol;
      if (small_residual && !first_order) {
        small_res_check(normprimal, normdual);
        sd = dual_scaling();
        first_order = fmax(normdual / sd, normprimal) <= epstol;
      }
      bool tired = neval_res + neval_cons > prm.max_eval;   // (:559; no wall clock on the device: max_time = Inf)
      bool broken = false;
      double epsk = 1e3;
      status = nls_get_status(first_order, small_residual, false, false, neval_res + neval_cons, prm.max_eval, 0,
                              prm.max_iter);

      while (status == NLS_UNKNOWN) {   // ---------------------------------------------- outer loop (:612)
        const double comb = normdual + normprimal;
        delta = fmax(prm.delta_min, fmin(prm.delta_dec * delta, comb));
        inner_iter = 0;
        double comb_hat = INFINITY;
        bool first_iteration = true;
        while (first_iteration || !(comb_hat <= 0.99 * comb + epsk || tired)) {   // inner loop (:622)
          first_iteration = false;
          if (inner_iter != 1 || prm.always_accept_extrapolation) {
            B2_T0(tq0);
            dn_fill_vals<NT>(I, x, lam, r, snx, csx, delta, wv, vs);
            B2_ACC(1, tq0);
            if (dbg_vals && first_system) {
              for (int q = tid; q < P.nnz; q += NT) dbg_vals[(size_t)b * P.nnz + q] = vs[q];
            }
            first_system = false;
            // newton_system! (:1008-1052): rho = 0 first, then the rho schedule; ONE call site of the
            // fused assemble + LDL' + inertia + solve so that its code exists once in the kernel
            int nfacti = 0, stage = 0;
            bool success = false;
            rho = 0.0;
            for (;;) {
              __syncthreads();
              B2_T0(tq1);
              batched_instance<NT>(P, raw, cnt, vs, stage > 0, rho, false, 0.0, prm.eig_tol, nullptr, nullptr, rhs, d,
                                   BF_SOLVE | BF_NEGATE);
              __syncthreads();
              B2_ACC(2, tq1);
              B2_ACC1(3);
              success = cnt[0] == P.nvar && cnt[1] == 0;
              __syncthreads();
              nfacti++;
              if (success) break;
              if (stage == 0) {
                rho = rho_old == 0.0 ? prm.rho0 : fmax(prm.rho_min, prm.kappa_dec * rho_old);
                stage = 1;
                continue;
              }
              if (!(rho <= prm.rho_max)) break;
              rho = (rho_old == 0.0 ? prm.kappa_largeinc : prm.kappa_inc) * rho;
              if (!(rho <= prm.rho_max)) break;
            }
            if (stage > 0 && rho <= prm.rho_max) rho_old = rho;
            staged = false;                       // the triangle went over the staged model
            I.Aw = I.At; I.Bw = I.Bt; I.Cw = I.Ct; I.ldw = m;
            nfact += nfacti;
            nlinsolve++;
            if (rho > prm.rho_max || !success || cta_nonfinite<NT>(d, N) || fx >= 1e60) {
              broken = true;
              break;
            }
            for (int k = tid; k < nc; k += NT) dlam[k] = -d[n + m + k];
            __syncthreads();
            ensure_staged();
          }
          B2_T0(tq3);
          if (inner_iter == 0) {   // extrapolation step (:656-670)
            epsk = fmax(fmin(1e3 * delta, 99 * epsk / 100), 9 * epsk / 10);
            for (int j = tid; j < n; j += NT) xt[j] = x[j] + dx[j];
            for (int i = tid; i < m; i += NT) rt[i] = r[i] + dr[i];
            const double ndl = nc > 0 ? sqrt(cta_dot<NT>(dlam, dlam, nc, red)) : 0.0;
            if (ndl > 1e4)
              for (int k = tid; k < nc; k += NT) dlam[k] = dlam[k] * 1e4 / ndl;
            __syncthreads();
            for (int k = tid; k < nc; k += NT) lamt[k] = lam[k] + dlam[k];
            __syncthreads();
            dn_residual<NT>(I, xt, Ft, snt, cst, scr);
            neval_res++;
            if (nc > 0) { dn_cons<NT>(I, xt, ct, scr); neval_cons++; }
          } else {                 // line search on the augmented Lagrangian (:672-703, :1054-1112)
            dn_jtprod_res<NT>(I, csx, Fx, t1);
            double Dphi = cta_dot<NT>(t1, dx, n, red);
            double eta = 0.0;
            if (nc > 0) {
              for (int k = tid; k < nc; k += NT) t3[k] = lam[k] - cx[k] / delta;
              __syncthreads();
              dn_jtprod_cons<NT>(I, x, t3, t1);
              Dphi -= cta_dot<NT>(dx, t1, n, red);
              eta = 1 / delta;
            }
            if (!(Dphi < 0)) { status = NLS_ERR_DPHI; break; }
            for (int j = tid; j < n; j += NT) xt[j] = x[j] + dx[j];
            __syncthreads();
            dn_residual<NT>(I, xt, Ft, snt, cst, scr);
            neval_res++;
            if (nc > 0) { dn_cons<NT>(I, xt, ct, scr); neval_cons++; }
            auto phi = [&](const double* Fv, const double* cv) -> double {   // :479-481
              const double a = cta_dot<NT>(Fv, Fv, m, red) / 2;
              if (nc == 0) return a;
              const double lc = cta_dot<NT>(lam, cv, nc, red), cc = cta_dot<NT>(cv, cv, nc, red);
              return a - lc + eta * cc / 2;
            };
            const double phix = phi(Fx, cx);
            double phit = phi(Ft, ct);
            double alpha = 1.0;
            bool alpha_err = false;
            while (!(phit <= phix + prm.gamma_A * alpha * Dphi)) {
              nbk++;
              alpha /= 4;
              for (int j = tid; j < n; j += NT) xt[j] = x[j] + alpha * dx[j];
              __syncthreads();
              dn_residual<NT>(I, xt, Ft, snt, cst, scr);
              neval_res++;
              if (nc > 0) { dn_cons<NT>(I, xt, ct, scr); neval_cons++; }
              phit = phi(Ft, ct);
              if (alpha < prm.eps2) { alpha_err = true; break; }
            }
            if (alpha_err) { status = NLS_ERR_ALPHA; break; }
            for (int i = tid; i < m; i += NT) rt[i] = Ft[i];
            for (int k = tid; k < nc; k += NT) lamt[k] = lam[k] - cx[k] / delta;
            __syncthreads();
          }
          // residuals at the trial point (:715-732); the Jacobians at xt are (cst, xt)
          B2_ACC(6, tq3);
          B2_T0(tq2);
          dn_jtprod_res<NT>(I, cst, rt, Jxtr);
          dn_jtprod_cons<NT>(I, xt, lamt, t1);
          for (int j = tid; j < n; j += NT) dual[j] = Jxtr[j] - t1[j];
          for (int i = tid; i < m; i += NT) primal[i] = Ft[i] - rt[i];
          for (int k = tid; k < nc; k += NT) primal[m + k] = ct[k];
          __syncthreads();
          normdualhat = cta_absmax<NT>(dual, n, red);
          normprimalhat = cta_absmax<NT>(primal, m + nc, red);
          comb_hat = normdualhat + normprimalhat;
          const bool good = comb_hat <= 0.99 * comb + epsk;
          if (inner_iter > 0 || prm.always_accept_extrapolation || good) {   // :734-748
            for (int j = tid; j < n; j += NT) x[j] = xt[j];
            for (int i = tid; i < m; i += NT) { r[i] = rt[i]; Fx[i] = Ft[i]; snx[i] = snt[i]; csx[i] = cst[i]; }
            for (int k = tid; k < nc; k += NT) cx[k] = ct[k];
            __syncthreads();
            fx = cta_dot<NT>(Fx, Fx, m, red) / 2;
          }
          if (good) {
            for (int k = tid; k < nc; k += NT) lam[k] = lamt[k];
            __syncthreads();
          } else {   // :753-755
            dn_jtprod_res<NT>(I, csx, r, Jxtr);
            dn_jtprod_cons<NT>(I, x, lam, t1);
            for (int j = tid; j < n; j += NT) dual[j] = Jxtr[j] - t1[j];
            __syncthreads();
          }
          if (nc > 0 && inner_iter > 0 && normdualhat <= 0.99 * normdual + epsk / 2 &&
              normprimalhat > 0.99 * normprimal + epsk / 2)
            delta = fmax(delta / 10, prm.delta_min);
          B2_ACC(4, tq2);
          inner_iter++;
          tired = neval_res + neval_cons > prm.max_eval || inner_iter > prm.max_inner;   // :765-767
        }
        if (status != NLS_UNKNOWN) break;   // an error the reference throws on
        normdual = normdualhat;
        normprimal = normprimalhat;
        sd = dual_scaling();
        first_order = fmax(normdual / sd, normprimal) <= epstol;
        small_residual = (2 * sqrt(fx) <= epsF) && norm_cx() <= epsc;
        if (small_residual && !first_order) {
          small_res_check(normprimal, normdual);
          sd = dual_scaling();
          first_order = fmax(normdual / sd, normprimal) <= epstol;
        }
        iter++;
        status = nls_get_status(first_order, small_residual, prm.max_inner >= 0 && inner_iter > prm.max_inner, broken,
                                neval_res + neval_cons, prm.max_eval, iter, prm.max_iter);
      }
    }
    // ---- record (:834-862)
    const double obj = cta_dot<NT>(Fx, Fx, m, red) / 2;
    const double pf = nc > 0 ? sqrt(cta_dot<NT>(cx, cx, nc, red)) : 0.0;
    double* R = rec + (size_t)b * rec_stride;
    if (tid == 0) {
      R[0] = status; R[1] = iter; R[2] = nfact; R[3] = nlinsolve; R[4] = nbk; R[5] = neval_res; R[6] = neval_cons;
      R[7] = obj; R[8] = pf; R[9] = normdual; R[10] = rho; R[11] = delta;
    }
    for (int j = tid; j < n; j += NT) R[NLS_REC_HEAD + j] = x[j];
    for (int k = tid; k < nc; k += NT) R[NLS_REC_HEAD + n + k] = lam[k];
    B2_ACC(0, tq_all);
    B2_ACC1(5);
  }
}

}  // namespace b2
