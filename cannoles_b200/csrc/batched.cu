// batched.cu -- batch of independent KKT systems sharing one sparsity pattern (BASELINE.json
// config 5: multi-start / per-sample estimation).  Host side of batched_kernels.cuh: one symbolic
// analysis for the shared pattern, then every verb is ONE kernel launch with one CTA per instance.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <new>
#include <vector>

#include "../../include/cannoles_b200.h"
#include "b2_cuda.h"
#include "batched_kernels.cuh"
#include "symbolic.h"

namespace b2 {

struct BatchEngine {
  Symbolic sym;
  int device = 0;
  int64_t batch = 0;
  cudaStream_t stream = nullptr;
  BatchPlanDev plan{};
  std::vector<int32_t> cbm_host;
  int smem = 0;
  std::vector<void*> dev_ptrs;
  double* d_vals = nullptr;
  double* d_rhs = nullptr;
  double* d_out = nullptr;
  double* d_L = nullptr;         // batch x npacked factors (allocated lazily)
  double *d_rho = nullptr, *d_delta = nullptr;
  uint8_t* d_active = nullptr;
  long long* d_counts = nullptr; // batch x 4
  long long* h_counts = nullptr; // pinned
  bool have_vals = false, factored = false;
  double bytes_device = 0, t_plan = 0;
  double last_ms = 0;
  cudaEvent_t ev[2] = {nullptr, nullptr};
  cudaEvent_t tev[2] = {nullptr, nullptr};

  template <typename T>
  int up(const std::vector<T>& v, const T** out) {
    T* p = nullptr;
    size_t n = std::max<size_t>(v.size(), 1);
    B2_CUDA_OK(cudaMalloc((void**)&p, n * sizeof(T)));
    if (!v.empty()) B2_CUDA_OK(cudaMemcpy(p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
    dev_ptrs.push_back(p);
    bytes_device += (double)(n * sizeof(T));
    *out = p;
    return 0;
  }
  template <typename T>
  int alloc(T** out, size_t n) {
    n = std::max<size_t>(n, 1);
    B2_CUDA_OK(cudaMalloc((void**)out, n * sizeof(T)));
    dev_ptrs.push_back(*out);
    bytes_device += (double)(n * sizeof(T));
    return 0;
  }

  int init(int dev, int64_t nbatch);
  void destroy();
  int launch(const double* dv, const double* rho, const double* delta, const uint8_t* act, double eig_tol,
             const double* rhs, double* out, int flags);
  int fetch_counts(int64_t* npos, int64_t* nzero, int64_t* nneg, int32_t* breakdown, const uint8_t* active);
};

int BatchEngine::init(int dev, int64_t nbatch) {
  device = dev;
  batch = nbatch;
  const Symbolic& S = sym;
  const int N = (int)S.N;
  B2_CUDA_OK(cudaSetDevice(dev));
  B2_CUDA_OK(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
  for (auto& e : ev) B2_CUDA_OK(cudaEventCreate(&e));
  for (auto& e : tev) B2_CUDA_OK(cudaEventCreate(&e));
  // packed lower triangle, column k holding rows k..N-1 at Pk[cbm[k] + row].  Column starts are
  // padded so that (cbm[k] + k) = 4 (k mod 4) (mod 16): the four columns a DMMA fragment touches
  // then fall into disjoint shared-memory bank groups (8 consecutive rows x 4 columns, 8-byte words)
  std::vector<int32_t> cbm(N);
  int64_t npacked = 0;
  for (int k = 0; k < N; k++) {
    while ((npacked & 15) != 4 * (k & 3)) npacked++;
    cbm[k] = (int32_t)(npacked - k);
    npacked += N - k;
  }
  cbm_host = cbm;
  smem = (int)batched_smem_bytes(N, npacked);
  if (smem > 227 * 1024 - 64) {
    snprintf(g_last_error, sizeof(g_last_error),
             "b2b_analyze: N = %d needs %d bytes of shared memory per instance (> 227 KB); use the "
             "single-system engine (b2_analyze) for systems this large", N, smem);
    return -1;
  }
  // rows below each supernode
  std::vector<int32_t> rb_ptr(S.nsuper + 1, 0), rb_idx;
  for (int s = 0; s < S.nsuper; s++) {
    int w = S.scol[s + 1] - S.scol[s];
    for (int64_t p = S.rptr[s] + w; p < S.rptr[s + 1]; p++) rb_idx.push_back(S.rowidx[p]);
    rb_ptr[s + 1] = (int32_t)rb_idx.size();
  }
  // contributing columns of supernode s: every column of a supernode d < s whose rows below
  // intersect the pivot columns of s
  std::vector<int32_t> ct_ptr(S.nsuper + 1, 0), ct_col;
  {
    std::vector<std::vector<int32_t>> ct(S.nsuper);
    for (int d = 0; d < S.nsuper; d++) {
      int last = -1;
      for (int32_t p = rb_ptr[d]; p < rb_ptr[d + 1]; p++) {
        int t = S.col2sn[rb_idx[p]];
        if (t == last) continue;
        last = t;  // rb is ascending, so supernodes appear in runs
        for (int k = S.scol[d]; k < S.scol[d + 1]; k++) ct[t].push_back(k);
      }
    }
    for (int s = 0; s < S.nsuper; s++) {
      std::sort(ct[s].begin(), ct[s].end());
      ct_col.insert(ct_col.end(), ct[s].begin(), ct[s].end());
      ct_ptr[s + 1] = (int32_t)ct_col.size();
    }
  }
  // COO -> packed scatter maps
  std::vector<int32_t> dst_single(S.nnz, -1), multi_dst, multi_ptr(1, 0), multi_coo;
  {
    std::vector<int32_t> slot_dst(S.nnzA);
    for (int64_t j = 0; j < N; j++)
      for (int64_t p = S.Ap[j]; p < S.Ap[j + 1]; p++) {
        int a = S.pinv[S.Ai[p]], b = S.pinv[j];
        int col = std::min(a, b), row = std::max(a, b);
        slot_dst[p] = cbm[col] + row;
      }
    for (int64_t s = 0; s < S.nnzA; s++) {
      int64_t a = S.slot_ptr[s], b = S.slot_ptr[s + 1];
      if (b - a == 1) {
        dst_single[S.coo_sorted[a]] = slot_dst[s];
      } else {
        multi_dst.push_back(slot_dst[s]);
        for (int64_t q = a; q < b; q++) multi_coo.push_back(S.coo_sorted[q]);
        multi_ptr.push_back((int32_t)multi_coo.size());
      }
    }
  }
  // per level: the supernodes the sequential loops of the kernel have work for
  std::vector<int32_t> phw_ptr(1, 0), phw_sn, phb_ptr(1, 0), phb_sn;
  for (int l = 0; l < S.nlevels; l++) {
    for (int q = S.level_ptr[l]; q < S.level_ptr[l + 1]; q++) {
      const int s = S.level_sn[q];
      const int w = S.scol[s + 1] - S.scol[s];
      if (!(w == 1 && ct_ptr[s + 1] == ct_ptr[s])) phw_sn.push_back(s);
      if (w > 1) phb_sn.push_back(s);
    }
    phw_ptr.push_back((int32_t)phw_sn.size());
    phb_ptr.push_back((int32_t)phb_sn.size());
  }
  plan.N = N; plan.nnz = (int)S.nnz; plan.nsuper = S.nsuper; plan.nphase = S.nlevels;
  plan.npacked = (int)npacked; plan.nvar = (int)S.nvar; plan.nequ = (int)S.nequ; plan.ncon = (int)S.ncon;
  plan.nmulti = (int)multi_dst.size();
  if (up(S.perm, &plan.perm) || up(cbm, &plan.cbm) || up(S.scol, &plan.sc0) || up(rb_ptr, &plan.rb_ptr) ||
      up(rb_idx, &plan.rb_idx) || up(ct_ptr, &plan.ct_ptr) || up(ct_col, &plan.ct_col) ||
      up(S.level_ptr, &plan.ph_ptr) || up(S.level_sn, &plan.ph_sn) || up(phw_ptr, &plan.phw_ptr) ||
      up(phw_sn, &plan.phw_sn) || up(phb_ptr, &plan.phb_ptr) || up(phb_sn, &plan.phb_sn) || up(dst_single, &plan.dst_single) ||
      up(multi_dst, &plan.multi_dst) || up(multi_ptr, &plan.multi_ptr) || up(multi_coo, &plan.multi_coo))
    return -1;
  if (alloc(&d_vals, (size_t)batch * S.nnz) || alloc(&d_rhs, (size_t)batch * N) ||
      alloc(&d_out, (size_t)batch * N) || alloc(&d_rho, (size_t)batch) || alloc(&d_delta, (size_t)batch) ||
      alloc(&d_active, (size_t)batch) || alloc(&d_counts, (size_t)batch * 4) ||
      alloc(&d_L, (size_t)batch * npacked))
    return -1;
  B2_CUDA_OK(cudaMemset(d_counts, 0, (size_t)batch * 4 * sizeof(long long)));
  B2_CUDA_OK(cudaMallocHost((void**)&h_counts, (size_t)batch * 4 * sizeof(long long)));
  B2_CUDA_OK(cudaFuncSetAttribute(k_batched<BATCH_NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  B2_CUDA_OK(cudaDeviceSynchronize());
  return 0;
}

void BatchEngine::destroy() {
  for (void* p : dev_ptrs) cudaFree(p);
  dev_ptrs.clear();
  if (h_counts) cudaFreeHost(h_counts);
  for (auto& e : ev) if (e) cudaEventDestroy(e);
  for (auto& e : tev) if (e) cudaEventDestroy(e);
  if (stream) cudaStreamDestroy(stream);
}

int BatchEngine::launch(const double* dv, const double* rho, const double* delta, const uint8_t* act,
                        double eig_tol, const double* rhs, double* out, int flags) {
  B2_CUDA_OK(cudaEventRecord(ev[0], stream));
  B2_LAUNCH(k_batched<BATCH_NT>, (unsigned)batch, BATCH_NT, smem, stream, plan, (int)batch, dv, rho, delta, act,
            eig_tol, d_counts, d_L, rhs, out, flags);
  B2_CUDA_OK(cudaGetLastError());
  B2_CUDA_OK(cudaEventRecord(ev[1], stream));
  return 0;
}

int BatchEngine::fetch_counts(int64_t* npos, int64_t* nzero, int64_t* nneg, int32_t* breakdown,
                              const uint8_t* active) {
  B2_CUDA_OK(cudaMemcpyAsync(h_counts, d_counts, (size_t)batch * 4 * sizeof(long long), cudaMemcpyDeviceToHost, stream));
  B2_CUDA_OK(cudaStreamSynchronize(stream));
  float ms = 0;
  cudaEventElapsedTime(&ms, ev[0], ev[1]);
  last_ms = ms;
  for (int64_t b = 0; b < batch; b++) {
    if (active && !active[b]) continue;
    if (npos) npos[b] = h_counts[4 * b];
    if (nzero) nzero[b] = h_counts[4 * b + 1];
    if (nneg) nneg[b] = h_counts[4 * b + 2];
    if (breakdown) breakdown[b] = h_counts[4 * b + 3] != 0;
  }
  return 0;
}

}  // namespace b2

struct b2b_handle {
  b2::BatchEngine eng;
};

namespace {
int failb(const char* msg) {
  snprintf(b2::g_last_error, sizeof(b2::g_last_error), "%s", msg);
  return -1;
}
}  // namespace

#ifdef B2_TIMING
extern "C" int b2b_debug_clocks(long long* out64) {
  return cudaMemcpyFromSymbol(out64, b2::b2_dbg, 64 * sizeof(long long)) == cudaSuccess ? 0 : -1;
}
#endif

extern "C" {

int b2b_analyze(int64_t N, int64_t nnz, const int64_t* rows1, const int64_t* cols1, int64_t nvar,
                int64_t nequ, int64_t ncon, int64_t batch, int ordering, const int64_t* user_perm,
                int device, b2b_handle** out) {
  if (!out) return failb("b2b_analyze: out == NULL");
  *out = nullptr;
  if (!rows1 || !cols1) return failb("b2b_analyze: NULL index arrays");
  if (batch <= 0) return failb("b2b_analyze: batch must be positive");
  int ndev = b2_device_count();
  if (ndev <= 0) return failb("b2b_analyze: no CUDA device (this backend has no CPU fallback)");
  if (device < 0 || device >= ndev) return failb("b2b_analyze: bad device ordinal");
  b2b_handle* h = new (std::nothrow) b2b_handle();
  if (!h) return failb("b2b_analyze: out of host memory");
  b2::SymbolicOptions opt;
  opt.ordering = ordering;
  opt.user_perm = user_perm;
  opt.build_spmv = false;
  // the packed dense triangle stores structural zeros for free, so amalgamate generously:
  // wide supernodes keep the tensor-core tiles full
  opt.relax_always = 16; opt.relax_z1 = 0.6; opt.relax_z2 = 0.4; opt.relax_z3 = 0.25;
  if (!b2::analyze(N, nnz, rows1, cols1, nvar, nequ, ncon, opt, h->eng.sym)) {
    snprintf(b2::g_last_error, sizeof(b2::g_last_error), "b2b_analyze: %s", h->eng.sym.error.c_str());
    delete h;
    return -1;
  }
  if (h->eng.init(device, batch)) {
    h->eng.destroy();
    delete h;
    return -1;
  }
  *out = h;
  return 0;
}

int b2b_factorize_dev(b2b_handle* h, const double* d_vals, const uint8_t* d_active, double eig_tol,
                      int64_t* d_counts4) {
  if (!h || !d_vals) return failb("b2b_factorize_dev: NULL argument");
  b2::BatchEngine& E = h->eng;
  if (cudaSetDevice(E.device) != cudaSuccess) return failb("cudaSetDevice failed");
  if (d_vals != E.d_vals)
    B2_CUDA_OK(cudaMemcpyAsync(E.d_vals, d_vals, (size_t)E.batch * E.sym.nnz * sizeof(double), cudaMemcpyDeviceToDevice, E.stream));
  E.have_vals = true;
  if (E.launch(E.d_vals, nullptr, nullptr, d_active, eig_tol, nullptr, nullptr, b2::BF_STORE)) return -1;
  if (d_counts4)
    B2_CUDA_OK(cudaMemcpyAsync(d_counts4, E.d_counts, (size_t)E.batch * 4 * sizeof(long long), cudaMemcpyDeviceToDevice, E.stream));
  B2_CUDA_OK(cudaStreamSynchronize(E.stream));
  float ms = 0;
  cudaEventElapsedTime(&ms, E.ev[0], E.ev[1]);
  E.last_ms = ms;
  E.factored = true;
  return 0;
}

static int upload_active(b2::BatchEngine& E, const uint8_t* active, const uint8_t** d_act) {
  *d_act = nullptr;
  if (!active) return 0;
  B2_CUDA_OK(cudaMemcpyAsync(E.d_active, active, (size_t)E.batch, cudaMemcpyHostToDevice, E.stream));
  *d_act = E.d_active;
  return 0;
}

int b2b_factorize(b2b_handle* h, const double* vals, const uint8_t* active, double eig_tol, int64_t* npos,
                  int64_t* nzero, int64_t* nneg, int32_t* breakdown) {
  if (!h || !vals) return failb("b2b_factorize: NULL argument");
  b2::BatchEngine& E = h->eng;
  if (cudaSetDevice(E.device) != cudaSuccess) return failb("cudaSetDevice failed");
  const size_t per = (size_t)E.sym.nnz;
  if (!active) {
    B2_CUDA_OK(cudaMemcpyAsync(E.d_vals, vals, (size_t)E.batch * per * sizeof(double), cudaMemcpyHostToDevice, E.stream));
  } else {  // only the active instances travel
    for (int64_t b = 0; b < E.batch; b++)
      if (active[b])
        B2_CUDA_OK(cudaMemcpyAsync(E.d_vals + b * per, vals + b * per, per * sizeof(double), cudaMemcpyHostToDevice, E.stream));
  }
  E.have_vals = true;
  const uint8_t* d_act;
  if (upload_active(E, active, &d_act)) return -1;
  if (E.launch(E.d_vals, nullptr, nullptr, d_act, eig_tol, nullptr, nullptr, b2::BF_STORE)) return -1;
  E.factored = true;
  return E.fetch_counts(npos, nzero, nneg, breakdown, active);
}

int b2b_refactorize_shift(b2b_handle* h, const double* rho, const double* delta_or_null, const uint8_t* active,
                          double eig_tol, int64_t* npos, int64_t* nzero, int64_t* nneg, int32_t* breakdown) {
  if (!h || !rho) return failb("b2b_refactorize_shift: NULL argument");
  b2::BatchEngine& E = h->eng;
  if (!E.sym.shift_ok) return failb("b2b_refactorize_shift: COO layout has no canonical rho/delta segments");
  if (!E.have_vals) return failb("b2b_refactorize_shift before any b2b_factorize");
  if (cudaSetDevice(E.device) != cudaSuccess) return failb("cudaSetDevice failed");
  B2_CUDA_OK(cudaMemcpyAsync(E.d_rho, rho, (size_t)E.batch * sizeof(double), cudaMemcpyHostToDevice, E.stream));
  if (delta_or_null)
    B2_CUDA_OK(cudaMemcpyAsync(E.d_delta, delta_or_null, (size_t)E.batch * sizeof(double), cudaMemcpyHostToDevice, E.stream));
  const uint8_t* d_act;
  if (upload_active(E, active, &d_act)) return -1;
  if (E.launch(E.d_vals, E.d_rho, delta_or_null ? E.d_delta : nullptr, d_act, eig_tol, nullptr, nullptr, b2::BF_STORE))
    return -1;
  return E.fetch_counts(npos, nzero, nneg, breakdown, active);
}

int b2b_solve_dev(b2b_handle* h, const double* d_rhs, double* d_out, const uint8_t* d_active, int negate) {
  if (!h || !d_rhs || !d_out) return failb("b2b_solve_dev: NULL argument");
  b2::BatchEngine& E = h->eng;
  if (!E.factored) return failb("b2b_solve before a factorization");
  if (cudaSetDevice(E.device) != cudaSuccess) return failb("cudaSetDevice failed");
  if (E.launch(nullptr, nullptr, nullptr, d_active, 0.0, d_rhs, d_out,
               b2::BF_LOAD | b2::BF_SOLVE | (negate ? b2::BF_NEGATE : 0)))
    return -1;
  B2_CUDA_OK(cudaStreamSynchronize(E.stream));
  float ms = 0;
  cudaEventElapsedTime(&ms, E.ev[0], E.ev[1]);
  E.last_ms = ms;
  return 0;
}

int b2b_solve(b2b_handle* h, const double* rhs, double* d_out, const uint8_t* active, int negate) {
  if (!h || !rhs || !d_out) return failb("b2b_solve: NULL argument");
  b2::BatchEngine& E = h->eng;
  if (!E.factored) return failb("b2b_solve before a factorization");
  if (cudaSetDevice(E.device) != cudaSuccess) return failb("cudaSetDevice failed");
  const size_t nb = (size_t)E.batch * E.sym.N * sizeof(double);
  B2_CUDA_OK(cudaMemcpyAsync(E.d_rhs, rhs, nb, cudaMemcpyHostToDevice, E.stream));
  const uint8_t* d_act;
  if (upload_active(E, active, &d_act)) return -1;
  if (active)  // inactive instances keep their previous output
    B2_CUDA_OK(cudaMemcpyAsync(E.d_out, d_out, nb, cudaMemcpyHostToDevice, E.stream));
  if (E.launch(nullptr, nullptr, nullptr, d_act, 0.0, E.d_rhs, E.d_out,
               b2::BF_LOAD | b2::BF_SOLVE | (negate ? b2::BF_NEGATE : 0)))
    return -1;
  B2_CUDA_OK(cudaMemcpyAsync(d_out, E.d_out, nb, cudaMemcpyDeviceToHost, E.stream));
  B2_CUDA_OK(cudaStreamSynchronize(E.stream));
  float ms = 0;
  cudaEventElapsedTime(&ms, E.ev[0], E.ev[1]);
  E.last_ms = ms;
  return 0;
}

/* fused: factorize every active instance and, where the inertia is the expected one
 * (npos == nvar, nzero == 0), solve straight from the shared-memory factor; instances that
 * failed leave d_out untouched and are retried by the caller with b2b_refactorize_shift. */
int b2b_factor_solve_dev(b2b_handle* h, const double* d_vals, const double* d_rhs, double* d_out,
                         const uint8_t* d_active, double eig_tol, int negate, int store_factor,
                         int64_t* d_counts4) {
  if (!h || !d_vals || !d_rhs || !d_out) return failb("b2b_factor_solve_dev: NULL argument");
  b2::BatchEngine& E = h->eng;
  if (cudaSetDevice(E.device) != cudaSuccess) return failb("cudaSetDevice failed");
  int flags = b2::BF_SOLVE | (negate ? b2::BF_NEGATE : 0) | (store_factor ? b2::BF_STORE : 0);
  if (E.launch(d_vals, nullptr, nullptr, d_active, eig_tol, d_rhs, d_out, flags)) return -1;
  if (d_counts4)
    B2_CUDA_OK(cudaMemcpyAsync(d_counts4, E.d_counts, (size_t)E.batch * 4 * sizeof(long long), cudaMemcpyDeviceToDevice, E.stream));
  B2_CUDA_OK(cudaStreamSynchronize(E.stream));
  float ms = 0;
  cudaEventElapsedTime(&ms, E.ev[0], E.ev[1]);
  E.last_ms = ms;
  if (store_factor) { E.factored = true; }
  return 0;
}

int b2b_factor_solve(b2b_handle* h, const double* vals, const double* rhs, double* d_out, const uint8_t* active,
                     double eig_tol, int negate, int64_t* npos, int64_t* nzero, int64_t* nneg, int32_t* breakdown) {
  if (!h || !vals || !rhs || !d_out) return failb("b2b_factor_solve: NULL argument");
  b2::BatchEngine& E = h->eng;
  if (cudaSetDevice(E.device) != cudaSuccess) return failb("cudaSetDevice failed");
  const size_t per = (size_t)E.sym.nnz, N = (size_t)E.sym.N;
  const size_t nb = (size_t)E.batch * N * sizeof(double);
  if (!active) {
    B2_CUDA_OK(cudaMemcpyAsync(E.d_vals, vals, (size_t)E.batch * per * sizeof(double), cudaMemcpyHostToDevice, E.stream));
  } else {
    for (int64_t b = 0; b < E.batch; b++)
      if (active[b])
        B2_CUDA_OK(cudaMemcpyAsync(E.d_vals + b * per, vals + b * per, per * sizeof(double), cudaMemcpyHostToDevice, E.stream));
  }
  E.have_vals = true;
  B2_CUDA_OK(cudaMemcpyAsync(E.d_rhs, rhs, nb, cudaMemcpyHostToDevice, E.stream));
  B2_CUDA_OK(cudaMemcpyAsync(E.d_out, d_out, nb, cudaMemcpyHostToDevice, E.stream));
  const uint8_t* d_act;
  if (upload_active(E, active, &d_act)) return -1;
  if (E.launch(E.d_vals, nullptr, nullptr, d_act, eig_tol, E.d_rhs, E.d_out,
               b2::BF_SOLVE | b2::BF_STORE | (negate ? b2::BF_NEGATE : 0)))
    return -1;
  E.factored = true;
  B2_CUDA_OK(cudaMemcpyAsync(d_out, E.d_out, nb, cudaMemcpyDeviceToHost, E.stream));
  return E.fetch_counts(npos, nzero, nneg, breakdown, active);
}

int b2b_last_ms(const b2b_handle* h, double* ms) {
  if (!h || !ms) return failb("b2b_last_ms: NULL argument");
  *ms = h->eng.last_ms;
  return 0;
}

int b2b_timer_start(b2b_handle* h) {
  if (!h) return failb("b2b_timer_start: NULL handle");
  if (cudaEventRecord(h->eng.tev[0], h->eng.stream) != cudaSuccess) return failb("b2b_timer_start: record failed");
  return 0;
}

int b2b_timer_stop(b2b_handle* h, double* ms) {
  if (!h || !ms) return failb("b2b_timer_stop: NULL argument");
  if (cudaEventRecord(h->eng.tev[1], h->eng.stream) != cudaSuccess) return failb("b2b_timer_stop: record failed");
  if (cudaEventSynchronize(h->eng.tev[1]) != cudaSuccess) return failb("b2b_timer_stop: sync failed");
  float f = 0;
  cudaEventElapsedTime(&f, h->eng.tev[0], h->eng.tev[1]);
  *ms = f;
  return 0;
}

int b2b_get_perm(const b2b_handle* h, int64_t* perm0) {
  if (!h || !perm0) return failb("b2b_get_perm: NULL argument");
  for (int64_t k = 0; k < h->eng.sym.N; k++) perm0[k] = h->eng.sym.perm[k];
  return 0;
}

/* pivots of instance b in elimination order (diagonal of its stored packed factor) */
int b2b_get_d(b2b_handle* h, int64_t b, double* d) {
  if (!h || !d) return failb("b2b_get_d: NULL argument");
  b2::BatchEngine& E = h->eng;
  if (!E.factored || b < 0 || b >= E.batch) return failb("b2b_get_d: no stored factor for that instance");
  const int N = (int)E.sym.N;
  std::vector<double> Lh((size_t)E.plan.npacked);
  cudaStreamSynchronize(E.stream);
  if (cudaMemcpy(Lh.data(), E.d_L + (size_t)b * E.plan.npacked, Lh.size() * sizeof(double), cudaMemcpyDeviceToHost) != cudaSuccess)
    return failb("b2b_get_d: copy failed");
  for (int k = 0; k < N; k++) d[k] = Lh[(size_t)E.cbm_host[k] + k];
  return 0;
}

int b2b_stats(const b2b_handle* h, b2_stats_t* o) {
  if (!h || !o) return failb("b2b_stats: NULL argument");
  const b2::Symbolic& S = h->eng.sym;
  memset(o, 0, sizeof(*o));
  o->N = S.N; o->nnz = S.nnz; o->nnzA = S.nnzA; o->nnzL = S.nnzL; o->nnzL_store = h->eng.plan.npacked;
  o->cb_store = 0; o->nsuper = S.nsuper; o->nlevels = S.nlevels; o->max_front = S.max_front;
  o->max_width = S.max_width; o->n_small = S.nsuper; o->n_large = 0;
  o->launches_factor = 1; o->launches_solve = 1;
  o->flops = S.flops; o->flops_store = S.flops_store; o->t_order = S.t_order; o->t_symbolic = S.t_symbolic;
  o->t_plan = h->eng.t_plan; o->bytes_device = h->eng.bytes_device;
  return 0;
}

int b2b_free(b2b_handle* h) {
  if (!h) return 0;
  h->eng.destroy();
  delete h;
  return 0;
}

}  // extern "C"
