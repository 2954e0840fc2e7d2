// capi.cu -- extern "C" surface declared in include/cannoles_b200.h (single-system verbs and
// device utilities; the batched verbs live in batched.cu).
#include <algorithm>
#include <cmath>
#include <cstring>
#include <new>

#include "../../include/cannoles_b200.h"
#include "engine.h"

struct b2_handle {
  b2::Engine eng;
};

namespace {
int fail(const char* msg) {
  snprintf(b2::g_last_error, sizeof(b2::g_last_error), "%s", msg);
  return -1;
}
}  // namespace

extern "C" {

const char* b2_last_error(void) { return b2::g_last_error; }
int b2_version(void) { return 100; }

int b2_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
  return n;
}

int b2_analyze(int64_t N, int64_t nnz, const int64_t* rows1, const int64_t* cols1, int64_t nvar,
               int64_t nequ, int64_t ncon, int ordering, const int64_t* user_perm, int device,
               b2_handle** out) {
  if (!out) return fail("b2_analyze: out == NULL");
  *out = nullptr;
  if (!rows1 || !cols1) return fail("b2_analyze: NULL index arrays");
  int ndev = b2_device_count();
  if (ndev <= 0) return fail("b2_analyze: no CUDA device (this backend has no CPU fallback)");
  if (device < 0 || device >= ndev) return fail("b2_analyze: bad device ordinal");
  b2_handle* h = new (std::nothrow) b2_handle();
  if (!h) return fail("b2_analyze: out of host memory");
  b2::SymbolicOptions opt;
  opt.ordering = ordering;
  opt.user_perm = user_perm;
  if (const char* e = getenv("B2_RELAX_SCALE")) {
    double sc = atof(e);
    if (sc >= 0) { opt.relax_z1 *= sc; opt.relax_z2 *= sc; opt.relax_z3 *= sc; if (sc == 0) opt.relax_always = 0; }
  }
  if (!b2::analyze(N, nnz, rows1, cols1, nvar, nequ, ncon, opt, h->eng.sym)) {
    snprintf(b2::g_last_error, sizeof(b2::g_last_error), "b2_analyze: %s", h->eng.sym.error.c_str());
    delete h;
    return -1;
  }
  if (const char* e = getenv("B2_SMALL_MAX_M")) h->eng.small_max_m = atof(e);
  if (const char* e = getenv("B2_NO_GRAPH")) h->eng.use_graph = atoi(e) == 0;
  if (const char* e = getenv("B2_NO_BRANCHES")) h->eng.use_branches = atoi(e) == 0;
  if (const char* e = getenv("B2_SOLVE_FORK")) h->eng.solve_fork = atoi(e);
  if (const char* e = getenv("B2_LOOKAHEAD")) h->eng.lookahead = atoi(e) != 0;
  if (const char* e = getenv("B2_DAG")) h->eng.use_dag = atoi(e) != 0;
  if (const char* e = getenv("B2_DAG_LEVEL_MAX")) h->eng.dag_level_max = atoi(e);
  if (const char* e = getenv("B2_DAG_EXCL_MAX")) h->eng.dag_excl_max = atoi(e);
  if (const char* e = getenv("B2_DAG_SCHED")) h->eng.dag_sched = atoi(e) != 0;
  if (const char* e = getenv("B2_SOLVE_BIG_M")) h->eng.solve_big_m = atof(e);
  if (const char* e = getenv("B2_INV_MIN_BLK")) h->eng.inv_min_blk = atoi(e);
  if (const char* e = getenv("B2_UPDATE_TMA")) h->eng.update_tma = atoi(e) != 0;
  if (const char* e = getenv("B2_TINY_MAX_M")) h->eng.tiny_max_m = std::min(8, atoi(e));
  if (const char* e = getenv("B2_TINY_SOLVE_MAX_M")) h->eng.tiny_solve_max_m = std::min(32, atoi(e));
  if (h->eng.init(device)) {
    h->eng.destroy();
    delete h;
    return -1;
  }
  *out = h;
  return 0;
}

int b2_factorize(b2_handle* h, const double* vals, double eig_tol, int64_t* npos, int64_t* nzero,
                 int64_t* nneg, int* breakdown) {
  if (!h || !vals) return fail("b2_factorize: NULL argument");
  return h->eng.factorize_host(vals, eig_tol, npos, nzero, nneg, breakdown);
}

int b2_factorize_dev(b2_handle* h, const double* d_vals, double eig_tol, int64_t* npos, int64_t* nzero,
                     int64_t* nneg, int* breakdown) {
  if (!h || !d_vals) return fail("b2_factorize_dev: NULL argument");
  return h->eng.factorize_dev(d_vals, eig_tol, npos, nzero, nneg, breakdown);
}

int b2_refactorize_shift(b2_handle* h, double rho, double delta_or_nan, double eig_tol, int64_t* npos,
                         int64_t* nzero, int64_t* nneg, int* breakdown) {
  if (!h) return fail("b2_refactorize_shift: NULL handle");
  return h->eng.refactorize_shift(rho, delta_or_nan, eig_tol, npos, nzero, nneg, breakdown);
}

int b2_factorize_retry(b2_handle* h, const double* vals, double rho, double eig_tol, int64_t* npos,
                       int64_t* nzero, int64_t* nneg, int* breakdown, int* speculation_held) {
  if (!h || !vals) return fail("b2_factorize_retry: NULL argument");
  return h->eng.factorize_retry(vals, rho, eig_tol, npos, nzero, nneg, breakdown, speculation_held);
}

int b2_solve(b2_handle* h, const double* rhs, double* d_out, int negate, int refine_steps, double* relres) {
  if (!h || !rhs || !d_out) return fail("b2_solve: NULL argument");
  return h->eng.solve_host(rhs, d_out, negate, refine_steps < 0 ? 0 : refine_steps, relres);
}

int b2_solve_dev(b2_handle* h, const double* d_rhs, double* d_out, int negate, int refine_steps, double* relres) {
  if (!h || !d_rhs || !d_out) return fail("b2_solve_dev: NULL argument");
  b2::Engine& E = h->eng;
  if (cudaSetDevice(E.device) != cudaSuccess) return fail("cudaSetDevice failed");
  cudaEventRecord(E.ev[1], E.stream);
  if (E.solve_core(d_rhs, d_out, negate, refine_steps < 0 ? 0 : refine_steps, relres)) return -1;
  cudaEventRecord(E.ev[4], E.stream);
  if (cudaStreamSynchronize(E.stream) != cudaSuccess) return fail("b2_solve_dev: stream sync failed");
  if (relres) *relres = E.h_scalars[1] > 0 ? std::sqrt(E.h_scalars[0] / E.h_scalars[1]) : std::sqrt(E.h_scalars[0]);
  float ms = 0;
  cudaEventElapsedTime(&ms, E.ev[1], E.ev[4]);
  E.last_ms[0] = 0; E.last_ms[3] = ms; E.last_ms[4] = 0;
  return 0;
}

int b2_register_host(b2_handle* h, void* ptr, size_t bytes) {
  if (!h || !ptr) return fail("b2_register_host: NULL argument");
  if (cudaHostRegister(ptr, bytes, cudaHostRegisterDefault) != cudaSuccess) {
    cudaGetLastError();
    return fail("b2_register_host: cudaHostRegister failed");
  }
  h->eng.registered.push_back(ptr);
  return 0;
}

int b2_unregister_host(b2_handle* h, void* ptr) {
  if (!h || !ptr) return fail("b2_unregister_host: NULL argument");
  for (size_t i = 0; i < h->eng.registered.size(); i++)
    if (h->eng.registered[i] == ptr) {
      cudaHostUnregister(ptr);
      h->eng.registered.erase(h->eng.registered.begin() + i);
      return 0;
    }
  return fail("b2_unregister_host: pointer was not registered");
}

int b2_stats(const b2_handle* h, b2_stats_t* o) {
  if (!h || !o) return fail("b2_stats: NULL argument");
  const b2::Symbolic& S = h->eng.sym;
  memset(o, 0, sizeof(*o));
  o->N = S.N; o->nnz = S.nnz; o->nnzA = S.nnzA; o->nnzL = S.nnzL; o->nnzL_store = S.nnzL_store;
  o->cb_store = S.cb_store; o->nsuper = S.nsuper; o->nlevels = S.nlevels; o->max_front = S.max_front;
  o->max_width = S.max_width; o->n_small = h->eng.n_small; o->n_large = h->eng.n_large;
  o->launches_factor = (int64_t)h->eng.fact_launches.size() + 4;
  o->launches_solve = (int64_t)(h->eng.fwd_launches.size() + h->eng.bwd_launches.size()) + 3;
  o->flops = S.flops; o->flops_store = S.flops_store; o->t_order = S.t_order; o->t_symbolic = S.t_symbolic;
  o->t_plan = h->eng.t_plan; o->bytes_device = h->eng.bytes_device;
  return 0;
}

int b2_last_timings(const b2_handle* h, double* ms5) {
  if (!h || !ms5) return fail("b2_last_timings: NULL argument");
  for (int i = 0; i < 5; i++) ms5[i] = h->eng.last_ms[i];
  return 0;
}

int b2_timer_start(b2_handle* h) {
  if (!h) return fail("b2_timer_start: NULL handle");
  if (cudaEventRecord(h->eng.tev[0], h->eng.stream) != cudaSuccess) return fail("b2_timer_start: record failed");
  return 0;
}

int b2_timer_stop(b2_handle* h, double* ms) {
  if (!h || !ms) return fail("b2_timer_stop: NULL argument");
  if (cudaEventRecord(h->eng.tev[1], h->eng.stream) != cudaSuccess) return fail("b2_timer_stop: record failed");
  if (cudaEventSynchronize(h->eng.tev[1]) != cudaSuccess) return fail("b2_timer_stop: sync failed");
  float f = 0;
  if (cudaEventElapsedTime(&f, h->eng.tev[0], h->eng.tev[1]) != cudaSuccess) return fail("b2_timer_stop: elapsed failed");
  *ms = f;
  return 0;
}

int b2_profile(b2_handle* h, int which, int max, int* kinds, int* cls, int* counts, double* ms, int* n) {
  if (!h || !kinds || !cls || !counts || !ms || !n) return fail("b2_profile: NULL argument");
  return h->eng.profile(which, max, kinds, cls, counts, ms, n);
}

int b2_front_sizes(const b2_handle* h, int64_t max, int32_t* width, int32_t* order, int32_t* level) {
  if (!h || !width || !order || !level) return fail("b2_front_sizes: NULL argument");
  const b2::Symbolic& S = h->eng.sym;
  for (int64_t s = 0; s < S.nsuper && s < max; s++) {
    width[s] = S.scol[s + 1] - S.scol[s];
    order[s] = (int32_t)(S.rptr[s + 1] - S.rptr[s]);
    level[s] = S.slevel[s];
  }
  return 0;
}

int b2_last_sweeps(const b2_handle* h) { return h ? h->eng.last_sweeps : -1; }

int b2_get_perm(const b2_handle* h, int64_t* perm0) {
  if (!h || !perm0) return fail("b2_get_perm: NULL argument");
  for (int64_t k = 0; k < h->eng.sym.N; k++) perm0[k] = h->eng.sym.perm[k];
  return 0;
}

int b2_get_csc(const b2_handle* h, int64_t* colptr0, int64_t* rowval0) {
  if (!h || !colptr0 || !rowval0) return fail("b2_get_csc: NULL argument");
  const b2::Symbolic& S = h->eng.sym;
  for (int64_t j = 0; j <= S.N; j++) colptr0[j] = S.Ap[j];
  for (int64_t p = 0; p < S.nnzA; p++) rowval0[p] = S.Ai[p];
  return 0;
}

int b2_get_nzval(b2_handle* h, double* nzval) {
  if (!h || !nzval) return fail("b2_get_nzval: NULL argument");
  cudaStreamSynchronize(h->eng.stream);
  if (cudaMemcpy(nzval, h->eng.d_nzval, (size_t)h->eng.sym.nnzA * sizeof(double), cudaMemcpyDeviceToHost) != cudaSuccess)
    return fail("b2_get_nzval: copy failed");
  return 0;
}

int b2_get_d(b2_handle* h, double* d) {
  if (!h || !d) return fail("b2_get_d: NULL argument");
  cudaStreamSynchronize(h->eng.stream);
  if (cudaMemcpy(d, h->eng.d_dvec, (size_t)h->eng.sym.N * sizeof(double), cudaMemcpyDeviceToHost) != cudaSuccess)
    return fail("b2_get_d: copy failed");
  return 0;
}

int b2_set_option(b2_handle* h, const char* key, double value) {
  if (!h || !key) return fail("b2_set_option: NULL argument");
  if (!strcmp(key, "use_graph")) { h->eng.use_graph = value != 0; return 0; }
  if (!strcmp(key, "refine_tol")) { h->eng.refine_tol = value; return 0; }
  return fail("b2_set_option: unknown key");
}

int b2_free(b2_handle* h) {
  if (!h) return 0;
  h->eng.destroy();
  delete h;
  return 0;
}

int b2_dev_malloc(void** dptr, size_t bytes) {
  if (!dptr) return fail("b2_dev_malloc: NULL");
  if (cudaMalloc(dptr, bytes ? bytes : 1) != cudaSuccess) return fail("b2_dev_malloc: cudaMalloc failed");
  return 0;
}
int b2_dev_free(void* dptr) { cudaFree(dptr); return 0; }
int b2_dev_upload(void* dptr, const void* hptr, size_t bytes) {
  if (cudaMemcpy(dptr, hptr, bytes, cudaMemcpyHostToDevice) != cudaSuccess) return fail("b2_dev_upload failed");
  return 0;
}
int b2_dev_download(void* hptr, const void* dptr, size_t bytes) {
  if (cudaMemcpy(hptr, dptr, bytes, cudaMemcpyDeviceToHost) != cudaSuccess) return fail("b2_dev_download failed");
  return 0;
}
int b2_host_register(void* ptr, size_t bytes) {
  if (!ptr) return fail("b2_host_register: NULL");
  if (cudaHostRegister(ptr, bytes, cudaHostRegisterDefault) != cudaSuccess) {
    cudaGetLastError();
    return fail("b2_host_register: cudaHostRegister failed");
  }
  return 0;
}
int b2_host_unregister(void* ptr) {
  if (cudaHostUnregister(ptr) != cudaSuccess) { cudaGetLastError(); return fail("b2_host_unregister failed"); }
  return 0;
}
int b2_dev_sync(void) {
  if (cudaDeviceSynchronize() != cudaSuccess) return fail("b2_dev_sync failed");
  return 0;
}

}  // extern "C"
