// batched.cu -- batch of independent KKT systems sharing one sparsity pattern (BASELINE.json
// config 5: multi-start / per-sample estimation).  Host side of batched_kernels.cuh: one symbolic
// analysis for the shared pattern, then every verb is ONE kernel launch with one CTA per instance.
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <new>
#include <vector>

#include "../../include/cannoles_b200.h"
#include "b2_cuda.h"
#include "batched_kernels.cuh"
#include "nls_kernels.cuh"
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
  // device-resident solver loop (nls_kernels.cuh)
  static constexpr int NLS_SLOTS = 3;          // chunk kernels in flight (one stream each)
  int nsm = 0, nls_smem = 0;
  bool nls_ready = false;
  cudaStream_t nls_stream[NLS_SLOTS] = {nullptr, nullptr, nullptr}, copy_stream = nullptr;
  cudaEvent_t nls_done[NLS_SLOTS] = {nullptr, nullptr, nullptr}, nls_start = nullptr;
  std::vector<cudaEvent_t> chunk_ev;
  double* nls_scr[NLS_SLOTS] = {nullptr, nullptr, nullptr};   // per slot: nsm x nnz COO values
  int* nls_tickets = nullptr;
  int nls_ntickets = 0;
  double* nls_rec = nullptr;                   // batch x record
  double* nls_model[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};   // At Bt Ct y e x0 (host-data verb)
  size_t nls_model_cap[6] = {0, 0, 0, 0, 0, 0};
  int nls_launches = 0;
  int nls_init(int n, int m, int nc);
  // asynchronous submissions (b2b_nls_dense_submit): every submission runs start to end on ONE lane
  // (its H2D, its kernel, its D2H in stream order, own model buffers / scratch / ticket), successive
  // submissions on successive lanes, so that batches overlap: the copy engines of one with the SMs of
  // another, and the SMs a batch leaves idle while its slowest instance finishes with the next batch
  static constexpr int NLS_LANES = 128;
  struct NlsLane {
    cudaStream_t q = nullptr;
    cudaEvent_t done = nullptr;
    double* scr = nullptr;
    int* ticket = nullptr;
    double* model[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    size_t cap[6] = {0, 0, 0, 0, 0, 0};
    double* rec = nullptr;
    size_t rec_cap = 0;
    bool pending = false;
  };
  NlsLane lanes[NLS_LANES];
  int next_lane = 0;
  cudaEvent_t lane_start = nullptr;

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
  // width-1 supernodes per level, flat (factorization leaves first)
  std::vector<int32_t> lw_ptr(1, 0), lw_nleaf;
  std::vector<int4> lw_meta;
  std::vector<int2> lf_ent;
  for (int l = 0; l < S.nlevels; l++) {
    int nleaf = 0;
    for (int pass = 0; pass < 2; pass++)
      for (int q = S.level_ptr[l]; q < S.level_ptr[l + 1]; q++) {
        const int s = S.level_sn[q];
        if (S.scol[s + 1] - S.scol[s] != 1) continue;
        const bool leaf = ct_ptr[s + 1] == ct_ptr[s];
        if (leaf != (pass == 0)) continue;
        const int k = S.scol[s];
        int4 mt;
        mt.x = k; mt.y = (int)lf_ent.size(); mt.z = rb_ptr[s + 1] - rb_ptr[s]; mt.w = cbm[k];
        lw_meta.push_back(mt);
        for (int32_t p = rb_ptr[s]; p < rb_ptr[s + 1]; p++) {
          int2 en;
          en.x = cbm[k] + rb_idx[p]; en.y = k | (rb_idx[p] << 16);
          lf_ent.push_back(en);
        }
        nleaf += leaf;
      }
    lw_nleaf.push_back(nleaf);
    lw_ptr.push_back((int32_t)lw_meta.size());
  }
  {
    int4 mt;
    mt.x = 0; mt.y = (int)lf_ent.size(); mt.z = 0; mt.w = 0;
    lw_meta.push_back(mt);   // sentinel: end of the entries
  }
  plan.N = N; plan.nnz = (int)S.nnz; plan.nsuper = S.nsuper; plan.nphase = S.nlevels;
  plan.npacked = (int)npacked; plan.nvar = (int)S.nvar; plan.nequ = (int)S.nequ; plan.ncon = (int)S.ncon;
  plan.nmulti = (int)multi_dst.size();
  if (up(S.perm, &plan.perm) || up(cbm, &plan.cbm) || up(S.scol, &plan.sc0) || up(rb_ptr, &plan.rb_ptr) ||
      up(rb_idx, &plan.rb_idx) || up(ct_ptr, &plan.ct_ptr) || up(ct_col, &plan.ct_col) ||
      up(S.level_ptr, &plan.ph_ptr) || up(S.level_sn, &plan.ph_sn) || up(phw_ptr, &plan.phw_ptr) ||
      up(phw_sn, &plan.phw_sn) || up(lw_ptr, &plan.lw_ptr) || up(lw_nleaf, &plan.lw_nleaf) || up(lw_meta, &plan.lw_meta) ||
      up(lf_ent, &plan.lf_ent) || up(phb_ptr, &plan.phb_ptr) || up(phb_sn, &plan.phb_sn) || up(dst_single, &plan.dst_single) ||
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
  for (auto& e : nls_done) if (e) cudaEventDestroy(e);
  for (auto& e : chunk_ev) if (e) cudaEventDestroy(e);
  if (nls_start) cudaEventDestroy(nls_start);
  if (lane_start) cudaEventDestroy(lane_start);
  for (auto& L : lanes) {
    if (L.done) cudaEventDestroy(L.done);
    if (L.q) cudaStreamDestroy(L.q);
  }
  for (auto& q : nls_stream) if (q) cudaStreamDestroy(q);
  if (copy_stream) cudaStreamDestroy(copy_stream);
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

// Buffers, streams and the shared-memory opt-in of k_nls_dense; idempotent.
int BatchEngine::nls_init(int n, int m, int nc) {
  if (nls_ready) return 0;
  const int N = (int)sym.N;
  if (n != sym.nvar || m != sym.nequ || nc != sym.ncon) {
    snprintf(g_last_error, sizeof(g_last_error), "b2b_nls: model dimensions (%d, %d, %d) differ from the analysed KKT "
             "layout (%lld, %lld, %lld)", n, m, nc, (long long)sym.nvar, (long long)sym.nequ, (long long)sym.ncon);
    return -1;
  }
  const long long nh = (long long)n * (n + 1) / 2;
  const long long expect = nh + (nc > 0 ? nh : 0) + (long long)n * m + (long long)n * nc + m + nc + n;
  if (expect != sym.nnz || !sym.shift_ok) {
    snprintf(g_last_error, sizeof(g_last_error), "b2b_nls: the analysed COO layout (nnz = %lld) is not the Newton-mode "
             "layout of a dense model with these dimensions (nnz = %lld, SURVEY App. B)", (long long)sym.nnz, expect);
    return -1;
  }
  if (m > BATCH_NT || n > BATCH_NT || (nc > BATCH_NT)) {
    snprintf(g_last_error, sizeof(g_last_error), "b2b_nls: n, m, ncon must not exceed %d", BATCH_NT);
    return -1;
  }
  nls_smem = (int)(((batched_smem_bytes(N, plan.npacked) + 15) & ~(size_t)15) +
                   nls_state_doubles(n, m, nc, BATCH_NT) * sizeof(double));
  if (nls_smem > 227 * 1024) {
    snprintf(g_last_error, sizeof(g_last_error), "b2b_nls: %d bytes of shared memory per instance (> 227 KB)", nls_smem);
    return -1;
  }
  B2_CUDA_OK(cudaSetDevice(device));
  cudaDeviceProp prop;
  B2_CUDA_OK(cudaGetDeviceProperties(&prop, device));
  nsm = prop.multiProcessorCount > 0 ? prop.multiProcessorCount : 1;
  for (auto& q : nls_stream) B2_CUDA_OK(cudaStreamCreateWithFlags(&q, cudaStreamNonBlocking));
  B2_CUDA_OK(cudaStreamCreateWithFlags(&copy_stream, cudaStreamNonBlocking));
  for (auto& e : nls_done) B2_CUDA_OK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  B2_CUDA_OK(cudaEventCreateWithFlags(&nls_start, cudaEventDisableTiming));
  for (auto& p : nls_scr)
    if (alloc(&p, (size_t)nsm * sym.nnz)) return -1;
  // the lanes of the asynchronous verb: everything a device-resident submission needs exists before the
  // first one (a cudaMalloc inside a stream of submissions synchronises the device and serialises them)
  B2_CUDA_OK(cudaEventCreateWithFlags(&lane_start, cudaEventDisableTiming));
  for (int i = 0; i < NLS_LANES; i++) {
    NlsLane& L = lanes[i];
    B2_CUDA_OK(cudaStreamCreateWithFlags(&L.q, cudaStreamNonBlocking));
    B2_CUDA_OK(cudaEventCreateWithFlags(&L.done, cudaEventDisableTiming));
    if (alloc(&L.scr, (size_t)nsm * sym.nnz) || alloc(&L.ticket, 1)) return -1;
  }
  nls_ntickets = 1024;
  if (alloc(&nls_tickets, (size_t)nls_ntickets)) return -1;
  if (alloc(&nls_rec, (size_t)batch * (NLS_REC_HEAD + n + nc))) return -1;
  B2_CUDA_OK(cudaFuncSetAttribute(k_nls_dense<BATCH_NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, nls_smem));
  nls_ready = true;
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
  if (cudaMemcpyFromSymbol(out64, b2::b2_dbg, 64 * sizeof(long long)) != cudaSuccess) return -1;
  static const long long zero[64] = {0};
  return cudaMemcpyToSymbol(b2::b2_dbg, zero, sizeof(zero)) == cudaSuccess ? 0 : -1;   // read and clear
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
  // the packed dense triangle stores structural zeros for free, but every column merged into the root
  // supernode lengthens its sequential pivot chain: only short chains are amalgamated (measured on
  // config 5: 129-column root 272 k cycles per instance, 65-column root + 143 leaves 197 k)
  opt.relax_always = 8; opt.relax_w1 = 16; opt.relax_z1 = 0.3; opt.relax_w2 = 32; opt.relax_z2 = 0.15; opt.relax_z3 = 0.0;
  if (const char* e = getenv("B2B_RELAX")) {   // developer knob: "always,w1,z1,w2,z2,z3"
    int a, w1, w2; double z1, z2, z3;
    if (sscanf(e, "%d,%d,%lf,%d,%lf,%lf", &a, &w1, &z1, &w2, &z2, &z3) == 6) {
      opt.relax_always = a; opt.relax_w1 = w1; opt.relax_z1 = z1; opt.relax_w2 = w2; opt.relax_z2 = z2; opt.relax_z3 = z3;
    }
  }
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

/* ---- device-resident CaNNOLeS loop for batches of small dense instances (nls_kernels.cuh) ---- */
void b2_nls_default_params(b2_nls_params_t* p) {   /* update!(params, eps(Float64)) : src/CaNNOLeS.jl:48-62 */
  if (!p) return;
  const double e = 2.220446049250313e-16;
  p->eig_tol = e; p->delta_min = std::sqrt(e); p->kappa_dec = 1.0 / 3.0; p->kappa_inc = 8.0; p->kappa_largeinc = 100.0;
  p->rho0 = std::pow(e, 1.0 / 3.0);   /* eps^T(1/3): pow, not cbrt -- they differ in the last bits and rho0 reaches the pivots */
  p->rho_max = std::pow(e, -2.0); p->rho_min = std::sqrt(e); p->gamma_A = std::pow(e, 0.25);
  p->atol = p->rtol = p->Fatol = std::sqrt(e); p->Frtol = e; p->delta_dec = 0.1; p->cgls_tol = std::sqrt(e);
  p->max_iter = -1; p->max_eval = 100000; p->max_inner = 10000; p->always_accept_extrapolation = 0;
  p->use_initial_multiplier = 0; p->reserved = 0;
}

int64_t b2b_nls_record_len(const b2b_handle* h) {
  if (!h) return -1;
  return b2::NLS_REC_HEAD + h->eng.sym.nvar + h->eng.sym.ncon;
}

namespace {
int nls_model_dev(const b2_dense_nls_t* md, b2::DenseNlsModel* M) {
  M->n = (int)md->n; M->m = (int)md->m; M->ncon = (int)md->ncon;
  const bool sh = md->shared_model != 0;
  M->stride_A = sh ? 0 : md->n * md->m; M->stride_C = sh ? 0 : md->n * md->ncon;
  M->stride_y = sh ? 0 : md->m; M->stride_e = sh ? 0 : md->ncon; M->stride_x0 = md->n;
  M->At = md->At; M->Bt = md->Bt; M->Ct = md->Ct; M->y = md->y; M->e = md->e; M->x0 = md->x0; M->y0 = md->y0;
  return 0;
}
b2::NlsParams nls_params(const b2_nls_params_t* p) {
  b2::NlsParams q;
  q.eig_tol = p->eig_tol; q.delta_min = p->delta_min; q.kappa_dec = p->kappa_dec; q.kappa_inc = p->kappa_inc;
  q.kappa_largeinc = p->kappa_largeinc; q.rho0 = p->rho0; q.rho_max = p->rho_max; q.rho_min = p->rho_min;
  q.gamma_A = p->gamma_A; q.atol = p->atol; q.rtol = p->rtol; q.Fatol = p->Fatol; q.Frtol = p->Frtol;
  q.delta_dec = p->delta_dec; q.cgls_tol = p->cgls_tol; q.eps2 = 2.220446049250313e-16 * 2.220446049250313e-16;
  q.max_iter = p->max_iter; q.max_eval = p->max_eval; q.max_inner = p->max_inner;
  q.always_accept_extrapolation = p->always_accept_extrapolation; q.use_initial_multiplier = p->use_initial_multiplier;
  return q;
}
}  // namespace

/* All pointers of *model_dev and d_records / d_dbg_vals are DEVICE pointers; instances 0 .. count-1. */
int b2b_nls_dense_solve_dev(b2b_handle* h, const b2_dense_nls_t* model_dev, int64_t count,
                            const b2_nls_params_t* params, double* d_records, double* d_dbg_vals) {
  if (!h || !model_dev || !params || !d_records) return failb("b2b_nls_dense_solve_dev: NULL argument");
  b2::BatchEngine& E = h->eng;
  if (count <= 0 || count > E.batch) return failb("b2b_nls_dense_solve_dev: count must be in 1 .. batch");
  if (E.nls_init((int)model_dev->n, (int)model_dev->m, (int)model_dev->ncon)) return -1;
  B2_CUDA_OK(cudaSetDevice(E.device));
  b2::DenseNlsModel M;
  nls_model_dev(model_dev, &M);
  const b2::NlsParams prm = nls_params(params);
  const int grid = (int)std::min<int64_t>(count, E.nsm);
  B2_CUDA_OK(cudaMemsetAsync(E.nls_tickets, 0, sizeof(int), E.stream));
  B2_CUDA_OK(cudaEventRecord(E.ev[0], E.stream));
  B2_LAUNCH(b2::k_nls_dense<b2::BATCH_NT>, (unsigned)grid, b2::BATCH_NT, E.nls_smem, E.stream, E.plan, M, prm, 0,
            (int)count, E.nls_tickets, E.nls_scr[0], d_records, (int)b2b_nls_record_len(h), d_dbg_vals);
  B2_CUDA_OK(cudaGetLastError());
  B2_CUDA_OK(cudaEventRecord(E.ev[1], E.stream));
  B2_CUDA_OK(cudaStreamSynchronize(E.stream));
  float ms = 0;
  cudaEventElapsedTime(&ms, E.ev[0], E.ev[1]);
  E.last_ms = ms;
  E.nls_launches = 1;
  return 0;
}

/* HOST model arrays (pin them with b2_host_register for full PCIe speed) and host records: the
 * model travels in chunks of `chunk` instances on a copy stream while the chunks that have arrived
 * are being solved (up to three chunk kernels in flight so that the tail of one chunk overlaps the
 * head of the next); records come back per chunk.  chunk <= 0 picks a size. */
int b2b_nls_dense_solve(b2b_handle* h, const b2_dense_nls_t* model_host, int64_t count,
                        const b2_nls_params_t* params, double* records, int64_t chunk) {
  if (!h || !model_host || !params || !records) return failb("b2b_nls_dense_solve: NULL argument");
  b2::BatchEngine& E = h->eng;
  if (count <= 0 || count > E.batch) return failb("b2b_nls_dense_solve: count must be in 1 .. batch");
  const b2_dense_nls_t& mh = *model_host;
  if (E.nls_init((int)mh.n, (int)mh.m, (int)mh.ncon)) return -1;
  B2_CUDA_OK(cudaSetDevice(E.device));
  const bool sh = mh.shared_model != 0;
  const size_t per[6] = {(size_t)(mh.n * mh.m), (size_t)(mh.n * mh.m), (size_t)(mh.n * mh.ncon), (size_t)mh.m,
                         (size_t)mh.ncon, (size_t)mh.n};
  const double* src[6] = {mh.At, mh.Bt, mh.Ct, mh.y, mh.e, mh.x0};
  for (int a = 0; a < 6; a++) {
    if (!src[a] && per[a]) return failb("b2b_nls_dense_solve: NULL model array");
    const size_t need = per[a] * (size_t)((sh && a < 5) ? 1 : count);
    if (E.nls_model_cap[a] < need) {
      if (E.alloc(&E.nls_model[a], need)) return -1;   // (an outgrown buffer stays owned by dev_ptrs until b2b_free)
      E.nls_model_cap[a] = need;
    }
  }
  if (chunk <= 0) chunk = std::max<int64_t>(2 * E.nsm, (count + 15) / 16);
  const int nchunk = (int)((count + chunk - 1) / chunk);
  if (nchunk > E.nls_ntickets) return failb("b2b_nls_dense_solve: too many chunks");
  while ((int)E.chunk_ev.size() < nchunk) {
    cudaEvent_t e;
    B2_CUDA_OK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    E.chunk_ev.push_back(e);
  }
  b2_dense_nls_t md = mh;
  md.At = E.nls_model[0]; md.Bt = E.nls_model[1]; md.Ct = E.nls_model[2]; md.y = E.nls_model[3];
  md.e = E.nls_model[4]; md.x0 = E.nls_model[5]; md.y0 = nullptr;
  if (mh.y0) return failb("b2b_nls_dense_solve: y0 is only supported by the _dev verb");
  b2::DenseNlsModel M;
  nls_model_dev(&md, &M);
  const b2::NlsParams prm = nls_params(params);
  const int rl = (int)b2b_nls_record_len(h);
  B2_CUDA_OK(cudaMemsetAsync(E.nls_tickets, 0, sizeof(int) * nchunk, E.stream));
  B2_CUDA_OK(cudaEventRecord(E.ev[0], E.stream));
  B2_CUDA_OK(cudaEventRecord(E.nls_start, E.stream));
  B2_CUDA_OK(cudaStreamWaitEvent(E.copy_stream, E.nls_start, 0));
  for (auto& q : E.nls_stream) B2_CUDA_OK(cudaStreamWaitEvent(q, E.nls_start, 0));
  if (sh)
    for (int a = 0; a < 5; a++)
      B2_CUDA_OK(cudaMemcpyAsync(E.nls_model[a], src[a], per[a] * sizeof(double), cudaMemcpyHostToDevice, E.copy_stream));
  for (int c = 0; c < nchunk; c++) {
    const int64_t c0 = (int64_t)c * chunk, cn = std::min<int64_t>(chunk, count - c0);
    for (int a = sh ? 5 : 0; a < 6; a++)
      B2_CUDA_OK(cudaMemcpyAsync(E.nls_model[a] + (size_t)c0 * per[a], src[a] + (size_t)c0 * per[a],
                                 (size_t)cn * per[a] * sizeof(double), cudaMemcpyHostToDevice, E.copy_stream));
    B2_CUDA_OK(cudaEventRecord(E.chunk_ev[c], E.copy_stream));
    cudaStream_t q = E.nls_stream[c % b2::BatchEngine::NLS_SLOTS];
    B2_CUDA_OK(cudaStreamWaitEvent(q, E.chunk_ev[c], 0));
    const int grid = (int)std::min<int64_t>(cn, E.nsm);
    B2_LAUNCH(b2::k_nls_dense<b2::BATCH_NT>, (unsigned)grid, b2::BATCH_NT, E.nls_smem, q, E.plan, M, prm, (int)c0, (int)cn,
              E.nls_tickets + c, E.nls_scr[c % b2::BatchEngine::NLS_SLOTS], E.nls_rec, rl, (double*)nullptr);
    B2_CUDA_OK(cudaGetLastError());
    B2_CUDA_OK(cudaMemcpyAsync(records + (size_t)c0 * rl, E.nls_rec + (size_t)c0 * rl, (size_t)cn * rl * sizeof(double),
                               cudaMemcpyDeviceToHost, q));
  }
  for (int q = 0; q < b2::BatchEngine::NLS_SLOTS; q++) {
    B2_CUDA_OK(cudaEventRecord(E.nls_done[q], E.nls_stream[q]));
    B2_CUDA_OK(cudaStreamWaitEvent(E.stream, E.nls_done[q], 0));
  }
  B2_CUDA_OK(cudaEventRecord(E.ev[1], E.stream));
  B2_CUDA_OK(cudaStreamSynchronize(E.stream));
  float ms = 0;
  cudaEventElapsedTime(&ms, E.ev[0], E.ev[1]);
  E.last_ms = ms;
  E.nls_launches = nchunk;
  return 0;
}

/* Asynchronous form: queues one batch (its upload, its solve, its download) on the next of the
 * handle's lanes and returns; b2b_nls_wait blocks until every queued batch is complete.  Batches in
 * flight overlap each other (PCIe of one with the SMs of another; the SMs one batch leaves idle
 * while its slowest instance finishes take the next batch).  where = 0: *model and records are HOST
 * memory (pinned for true asynchrony); where = 1: both are DEVICE memory.  The caller keeps model
 * arrays and records alive and untouched until b2b_nls_wait returns. */
int b2b_nls_dense_submit(b2b_handle* h, const b2_dense_nls_t* model, int64_t count, const b2_nls_params_t* params,
                         double* records, int where) {
  if (!h || !model || !params || !records) return failb("b2b_nls_dense_submit: NULL argument");
  b2::BatchEngine& E = h->eng;
  if (count <= 0 || count > E.batch) return failb("b2b_nls_dense_submit: count must be in 1 .. batch");
  const b2_dense_nls_t& mh = *model;
  if (E.nls_init((int)mh.n, (int)mh.m, (int)mh.ncon)) return -1;
  B2_CUDA_OK(cudaSetDevice(E.device));
  b2::BatchEngine::NlsLane& L = E.lanes[E.next_lane];
  E.next_lane = (E.next_lane + 1) % b2::BatchEngine::NLS_LANES;
  const int rl = (int)b2b_nls_record_len(h);
  const bool sh = mh.shared_model != 0;
  b2_dense_nls_t md = mh;
  double* drec = records;
  // the lane starts after whatever the caller queued on the handle's main stream (b2b_timer_start)
  B2_CUDA_OK(cudaEventRecord(E.lane_start, E.stream));
  B2_CUDA_OK(cudaStreamWaitEvent(L.q, E.lane_start, 0));
  if (where == 0) {
    if (mh.y0) return failb("b2b_nls_dense_submit: y0 needs device-resident data (where = 1)");
    const size_t per[6] = {(size_t)(mh.n * mh.m), (size_t)(mh.n * mh.m), (size_t)(mh.n * mh.ncon), (size_t)mh.m,
                           (size_t)mh.ncon, (size_t)mh.n};
    const double* src[6] = {mh.At, mh.Bt, mh.Ct, mh.y, mh.e, mh.x0};
    for (int a = 0; a < 6; a++) {
      if (!src[a] && per[a]) return failb("b2b_nls_dense_submit: NULL model array");
      const size_t need = per[a] * (size_t)((sh && a < 5) ? 1 : count);
      if (L.cap[a] < need) {
        if (E.alloc(&L.model[a], need)) return -1;
        L.cap[a] = need;
      }
      B2_CUDA_OK(cudaMemcpyAsync(L.model[a], src[a], need * sizeof(double), cudaMemcpyHostToDevice, L.q));
    }
    if (L.rec_cap < (size_t)count * rl) {
      if (E.alloc(&L.rec, (size_t)count * rl)) return -1;
      L.rec_cap = (size_t)count * rl;
    }
    md.At = L.model[0]; md.Bt = L.model[1]; md.Ct = L.model[2]; md.y = L.model[3]; md.e = L.model[4];
    md.x0 = L.model[5]; md.y0 = nullptr;
    drec = L.rec;
  }
  b2::DenseNlsModel M;
  nls_model_dev(&md, &M);
  const b2::NlsParams prm = nls_params(params);
  const int grid = (int)std::min<int64_t>(count, E.nsm);
  B2_CUDA_OK(cudaMemsetAsync(L.ticket, 0, sizeof(int), L.q));
  B2_LAUNCH(b2::k_nls_dense<b2::BATCH_NT>, (unsigned)grid, b2::BATCH_NT, E.nls_smem, L.q, E.plan, M, prm, 0, (int)count,
            L.ticket, L.scr, drec, rl, (double*)nullptr);
  B2_CUDA_OK(cudaGetLastError());
  if (where == 0)
    B2_CUDA_OK(cudaMemcpyAsync(records, L.rec, (size_t)count * rl * sizeof(double), cudaMemcpyDeviceToHost, L.q));
  B2_CUDA_OK(cudaEventRecord(L.done, L.q));
  L.pending = true;
  E.nls_launches++;
  return 0;
}

int b2b_nls_wait(b2b_handle* h) {
  if (!h) return failb("b2b_nls_wait: NULL handle");
  b2::BatchEngine& E = h->eng;
  B2_CUDA_OK(cudaSetDevice(E.device));
  for (auto& L : E.lanes)
    if (L.pending) {
      B2_CUDA_OK(cudaStreamWaitEvent(E.stream, L.done, 0));   // so that b2b_timer_stop brackets the lanes
      L.pending = false;
    }
  B2_CUDA_OK(cudaStreamSynchronize(E.stream));
  E.next_lane = 0;    // everything is complete: the next stream of submissions starts on the first lane again
  return 0;           // (so that a repeated stream of K batches reuses the K model buffers it has allocated)
}

int b2b_free(b2b_handle* h) {
  if (!h) return 0;
  h->eng.destroy();
  delete h;
  return 0;
}

}  // extern "C"
