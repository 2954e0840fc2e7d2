// symbolic.cpp -- see symbolic.h.  Host-side, once per solver (the `ldl_analyze` slot of
// reference/src/solver_types.jl:63).
#include "symbolic.h"

#include <algorithm>
#include <chrono>
#include <cstring>
#include <numeric>

namespace b2 {

namespace {

double now_s() {
  return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

// Liu's elimination tree of an upper-triangular CSC pattern (column k holds rows i <= k).
void etree_upper(int64_t n, const std::vector<int64_t>& Cp, const std::vector<int32_t>& Ci,
                 std::vector<int32_t>& parent) {
  parent.assign(n, -1);
  std::vector<int32_t> ancestor(n, -1);
  for (int64_t k = 0; k < n; k++) {
    for (int64_t p = Cp[k]; p < Cp[k + 1]; p++) {
      int32_t i = Ci[p];
      while (i != -1 && i < k) {
        int32_t inext = ancestor[i];
        ancestor[i] = (int32_t)k;
        if (inext == -1) parent[i] = (int32_t)k;
        i = inext;
      }
    }
  }
}

// postorder of a forest; children visited in ascending index order
void postorder(int64_t n, const std::vector<int32_t>& parent, std::vector<int32_t>& post) {
  std::vector<int32_t> head(n, -1), next(n, -1), stack(n);
  for (int64_t j = n - 1; j >= 0; j--) {
    if (parent[j] == -1) continue;
    next[j] = head[parent[j]];
    head[parent[j]] = (int32_t)j;
  }
  post.resize(n);
  int64_t k = 0;
  for (int64_t root = 0; root < n; root++) {
    if (parent[root] != -1) continue;
    int64_t top = 0;
    stack[0] = (int32_t)root;
    while (top >= 0) {
      int32_t p = stack[top];
      int32_t i = head[p];
      if (i == -1) {
        top--;
        post[k++] = p;
      } else {
        head[p] = next[i];
        stack[++top] = i;
      }
    }
  }
}

// C = triu(P A P') from the upper CSC of A.  Column max(pi,pj) receives row min(pi,pj).
void permute_upper(int64_t n, const std::vector<int64_t>& Ap, const std::vector<int32_t>& Ai,
                   const std::vector<int32_t>& pinv, std::vector<int64_t>& Cp,
                   std::vector<int32_t>& Ci) {
  Cp.assign(n + 1, 0);
  for (int64_t j = 0; j < n; j++)
    for (int64_t p = Ap[j]; p < Ap[j + 1]; p++) {
      int32_t a = pinv[Ai[p]], b = pinv[j];
      Cp[std::max(a, b) + 1]++;
    }
  for (int64_t j = 0; j < n; j++) Cp[j + 1] += Cp[j];
  Ci.resize(Cp[n]);
  std::vector<int64_t> cur(Cp.begin(), Cp.end() - 1);
  for (int64_t j = 0; j < n; j++)
    for (int64_t p = Ap[j]; p < Ap[j + 1]; p++) {
      int32_t a = pinv[Ai[p]], b = pinv[j];
      Ci[cur[std::max(a, b)]++] = std::min(a, b);
    }
}

}  // namespace

bool analyze(int64_t N, int64_t nnz, const int64_t* rows1, const int64_t* cols1, int64_t nvar,
             int64_t nequ, int64_t ncon, const SymbolicOptions& opt, Symbolic& S) {
  const double t_begin = now_s();
  S = Symbolic();
  S.N = N; S.nnz = nnz; S.nvar = nvar; S.nequ = nequ; S.ncon = ncon;
  if (N <= 0 || nnz < 0) { S.error = "empty system"; return false; }
  if (N >= (int64_t)INT32_MAX || nnz >= (int64_t)INT32_MAX) {
    S.error = "N and nnz must fit in int32"; return false;
  }
  if (nvar + nequ + ncon != N) { S.error = "nvar + nequ + ncon != N"; return false; }

  // ------------------------------------------------------------------ 1. A pattern + slots
  // Entry t lives in column j = rows[t]-1, row i = cols[t]-1 of the upper CSC
  // (sparse(cols, rows, vals) at src/solver_types.jl:62).
  for (int64_t t = 0; t < nnz; t++) {
    int64_t r = rows1[t] - 1, c = cols1[t] - 1;
    if (r < 0 || r >= N || c < 0 || c >= N) { S.error = "COO index out of range"; return false; }
    if (r < c) {
      S.error = "strictly upper-triangular COO entry (the reference's triu would drop it)";
      return false;
    }
  }
  {
    // stable two-pass counting sort by (column j, row i); ties keep increasing t
    std::vector<int64_t> cnt(N + 1, 0);
    std::vector<int32_t> tmp(nnz);
    for (int64_t t = 0; t < nnz; t++) cnt[cols1[t]]++;          // key i = cols1-1 -> cnt[i+1]
    for (int64_t i = 0; i < N; i++) cnt[i + 1] += cnt[i];
    for (int64_t t = 0; t < nnz; t++) tmp[cnt[cols1[t] - 1]++] = (int32_t)t;
    std::fill(cnt.begin(), cnt.end(), 0);
    for (int64_t t = 0; t < nnz; t++) cnt[rows1[t]]++;
    for (int64_t j = 0; j < N; j++) cnt[j + 1] += cnt[j];
    S.coo_sorted.resize(nnz);
    for (int64_t q = 0; q < nnz; q++) {
      int32_t t = tmp[q];
      S.coo_sorted[cnt[rows1[t] - 1]++] = t;
    }
    S.Ap.assign(N + 1, 0);
    S.Ai.clear(); S.Ai.reserve(nnz);
    S.slot_ptr.clear(); S.slot_ptr.reserve(nnz + 1);
    S.coo_slot.resize(nnz);
    int64_t pj = -1, pi = -1;
    for (int64_t q = 0; q < nnz; q++) {
      int32_t t = S.coo_sorted[q];
      int64_t j = rows1[t] - 1, i = cols1[t] - 1;
      if (j != pj || i != pi) {
        S.slot_ptr.push_back(q);
        S.Ai.push_back((int32_t)i);
        S.Ap[j + 1]++;
        pj = j; pi = i;
      }
      S.coo_slot[t] = (int32_t)(S.Ai.size() - 1);
    }
    S.slot_ptr.push_back(nnz);
    S.nnzA = (int64_t)S.Ai.size();
    for (int64_t j = 0; j < N; j++) S.Ap[j + 1] += S.Ap[j];
  }
  // rho / delta segments: trailing ncon + nvar diagonal entries
  S.shift_ok = (nnz >= nvar + ncon);
  if (S.shift_ok) {
    S.rho_slot.resize(nvar);
    S.delta_slot.resize(ncon);
    for (int64_t i = 0; i < nvar && S.shift_ok; i++) {
      int64_t t = nnz - nvar + i;
      if (rows1[t] != i + 1 || cols1[t] != i + 1) { S.shift_ok = false; break; }
      int32_t s = S.coo_slot[t];
      if (S.coo_sorted[S.slot_ptr[s + 1] - 1] != t) { S.shift_ok = false; break; }
      S.rho_slot[i] = s;
    }
    for (int64_t j = 0; j < ncon && S.shift_ok; j++) {
      int64_t t = nnz - nvar - ncon + j;
      int64_t dgi = nvar + nequ + j + 1;
      if (rows1[t] != dgi || cols1[t] != dgi) { S.shift_ok = false; break; }
      int32_t s = S.coo_slot[t];
      if (S.coo_sorted[S.slot_ptr[s + 1] - 1] != t) { S.shift_ok = false; break; }
      S.delta_slot[j] = s;
    }
    if (!S.shift_ok) { S.rho_slot.clear(); S.delta_slot.clear(); }
  }

  // ------------------------------------------------------------------ 2. ordering
  const double t_ord0 = now_s();
  std::vector<int32_t> perm0(N);
  if (opt.ordering == ORDER_NATURAL) {
    std::iota(perm0.begin(), perm0.end(), 0);
  } else if (opt.ordering == ORDER_USER) {
    if (!opt.user_perm) { S.error = "ORDER_USER without a permutation"; return false; }
    std::vector<char> seen(N, 0);
    for (int64_t k = 0; k < N; k++) {
      int64_t v = opt.user_perm[k];
      if (v < 0 || v >= N || seen[v]) { S.error = "user permutation is not a permutation"; return false; }
      seen[v] = 1;
      perm0[k] = (int32_t)v;
    }
  } else {
    std::vector<int64_t> xadj(N + 1, 0);
    for (int64_t j = 0; j < N; j++)
      for (int64_t p = S.Ap[j]; p < S.Ap[j + 1]; p++) {
        int32_t i = S.Ai[p];
        if (i != j) { xadj[i + 1]++; xadj[j + 1]++; }
      }
    for (int64_t j = 0; j < N; j++) xadj[j + 1] += xadj[j];
    std::vector<int64_t> adj(xadj[N]);
    std::vector<int64_t> cur(xadj.begin(), xadj.end() - 1);
    for (int64_t j = 0; j < N; j++)
      for (int64_t p = S.Ap[j]; p < S.Ap[j + 1]; p++) {
        int32_t i = S.Ai[p];
        if (i != j) { adj[cur[i]++] = j; adj[cur[j]++] = i; }
      }
    bool ok;
    if (opt.ordering == ORDER_AMD) ok = order_amd(N, xadj, adj, perm0, S.error);
    else if (opt.ordering == ORDER_ND && order_kkt_compressed_nd(N, nvar, xadj, adj, perm0, S.error)) ok = true;
    else ok = S.error.empty() && order_metis_nd(N, xadj, adj, perm0, S.error);
    if (!ok) return false;
  }
  S.t_order = now_s() - t_ord0;

  // ------------------------------------------------------------------ 3. etree + postorder
  std::vector<int32_t> pinv0(N);
  for (int64_t k = 0; k < N; k++) pinv0[perm0[k]] = (int32_t)k;
  std::vector<int64_t> Cp;
  std::vector<int32_t> Ci;
  {
    permute_upper(N, S.Ap, S.Ai, pinv0, Cp, Ci);
    std::vector<int32_t> par0, post;
    etree_upper(N, Cp, Ci, par0);
    postorder(N, par0, post);
    S.perm.resize(N);
    S.pinv.resize(N);
    for (int64_t k = 0; k < N; k++) S.perm[k] = perm0[post[k]];
    for (int64_t k = 0; k < N; k++) S.pinv[S.perm[k]] = (int32_t)k;
  }
  permute_upper(N, S.Ap, S.Ai, S.pinv, Cp, Ci);
  etree_upper(N, Cp, Ci, S.parent);
  const std::vector<int32_t>& parent = S.parent;
  for (int64_t k = 0; k < N; k++)
    if (parent[k] != -1 && parent[k] <= k) { S.error = "internal: etree not topologically ordered"; return false; }

  // strict lower pattern B = (strict upper C)': column j holds rows i > j
  std::vector<int64_t> Bp(N + 1, 0);
  std::vector<int32_t> Bi;
  {
    for (int64_t k = 0; k < N; k++)
      for (int64_t p = Cp[k]; p < Cp[k + 1]; p++)
        if (Ci[p] != k) Bp[Ci[p] + 1]++;
    for (int64_t j = 0; j < N; j++) Bp[j + 1] += Bp[j];
    Bi.resize(Bp[N]);
    std::vector<int64_t> cur(Bp.begin(), Bp.end() - 1);
    for (int64_t k = 0; k < N; k++)
      for (int64_t p = Cp[k]; p < Cp[k + 1]; p++)
        if (Ci[p] != k) Bi[cur[Ci[p]]++] = (int32_t)k;
  }

  // ------------------------------------------------------------------ 4. column counts
  // exact |L(:,k)| by the row-subtree walk (O(nnz L) steps, once per solver)
  {
    std::vector<int32_t>& cc = S.colcount;
    cc.assign(N, 1);
    std::vector<int32_t> flag(N, -1);
    for (int64_t k = 0; k < N; k++) {
      flag[k] = (int32_t)k;
      for (int64_t p = Cp[k]; p < Cp[k + 1]; p++) {
        int32_t i = Ci[p];
        for (; i < k && flag[i] != k; i = parent[i]) {
          cc[i]++;
          flag[i] = (int32_t)k;
        }
      }
    }
  }
  S.nnzL = 0; S.flops = 0;
  for (int64_t k = 0; k < N; k++) {
    double c = (double)(S.colcount[k] - 1);
    S.nnzL += S.colcount[k] - 1;
    S.flops += c * c + 3.0 * c;
  }

  // ------------------------------------------------------------------ 5. supernodes
  const std::vector<int32_t>& cc = S.colcount;
  struct SN { int32_t start, w, m; double zeros; };
  std::vector<SN> stack;
  stack.reserve(N);
  {
    int64_t k = 0;
    while (k < N) {
      int64_t e = k + 1;
      while (e < N && parent[e - 1] == e && cc[e - 1] == cc[e] + 1) e++;
      SN p{(int32_t)k, (int32_t)(e - k), cc[k], 0.0};
      // relaxed amalgamation with the supernode that ends right before p, if p is its parent
      while (!stack.empty()) {
        const SN& c = stack.back();
        int32_t clast = c.start + c.w - 1;
        if (parent[clast] < p.start || parent[clast] >= p.start + p.w) break;  // not a child
        // (parent[clast] is then the first column of p's own chain or inside the merged block)
        double wc = c.w, mc = c.m, wp = p.w, mp = p.m;
        double w = wc + wp, m = wc + mp;
        double ztot = c.zeros + p.zeros + wc * (m - mc);
        double lnz = w * m - w * (w - 1) / 2;
        double zf = ztot / lnz;
        bool merge = (w <= opt.relax_always) || (w <= opt.relax_w1 && zf < opt.relax_z1) ||
                     (w <= opt.relax_w2 && zf < opt.relax_z2) || (zf < opt.relax_z3);
        if (!merge) break;
        p.start = c.start; p.w = (int32_t)w; p.m = (int32_t)m; p.zeros = ztot;
        stack.pop_back();
      }
      stack.push_back(p);
      k = e;
    }
  }
  S.nsuper = (int32_t)stack.size();
  const int32_t ns = S.nsuper;
  S.scol.resize(ns + 1);
  for (int32_t s = 0; s < ns; s++) S.scol[s] = stack[s].start;
  S.scol[ns] = (int32_t)N;
  S.col2sn.resize(N);
  for (int32_t s = 0; s < ns; s++)
    for (int32_t k = S.scol[s]; k < S.scol[s + 1]; k++) S.col2sn[k] = s;
  S.sparent.assign(ns, -1);
  for (int32_t s = 0; s < ns; s++) {
    int32_t last = S.scol[s + 1] - 1;
    if (parent[last] != -1) S.sparent[s] = S.col2sn[parent[last]];
  }
  // children lists
  S.child_ptr.assign(ns + 1, 0);
  for (int32_t s = 0; s < ns; s++)
    if (S.sparent[s] != -1) S.child_ptr[S.sparent[s] + 1]++;
  for (int32_t s = 0; s < ns; s++) S.child_ptr[s + 1] += S.child_ptr[s];
  S.child_idx.resize(S.child_ptr[ns]);
  {
    std::vector<int32_t> cur(S.child_ptr.begin(), S.child_ptr.end() - 1);
    for (int32_t s = 0; s < ns; s++)
      if (S.sparent[s] != -1) S.child_idx[cur[S.sparent[s]]++] = s;
  }

  // ------------------------------------------------------------------ 6. front row lists
  S.rptr.assign(ns + 1, 0);
  S.rowidx.clear();
  S.rowidx.reserve((size_t)N * 2);
  {
    std::vector<int32_t> mark(N, -1);
    std::vector<int32_t> buf;
    for (int32_t s = 0; s < ns; s++) {
      int32_t c0 = S.scol[s], c1 = S.scol[s + 1];
      buf.clear();
      for (int32_t j = c0; j < c1; j++)
        for (int64_t p = Bp[j]; p < Bp[j + 1]; p++) {
          int32_t i = Bi[p];
          if (i >= c1 && mark[i] != s) { mark[i] = s; buf.push_back(i); }
        }
      for (int32_t q = S.child_ptr[s]; q < S.child_ptr[s + 1]; q++) {
        int32_t c = S.child_idx[q];
        int32_t wc = S.scol[c + 1] - S.scol[c];
        for (int64_t p = S.rptr[c] + wc; p < S.rptr[c + 1]; p++) {
          int32_t i = S.rowidx[p];
          if (i >= c1 && mark[i] != s) { mark[i] = s; buf.push_back(i); }
        }
      }
      std::sort(buf.begin(), buf.end());
      for (int32_t j = c0; j < c1; j++) S.rowidx.push_back(j);
      S.rowidx.insert(S.rowidx.end(), buf.begin(), buf.end());
      S.rptr[s + 1] = (int64_t)S.rowidx.size();
    }
  }
  // relative indices into the parent front, storage offsets, stats
  S.rel.assign(S.rowidx.size(), -1);
  S.lptr.assign(ns + 1, 0);
  S.cbptr.assign(ns + 1, 0);
  S.uptr.assign(ns + 1, 0);
  S.flops_store = 0;
  for (int32_t s = 0; s < ns; s++) {
    int64_t w = S.scol[s + 1] - S.scol[s];
    int64_t m = S.rptr[s + 1] - S.rptr[s];
    int64_t r = m - w;
    if (m * w >= (int64_t)INT32_MAX) { S.error = "front panel too large for int32 offsets"; return false; }
    S.lptr[s + 1] = S.lptr[s] + m * w;
    S.cbptr[s + 1] = S.cbptr[s] + r * r;
    S.uptr[s + 1] = S.uptr[s] + r;
    S.max_front = std::max<int32_t>(S.max_front, (int32_t)m);
    S.max_width = std::max<int32_t>(S.max_width, (int32_t)w);
    for (int64_t k = 0; k < w; k++) {
      double c = (double)(m - k - 1);
      S.flops_store += c * c + 3.0 * c;
    }
    int32_t p = S.sparent[s];
    if (r > 0) {
      if (p < 0) { S.error = "internal: root front with rows below"; return false; }
      const int32_t* prow = &S.rowidx[S.rptr[p]];
      int64_t pm = S.rptr[p + 1] - S.rptr[p];
      int64_t q = 0;
      for (int64_t i = w; i < m; i++) {
        int32_t g = S.rowidx[S.rptr[s] + i];
        while (q < pm && prow[q] < g) q++;
        if (q >= pm || prow[q] != g) { S.error = "internal: child row missing in parent front"; return false; }
        S.rel[S.rptr[s] + i] = (int32_t)q;
      }
    }
  }
  S.nnzL_store = S.lptr[ns];
  S.cb_store = S.cbptr[ns];

  // ------------------------------------------------------------------ 7. CSC slot -> front map
  {
    S.amap_ptr.assign(ns + 1, 0);
    std::vector<int32_t> sn_of(S.nnzA), pos_of(S.nnzA);
    for (int64_t j = 0; j < N; j++)
      for (int64_t p = S.Ap[j]; p < S.Ap[j + 1]; p++) {
        int32_t a = S.pinv[S.Ai[p]], b = S.pinv[j];
        int32_t col = std::min(a, b), row = std::max(a, b);
        int32_t s = S.col2sn[col];
        int64_t m = S.rptr[s + 1] - S.rptr[s];
        int32_t c1 = S.scol[s + 1];
        int64_t rpos;
        if (row < c1) {
          rpos = row - S.scol[s];
        } else {
          const int32_t* b0 = &S.rowidx[S.rptr[s] + (c1 - S.scol[s])];
          const int32_t* b1 = &S.rowidx[S.rptr[s + 1]];
          const int32_t* it = std::lower_bound(b0, b1, row);
          if (it == b1 || *it != row) { S.error = "internal: A entry outside front structure"; return false; }
          rpos = (c1 - S.scol[s]) + (it - b0);
        }
        sn_of[p] = s;
        pos_of[p] = (int32_t)(rpos + (int64_t)(col - S.scol[s]) * m);
        S.amap_ptr[s + 1]++;
      }
    for (int32_t s = 0; s < ns; s++) S.amap_ptr[s + 1] += S.amap_ptr[s];
    S.amap_slot.resize(S.nnzA);
    S.amap_pos.resize(S.nnzA);
    std::vector<int64_t> cur(S.amap_ptr.begin(), S.amap_ptr.end() - 1);
    for (int64_t p = 0; p < S.nnzA; p++) {
      int64_t dst = cur[sn_of[p]]++;
      S.amap_slot[dst] = (int32_t)p;
      S.amap_pos[dst] = pos_of[p];
    }
    // sort each front's entries by position (k_assemble_large bisects on it)
    std::vector<std::pair<int32_t, int32_t>> tmp;
    for (int32_t s = 0; s < ns; s++) {
      int64_t a = S.amap_ptr[s], b = S.amap_ptr[s + 1];
      bool sorted = true;
      for (int64_t q = a + 1; q < b; q++)
        if (S.amap_pos[q] < S.amap_pos[q - 1]) { sorted = false; break; }
      if (sorted) continue;
      tmp.resize(b - a);
      for (int64_t q = a; q < b; q++) tmp[q - a] = {S.amap_pos[q], S.amap_slot[q]};
      std::sort(tmp.begin(), tmp.end());
      for (int64_t q = a; q < b; q++) { S.amap_pos[q] = tmp[q - a].first; S.amap_slot[q] = tmp[q - a].second; }
    }
  }

  // ------------------------------------------------------------------ 8. level sets
  S.slevel.assign(ns, 0);
  for (int32_t s = 0; s < ns; s++) {
    int32_t p = S.sparent[s];
    if (p != -1) S.slevel[p] = std::max(S.slevel[p], S.slevel[s] + 1);
  }
  S.nlevels = 0;
  for (int32_t s = 0; s < ns; s++) S.nlevels = std::max(S.nlevels, S.slevel[s] + 1);
  S.level_ptr.assign(S.nlevels + 1, 0);
  for (int32_t s = 0; s < ns; s++) S.level_ptr[S.slevel[s] + 1]++;
  for (int32_t l = 0; l < S.nlevels; l++) S.level_ptr[l + 1] += S.level_ptr[l];
  S.level_sn.resize(ns);
  {
    std::vector<int32_t> cur(S.level_ptr.begin(), S.level_ptr.end() - 1);
    for (int32_t s = 0; s < ns; s++) S.level_sn[cur[S.slevel[s]]++] = s;
  }

  // ------------------------------------------------------------------ 9. symmetric CSR for K x
  if (opt.build_spmv) {
    S.Sp.assign(N + 1, 0);
    for (int64_t j = 0; j < N; j++)
      for (int64_t p = S.Ap[j]; p < S.Ap[j + 1]; p++) {
        int32_t i = S.Ai[p];
        S.Sp[i + 1]++;
        if (i != j) S.Sp[j + 1]++;
      }
    for (int64_t j = 0; j < N; j++) S.Sp[j + 1] += S.Sp[j];
    S.Sj.resize(S.Sp[N]);
    S.Sslot.resize(S.Sp[N]);
    std::vector<int64_t> cur(S.Sp.begin(), S.Sp.end() - 1);
    for (int64_t j = 0; j < N; j++)
      for (int64_t p = S.Ap[j]; p < S.Ap[j + 1]; p++) {
        int32_t i = S.Ai[p];
        int64_t d = cur[i]++;
        S.Sj[d] = (int32_t)j; S.Sslot[d] = (int32_t)p;
        if (i != j) {
          d = cur[j]++;
          S.Sj[d] = i; S.Sslot[d] = (int32_t)p;
        }
      }
  }
  S.t_symbolic = now_s() - t_begin - S.t_order;
  return true;
}

}  // namespace b2
