// ordering.cpp -- fill-reducing orderings for the host symbolic phase.
//
// Nested dissection comes from METIS (NodeND), shipped with the CUDA toolkit as
// libmetis_static.a (64-bit idx_t build; same library cusolverSpXcsrmetisndHost wraps).  It is
// a host-side, once-per-solver step standing where the reference calls AMD inside
// `ldl_analyze` (reference/src/solver_types.jl:63).
#include "symbolic.h"

#include <algorithm>
#include <cmath>
#include <cstdlib>

extern "C" {
int METIS_NodeND(int64_t* nvtxs, int64_t* xadj, int64_t* adjncy, int64_t* vwgt, int64_t* options,
                 int64_t* perm, int64_t* iperm);
int METIS_SetDefaultOptions(int64_t* options);
}

namespace b2 {

bool order_metis_nd(int64_t n, const std::vector<int64_t>& xadj, const std::vector<int64_t>& adj,
                    std::vector<int32_t>& perm, std::string& err) {
  perm.resize(n);
  if (n == 1 || adj.empty()) {
    for (int64_t i = 0; i < n; i++) perm[i] = (int32_t)i;
    return true;
  }
  std::vector<int64_t> xa(xadj), ad(adj), p(n), ip(n);
  int64_t options[40];
  METIS_SetDefaultOptions(options);
  options[17] = 0;  // METIS_OPTION_NUMBERING = C-style
  if (const char* e = getenv("B2_METIS_OPTS")) {   // developer knob: "index=value,index=value" (METIS 5.1 option indices)
    const char* q = e;
    while (*q) {
      char* end = nullptr;
      const long idx = strtol(q, &end, 10);
      if (end == q || *end != '=') break;
      q = end + 1;
      const long val = strtol(q, &end, 10);
      if (end == q) break;
      if (idx >= 0 && idx < 40) options[idx] = val;
      q = (*end == ',') ? end + 1 : end;
    }
  }
  int64_t nv = n;
  int rc = METIS_NodeND(&nv, xa.data(), ad.data(), nullptr, options, p.data(), ip.data());
  if (rc != 1) { err = "METIS_NodeND failed"; return false; }
  // METIS: A(perm, perm) is the reordered matrix, i.e. perm[k] = original index of pivot k
  for (int64_t k = 0; k < n; k++) perm[k] = (int32_t)p[k];
  return true;
}


// Nested dissection of the COMPRESSED graph of a KKT matrix [H J'; J -D] (D diagonal): the r / lambda
// vertices (index >= nvar) only touch x vertices, their pivots are -1 / -delta whatever the order
// (SURVEY App. B), and eliminating one of them first makes a clique of its x-neighbours -- which is
// exactly an edge set of the graph G' of H + J'J on the nvar x-vertices.  So: dissect G' (a third
// of the vertices of the raw graph for config 4, and separators counted in x-vertices only: the
// critical path of the factorization is the sum of the separator sizes), then order all r / lambda
// vertices first and the x-vertices in the dissection order; the postorder of the elimination tree
// hangs every r / lambda leaf under the first of its x-neighbours.
// Returns false (perm untouched) when the trailing block is not diagonal: the caller dissects the
// raw graph instead.
bool order_kkt_compressed_nd(int64_t n, int64_t nvar, const std::vector<int64_t>& xadj,
                             const std::vector<int64_t>& adj, std::vector<int32_t>& perm, std::string& err) {
  if (nvar <= 0 || nvar >= n) return false;
  for (int64_t v = nvar; v < n; v++)
    for (int64_t p = xadj[v]; p < xadj[v + 1]; p++)
      if (adj[p] >= nvar) return false;
  // adjacency of G' with a marker array: sum over rows of |N(v)|^2 steps, no sort
  std::vector<int64_t> gx(nvar + 1, 0), ga;
  std::vector<int64_t> mark(nvar, -1);
  {
    double work = 0;
    for (int64_t v = nvar; v < n; v++) { const double d = (double)(xadj[v + 1] - xadj[v]); work += d * d; }
    if (work > 4e9) return false;                // a dense constraint block: the clique expansion is not worth it
    ga.reserve((size_t)std::min(work + (double)xadj[nvar], 2e9));
  }
  for (int64_t a = 0; a < nvar; a++) {
    mark[a] = a;
    for (int64_t p = xadj[a]; p < xadj[a + 1]; p++) {
      const int64_t v = adj[p];
      if (v < nvar) {
        if (mark[v] != a) { mark[v] = a; ga.push_back(v); }
      } else {
        for (int64_t q = xadj[v]; q < xadj[v + 1]; q++) {
          const int64_t b = adj[q];
          if (mark[b] != a) { mark[b] = a; ga.push_back(b); }
        }
      }
    }
    gx[a + 1] = (int64_t)ga.size();
  }
  std::vector<int32_t> px;
  if (!order_metis_nd(nvar, gx, ga, px, err)) return false;
  perm.resize(n);
  int64_t k = 0;
  for (int64_t v = nvar; v < n; v++) perm[k++] = (int32_t)v;
  for (int64_t i = 0; i < nvar; i++) perm[k++] = px[i];
  return true;
}

// ------------------------------------------------------------------------------------------
// Approximate minimum degree (Amestoy, Davis, Duff): quotient-graph elimination with
// approximate external degrees, element absorption, mass elimination and hashed detection of
// indistinguishable variables.  Written for this backend (int32 indices, std::vector
// workspace); the result is then postordered by the caller together with the etree.
// ------------------------------------------------------------------------------------------
namespace {

class AmdOrdering {
 public:
  AmdOrdering(int64_t n, const std::vector<int64_t>& xadj, const std::vector<int64_t>& adj)
      : n_((int32_t)n) {
    const int64_t nz = (int64_t)adj.size();
    cap_ = nz + nz / 5 + 2 * n + 32;
    iw_.resize(cap_);
    pe_.resize(n + 1); len_.assign(n + 1, 0); elen_.assign(n + 1, 0); nv_.assign(n + 1, 1);
    deg_.assign(n + 1, 0); w_.assign(n + 1, 1); head_.assign(n + 1, -1); next_.assign(n + 1, -1);
    prev_.assign(n + 1, -1); hash_head_.assign(n + 1, -1);
    for (int64_t j = 0; j < n; j++) {
      pe_[j] = xadj[j];
      len_[j] = (int32_t)(xadj[j + 1] - xadj[j]);
      deg_[j] = len_[j];
      for (int64_t p = xadj[j]; p < xadj[j + 1]; p++) iw_[p] = (int32_t)adj[p];
    }
    free_ = nz;
  }

  void run(std::vector<int32_t>& order) {
    const int32_t n = n_;
    int32_t dense = (int32_t)(10.0 * std::sqrt((double)n));
    dense = std::max(16, dense);
    dense = std::min(n - 2, dense);
    pe_[n] = kDead; elen_[n] = -2; w_[n] = 0; len_[n] = 0; nv_[n] = 1;
    mark_ = 2;
    int32_t nel = 0;
    for (int32_t i = 0; i < n; i++) {
      const int32_t d = deg_[i];
      if (d == 0) { elen_[i] = -2; nel++; pe_[i] = kDead; w_[i] = 0; }
      else if (d > dense) { nv_[i] = 0; elen_[i] = -1; nel++; pe_[i] = flip(n); nv_[n]++; }
      else list_insert(i, d);
    }
    int32_t mindeg = 0, lemax = 0;
    while (nel < n) {
      int32_t k = -1;
      while (mindeg < n && (k = head_[mindeg]) == -1) mindeg++;
      list_remove_head(mindeg, k);
      const int32_t elenk = elen_[k];
      int32_t nvk = nv_[k];
      nel += nvk;
      if (elenk > 0 && free_ + mindeg >= cap_) compact();
      // ---- new element Lk
      int32_t dk = 0;
      nv_[k] = -nvk;
      int64_t p = pe_[k];
      const int64_t pk1 = (elenk == 0) ? p : free_;
      int64_t pk2 = pk1;
      for (int32_t t = 0; t <= elenk; t++) {
        int32_t e; int64_t pj; int32_t ln;
        if (t == elenk) { e = k; pj = p; ln = len_[k] - elenk; }
        else { e = iw_[p++]; pj = pe_[e]; ln = len_[e]; }
        for (int32_t q = 0; q < ln; q++) {
          const int32_t i = iw_[pj++];
          const int32_t nvi = nv_[i];
          if (nvi <= 0) continue;
          dk += nvi;
          nv_[i] = -nvi;
          iw_[pk2++] = i;
          list_unlink(i);
        }
        if (e != k) { pe_[e] = flip(k); w_[e] = 0; }
      }
      if (elenk != 0) free_ = pk2;
      deg_[k] = dk; pe_[k] = pk1; len_[k] = (int32_t)(pk2 - pk1); elen_[k] = -2;
      // ---- scan 1: external sizes |Le \ Lk|
      clear_marks(lemax);
      for (int64_t pk = pk1; pk < pk2; pk++) {
        const int32_t i = iw_[pk];
        const int32_t eln = elen_[i];
        if (eln <= 0) continue;
        const int32_t nvi = -nv_[i];
        const int64_t wnvi = mark_ - nvi;
        for (int64_t q = pe_[i]; q < pe_[i] + eln; q++) {
          const int32_t e = iw_[q];
          if (w_[e] >= mark_) w_[e] -= nvi;
          else if (w_[e] != 0) w_[e] = deg_[e] + wnvi;
        }
      }
      // ---- scan 2: degrees, absorption, hashing
      for (int64_t pk = pk1; pk < pk2; pk++) {
        const int32_t i = iw_[pk];
        const int64_t p1 = pe_[i], p2 = p1 + elen_[i] - 1;
        int64_t pn = p1;
        int64_t h = 0; int32_t d = 0;
        for (int64_t q = p1; q <= p2; q++) {
          const int32_t e = iw_[q];
          if (w_[e] == 0) continue;
          const int64_t dext = w_[e] - mark_;
          if (dext > 0) { d += (int32_t)dext; iw_[pn++] = e; h += e; }
          else { pe_[e] = flip(k); w_[e] = 0; }
        }
        elen_[i] = (int32_t)(pn - p1 + 1);
        const int64_t p3 = pn, p4 = p1 + len_[i];
        for (int64_t q = p2 + 1; q < p4; q++) {
          const int32_t j = iw_[q];
          const int32_t nvj = nv_[j];
          if (nvj <= 0) continue;
          d += nvj; iw_[pn++] = j; h += j;
        }
        if (d == 0) {
          pe_[i] = flip(k);
          const int32_t nvi = -nv_[i];
          dk -= nvi; nvk += nvi; nel += nvi; nv_[i] = 0; elen_[i] = -1;
        } else {
          deg_[i] = std::min(deg_[i], d);
          iw_[pn] = iw_[p3]; iw_[p3] = iw_[p1]; iw_[p1] = k;
          len_[i] = (int32_t)(pn - p1 + 1);
          const int32_t hb = (int32_t)(h % n);
          next_[i] = hash_head_[hb]; hash_head_[hb] = i; prev_[i] = hb;
        }
      }
      deg_[k] = dk;
      lemax = std::max(lemax, dk);
      mark_ += lemax;
      clear_marks(lemax);
      // ---- indistinguishable variables
      for (int64_t pk = pk1; pk < pk2; pk++) {
        int32_t i = iw_[pk];
        if (nv_[i] >= 0) continue;
        const int32_t hb = prev_[i];
        i = hash_head_[hb];
        hash_head_[hb] = -1;
        for (; i != -1 && next_[i] != -1; i = next_[i], mark_++) {
          const int32_t ln = len_[i], eln = elen_[i];
          for (int64_t q = pe_[i] + 1; q < pe_[i] + ln; q++) w_[iw_[q]] = mark_;
          int32_t jlast = i;
          for (int32_t j = next_[i]; j != -1;) {
            bool same = len_[j] == ln && elen_[j] == eln;
            for (int64_t q = pe_[j] + 1; same && q < pe_[j] + ln; q++)
              if (w_[iw_[q]] != mark_) same = false;
            if (same) {
              pe_[j] = flip(i); nv_[i] += nv_[j]; nv_[j] = 0; elen_[j] = -1;
              j = next_[j]; next_[jlast] = j;
            } else { jlast = j; j = next_[j]; }
          }
        }
      }
      // ---- finalise Lk
      int64_t pf = pk1;
      for (int64_t pk = pk1; pk < pk2; pk++) {
        const int32_t i = iw_[pk];
        const int32_t nvi = -nv_[i];
        if (nvi <= 0) continue;
        nv_[i] = nvi;
        int32_t d = deg_[i] + dk - nvi;
        d = std::min(d, n - nel - nvi);
        list_insert(i, d);
        mindeg = std::min(mindeg, d);
        deg_[i] = d;
        iw_[pf++] = i;
      }
      nv_[k] = nvk;
      if ((len_[k] = (int32_t)(pf - pk1)) == 0) { pe_[k] = kDead; w_[k] = 0; }
      if (elenk != 0) free_ = pf;
    }
    // ---- elimination order: pivots in the order they were chosen is implicit in the tree
    // of absorptions; emit a postorder of that tree (children before parents).
    std::vector<int32_t> par(n + 1), kid(n + 1, -1), sib(n + 1, -1);
    for (int32_t i = 0; i <= n; i++) par[i] = (pe_[i] == kDead) ? -1 : (int32_t)flip(pe_[i]);
    for (int32_t j = n; j >= 0; j--) {          // absorbed variables first ...
      if (nv_[j] > 0 || par[j] < 0) continue;
      sib[j] = kid[par[j]]; kid[par[j]] = j;
    }
    for (int32_t e = n; e >= 0; e--) {          // ... then elements, so elements are visited first
      if (nv_[e] <= 0 || par[e] < 0) continue;
      sib[e] = kid[par[e]]; kid[par[e]] = e;
    }
    order.clear(); order.reserve(n);
    std::vector<int32_t> stack;
    for (int32_t root = 0; root <= n; root++) {
      if (par[root] != -1) continue;
      stack.push_back(root);
      while (!stack.empty()) {
        const int32_t v = stack.back();
        const int32_t c = kid[v];
        if (c == -1) { stack.pop_back(); if (v != n) order.push_back(v); }
        else { kid[v] = sib[c]; stack.push_back(c); }
      }
    }
  }

 private:
  static constexpr int64_t kDead = -1;
  static int64_t flip(int64_t i) { return -i - 2; }
  void list_insert(int32_t i, int32_t d) {
    if (head_[d] != -1) prev_[head_[d]] = i;
    next_[i] = head_[d]; prev_[i] = -1; head_[d] = i;
  }
  void list_remove_head(int32_t d, int32_t k) {
    if (next_[k] != -1) prev_[next_[k]] = -1;
    head_[d] = next_[k];
  }
  void list_unlink(int32_t i) {
    if (next_[i] != -1) prev_[next_[i]] = prev_[i];
    if (prev_[i] != -1) next_[prev_[i]] = next_[i];
    else head_[deg_[i]] = next_[i];
  }
  void clear_marks(int32_t lemax) {
    if (mark_ < 2 || mark_ + lemax > (int64_t)1 << 60) {
      for (int32_t k = 0; k < n_; k++) if (w_[k] != 0) w_[k] = 1;
      mark_ = 2;
    }
  }
  void compact() {
    for (int32_t j = 0; j < n_; j++) {
      const int64_t p = pe_[j];
      if (p >= 0) { pe_[j] = iw_[p]; iw_[p] = (int32_t)flip(j); }
    }
    int64_t q = 0;
    for (int64_t p = 0; p < free_;) {
      const int32_t j = (int32_t)flip(iw_[p++]);
      if (j >= 0) {
        iw_[q] = (int32_t)pe_[j];
        pe_[j] = q++;
        for (int32_t t = 0; t < len_[j] - 1; t++) iw_[q++] = iw_[p++];
      }
    }
    free_ = q;
  }
  int32_t n_;
  int64_t cap_, free_ = 0, mark_ = 2;
  std::vector<int32_t> iw_, len_, elen_, nv_, deg_, head_, next_, prev_, hash_head_;
  std::vector<int64_t> pe_, w_;
};

}  // namespace

bool order_amd(int64_t n, const std::vector<int64_t>& xadj, const std::vector<int64_t>& adj,
               std::vector<int32_t>& perm, std::string& err) {
  perm.resize(n);
  if (n <= 2 || adj.empty()) {
    for (int64_t i = 0; i < n; i++) perm[i] = (int32_t)i;
    return true;
  }
  AmdOrdering amd(n, xadj, adj);
  std::vector<int32_t> order;
  amd.run(order);
  if ((int64_t)order.size() != n) { err = "internal: AMD produced an incomplete ordering"; return false; }
  perm = order;
  return true;
}

}  // namespace b2
