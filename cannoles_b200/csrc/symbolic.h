// symbolic.h -- host-side symbolic analysis for the B200 KKT LDL^T backend.
//
// Replaces what `LDLFactStruct(N, rows, cols, vals)` does once per solver
// (reference/src/solver_types.jl:61-65: `sparse(cols, rows, vals)`, `triu`, `ldl_analyze`):
// merge the COO lower triangle into a fixed upper CSC pattern, pick a fill-reducing ordering,
// build the elimination tree, and lay out everything the device kernels need -- here a
// supernodal multifrontal plan (assembly tree, per-front row lists, extend-add maps, the
// COO->CSC and CSC->front scatter maps, level sets).  Pure C++17, no CUDA, so that the same
// code can be exercised on a CPU-only box by tests/hostsim.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace b2 {

enum Ordering : int {
  ORDER_ND = 0,       // nested dissection (METIS NodeND from the CUDA toolkit's libmetis_static.a)
  ORDER_NATURAL = 1,  // identity
  ORDER_USER = 2,     // caller-supplied permutation (perm[k] = 0-based original index of pivot k)
  ORDER_AMD = 3,      // approximate minimum degree (own implementation, ordering.cpp)
  ORDER_ND_RAW = 4,   // nested dissection of the raw N-vertex graph (ORDER_ND dissects the compressed
                      // x-vertex graph of a KKT matrix when the trailing block is diagonal)
};

struct SymbolicOptions {
  int ordering = ORDER_ND;
  const int64_t* user_perm = nullptr;
  // relaxed supernode amalgamation (merge last child into parent when few explicit zeros appear)
  int relax_always = 8;        // merged width <= this: always merge
  int relax_w1 = 32;  double relax_z1 = 0.50;
  int relax_w2 = 96;  double relax_z2 = 0.15;
  double relax_z3 = 0.05;      // any width
  bool build_spmv = true;      // full symmetric CSR for residual / refinement
};

struct Symbolic {
  int64_t N = 0, nnz = 0, nnzA = 0;
  int64_t nvar = 0, nequ = 0, ncon = 0;

  // ---- A = triu(sparse(cols, rows, vals)): CSC, column j holds rows i <= j (original indices)
  std::vector<int64_t> Ap;       // N+1
  std::vector<int32_t> Ai;       // nnzA
  // ---- COO -> CSC accumulate plan: slot s sums vals[coo_sorted[slot_ptr[s] .. slot_ptr[s+1])]
  //      in increasing COO index (the order set_vals! adds them, src/solver_types.jl:53-59)
  std::vector<int64_t> slot_ptr;    // nnzA+1
  std::vector<int32_t> coo_sorted;  // nnz
  std::vector<int32_t> coo_slot;    // nnz: COO entry -> slot
  // diagonal bookkeeping for device-side rho / delta shifts (segments S6/S7)
  std::vector<int32_t> rho_slot;    // nvar: slot of (i,i), i < nvar
  std::vector<int32_t> delta_slot;  // ncon: slot of the (nvar+nequ+j) diagonal
  bool shift_ok = false;            // trailing COO entries are the canonical -delta / rho segments

  // ---- ordering: perm[k] = original index of pivot k; pinv = inverse
  std::vector<int32_t> perm, pinv;
  std::vector<int32_t> parent;      // elimination tree on permuted indices (-1 = root)
  std::vector<int32_t> colcount;    // |L(:,k)| including the diagonal (exact, before relaxation)

  // ---- supernodes / fronts (permuted index space)
  int32_t nsuper = 0;
  std::vector<int32_t> scol;        // nsuper+1: first pivot column of each supernode
  std::vector<int32_t> sparent;     // nsuper (-1 = root)
  std::vector<int32_t> col2sn;      // N
  std::vector<int64_t> rptr;        // nsuper+1 into rowidx / rel
  std::vector<int32_t> rowidx;      // front row list: w pivots then r sorted rows below
  std::vector<int32_t> rel;         // aligned with rowidx: position of that row in the parent front
  std::vector<int64_t> lptr;        // nsuper+1: panel offset in Lx (m x w, column-major, lda = m)
  std::vector<int64_t> cbptr;       // nsuper+1: contribution block offset (r x r, column-major)
  std::vector<int64_t> uptr;        // nsuper+1: solve-phase update vector offset (r)
  std::vector<int32_t> child_ptr, child_idx;  // children lists (ascending supernode index)
  // CSC slot -> front scatter, grouped by supernode
  std::vector<int64_t> amap_ptr;    // nsuper+1
  std::vector<int32_t> amap_slot;   // nnzA
  std::vector<int32_t> amap_pos;    // nnzA: row + col*m inside the panel
  // level sets of the assembly tree (level 0 = leaves)
  int32_t nlevels = 0;
  std::vector<int32_t> level_ptr;   // nlevels+1
  std::vector<int32_t> level_sn;    // nsuper, grouped by level
  std::vector<int32_t> slevel;      // nsuper

  // ---- full symmetric CSR of A for y = K x (residual / refinement), values by slot
  std::vector<int64_t> Sp;          // N+1
  std::vector<int32_t> Sj;          // 2 nnzA - ndiag
  std::vector<int32_t> Sslot;

  // ---- statistics
  int64_t nnzL = 0;          // exact off-diagonal count of L (no padding)
  double flops = 0;          // sum_j (c_j^2 + 3 c_j), exact symbolic
  int64_t nnzL_store = 0;    // doubles in panel storage (with relaxation zeros + upper diag blocks)
  int64_t cb_store = 0;      // doubles in contribution-block storage
  double flops_store = 0;    // flops actually executed on the dense fronts
  int32_t max_front = 0, max_width = 0;
  double t_order = 0, t_symbolic = 0;

  std::string error;         // non-empty => analysis failed
};

// rows1/cols1: 1-based COO of the lower triangle (rows >= cols), duplicates allowed.
// Returns false (and sets sym.error) on malformed input.
bool analyze(int64_t N, int64_t nnz, const int64_t* rows1, const int64_t* cols1, int64_t nvar,
             int64_t nequ, int64_t ncon, const SymbolicOptions& opt, Symbolic& sym);

// orderings (ordering.cpp). adjacency = full symmetric pattern without diagonal.
bool order_metis_nd(int64_t n, const std::vector<int64_t>& xadj, const std::vector<int64_t>& adj,
                    std::vector<int32_t>& perm, std::string& err);
bool order_kkt_compressed_nd(int64_t n, int64_t nvar, const std::vector<int64_t>& xadj,
                             const std::vector<int64_t>& adj, std::vector<int32_t>& perm, std::string& err);
bool order_amd(int64_t n, const std::vector<int64_t>& xadj, const std::vector<int64_t>& adj,
               std::vector<int32_t>& perm, std::string& err);

}  // namespace b2
