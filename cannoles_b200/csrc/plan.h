// plan.h -- device-side view of the multifrontal plan (raw pointers passed to the kernels by value)
#pragma once
#include <cstdint>

namespace b2 {

struct PlanDev {
  const int32_t* scol;
  const int64_t* rptr;
  const int64_t* lptr;
  const int64_t* cbptr;
  const int64_t* uptr;
  const int32_t* rowidx;
  const int32_t* rel;
  const int32_t* child_ptr;
  const int32_t* child_idx;
  const int64_t* amap_ptr;
  const int32_t* amap_slot;
  const int32_t* amap_pos;
  const double* nzval;
  double* Lx;
  double* CB;
  double* dvec;
  double* dstage;         // factored diagonal blocks of the tiled fronts (NB x NB each)
  const int64_t* dsptr;   // per front: offset into dstage (tiled fronts only)
  const int64_t* asm_cptr; // per destination column of the tiled fronts: range in asm_ent
  const int32_t* asm_ent;  // pairs (child descriptor, child column), in child order
  const int32_t* asm_rc;   // per child descriptor: rows below the child's pivots (order of its contribution block)
  const int64_t* asm_off;  // per child descriptor: offset of the child's rel[] and of its contribution block
  const int64_t* sb_ptr;   // big-front solve: per (chunk, row) CSR pointers into sb_src (SB + 1 per chunk)
  const int32_t* sb_src;   // offsets into the update-vector storage, in child order
  const int32_t* sb_flag;  // per front: offset of its block flags (big fronts only)
  const int32_t* linv_idx; // per front: first slot in Linv of the explicit inverses of its unit-lower diagonal blocks, -1: none
  double* Linv;            // SB x SB per slot, column-major, full lower triangle (unit diagonal), zeros above
  const int32_t* ug_ptr;   // forward solve of the other fronts: per front row (rptr[s] + i) the range in ug_src
  const int32_t* ug_src;   //   of the child update entries (offsets into upd) that land on it, in child order
  const uint8_t* ug_row;   //   and the row of the front each entry lands on (fronts of order <= 255: the thread-per-front kernels)
  int* flags;  // [0] = breakdown (exact zero pivot seen)
  // cross-level dataflow factorization (k_front_dag): per-front records (4 ints each, see the
  // kernel) and the counters (zeroed with the tile flags before every factorization):
  // contribution-block tiles done per front at dcnt_cb, children done per front at dcnt_ch
  const int32_t* dfr;
  int* dcnt;
  int dcnt_cb, dcnt_ch;
  const int32_t* tl_ptr;   // per tile of the dataflow fronts (first list + J * nrb + I): range in tl_ent
  const int32_t* tl_ent;   // 72 ints: order of the child's contribution block, offset of the block (lo, hi), 5 x -, then 64 + 64 uint16: child row / column landing on each row / column of the tile (0xffff: none)
  const int64_t* sf_ptr;   // fronts of the shared-memory path: range in sf_ent of the flat extend-add list (nullptr: per-child loop)
  const int32_t* sf_ent;   // 2 ints: position in the front (15 bits) | A flag (bit 15) | run length << 16 (first of a run), source offset
  const int32_t* fl_ptr;   // per tile: range in fl_ent of the flat extend-add list (A entries + small children)
  const int32_t* fl_ent;   // 2 ints: destination in the tile (13 bits) | A flag (bit 13) | run length << 14 (first of a run), source offset (nzval slot / CB)
};

constexpr int NB = 64;        // pivot block width of the tiled path (== TILE: tile (0,0) is the next diagonal block)
constexpr int TILE = 64;      // update tile (TILE x TILE per CTA)
constexpr int UPD_KC = 16;    // K-chunk of k_update's shared-memory pipeline
constexpr int DIAG_LD = 65;   // leading dimension of a diagonal block in shared memory
constexpr int TRSM_THREADS = 128;
#ifndef B2_TRSM_RPT
#define B2_TRSM_RPT 1
#endif
constexpr int TRSM_RPT = B2_TRSM_RPT;                 // rows of the panel per k_trsm thread
constexpr int TRSM_ROWS = TRSM_THREADS * TRSM_RPT;    // ... and per CTA
constexpr int ASM_TILE = 2048; // doubles in the destination tile of k_assemble_large (8 x 256 or 16 x 128)
// threads (= fronts) per CTA of the one-thread-per-front kernels for fronts of order <= mm
#ifdef __CUDACC__
__host__ __device__
#endif
constexpr int tiny_nt(int mm) { return mm <= 4 ? 128 : 64; }
constexpr int FPB32 = 4;      // fronts per CTA in the 32-thread class of the per-front kernels (one warp each)
constexpr int SNB = 32;       // column block of the one-CTA-per-front solve kernels
constexpr int SB = 64;        // row chunk / column block of the multi-CTA solve of big fronts

}  // namespace b2
