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
  int* flags;  // [0] = breakdown (exact zero pivot seen)
};

constexpr int NB = 32;      // pivot block width of the tiled path
constexpr int TILE = 64;    // update tile (TILE x TILE per CTA)
constexpr int TRSM_ROWS = 128;
constexpr int ASM_COLS = 8; // destination columns per CTA in k_assemble_large

}  // namespace b2
