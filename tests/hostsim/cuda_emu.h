// cuda_emu.h -- TEST INFRASTRUCTURE ONLY.  A minimal CPU emulation of the CUDA execution model
// (grid of blocks, threads as ucontext fibers, __syncthreads / __syncwarp barriers, warp
// shuffles, atomics, a malloc-backed runtime API) so that the product's kernel sources
// (cannoles_b200/csrc/*.cu) can be compiled with g++ -DB2_EMULATE and exercised on a box
// without a GPU (`pytest -m "not gpu"`).  It is never linked into libcannoles_b200.so; the
// product path has no CPU fallback.  Blocks run one after another on one OS thread.
#pragma once
#include <ucontext.h>

#include <cassert>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <vector>
#include <algorithm>
using std::min;
using std::max;
using std::isfinite;
using std::isnan;

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __noinline__ inline
#define __restrict__
#define __launch_bounds__(...)
#define __shared__ static
#define __align__(x)

struct uint3 { unsigned x, y, z; };
struct int2 { int x, y; };
struct int4 { int x, y, z, w; };
struct dim3 {
  unsigned x, y, z;
  dim3(unsigned a = 1, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {}
};

namespace emu {

struct Fiber {
  ucontext_t ctx;
  char* stack = nullptr;
  int state = 0;  // 0 runnable, 1 at block barrier, 2 at warp barrier, 3 done
};

struct State {
  ucontext_t sched;
  std::vector<Fiber> fibers;
  int cur = -1;
  int nthreads = 0;
  std::function<void()> body;
  unsigned char* dyn_smem = nullptr;
  size_t dyn_smem_cap = 0;
  unsigned long long xbuf[32 * 64];  // warp exchange buffers (per warp: 32 x u64), <= 64 warps
};

inline State& S() { static State s; return s; }

extern thread_local uint3 g_threadIdx, g_blockIdx;
extern thread_local dim3 g_blockDim, g_gridDim;

inline void set_tid(int t) {
  dim3 b = g_blockDim;
  g_threadIdx.x = t % b.x;
  g_threadIdx.y = (t / b.x) % b.y;
  g_threadIdx.z = t / (b.x * b.y);
}

inline void yield_to_sched() {
  State& s = S();
  int me = s.cur;
  swapcontext(&s.fibers[me].ctx, &s.sched);
  set_tid(me);
}

inline void fiber_entry() {
  State& s = S();
  s.body();
  s.fibers[s.cur].state = 3;
  swapcontext(&s.fibers[s.cur].ctx, &s.sched);
}

inline void run_block(int nthreads, const std::function<void()>& body) {
  State& s = S();
  s.nthreads = nthreads;
  s.body = body;
  if ((int)s.fibers.size() < nthreads) {
    size_t old = s.fibers.size();
    s.fibers.resize(nthreads);
    for (size_t i = old; i < s.fibers.size(); i++) s.fibers[i].stack = (char*)malloc(1 << 16);
  }
  for (int t = 0; t < nthreads; t++) {
    Fiber& f = s.fibers[t];
    getcontext(&f.ctx);
    f.ctx.uc_stack.ss_sp = f.stack;
    f.ctx.uc_stack.ss_size = 1 << 16;
    f.ctx.uc_link = nullptr;
    makecontext(&f.ctx, (void (*)())fiber_entry, 0);
    f.state = 0;
  }
  int done = 0;
  while (done < nthreads) {
    bool progressed = false;
    for (int t = 0; t < nthreads; t++) {
      Fiber& f = s.fibers[t];
      if (f.state != 0) continue;
      s.cur = t;
      set_tid(t);
      swapcontext(&s.sched, &f.ctx);
      progressed = true;
      if (f.state == 3) done++;
    }
    // release block barrier when every live fiber waits on it
    int nb = 0, live = 0;
    for (int t = 0; t < nthreads; t++) {
      if (s.fibers[t].state != 3) live++;
      if (s.fibers[t].state == 1) nb++;
    }
    if (live > 0 && nb == live) {
      for (int t = 0; t < nthreads; t++)
        if (s.fibers[t].state == 1) s.fibers[t].state = 0;
      progressed = true;
    }
    // release warp barriers: all live lanes of a warp wait on it
    for (int w0 = 0; w0 < nthreads; w0 += 32) {
      int nw = 0, lw = 0;
      for (int t = w0; t < w0 + 32 && t < nthreads; t++) {
        if (s.fibers[t].state != 3) lw++;
        if (s.fibers[t].state == 2) nw++;
      }
      if (lw > 0 && nw == lw) {
        for (int t = w0; t < w0 + 32 && t < nthreads; t++)
          if (s.fibers[t].state == 2) s.fibers[t].state = 0;
        progressed = true;
      }
    }
    if (!progressed && done < nthreads) {
      fprintf(stderr, "cuda_emu: deadlock (divergent barrier) in block (%u,%u)\n", g_blockIdx.x,
              g_blockIdx.y);
      abort();
    }
  }
}

inline void block_barrier() {
  State& s = S();
  s.fibers[s.cur].state = 1;
  yield_to_sched();
}
inline void warp_barrier() {
  State& s = S();
  s.fibers[s.cur].state = 2;
  yield_to_sched();
}
inline int lane_id() { return S().cur & 31; }
inline int warp_id() { return S().cur >> 5; }

template <typename T>
inline T shfl_idx(T v, int src) {
  static_assert(sizeof(T) <= 8, "shfl of >8 bytes");
  State& s = S();
  unsigned long long* buf = &s.xbuf[warp_id() * 32];
  unsigned long long raw = 0;
  memcpy(&raw, &v, sizeof(T));
  buf[lane_id()] = raw;
  warp_barrier();
  unsigned long long got = buf[src & 31];
  warp_barrier();
  T out;
  memcpy(&out, &got, sizeof(T));
  return out;
}

template <typename F, typename... Args>
inline void launch(F kernel, dim3 grid, dim3 block, size_t smem, Args... args) {
  State& s = S();
  if (smem > s.dyn_smem_cap) {
    free(s.dyn_smem);
    s.dyn_smem = (unsigned char*)aligned_alloc(128, (smem + 127) / 128 * 128);
    s.dyn_smem_cap = smem;
  }
  g_gridDim = grid;
  g_blockDim = block;
  int nthreads = block.x * block.y * block.z;
  for (unsigned bz = 0; bz < grid.z; bz++)
    for (unsigned by = 0; by < grid.y; by++)
      for (unsigned bx = 0; bx < grid.x; bx++) {
        g_blockIdx = {bx, by, bz};
        run_block(nthreads, [&]() { kernel(args...); });
      }
}

}  // namespace emu

#define threadIdx (emu::g_threadIdx)
#define blockIdx (emu::g_blockIdx)
#define blockDim (emu::g_blockDim)
#define gridDim (emu::g_gridDim)

inline void __syncthreads() { emu::block_barrier(); }
inline void __syncwarp(unsigned = 0xffffffffu) { emu::warp_barrier(); }
inline int __syncthreads_and(int pred) {
  static int acc;
  emu::block_barrier();
  if (emu::S().cur == 0) acc = 1;
  emu::block_barrier();
  if (!pred) acc = 0;
  emu::block_barrier();
  return acc;
}
inline int __syncthreads_or(int pred) {
  static int acc;
  emu::block_barrier();
  if (emu::S().cur == 0) acc = 0;
  emu::block_barrier();
  if (pred) acc = 1;
  emu::block_barrier();
  return acc;
}
inline void __nanosleep(unsigned) {}
inline void __threadfence() {}
inline void __threadfence_block() {}
template <typename T> inline T __ldg(const T* p) { return *p; }
template <typename T> inline T __shfl_sync(unsigned, T v, int src, int = 32) { return emu::shfl_idx(v, src); }
template <typename T> inline T __shfl_xor_sync(unsigned, T v, int m, int = 32) {
  return emu::shfl_idx(v, emu::lane_id() ^ m);
}
template <typename T> inline T __shfl_down_sync(unsigned, T v, unsigned d, int = 32) {
  int src = emu::lane_id() + (int)d;
  T o = emu::shfl_idx(v, src > 31 ? emu::lane_id() : src);
  return o;
}
template <typename T> inline T __shfl_up_sync(unsigned, T v, unsigned d, int = 32) {
  int src = emu::lane_id() - (int)d;
  T o = emu::shfl_idx(v, src < 0 ? emu::lane_id() : src);
  return o;
}
inline unsigned __ballot_sync(unsigned, int pred) {
  unsigned r = 0;
  for (int l = 0; l < 32; l++) {
    int p = emu::shfl_idx(pred, l);
    if (p) r |= (1u << l);
  }
  return r;
}
inline int __ffs(unsigned v) { return v ? __builtin_ctz(v) + 1 : 0; }
inline int __any_sync(unsigned m, int pred) { return __ballot_sync(m, pred) != 0; }
template <typename T> inline T atomicAdd(T* p, T v) { T o = *p; *p = o + v; return o; }
inline int atomicMax(int* p, int v) { int o = *p; if (v > o) *p = v; return o; }
inline int atomicOr(int* p, int v) { int o = *p; *p = o | v; return o; }
inline int atomicExch(int* p, int v) { int o = *p; *p = v; return o; }
inline long long __double_as_longlong(double a) { long long o; memcpy(&o, &a, 8); return o; }
inline double __longlong_as_double(long long a) { double o; memcpy(&o, &a, 8); return o; }
inline double __drcp_rn(double a) { return 1.0 / a; }
inline double __fma_rn(double a, double b, double c) { return std::fma(a, b, c); }
inline double __dmul_rn(double a, double b) { return a * b; }
inline double __ddiv_rn(double a, double b) { return a / b; }
inline double __dsub_rn(double a, double b) { return a - b; }
inline double __dadd_rn(double a, double b) { return a + b; }

// ---------------------------------------------------------------- runtime API shim
typedef int cudaError_t;
typedef void* cudaStream_t;
typedef void* cudaEvent_t;
typedef void* cudaGraph_t;
typedef void* cudaGraphExec_t;
enum { cudaSuccess = 0 };
enum cudaMemcpyKind { cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice };
enum { cudaStreamCaptureModeGlobal = 0, cudaStreamNonBlocking = 1, cudaHostRegisterDefault = 0,
       cudaEventDisableTiming = 2, cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };
struct cudaDeviceProp { int multiProcessorCount; size_t sharedMemPerBlockOptin; char name[64]; };
inline const char* cudaGetErrorString(cudaError_t) { return "emu"; }
inline cudaError_t cudaGetLastError() { return 0; }
inline cudaError_t cudaMalloc(void** p, size_t n) { *p = calloc(n ? n : 1, 1); return *p ? 0 : 2; }
inline cudaError_t cudaFree(void* p) { free(p); return 0; }
inline cudaError_t cudaMallocHost(void** p, size_t n) { *p = calloc(n ? n : 1, 1); return 0; }
inline cudaError_t cudaFreeHost(void* p) { free(p); return 0; }
inline cudaError_t cudaHostRegister(void*, size_t, unsigned) { return 0; }
inline cudaError_t cudaHostUnregister(void*) { return 0; }
inline cudaError_t cudaMemcpy(void* d, const void* s, size_t n, cudaMemcpyKind) { memcpy(d, s, n); return 0; }
inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t = 0) { memcpy(d, s, n); return 0; }
inline cudaError_t cudaMemset(void* d, int v, size_t n) { memset(d, v, n); return 0; }
inline cudaError_t cudaMemsetAsync(void* d, int v, size_t n, cudaStream_t = 0) { memset(d, v, n); return 0; }
inline cudaError_t cudaStreamCreate(cudaStream_t* s) { *s = nullptr; return 0; }
inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, unsigned) { *s = nullptr; return 0; }
inline cudaError_t cudaStreamDestroy(cudaStream_t) { return 0; }
inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return 0; }
inline cudaError_t cudaDeviceSynchronize() { return 0; }
inline cudaError_t cudaEventCreate(cudaEvent_t* e) { *e = nullptr; return 0; }
inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t* e, unsigned) { *e = nullptr; return 0; }
inline cudaError_t cudaEventDestroy(cudaEvent_t) { return 0; }
inline cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t = 0) { return 0; }
inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return 0; }
inline cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t, cudaEvent_t) { *ms = 0.f; return 0; }
inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned = 0) { return 0; }
inline cudaError_t cudaGetDevice(int* d) { *d = 0; return 0; }
inline cudaError_t cudaSetDevice(int) { return 0; }
inline cudaError_t cudaGetDeviceCount(int* n) { *n = 1; return 0; }
inline cudaError_t cudaGetDeviceProperties(cudaDeviceProp* p, int) {
  p->multiProcessorCount = 4; p->sharedMemPerBlockOptin = 227 * 1024; strcpy(p->name, "cuda_emu");
  return 0;
}
template <typename F> inline cudaError_t cudaFuncSetAttribute(F, int, int) { return 0; }
