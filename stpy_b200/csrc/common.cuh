// Shared device helpers for the stpy_b200 kernels (sm_100a only).
//
// FP64 tensor-core math on sm_100a is the warp-level DMMA.8x8x4 instruction
// (PTX mma.sync.m8n8k4.f64); tcgen05/wgmma have no f64 kind.  Operands are
// staged global -> shared with cp.async (LDGSTS) in a "k4-packed" layout
// [k/4][row][4] so that every DMMA fragment load is one contiguous 256-byte,
// bank-conflict-free shared-memory read (fragment element of lane l is word l).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace stpyb {

typedef long long i64;

#define STPYB_OK 0
#define STPYB_ERR_CUDA 1000   // + cudaError_t

#define STPYB_CUDA(expr)                                     \
  do {                                                       \
    cudaError_t _e = (expr);                                 \
    if (_e != cudaSuccess) return STPYB_ERR_CUDA + (int)_e;  \
  } while (0)

#define STPYB_TRY(expr)          \
  do {                           \
    int _r = (expr);             \
    if (_r != 0) return _r;      \
  } while (0)

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
      : "+d"(c0), "+d"(c1)
      : "d"(a), "d"(b));
}

// 16-byte async copy global->shared; bytes beyond src_bytes are zero-filled.
__device__ __forceinline__ void cp_async16(void* smem, const void* gptr, int src_bytes) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(s), "l"(gptr), "r"(src_bytes)
               : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Block-wide sum; result valid in thread 0 (and broadcast through smem to all).
template <int THREADS>
__device__ __forceinline__ double block_sum(double v, double* red /* >= THREADS/32 doubles */) {
  v = warp_sum(v);
  int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  if (l == 0) red[w] = v;
  __syncthreads();
  double t = 0.0;
  if (threadIdx.x < 32) {
    t = (l < THREADS / 32) ? red[l] : 0.0;
    t = warp_sum(t);
    if (l == 0) red[0] = t;
  }
  __syncthreads();
  t = red[0];
  __syncthreads();
  return t;
}

// ---- instrumentation (api.cu): launch counter and per-category CUDA-event timing -------------
enum { PROF_DIAG = 0, PROF_TRSM = 1, PROF_PANEL_UPD = 2, PROF_SYRK = 3, PROF_GRAM = 4, PROF_OTHER = 5, PROF_SOLVE = 6, PROF_NCAT = 7 };
extern long long g_launches;
extern int g_prof_on;
void prof_begin(int cat, double flops, cudaStream_t st);
void prof_end(cudaStream_t st);
#define STPYB_COUNT_LAUNCH() (++::stpyb::g_launches)

static inline int ceil_div(i64 a, i64 b) { return (int)((a + b - 1) / b); }

// Function attributes (opt-in dynamic shared memory) are per DEVICE: a process that drives several
// GPUs must set them once on each.  Returns true the first time it is called for the current device.
static inline bool first_use_on_device(bool (&seen)[64]) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return true;
  if (seen[dev]) return false;
  seen[dev] = true;
  return true;
}

}  // namespace stpyb
