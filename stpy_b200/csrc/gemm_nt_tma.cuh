// TMA-staged variant of the NT contraction: the operand tiles are brought into shared memory by
// cp.async.bulk.tensor (one elected thread issues the copies, completion is tracked with an
// mbarrier transaction count) instead of per-thread cp.async.  A K slice of a tile is stored as
// sub-tiles [ROWS][16 doubles] (128-byte rows = one box row, so the copy engine moves full lines)
// with the SWIZZLE_128B_ATOM_32B pattern: the 32-byte chunk c of row r lands at chunk c ^ (r & 3)
// (probed on the device: tools/tma_swizzle_probe.cu).  A DMMA fragment load -- lane (g, t) reads
// row g, columns 4*ks + t -- then touches chunks ks ^ (g & 3): the 16 lanes of a half-warp cover
// 16 distinct 8-byte words of one 128-byte line, i.e. it stays bank-conflict-free.  Ragged M / N /
// K edges need no predicates: the tensor map carries the true extents and TMA zero-fills.
#pragma once
#include <cuda.h>
#include "gemm_nt.cuh"

namespace stpyb {

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(unsigned long long* bar, unsigned parity) {
  unsigned ok;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// bounded wait: a copy that never completes (bad descriptor) must abort the kernel, not hang the GPU
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 2000000000LL) __trap();
  }
}
// 2-D tiled load: box origin (c0 = column, c1 = row) of the tensor described by `map`
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, int c0, int c1, unsigned long long* bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
          smem_u32(dst)),
      "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}

template <class Cfg, class Epi>
__global__ void __launch_bounds__(Cfg::THREADS, Cfg::MINB)
gemm_nt_tma_kernel(GemmArgs g, Epi epi, const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB) {
  extern __shared__ __align__(1024) double smem[];
  __shared__ __align__(8) unsigned long long full_bar[Cfg::STAGES];
  int tm, tn;
  decode_tile<Cfg>(g, (i64)blockIdx.x, tm, tn);
  const int m0 = tm * Cfg::BM, n0 = tn * Cfg::BN;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int wm = warp / Cfg::WARPS_N, wn = warp % Cfg::WARPS_N;
  const int gr = lane >> 2, tc = (lane & 3) * 2;
  const int row_base = m0 + wm * Cfg::WM + gr, col_base = n0 + wn * Cfg::WN + tc;
  const int kbeg = g.kskip ? m0 : 0;
  const int Keff = g.K - kbeg;
  const int KT = (Keff + Cfg::BK - 1) / Cfg::BK;
  constexpr unsigned STAGE_BYTES = (unsigned)(Cfg::STAGE * sizeof(double));

  if (threadIdx.x == 0) {
#pragma unroll
    for (int s = 0; s < Cfg::STAGES; ++s) mbar_init(&full_bar[s], 1);
    mbar_fence_init();
  }
  __syncthreads();

  constexpr int SUBS = Cfg::BK / 16;  // 16-column sub-tiles per K slice
  auto issue = [&](int slot, int kt) {  // elected thread only
    double* st = smem + slot * Cfg::STAGE;
    const int k0 = kbeg + kt * Cfg::BK;
    mbar_expect_tx(&full_bar[slot], STAGE_BYTES);
#pragma unroll
    for (int u = 0; u < SUBS; ++u) tma_load_2d(st + u * Cfg::BM * 16, &mapA, k0 + 16 * u, m0, &full_bar[slot]);
#pragma unroll
    for (int u = 0; u < SUBS; ++u)
      tma_load_2d(st + Cfg::A_STAGE + u * Cfg::BN * 16, &mapB, k0 + 16 * u, n0, &full_bar[slot]);
  };

  if (threadIdx.x == 0) {
#pragma unroll
    for (int s = 0; s < Cfg::STAGES - 1; ++s)
      if (s < KT) issue(s, s);
  }

  double acc[Cfg::MI][Cfg::NI][2];
#pragma unroll
  for (int i = 0; i < Cfg::MI; ++i) {
#pragma unroll
    for (int j = 0; j < Cfg::NI; ++j) {
      acc[i][j][0] = acc[i][j][1] = 0.0;
      if (Epi::kPreload) {
        const int row = row_base + i * 8, col = col_base + j * 8;
        if (row < g.M && col < g.N) epi.preload(row, col, (col + 1 < g.N) ? 2 : 1, acc[i][j][0], acc[i][j][1]);
      }
    }
  }
  if (Epi::kPreload) {
#pragma unroll
    for (int i = 0; i < Cfg::MI; ++i)
#pragma unroll
      for (int j = 0; j < Cfg::NI; ++j) epi.preload_finish(acc[i][j][0], acc[i][j][1]);
  }

  for (int kt = 0; kt < KT; ++kt) {
    const int slot = kt % Cfg::STAGES;
    // every warp is done with the slot that was read in the previous iteration: refill it
    __syncthreads();
    const int nk = kt + Cfg::STAGES - 1;
    if (threadIdx.x == 0 && nk < KT) issue(nk % Cfg::STAGES, nk);
    mbar_wait(&full_bar[slot], (unsigned)((kt / Cfg::STAGES) & 1));
    // lane (gr, t = lane & 3): row gr of each 8-row group, columns 4*ks + t of the slice; swizzled chunk
    const double* sA = smem + slot * Cfg::STAGE + (wm * Cfg::WM + gr) * 16 + (lane & 3);
    const double* sB = smem + slot * Cfg::STAGE + Cfg::A_STAGE + (wn * Cfg::WN + gr) * 16 + (lane & 3);
    const int g3 = gr & 3;
#pragma unroll
    for (int k4 = 0; k4 < Cfg::K4; ++k4) {
      const int u = k4 >> 2, chunk = ((k4 & 3) ^ g3) << 2;
      double a[Cfg::MI], b[Cfg::NI];
#pragma unroll
      for (int i = 0; i < Cfg::MI; ++i) a[i] = sA[u * Cfg::BM * 16 + i * 128 + chunk];
#pragma unroll
      for (int j = 0; j < Cfg::NI; ++j) b[j] = sB[u * Cfg::BN * 16 + j * 128 + chunk];
#pragma unroll
      for (int i = 0; i < Cfg::MI; ++i)
#pragma unroll
        for (int j = 0; j < Cfg::NI; ++j) dmma884(acc[i][j][0], acc[i][j][1], a[i], b[j]);
    }
  }

#pragma unroll
  for (int i = 0; i < Cfg::MI; ++i) {
    const int row = row_base + i * 8;
    if (row >= g.M) continue;
    if constexpr (Epi::kRowBatch) {
      static_assert(Cfg::NI == 4, "row-batched epilogues take 8 values");
      epi.apply_row(row, col_base, g.N, acc[i][0][0], acc[i][0][1], acc[i][1][0], acc[i][1][1], acc[i][2][0],
                    acc[i][2][1], acc[i][3][0], acc[i][3][1]);
    } else {
#pragma unroll
      for (int j = 0; j < Cfg::NI; ++j) {
        const int col = col_base + j * 8;
        if (col >= g.N) continue;
        epi.apply(row, col, acc[i][j][0], acc[i][j][1], (col + 1 < g.N) ? 2 : 1);
      }
    }
  }
}

typedef CUresult (*TensorMapEncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                      const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                      CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline TensorMapEncodeFn tensor_map_encoder() {
  static TensorMapEncodeFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = (TensorMapEncodeFn)p;
  }
  return fn;
}

// row-major matrix P (rows x cols, leading dimension ld doubles): boxes of {16 columns, box_rows rows}
inline int make_operand_map(CUtensorMap* map, const double* P, i64 rows, i64 cols, i64 ld, int box_rows) {
  TensorMapEncodeFn enc = tensor_map_encoder();
  if (!enc) return -1;
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * sizeof(double)};
  cuuint32_t box[2] = {16u, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1u, 1u};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, (void*)P, dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : -1;
}

template <class Cfg, class Epi>
int launch_gemm_nt_tma(GemmArgs g, const Epi& epi, cudaStream_t st) {
  if (g.M <= 0 || g.N <= 0) return 0;
  static_assert(Cfg::BK % 16 == 0, "K slices are made of 16-column sub-tiles");
  if (g.K < 1 || (g.lda & 1) || (g.ldb & 1)) return -1;
  if ((((uintptr_t)g.A) & 15) || (((uintptr_t)g.B) & 15)) return -1;
  alignas(64) CUtensorMap mapA, mapB;
  if (make_operand_map(&mapA, g.A, g.M, g.K, g.lda, Cfg::BM) != 0) return -20;
  if (make_operand_map(&mapB, g.B, g.N, g.K, g.ldb, Cfg::BN) != 0) return -20;
  static bool configured[64] = {false};
  if (first_use_on_device(configured)) {
    STPYB_CUDA(cudaFuncSetAttribute(gemm_nt_tma_kernel<Cfg, Epi>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    (int)Cfg::SMEM));
  }
  i64 grid = plan_grid<Cfg>(g);
  if (grid <= 0) return 0;
  if (grid > 2147483647LL) return -2;
  gemm_nt_tma_kernel<Cfg, Epi><<<(unsigned)grid, Cfg::THREADS, Cfg::SMEM, st>>>(g, epi, mapA, mapB);
  STPYB_COUNT_LAUNCH();
  STPYB_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace stpyb
