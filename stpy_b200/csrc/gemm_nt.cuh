// Tiled FP64 tensor-core contraction  acc[M x N] = A[M x K] * B[N x K]^T
// (both operands row-major with K contiguous: "NT"), with a pluggable register
// epilogue.  Every dense stage of the GP hot path is phrased in this form:
//   Gram cross term   b a^T              (stpy/kernels.py:393, 579, 782)
//   SYRK / GEMM trailing update, TRSM against an inverted diagonal block
//                                        (torch.linalg.cholesky, estimator.py:35)
//   RFF projection    X W^T              (stpy/embeddings/embedding.py:232-234)
//   normal equations  Phi^T Phi          (kernelized_features.py:237)
//
// Mapping to sm_100a: CTA tile BM x BN, warp tile WM x WN built from
// DMMA.8x8x4 atoms (accumulators in registers; there is no f64 tcgen05 kind),
// K streamed in 16-wide slices through a STAGES-deep cp.async ring in shared
// memory using the k4-packed layout described in common.cuh.
#pragma once
#include "common.cuh"

namespace stpyb {

template <int BM_, int BN_, int WM_, int WN_, int STAGES_, int MINB_, int BK_ = 16>
struct TileCfg {
  static constexpr int BM = BM_, BN = BN_, WM = WM_, WN = WN_, STAGES = STAGES_, MINB = MINB_;
  static constexpr int BK = BK_, K4 = BK / 4;
  static constexpr int WARPS_M = BM / WM, WARPS_N = BN / WN;
  static constexpr int THREADS = WARPS_M * WARPS_N * 32;
  static constexpr int A_STAGE = BM * BK, B_STAGE = BN * BK;  // doubles
  static constexpr int STAGE = A_STAGE + B_STAGE;
  static constexpr size_t SMEM = (size_t)STAGES * STAGE * sizeof(double);
  static constexpr int MI = WM / 8, NI = WN / 8;
};

// 128x64 tile, 4 warps of 64x32, two CTAs per SM: the second CTA's main loop
// hides the first one's prologue fill and read-modify-write epilogue.
typedef TileCfg<128, 64, 64, 32, 4, 2> CfgStream;
// the same tile with 32-wide K slices and a 2-stage ring (same 96 KB of shared memory): one
// barrier per 256 DMMAs per warp.  Used by the update kernels, whose K is 128..512 and whose
// slices are long enough (>= 8k cycles) for double buffering to cover the global-load latency.
typedef TileCfg<128, 64, 64, 32, 2, 2, 32> CfgStreamK32;
// 128x128 tile, 8 warps, one CTA per SM: a CTA owns full 128-wide rows, which
// makes the in-place panel TRSM (output rows == input rows) race-free.
typedef TileCfg<128, 128, 64, 32, 3, 1> CfgSquare;

enum { TRI_FULL = 0, TRI_LOWER = 1 };

struct GemmArgs {
  const double* A;
  const double* B;
  i64 lda, ldb;   // in doubles; must be even, base pointers 16-byte aligned
  int M, N, K;    // any K >= 1 (tails are zero-filled in shared memory)
  int tri;        // TRI_LOWER: skip tiles strictly above the diagonal of the M x N block
  int kskip;      // 1: rows >= m0 of A and B are zero before column m0 (U U^T): start the K loop at m0
  int tiles_m, tiles_n;
  int tri_rows;   // number of tile rows in the triangular (uncapped) part
  i64 tri_count;  // CTAs in the triangular part
};

// Number of column tiles kept in tile-row r for TRI_LOWER (before capping).
template <class Cfg>
__host__ __device__ inline i64 tri_tiles_before(i64 r) {
  // sum_{q<r} c*(q+1) with c = BM/BN column tiles per row tile
  return (i64)(Cfg::BM / Cfg::BN) * r * (r + 1) / 2;
}

template <class Cfg>
inline i64 plan_grid(GemmArgs& g) {
  static_assert(Cfg::BM % Cfg::BN == 0, "BM must be a multiple of BN");
  g.tiles_m = ceil_div(g.M, Cfg::BM);
  g.tiles_n = ceil_div(g.N, Cfg::BN);
  if (g.tri == TRI_FULL) {
    g.tri_rows = 0;
    g.tri_count = 0;
    return (i64)g.tiles_m * g.tiles_n;
  }
  const int c = Cfg::BM / Cfg::BN;
  int r0 = g.tiles_n / c;  // rows r with c*(r+1) <= tiles_n
  if (r0 > g.tiles_m) r0 = g.tiles_m;
  g.tri_rows = r0;
  g.tri_count = tri_tiles_before<Cfg>(r0);
  return g.tri_count + (i64)(g.tiles_m - r0) * g.tiles_n;
}

// Linear block index -> tile.  Tiles are walked band by band (BAND tile rows = 1024 matrix rows)
// and column-major inside a band: the ~300 CTAs resident at any time then share a handful of
// B tile-columns and one A band (a few MB), so operand re-reads hit L2 instead of DRAM (the
// row-major walk streamed the whole panel per tile row: 9.5 GB of DRAM re-reads per launch at
// order 28k, L2 hit rate 66 %).
constexpr int BAND = 8;

template <class Cfg>
__device__ __forceinline__ i64 tiles_before_row(const GemmArgs& g, i64 r) {
  if (g.tri == TRI_FULL) return r * g.tiles_n;
  const i64 rt = r < g.tri_rows ? r : g.tri_rows;
  return tri_tiles_before<Cfg>(rt) + (r - rt) * g.tiles_n;
}

template <class Cfg>
__device__ __forceinline__ int tiles_in_row(const GemmArgs& g, int r) {
  if (g.tri == TRI_FULL) return g.tiles_n;
  const int c = (Cfg::BM / Cfg::BN) * (r + 1);
  return c < g.tiles_n ? c : g.tiles_n;
}

template <class Cfg>
__device__ __forceinline__ void decode_tile(const GemmArgs& g, i64 bid, int& tm, int& tn) {
  // 1. the tile row that contains bid in the row-major enumeration fixes the band
  i64 r;
  if (g.tri == TRI_FULL) {
    r = bid / g.tiles_n;
  } else if (bid < g.tri_count) {
    const double c = (double)(Cfg::BM / Cfg::BN);
    r = (i64)((sqrt(8.0 * (double)bid / c + 1.0) - 1.0) * 0.5);
    while (tri_tiles_before<Cfg>(r + 1) <= bid) ++r;
    while (tri_tiles_before<Cfg>(r) > bid) --r;
  } else {
    r = g.tri_rows + (bid - g.tri_count) / g.tiles_n;
  }
  const int r0 = (int)(r / BAND) * BAND;
  const int nr = (g.tiles_m - r0 < BAND) ? (g.tiles_m - r0) : BAND;
  i64 l = bid - tiles_before_row<Cfg>(g, r0);
  // 2. inside the band: the first cnt(r0) tile columns are full (nr rows each) ...
  const int cnt0 = tiles_in_row<Cfg>(g, r0);
  if (l < (i64)nr * cnt0) {
    tn = (int)(l / nr);
    tm = r0 + (int)(l - (i64)tn * nr);
    return;
  }
  // ... then the staircase of the lower-triangular part: column t is present in rows whose count exceeds t
  l -= (i64)nr * cnt0;
  int t = cnt0;
  int first = 1;  // first row (relative to r0) that still has column t
  for (;;) {
    while (first < nr && tiles_in_row<Cfg>(g, r0 + first) <= t) ++first;
    const int rows = nr - first;
    if (l < rows) {
      tn = t;
      tm = r0 + first + (int)l;
      return;
    }
    l -= rows;
    ++t;
  }
}

// Per-thread copy plan for one operand tile of ROWS rows.  Chunk c = tid + THREADS*i moves 16
// bytes: half = c & 1, row = (c >> 1) % ROWS, k4 = (c >> 1) / ROWS.  With ROWS a multiple of
// THREADS/2 a thread keeps its k-half and touches only ROWS/(THREADS/2) distinct rows, so the
// global row pointers are formed once and the K loop only adds the slice offset.  A warp covers
// 16 rows x one 32-byte k4 group: full sectors in global memory, one contiguous 512-byte run in
// shared memory (conflict-free).
template <class Cfg, int ROWS>
struct OperandPlan {
  static constexpr int ROW_STEP = Cfg::THREADS / 2;
  static_assert(ROWS % ROW_STEP == 0, "ROWS must be a multiple of THREADS/2");
  static constexpr int NROW = ROWS / ROW_STEP;  // distinct rows per thread
  const double* ptr[NROW];                      // row base + half*2
  bool ok[NROW];
  int soff;                                     // shared offset (doubles) of (row, half) in k4 group 0
  int half2;

  __device__ __forceinline__ void init(const double* P, i64 ld, int r0, int rmax) {
    half2 = (threadIdx.x & 1) * 2;
    const int r = threadIdx.x >> 1;
    soff = r * 4 + half2;
#pragma unroll
    for (int q = 0; q < NROW; ++q) {
      const int grow = r0 + r + q * ROW_STEP;
      ok[q] = grow < rmax;
      ptr[q] = P + (ok[q] ? (i64)grow * ld : 0) + half2;
    }
  }
  // copy the k4-th group of K-slice [k0, k0+16) into sdst; kleft = K - k0 (< 16 only on a ragged last slice)
  __device__ __forceinline__ void issue_k4(double* sdst, int k0, int kleft, int k4) const {
    int bytes = (kleft - k4 * 4 - half2) * 8;
    bytes = bytes > 16 ? 16 : (bytes < 0 ? 0 : bytes);
#pragma unroll
    for (int q = 0; q < NROW; ++q) {
      cp_async16(sdst + soff + (k4 * ROWS + q * ROW_STEP) * 4, ptr[q] + k0 + k4 * 4, ok[q] ? bytes : 0);
    }
  }
  __device__ __forceinline__ void issue(double* sdst, int k0, int kleft) const {
#pragma unroll
    for (int k4 = 0; k4 < Cfg::K4; ++k4) issue_k4(sdst, k0, kleft, k4);
  }
};

template <class Cfg, class Epi>
__global__ void __launch_bounds__(Cfg::THREADS, Cfg::MINB) gemm_nt_kernel(GemmArgs g, Epi epi) {
  extern __shared__ __align__(16) double smem[];
  int tm, tn;
  decode_tile<Cfg>(g, (i64)blockIdx.x, tm, tn);
  const int m0 = tm * Cfg::BM, n0 = tn * Cfg::BN;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int wm = warp / Cfg::WARPS_N, wn = warp % Cfg::WARPS_N;
  // lane (gr = lane/4, tc = 2*(lane%4)) owns rows gr+8i and column pairs tc+8j of the warp tile
  const int gr = lane >> 2, tc = (lane & 3) * 2;
  const int row_base = m0 + wm * Cfg::WM + gr, col_base = n0 + wn * Cfg::WN + tc;

  const int kbeg = g.kskip ? m0 : 0;  // m0 is a multiple of BK
  const int Keff = g.K - kbeg;
  const int KT = (Keff + Cfg::BK - 1) / Cfg::BK;
  OperandPlan<Cfg, Cfg::BM> pa;
  OperandPlan<Cfg, Cfg::BN> pb;
  pa.init(g.A + kbeg, g.lda, m0, g.M);
  pb.init(g.B + kbeg, g.ldb, n0, g.N);

  // start filling the ring first, then (for accumulate epilogues) pull the C tile straight into
  // the accumulator registers: all loads are in flight together and nothing is left to read in
  // the epilogue.
#pragma unroll
  for (int s = 0; s < Cfg::STAGES - 1; ++s) {
    if (s < KT) {
      double* st = smem + s * Cfg::STAGE;
      pa.issue(st, s * Cfg::BK, Keff - s * Cfg::BK);
      pb.issue(st + Cfg::A_STAGE, s * Cfg::BK, Keff - s * Cfg::BK);
    }
    cp_async_commit();
  }

  double acc[Cfg::MI][Cfg::NI][2];
#pragma unroll
  for (int i = 0; i < Cfg::MI; ++i) {
#pragma unroll
    for (int j = 0; j < Cfg::NI; ++j) {
      acc[i][j][0] = acc[i][j][1] = 0.0;
      if (Epi::kPreload) {
        // raw loads only: nothing here may consume a loaded value, or the in-order warp would
        // serialise the 2*MI*NI round trips
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
    cp_async_wait<Cfg::STAGES - 2>();
    __syncthreads();
    const int nk = kt + Cfg::STAGES - 1;
    double* nst = smem + (nk % Cfg::STAGES) * Cfg::STAGE;
    const bool more = nk < KT;
    const double* sA = smem + (kt % Cfg::STAGES) * Cfg::STAGE + (wm * Cfg::WM) * 4 + lane;
    const double* sB = smem + (kt % Cfg::STAGES) * Cfg::STAGE + Cfg::A_STAGE + (wn * Cfg::WN) * 4 + lane;
#pragma unroll
    for (int k4 = 0; k4 < Cfg::K4; ++k4) {
      // the refill of the slot freed by the previous iteration is spread over the four k4 steps
      // so that its address arithmetic rides in the issue slots between DMMAs
      if (more) {
        pa.issue_k4(nst, nk * Cfg::BK, Keff - nk * Cfg::BK, k4);
        pb.issue_k4(nst + Cfg::A_STAGE, nk * Cfg::BK, Keff - nk * Cfg::BK, k4);
      }
      double a[Cfg::MI], b[Cfg::NI];
#pragma unroll
      for (int i = 0; i < Cfg::MI; ++i) a[i] = sA[(k4 * Cfg::BM + i * 8) * 4];
#pragma unroll
      for (int j = 0; j < Cfg::NI; ++j) b[j] = sB[(k4 * Cfg::BN + j * 8) * 4];
#pragma unroll
      for (int i = 0; i < Cfg::MI; ++i)
#pragma unroll
        for (int j = 0; j < Cfg::NI; ++j) dmma884(acc[i][j][0], acc[i][j][1], a[i], b[j]);
    }
    cp_async_commit();
  }
  cp_async_wait<0>();

#pragma unroll
  for (int i = 0; i < Cfg::MI; ++i) {
    const int row = row_base + i * 8;
    if (row >= g.M) continue;
    if constexpr (Epi::kRowBatch) {
      // the epilogue maps all 2*NI values of this row before storing (independent chains -> ILP)
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

template <class Cfg, class Epi>
int launch_gemm_nt(GemmArgs g, const Epi& epi, cudaStream_t st) {
  if (g.M <= 0 || g.N <= 0) return 0;
  if (g.K <= 0 || (g.lda & 1) || (g.ldb & 1)) return -1;
  if ((((uintptr_t)g.A) & 15) || (((uintptr_t)g.B) & 15)) return -1;
  static bool configured[64] = {false};
  if (first_use_on_device(configured)) {
    STPYB_CUDA(cudaFuncSetAttribute(gemm_nt_kernel<Cfg, Epi>,
                                    cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM));
  }
  i64 grid = plan_grid<Cfg>(g);
  if (grid <= 0) return 0;
  if (grid > 2147483647LL) return -2;
  gemm_nt_kernel<Cfg, Epi><<<(unsigned)grid, Cfg::THREADS, Cfg::SMEM, st>>>(g, epi);
  STPYB_COUNT_LAUNCH();
  STPYB_CUDA(cudaGetLastError());
  return 0;
}

// C = alpha * acc + beta * C   (beta == 0 never reads C)
struct EpiAxpby {
  static constexpr bool kPreload = false;
  static constexpr bool kRowBatch = false;
  __device__ __forceinline__ void preload(int, int, int, double&, double&) const {}
  __device__ __forceinline__ void preload_finish(double&, double&) const {}
  double* C;
  i64 ldc;
  double alpha, beta;
  int vec;  // 1 if ldc even and C 16-byte aligned
  double* C2 = nullptr;  // optional mirror of the result (same rows / columns, own leading dimension): the
  i64 ldc2 = 0;          // distributed panel solve writes the slab and the broadcast buffer in one pass
  __device__ __forceinline__ void apply(int row, int col, double v0, double v1, int nc) const {
    double* p = C + (i64)row * ldc + col;
    double2 o;
    if (vec && nc == 2) {
      if (beta != 0.0) {
        double2 c = *reinterpret_cast<const double2*>(p);
        o.x = alpha * v0 + beta * c.x;
        o.y = alpha * v1 + beta * c.y;
      } else {
        o.x = alpha * v0;
        o.y = alpha * v1;
      }
      *reinterpret_cast<double2*>(p) = o;
    } else {
      o.x = (beta != 0.0) ? alpha * v0 + beta * p[0] : alpha * v0;
      p[0] = o.x;
      if (nc == 2) {
        o.y = (beta != 0.0) ? alpha * v1 + beta * p[1] : alpha * v1;
        p[1] = o.y;
      }
    }
    if (C2) {
      double* q = C2 + (i64)row * ldc2 + col;
      q[0] = o.x;
      if (nc == 2) q[1] = o.y;
    }
  }
};

// C = C + sign * acc (sign = +1 / -1): the update form of every blocked stage.  The C tile is
// loaded into the accumulators before the main loop (negated for sign = -1, which is exact) and
// the epilogue only stores: no read-modify-write dependency at the end of the tile.
struct EpiAccum {
  static constexpr bool kPreload = true;
  static constexpr bool kRowBatch = false;
  double* C;
  i64 ldc;
  int negate;  // 1: C - A B^T
  int vec;
  __device__ __forceinline__ void preload(int row, int col, int nc, double& v0, double& v1) const {
    const double* p = C + (i64)row * ldc + col;
    if (vec && nc == 2) {
      const double2 c = *reinterpret_cast<const double2*>(p);
      v0 = c.x;
      v1 = c.y;
    } else {
      v0 = p[0];
      v1 = (nc == 2) ? p[1] : 0.0;
    }
  }
  // sign flips are integer XORs on the high word (exact, off the FP64 pipe)
  __device__ __forceinline__ void flip(double& v) const {
    v = __longlong_as_double(__double_as_longlong(v) ^ (negate ? (long long)0x8000000000000000ULL : 0LL));
  }
  __device__ __forceinline__ void preload_finish(double& v0, double& v1) const {
    flip(v0);
    flip(v1);
  }
  __device__ __forceinline__ void apply(int row, int col, double v0, double v1, int nc) const {
    double* p = C + (i64)row * ldc + col;
    flip(v0);
    flip(v1);
    if (vec && nc == 2) {
      *reinterpret_cast<double2*>(p) = make_double2(v0, v1);
    } else {
      p[0] = v0;
      if (nc == 2) p[1] = v1;
    }
  }
};

inline EpiAxpby make_axpby(double* C, i64 ldc, double alpha, double beta) {
  EpiAxpby e;
  e.C = C;
  e.ldc = ldc;
  e.alpha = alpha;
  e.beta = beta;
  e.vec = ((ldc & 1) == 0 && (((uintptr_t)C) & 15) == 0) ? 1 : 0;
  return e;
}

}  // namespace stpyb
