// Blocked right-looking Cholesky factorisation (lower, row-major, in place).
//
// Replaces torch.linalg.cholesky / lstsq / lu_factor / slogdet+solve on the GP
// path: stpy/estimator.py:35, stpy/continuous_processes/gauss_procc.py:367-378,
// 633-635.  Structure per 128-wide block column j:
//   C1 potrf_diag_kernel : factor the 128x128 diagonal block in shared memory
//                          and form inv(L_jj) (kept for all later solves)
//   C2 panel TRSM        : A21 <- A21 * inv(L_jj)^T  as a DMMA GEMM (in place)
//   C3 trailing update   : A22 <- A22 - L21 * L21^T  as a DMMA SYRK over the
//                          lower tiles, with K = outer panel width (128..512)
#include "gemm_nt.cuh"
#include "stpyb_internal.h"

namespace stpyb {

constexpr int SLD = 129;  // odd row stride: consecutive rows fall in distinct shared-memory banks
constexpr int LEAF = 8;

// One CTA (512 threads) factors a diagonal block of order b <= 128 held in shared memory and
// forms its triangular inverse.  The serial dependency chain of a Cholesky (pivot -> scale ->
// update, 128 times) is what bounds this kernel, so it is organised to keep that chain short:
//   * 16 leaf panels of 8 columns; the 8x8 leaf and its inverse are done by ONE thread entirely
//     in registers (rsqrt instead of sqrt + divide, no barriers inside the leaf);
//   * the rows below the leaf are solved one thread per row against the leaf inverse (36 FMAs);
//   * the rank-8 trailing update uses one 4x4 register tile per thread;
//   * the off-leaf blocks of the inverse are then built by 16 warps, one per 8-column block
//     column, each sweeping down its own column with contiguous dot products (no block barrier).
// inv(L) is parked transposed in the unused strict upper triangle of the same buffer, its
// diagonal in dinv[].
__global__ void __launch_bounds__(512, 1)
potrf_diag_kernel(double* __restrict__ A, i64 lda, int b, double* __restrict__ Linv, int* info, int j0) {
  extern __shared__ __align__(16) double S[];  // [128][SLD] + dinv[128] + scratch[16][72]
  double* dinv = S + DB * SLD;
  double* scratch = dinv + DB;
  const int tid = threadIdx.x;

  for (int idx = tid; idx < DB * DB; idx += 512) {
    const int r = idx >> 7, c = idx & 127;
    double v = 0.0;
    if (r < b && c <= r) v = A[(i64)r * lda + c];
    else if (r >= b && c == r) v = 1.0;  // identity padding keeps tail blocks on the same code path
    S[r * SLD + c] = v;
  }
  __syncthreads();

  for (int p = 0; p < DB / LEAF; ++p) {
    const int o = p * LEAF;
    // ---- leaf: Cholesky of the 8x8 diagonal block and its inverse, one thread, registers only
    if (tid == 0) {
      double a[LEAF][LEAF], x[LEAF][LEAF], rs[LEAF];
#pragma unroll
      for (int i = 0; i < LEAF; ++i)
#pragma unroll
        for (int j = 0; j <= i; ++j) a[i][j] = S[(o + i) * SLD + o + j];
#pragma unroll
      for (int j = 0; j < LEAF; ++j) {
        const double d = a[j][j];
        if (!(d > 0.0) && o + j < b) atomicCAS(info, 0, j0 + o + j + 1);
        rs[j] = rsqrt(d);
        a[j][j] = d * rs[j];
#pragma unroll
        for (int i = j + 1; i < LEAF; ++i) a[i][j] *= rs[j];
#pragma unroll
        for (int k = j + 1; k < LEAF; ++k)
#pragma unroll
          for (int i = k; i < LEAF; ++i) a[i][k] = fma(-a[i][j], a[k][j], a[i][k]);
      }
#pragma unroll
      for (int j = 0; j < LEAF; ++j) {
        x[j][j] = rs[j];
#pragma unroll
        for (int i = j + 1; i < LEAF; ++i) {
          double t = 0.0;
#pragma unroll
          for (int k = j; k < i; ++k) t = fma(a[i][k], x[k][j], t);
          x[i][j] = -t * rs[i];
        }
      }
#pragma unroll
      for (int i = 0; i < LEAF; ++i) {
        dinv[o + i] = rs[i];
#pragma unroll
        for (int j = 0; j <= i; ++j) {
          S[(o + i) * SLD + o + j] = a[i][j];
          if (j < i) S[(o + j) * SLD + o + i] = x[i][j];  // inverse, transposed into the upper triangle
        }
      }
    }
    __syncthreads();
    // ---- panel: rows below the leaf, one thread per row:  w = v * inv(leaf)^T
    if (tid < DB && tid >= o + LEAF) {
      double* row = S + tid * SLD + o;
      double v[LEAF], w[LEAF];
#pragma unroll
      for (int k = 0; k < LEAF; ++k) v[k] = row[k];
#pragma unroll
      for (int c = 0; c < LEAF; ++c) {
        double t = v[c] * dinv[o + c];
#pragma unroll
        for (int k = 0; k < c; ++k) t = fma(v[k], S[(o + k) * SLD + o + c], t);
        w[c] = t;
      }
#pragma unroll
      for (int c = 0; c < LEAF; ++c) row[c] = w[c];
    }
    __syncthreads();
    // ---- rank-8 update of the trailing block, one 4x4 tile per thread over the lower triangle
    {
      const int t0 = o + LEAF;
      const int nt = (DB - t0) >> 2;  // tiles per side
      if (tid < nt * (nt + 1) / 2) {
        int I = (int)((sqrtf(8.0f * (float)tid + 1.0f) - 1.0f) * 0.5f);
        while ((I + 1) * (I + 2) / 2 <= tid) ++I;
        while (I * (I + 1) / 2 > tid) --I;
        const int J = tid - I * (I + 1) / 2;
        const int i0 = t0 + 4 * I, c0 = t0 + 4 * J;
        double acc[4][4];
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
          for (int v = 0; v < 4; ++v) acc[u][v] = 0.0;
#pragma unroll
        for (int k = 0; k < LEAF; ++k) {
          double ri[4], rj[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) ri[u] = S[(i0 + u) * SLD + o + k];
#pragma unroll
          for (int v = 0; v < 4; ++v) rj[v] = S[(c0 + v) * SLD + o + k];
#pragma unroll
          for (int u = 0; u < 4; ++u)
#pragma unroll
            for (int v = 0; v < 4; ++v) acc[u][v] = fma(ri[u], rj[v], acc[u][v]);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
          for (int v = 0; v < 4; ++v)
            if (c0 + v <= i0 + u) S[(i0 + u) * SLD + c0 + v] -= acc[u][v];
      }
    }
    __syncthreads();
  }

  // ---- inverse assembly: warp c owns block column c of inv(L); sweep the block rows below
  {
    const int c = tid >> 5, lane = tid & 31;
    const int i = lane >> 2, jj = (lane & 3) * 2;  // outputs (i, jj) and (i, jj + 1) of the 8x8 block
    double* sc = scratch + c * 72;                  // [8][9]
    const int col0 = c * LEAF + jj, col1 = col0 + 1;
    for (int rb = c + 1; rb < DB / LEAF; ++rb) {
      const double* lrow = S + (rb * LEAF + i) * SLD;
      const double* x0 = S + col0 * SLD;
      const double* x1 = S + col1 * SLD;
      const int kend = rb * LEAF;
      // s = sum_{k >= col} L[row][k] * X[k][col]; X[col][col] = dinv[col], X[k][col] = S[col][k] (k > col)
      double s0 = lrow[col0] * dinv[col0], s1 = lrow[col1] * dinv[col1];
      s0 = fma(lrow[col1], x0[col1], s0);
      double t0 = 0.0, t1 = 0.0;
      int k = col1 + 1;
      for (; k + 1 < kend; k += 2) {
        const double l0 = lrow[k], l1 = lrow[k + 1];
        s0 = fma(l0, x0[k], s0);
        s1 = fma(l0, x1[k], s1);
        t0 = fma(l1, x0[k + 1], t0);
        t1 = fma(l1, x1[k + 1], t1);
      }
      if (k < kend) {
        s0 = fma(lrow[k], x0[k], s0);
        s1 = fma(lrow[k], x1[k], s1);
      }
      sc[i * 9 + jj] = s0 + t0;
      sc[i * 9 + jj + 1] = s1 + t1;
      __syncwarp();
      // X[rb][c] = -inv(leaf_rb) * s ; inv(leaf_rb)[i][a] = S[(rb*8+a)][rb*8+i] for a < i, dinv on the diagonal
      double o0 = dinv[rb * LEAF + i] * sc[i * 9 + jj], o1 = dinv[rb * LEAF + i] * sc[i * 9 + jj + 1];
      for (int a = 0; a < i; ++a) {
        const double xi = S[(rb * LEAF + a) * SLD + rb * LEAF + i];
        o0 = fma(xi, sc[a * 9 + jj], o0);
        o1 = fma(xi, sc[a * 9 + jj + 1], o1);
      }
      __syncwarp();
      S[col0 * SLD + rb * LEAF + i] = -o0;
      S[col1 * SLD + rb * LEAF + i] = -o1;
      __syncwarp();
    }
  }
  __syncthreads();

  for (int idx = tid; idx < DB * DB; idx += 512) {
    const int r = idx >> 7, c = idx & 127;
    if (r < b && c <= r) A[(i64)r * lda + c] = S[r * SLD + c];
    double li = 0.0;
    if (r < b && c < r) li = S[c * SLD + r];
    else if (r < b && c == r) li = dinv[r];
    Linv[idx] = li;
  }
}

int potrf_diag(double* A, i64 lda, int b, double* Linv, int* info, int j0, cudaStream_t st) {
  static bool configured = false;
  const int smem = (DB * SLD + DB + 16 * 72) * (int)sizeof(double);
  if (!configured) {
    STPYB_CUDA(cudaFuncSetAttribute(potrf_diag_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    configured = true;
  }
  prof_begin(PROF_DIAG, (double)b * b * b / 3.0, st);
  potrf_diag_kernel<<<1, 512, smem, st>>>(A, lda, b, Linv, info, j0);
  prof_end(st);
  STPYB_COUNT_LAUNCH();
  STPYB_CUDA(cudaGetLastError());
  return 0;
}

// C[M x N] = alpha * A B^T + beta * C, generic entry used by every blocked stage.
int gemm_nt(int M, int N, int K, const double* A, i64 lda, const double* B, i64 ldb, double* C, i64 ldc,
            double alpha, double beta, int tri, int square_cfg, cudaStream_t st, int kskip) {
  GemmArgs g;
  g.A = A; g.B = B; g.lda = lda; g.ldb = ldb; g.M = M; g.N = N; g.K = K; g.tri = tri; g.kskip = kskip;
  if (beta == 1.0 && (alpha == 1.0 || alpha == -1.0) && !square_cfg) {
    EpiAccum e;
    e.C = C; e.ldc = ldc; e.negate = (alpha < 0.0) ? 1 : 0;
    e.vec = ((ldc & 1) == 0 && (((uintptr_t)C) & 15) == 0) ? 1 : 0;
    return launch_gemm_nt<CfgStream, EpiAccum>(g, e, st);
  }
  EpiAxpby e = make_axpby(C, ldc, alpha, beta);
  if (square_cfg) return launch_gemm_nt<CfgSquare, EpiAxpby>(g, e, st);
  return launch_gemm_nt<CfgStream, EpiAxpby>(g, e, st);
}

// Factor the tall panel P (rows x w, top w x w block = diagonal block) in place.
int potrf_panel(double* P, i64 rows, int w, i64 ldp, double* dinv, int* info, i64 j0, cudaStream_t st) {
  for (int j = 0; j < w; j += DB) {
    const int b = (w - j < DB) ? (w - j) : DB;
    double* Pjj = P + (i64)j * ldp + j;
    double* Li = dinv + (i64)(j / DB) * (DB * DB);
    STPYB_TRY(potrf_diag(Pjj, ldp, b, Li, info, (int)(j0 + j), st));
    const i64 below = rows - (j + b);
    if (below > 0) {
      // in-place panel TRSM: rows below the diagonal block
      double* P21 = P + (i64)(j + b) * ldp + j;
      prof_begin(PROF_TRSM, (double)below * b * b, st);
      STPYB_TRY(gemm_nt((int)below, b, b, P21, ldp, Li, DB, P21, ldp, 1.0, 0.0, TRI_FULL, 1, st));
      prof_end(st);
      const int rest = w - (j + b);  // remaining columns inside the panel
      if (rest > 0) {
        double* C = P + (i64)(j + b) * ldp + (j + b);
        prof_begin(PROF_PANEL_UPD, 2.0 * ((double)below * rest - 0.5 * (double)rest * rest) * b, st);
        STPYB_TRY(gemm_nt((int)below, rest, b, P21, ldp, P21, ldp, C, ldp, -1.0, 1.0, TRI_LOWER, 0, st));
        prof_end(st);
      }
    }
  }
  return 0;
}

int potrf_lower(double* A, i64 n, i64 lda, double* dinv, int* info, int outer, cudaStream_t st) {
  if (n <= 0) return 0;
  if ((lda & 1) || (((uintptr_t)A) & 15) || (((uintptr_t)dinv) & 15)) return -3;
  if (outer < DB) outer = DB;
  outer = (outer / DB) * DB;
  STPYB_CUDA(cudaMemsetAsync(info, 0, sizeof(int), st));
  for (i64 J = 0; J < n; J += outer) {
    const int jb = (int)((n - J < outer) ? (n - J) : outer);
    STPYB_TRY(potrf_panel(A + J * lda + J, n - J, jb, lda, dinv + (J / DB) * (i64)(DB * DB), info, J, st));
    const i64 trail = n - (J + jb);
    if (trail > 0) {
      const double* P = A + (J + jb) * lda + J;
      double* C = A + (J + jb) * lda + (J + jb);
      prof_begin(PROF_SYRK, (double)trail * (double)trail * jb, st);
      STPYB_TRY(gemm_nt((int)trail, (int)trail, jb, P, lda, P, lda, C, lda, -1.0, 1.0, TRI_LOWER, 0, st));
      prof_end(st);
    }
  }
  return 0;
}

}  // namespace stpyb

using namespace stpyb;

extern "C" int stpyb_potrf(double* K_inout, long long n, long long ld, double* dinv, int* info_dev,
                           int outer_block, void* stream) {
  return potrf_lower(K_inout, n, ld, dinv, info_dev, outer_block, (cudaStream_t)stream);
}

extern "C" int stpyb_potrf_panel(double* P, long long rows, int w, long long ldp, double* dinv, int* info_dev,
                                 long long j0, void* stream) {
  if (rows < w || w <= 0) return -2;
  if ((ldp & 1) || (((uintptr_t)P) & 15)) return -4;
  return potrf_panel(P, rows, w, ldp, dinv, info_dev, j0, (cudaStream_t)stream);
}

extern "C" int stpyb_gemm_nt(int M, int N, int K, const double* A, long long lda, const double* B,
                             long long ldb, double* C, long long ldc, double alpha, double beta, int lower,
                             void* stream) {
  return gemm_nt(M, N, K, A, lda, B, ldb, C, ldc, alpha, beta, lower ? TRI_LOWER : TRI_FULL, 0,
                 (cudaStream_t)stream);
}
