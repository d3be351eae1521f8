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

constexpr int SLD = 132;  // padded smem row stride: 132 mod 16 == 4 -> conflict-free 4-lane rows

// One CTA, 512 threads; 4 lanes cooperate on one matrix row / inverse column.
__global__ void __launch_bounds__(512, 1)
potrf_diag_kernel(double* __restrict__ A, i64 lda, int b, double* __restrict__ Linv, int* info, int j0) {
  extern __shared__ __align__(16) double S[];  // [128][SLD] + dinv[128]
  double* dinv = S + DB * SLD;
  const int tid = threadIdx.x;
  const int i = tid >> 2, q = tid & 3;

  for (int idx = tid; idx < DB * DB; idx += 512) {
    int r = idx >> 7, c = idx & 127;
    double v = 0.0;
    if (r < b && c <= r) v = A[(i64)r * lda + c];
    S[r * SLD + c] = v;
  }
  __syncthreads();

  // Left-looking column Cholesky.
  for (int j = 0; j < b; ++j) {
    double p0 = 0.0, p1 = 0.0;
    if (i >= j && i < b) {
      const double* ri = S + i * SLD;
      const double* rj = S + j * SLD;
      int k = q;
      for (; k + 4 < j; k += 8) {
        p0 = fma(ri[k], rj[k], p0);
        p1 = fma(ri[k + 4], rj[k + 4], p1);
      }
      if (k < j) p0 = fma(ri[k], rj[k], p0);
    }
    double p = p0 + p1;
    p += __shfl_xor_sync(0xffffffffu, p, 1);
    p += __shfl_xor_sync(0xffffffffu, p, 2);
    double v = 0.0;
    if (i >= j && i < b) v = S[i * SLD + j] - p;
    if (i == j && q == 0) {
      if (!(v > 0.0)) {
        atomicCAS(info, 0, j0 + j + 1);
      }
      S[j * SLD + j] = sqrt(v);
    }
    __syncthreads();
    if (i > j && i < b && q == 0) S[i * SLD + j] = v / S[j * SLD + j];
    __syncthreads();
  }

  if (tid < DB) dinv[tid] = (tid < b) ? 1.0 / S[tid * SLD + tid] : 0.0;
  __syncthreads();

  // Triangular inverse: lane group j builds column j of inv(L) by forward
  // substitution and parks it, transposed, in the unused upper triangle of S.
  // Groups run different trip counts, so shuffles name only the group's lanes.
  {
    const int j = i;
    const unsigned gmask = 0xFu << ((tid & 31) & ~3);
    if (j < b) {
      const double* rj = S + j * SLD;
      const double dj = dinv[j];
      for (int r = j + 1; r < b; ++r) {
        const double* rr = S + r * SLD;
        double s0 = (q == 0) ? rr[j] * dj : 0.0, s1 = 0.0;
        int k = j + 1 + q;
        for (; k + 4 < r; k += 8) {
          s0 = fma(rr[k], rj[k], s0);
          s1 = fma(rr[k + 4], rj[k + 4], s1);
        }
        if (k < r) s0 = fma(rr[k], rj[k], s0);
        double s = s0 + s1;
        s += __shfl_xor_sync(gmask, s, 1);
        s += __shfl_xor_sync(gmask, s, 2);
        if (q == 0) S[j * SLD + r] = -s * dinv[r];
        __syncwarp(gmask);
      }
    }
  }
  __syncthreads();

  for (int idx = tid; idx < DB * DB; idx += 512) {
    int r = idx >> 7, c = idx & 127;
    if (r < b && c <= r) A[(i64)r * lda + c] = S[r * SLD + c];
    double li = 0.0;
    if (r < b && c < r) li = S[c * SLD + r];
    else if (r < b && c == r) li = dinv[r];
    Linv[idx] = li;
  }
}

int potrf_diag(double* A, i64 lda, int b, double* Linv, int* info, int j0, cudaStream_t st) {
  static bool configured = false;
  const int smem = (DB * SLD + DB) * (int)sizeof(double);
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
