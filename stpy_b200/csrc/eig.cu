// Symmetric eigendecomposition by one-sided (Hestenes) Jacobi rotations, for the Nystrom feature map
// (stpy/continuous_processes/nystrom_fea.py:116-136, 188-196: torch.linalg.eigh of a Gram matrix).
//
// W holds the eigenvector candidates as ROWS (initially I) and H = W A.  Rotating rows p and q of H and W by
// the same plane rotation keeps H = W A; the rotation is chosen to make H_p and H_q orthogonal.  When all rows
// of H are mutually orthogonal, H H^T = W A^2 W^T is diagonal: the rows of W are eigenvectors of A and
// lambda_i = H_i . W_i (the Rayleigh quotient, which also carries the sign).  One ROUND rotates n/2 disjoint
// pairs -- one CTA per pair, the round-robin "circle" ordering -- and n-1 rounds (one SWEEP) visit every pair;
// the host stops after the first sweep in which no pair needed a rotation.  The method is backward stable and
// computes small eigenvalues of a positive semi-definite Gram matrix to high relative accuracy, which is what
// the 1/sqrt(lambda) scaling of the feature map needs.
#include "common.cuh"
#include "stpyb_internal.h"
#include "../../include/stpyb.h"

namespace stpyb {

__global__ void __launch_bounds__(256) jacobi_init_kernel(const double* __restrict__ A, i64 lda, double* H, double* W,
                                                         i64 n, i64 np, i64 ld) {
  // H = A (zero padded to np rows / columns), W = I
  const i64 i = blockIdx.x;
  for (i64 j = threadIdx.x; j < np; j += 256) {
    H[i * ld + j] = (i < n && j < n) ? A[i * lda + j] : 0.0;
    W[i * ld + j] = (i == j) ? 1.0 : 0.0;
  }
}

__global__ void __launch_bounds__(256) jacobi_round_kernel(double* H, double* W, i64 np, i64 ld, int round, double tol,
                                                          int* rotated) {
  __shared__ double red[8];
  __shared__ double cs[2];
  // circle method: player np-1 stays, the others rotate; pair i of this round
  const int m = (int)np - 1;
  const int i = blockIdx.x;
  int p, q;
  if (i == 0) {
    p = m;
    q = round % m;
  } else {
    p = (round + i) % m;
    q = (round - i + m) % m;
  }
  double* hp = H + (i64)p * ld;
  double* hq = H + (i64)q * ld;
  double a = 0.0, b = 0.0, g = 0.0;
  for (i64 j = threadIdx.x; j < np; j += 256) {
    const double x = hp[j], y = hq[j];
    a = fma(x, x, a);
    b = fma(y, y, b);
    g = fma(x, y, g);
  }
  a = block_sum<256>(a, red);
  b = block_sum<256>(b, red);
  g = block_sum<256>(g, red);
  if (threadIdx.x == 0) {
    double c = 1.0, s = 0.0;
    if (fabs(g) > tol * sqrt(a * b) && fabs(g) > 0.0) {
      const double zeta = (b - a) / (2.0 * g);
      const double t = (zeta >= 0.0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
      c = 1.0 / sqrt(1.0 + t * t);
      s = c * t;
      atomicAdd(rotated, 1);
    }
    cs[0] = c;
    cs[1] = s;
  }
  __syncthreads();
  const double c = cs[0], s = cs[1];
  if (s == 0.0) return;
  double* wp = W + (i64)p * ld;
  double* wq = W + (i64)q * ld;
  for (i64 j = threadIdx.x; j < np; j += 256) {
    const double x = hp[j], y = hq[j];
    hp[j] = c * x - s * y;
    hq[j] = s * x + c * y;
    const double u = wp[j], v = wq[j];
    wp[j] = c * u - s * v;
    wq[j] = s * u + c * v;
  }
}

// lambda_i = H_i . W_i
__global__ void __launch_bounds__(256) jacobi_eigs_kernel(const double* __restrict__ H, const double* __restrict__ W,
                                                         i64 np, i64 ld, double* lam) {
  __shared__ double red[8];
  const double* h = H + (i64)blockIdx.x * ld;
  const double* w = W + (i64)blockIdx.x * ld;
  double s = 0.0;
  for (i64 j = threadIdx.x; j < np; j += 256) s = fma(h[j], w[j], s);
  s = block_sum<256>(s, red);
  if (threadIdx.x == 0) lam[blockIdx.x] = s;
}

}  // namespace stpyb

using namespace stpyb;

extern "C" int stpyb_jacobi_init(const double* A, long long lda, double* H, double* W, long long n, long long np,
                                 long long ld, void* stream) {
  if (n <= 0 || np < n || (np & 1) || ld < np) return -5;
  jacobi_init_kernel<<<(unsigned)np, 256, 0, (cudaStream_t)stream>>>(A, lda, H, W, n, np, ld);
  STPYB_COUNT_LAUNCH();
  STPYB_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int stpyb_jacobi_sweep(double* H, double* W, long long np, long long ld, double tol, int* rotated_dev,
                                  void* stream) {
  if (np < 2 || (np & 1) || ld < np) return -3;
  cudaStream_t st = (cudaStream_t)stream;
  STPYB_CUDA(cudaMemsetAsync(rotated_dev, 0, sizeof(int), st));
  for (int r = 0; r < (int)np - 1; ++r) {
    jacobi_round_kernel<<<(unsigned)(np / 2), 256, 0, st>>>(H, W, np, ld, r, tol, rotated_dev);
    STPYB_COUNT_LAUNCH();
  }
  STPYB_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int stpyb_jacobi_eigenvalues(const double* H, const double* W, long long np, long long ld, double* lam,
                                        void* stream) {
  if (np <= 0) return -3;
  jacobi_eigs_kernel<<<(unsigned)np, 256, 0, (cudaStream_t)stream>>>(H, W, np, ld, lam);
  STPYB_COUNT_LAUNCH();
  STPYB_CUDA(cudaGetLastError());
  return 0;
}
