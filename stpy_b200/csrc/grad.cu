// Analytic gradient of the log marginal likelihood for squared-exponential-type
// kernels (isotropic SE and ARD-SE), replacing autograd's backward pass through
// torch.linalg.solve / slogdet / exp / mm (gauss_procc.py:631-638 + .backward(),
// estimator.py:32-40), i.e. another O(n^3) of LU solves on the CPU.
//
// With Wm = w K^-1 - alpha alpha^T :  dLML/dtheta = 0.5 tr(Wm dK/dtheta).
//   K^-1 = U U^T,  U = L^-T            two DMMA passes of n^3/3 flops each
//   dK_ij/dgamma_k = K_ij (x_ik - x_jk)^2 / gamma_k^3
// so one fused pass over the lower triangle of (Wm o K) yields every
// lengthscale derivative, d/dkappa and d/ds at once.
#include "gemm_nt.cuh"
#include "stpyb_internal.h"
#include "../../include/stpyb.h"

namespace stpyb {

__global__ void set_identity_kernel(double* W, i64 n, i64 ldw) {
  const i64 idx = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n * ldw) return;
  const i64 r = idx / ldw, c = idx - r * ldw;
  W[idx] = (r == c) ? 1.0 : 0.0;
}

int potri(const double* L, i64 n, i64 ld, const double* dinv, double* work, i64 ldw, double* Kinv, i64 ldki,
          cudaStream_t st) {
  if (n <= 0) return 0;
  if ((ld & 1) || (ldw & 1) || (ldki & 1) || ldw < n) return -6;
  const i64 total = n * ldw;
  set_identity_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(work, n, ldw);
  STPYB_COUNT_LAUNCH();
  STPYB_CUDA(cudaGetLastError());
  const i64 nblk = (n + DB - 1) / DB;
  // U = I L^-T, touching only the rows that are already non-zero (U is upper triangular)
  for (i64 k = 0; k < nblk; ++k) {
    const int b = (int)((n - k * DB < DB) ? (n - k * DB) : DB);
    const i64 rows = ((k + 1) * DB < n) ? (k + 1) * DB : n;
    double* Bk = work + k * DB;
    STPYB_TRY(gemm_nt((int)rows, b, b, Bk, ldw, dinv + k * (i64)(DB * DB), DB, Bk, ldw, 1.0, 0.0, TRI_FULL, 1, st));
    const i64 c0 = (k + 1) * DB;
    if (c0 < n) {
      STPYB_TRY(gemm_nt((int)rows, (int)(n - c0), b, Bk, ldw, L + c0 * ld + k * DB, ld, work + c0, ldw, -1.0, 1.0,
                        TRI_FULL, 0, st));
    }
  }
  // K^-1 = U U^T (lower tiles), K loop starting at the tile's first row
  return gemm_nt((int)n, (int)n, (int)n, work, ldw, work, ldw, Kinv, ldki, 1.0, 0.0, TRI_LOWER, 0, st, 1);
}

constexpr int GT = 64;  // tile edge of the gradient pass

// One CTA per lower 64x64 tile, 256 threads, 4x4 elements per thread.
__global__ void __launch_bounds__(256) lml_grad_se_kernel(const double* __restrict__ Kinv, i64 ldki,
                                                         const double* __restrict__ alpha,
                                                         const double* __restrict__ Xp,
                                                         const double* __restrict__ norms, i64 n, int dpad, int dg,
                                                         double arg_scale, double kappa, double weight,
                                                         double* out) {
  extern __shared__ double sm[];  // xi[GT][dpad], xj[GT][dpad], red[8]
  double* xi = sm;
  double* xj = sm + GT * dpad;
  double* red = xj + GT * dpad;
  // decode lower-triangular tile index
  const i64 bid = blockIdx.x;
  i64 ti = (i64)((sqrt(8.0 * (double)bid + 1.0) - 1.0) * 0.5);
  while ((ti + 1) * (ti + 2) / 2 <= bid) ++ti;
  while (ti * (ti + 1) / 2 > bid) --ti;
  const i64 tj = bid - ti * (ti + 1) / 2;
  const i64 i0 = ti * GT, j0 = tj * GT;
  for (int idx = threadIdx.x; idx < GT * dpad; idx += 256) {
    const int r = idx / dpad, c = idx - r * dpad;
    xi[idx] = (i0 + r < n) ? Xp[(i0 + r) * dpad + c] : 0.0;
    xj[idx] = (j0 + r < n) ? Xp[(j0 + r) * dpad + c] : 0.0;
  }
  __syncthreads();
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  double gk = 0.0, gd = 0.0;
  // per-dimension partial sums live in shared memory-free registers only up to 16 dims at a
  // time; loop over dimension chunks to bound register use
  for (int kc = 0; kc < dg; kc += 8) {
    double g[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) g[u] = 0.0;
    for (int a = 0; a < 4; ++a) {
      const i64 i = i0 + ty * 4 + a;
      if (i >= n) continue;
      const double ai = alpha[i], ni = norms[i];
      for (int b = 0; b < 4; ++b) {
        const i64 j = j0 + tx * 4 + b;
        if (j >= n || j > i) continue;
        const double* pi = xi + (ty * 4 + a) * dpad;
        const double* pj = xj + (tx * 4 + b) * dpad;
        double dot = 0.0;
        for (int k = 0; k < dpad; ++k) dot = fma(pi[k], pj[k], dot);
        const double sq = (-2.0 * dot + norms[j]) + ni;
        const double kij = kappa * exp(arg_scale * sq);
        const double wm = weight * Kinv[i * ldki + j] - ai * alpha[j];
        const double f = (i == j ? 0.5 : 1.0) * wm;  // 0.5 * (2 for the mirrored entry)
        const double fk = f * kij;
        if (kc == 0) {
          gk += fk;
          if (i == j) gd += f;
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          if (kc + u < dg) {
            const double df = pi[kc + u] - pj[kc + u];
            g[u] = fma(fk, df * df, g[u]);
          }
        }
      }
    }
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      if (kc + u < dg) {
        const double t = block_sum<256>(g[u], red);
        if (threadIdx.x == 0) atomicAdd(out + kc + u, t);
      }
    }
  }
  {
    const double t1 = block_sum<256>(gk, red);
    const double t2 = block_sum<256>(gd, red);
    if (threadIdx.x == 0) {
      atomicAdd(out + dg, t1 / kappa);
      atomicAdd(out + dg + 1, t2);
    }
  }
}

}  // namespace stpyb

using namespace stpyb;

extern "C" int stpyb_potri(const double* L, long long n, long long ld, const double* dinv, double* work,
                           long long ldw, double* Kinv, long long ldki, void* stream) {
  return potri(L, n, ld, dinv, work, ldw, Kinv, ldki, (cudaStream_t)stream);
}

extern "C" int stpyb_lml_grad_se(const double* Kinv, long long ldki, const double* alpha, const double* Xp,
                                 const double* norms, long long n, int dpad, int dg, double arg_scale,
                                 double kappa, double weight, double* out, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  if (n <= 0) return -6;
  if (dg <= 0 || dg > dpad || dpad > STPYB_MAX_DIM) return -8;
  STPYB_CUDA(cudaMemsetAsync(out, 0, (size_t)(dg + 2) * sizeof(double), st));
  const long long T = (n + GT - 1) / GT;
  const long long tiles = T * (T + 1) / 2;
  if (tiles > 2147483647LL) return -6;
  const size_t smem = (size_t)(2 * GT * dpad + 8) * sizeof(double);
  static bool configured[64] = {false};
  if (first_use_on_device(configured)) {
    STPYB_CUDA(cudaFuncSetAttribute(lml_grad_se_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 80 * 1024));
  }
  lml_grad_se_kernel<<<(unsigned)tiles, 256, smem, st>>>(Kinv, ldki, alpha, Xp, norms, n, dpad, dg, arg_scale,
                                                        kappa, weight, out);
  STPYB_COUNT_LAUNCH();
  STPYB_CUDA(cudaGetLastError());
  return 0;
}
