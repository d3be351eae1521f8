// Operations on a STACK of Gram matrices of one dataset (k kernels, matrix q at stack + q * stride):
// the shape of the reference's multiple-kernel learner, which builds one Gram per kernel in a Python loop
// (stpy/continuous_processes/mkl_estimator.py:35-37), combines them with the learned weights
//   K = sum_q alpha_q K_q + lam s^2 I                                   (mkl_estimator.py:90)
// and whose weight objective y^T K(alpha)^-1 y (mkl_estimator.py:60-64, a cvxpy matrix_frac) has the gradient
//   d/d alpha_q = - beta^T K_q beta,  beta = K(alpha)^-1 y.
// Both passes are HBM-bound streams over the stack (k * 4 n^2 bytes with lower-triangular storage).
#include "common.cuh"
#include "stpyb_internal.h"
#include "../../include/stpyb.h"

namespace stpyb {

struct StackWeights {
  double w[64];
  int k;
};

// out[i][j] = sum_q w_q K_q[i][j] (+ diag_add on i == j); one CTA per (256-column strip, row);
// lower = 1 touches only j <= i (what the factorisation reads)
__global__ void __launch_bounds__(128) stack_combine_kernel(const double* __restrict__ stack, i64 n, i64 ld, i64 stride,
                                                           StackWeights sw, double diag_add, int lower,
                                                           double* __restrict__ out, i64 ldo) {
  const i64 i = blockIdx.y;
  const i64 j = ((i64)blockIdx.x * 128 + threadIdx.x) * 2;
  if (j >= n || (lower && j > i)) return;
  const bool two = (j + 1 < n);
  const double* p = stack + i * ld + j;
  double s0 = 0.0, s1 = 0.0;
  for (int q = 0; q < sw.k; ++q) {
    if (two) {
      const double2 v = *reinterpret_cast<const double2*>(p + (i64)q * stride);
      s0 = fma(sw.w[q], v.x, s0);
      s1 = fma(sw.w[q], v.y, s1);
    } else {
      s0 = fma(sw.w[q], p[(i64)q * stride], s0);
    }
  }
  if (i == j) s0 += diag_add;
  if (i == j + 1) s1 += diag_add;
  double* o = out + i * ldo + j;
  if (two) *reinterpret_cast<double2*>(o) = make_double2(s0, s1);
  else o[0] = s0;
}

// out[q] += sum over this CTA's rows i of beta_i * sum_j K_q[i][j] beta_j, the symmetric matrix read from its
// lower triangle (j < i counted twice); grid (row blocks of 4, k)
__global__ void __launch_bounds__(256) stack_quadform_kernel(const double* __restrict__ stack, i64 n, i64 ld,
                                                            i64 stride, int lower, const double* __restrict__ beta,
                                                            double* out) {
  __shared__ double red[8];
  const double* K = stack + (i64)blockIdx.y * stride;
  double tot = 0.0;
  for (int rr = 0; rr < 4; ++rr) {
    const i64 i = (i64)blockIdx.x * 4 + rr;
    if (i >= n) break;
    const double* row = K + i * ld;
    const i64 jend = lower ? i + 1 : n;
    double s = 0.0;
    for (i64 j = threadIdx.x; j < jend; j += 256) {
      const double f = (lower && j < i) ? 2.0 : 1.0;
      s = fma(f * row[j], beta[j], s);
    }
    tot = fma(beta[i], s, tot);
  }
  tot = block_sum<256>(tot, red);
  if (threadIdx.x == 0) atomicAdd(out + blockIdx.y, tot);
}

}  // namespace stpyb

using namespace stpyb;

extern "C" int stpyb_stack_combine(const double* stack, int k, const double* weights_host, long long n, long long ld,
                                   long long stride, double diag_add, int lower, double* out, long long ldo,
                                   void* stream) {
  if (k <= 0 || k > 64) return -2;
  if (n <= 0) return 0;
  if ((ld & 1) || (ldo & 1) || (stride & 1) || (((uintptr_t)stack) & 15) || (((uintptr_t)out) & 15)) return -5;
  if (n > 65535) return -4;  // one grid row per matrix row (gridDim.y limit); a stack of such matrices would not fit in HBM anyway
  StackWeights sw;
  sw.k = k;
  for (int q = 0; q < k; ++q) sw.w[q] = weights_host[q];
  dim3 grid((unsigned)((n + 255) / 256), (unsigned)n);
  stack_combine_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(stack, n, ld, stride, sw, diag_add, lower, out, ldo);
  STPYB_COUNT_LAUNCH();
  STPYB_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int stpyb_stack_quadform(const double* stack, int k, long long n, long long ld, long long stride, int lower,
                                    const double* beta, double* out_k, void* stream) {
  if (k <= 0 || k > 65535) return -2;
  if (n <= 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  STPYB_CUDA(cudaMemsetAsync(out_k, 0, (size_t)k * sizeof(double), st));
  dim3 grid((unsigned)((n + 3) / 4), (unsigned)k);
  stack_quadform_kernel<<<grid, 256, 0, st>>>(stack, n, ld, stride, lower, beta, out_k);
  STPYB_COUNT_LAUNCH();
  STPYB_CUDA(cudaGetLastError());
  return 0;
}
