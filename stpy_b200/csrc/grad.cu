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
#include <cstring>
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
  // U = I L^-T (upper triangular), column panel by column panel (left-looking over panels of 1024 columns, as the
  // factorisation): panel [J0, J1) first receives the contribution of ALL earlier columns in one contraction of
  // depth J0 -- U[0:J0, J0:J1] = -U[0:J0, 0:J0] L[J0:J1, 0:J0]^T, the K loop of a tile starting at its first row
  // because U is upper triangular -- and is then solved against the 1024 x 1024 diagonal block of L in 128-column
  // steps that touch only the panel.  (The first version applied every 128-block to all remaining columns: n/128
  // rank-128 updates that re-read and re-wrote the whole right part of U, 25 TFLOP/s; this form runs the bulk at
  // the depth of the trailing update of the factorisation.)
  const i64 OUTER = 1024;
  for (i64 J0 = 0; J0 < n; J0 += OUTER) {
    const i64 J1 = (J0 + OUTER < n) ? J0 + OUTER : n;
    if (J0 > 0) {
      STPYB_TRY(gemm_nt((int)J0, (int)(J1 - J0), (int)J0, work, ldw, L + J0 * ld, ld, work + J0, ldw, -1.0, 1.0, TRI_FULL,
                        0, st, 1));
    }
    for (i64 c = J0; c < J1; c += DB) {
      const i64 k = c / DB;
      const int b = (int)((n - c < DB) ? (n - c) : DB);
      const i64 rows = (c + DB < n) ? c + DB : n;  // rows of U that are non-zero in this block column
      double* Bk = work + c;
      STPYB_TRY(gemm_nt((int)rows, b, b, Bk, ldw, dinv + k * (i64)(DB * DB), DB, Bk, ldw, 1.0, 0.0, TRI_FULL, 1, st));
      const i64 c0 = c + DB;
      if (c0 < J1) {
        STPYB_TRY(gemm_nt((int)rows, (int)(J1 - c0), b, Bk, ldw, L + c0 * ld + c, ld, work + c0, ldw, -1.0, 1.0,
                          TRI_FULL, 0, st));
      }
    }
  }
  // K^-1 = U U^T (lower tiles), K loop starting at the tile's first row
  return gemm_nt((int)n, (int)n, (int)n, work, ldw, work, ldw, Kinv, ldki, 1.0, 0.0, TRI_LOWER, 0, st, 1);
}

constexpr int GT = 64;  // tile edge of the gradient pass

// ---- derivative pass over a composite kernel ---------------------------------------------------
// K is the left fold of sub-kernel Grams, out_0 = G_0, out_p = out_{p-1} (+|*) G_p
// (stpy/kernels.py:146-157), each G_p a sum of ITEMS (one item per additive group,
// kernels.py:700-729; a plain kernel is one item).  An item is kappa * f(sq) with
// sq = sum_c ((x_ic - x_jc) sc_c)^2 over its columns (SE, Matern) or kappa * f(<x_i, x_j>) (+ offset)
// (polynomial, linear).  For a parameter of item q in sub-kernel p
//   dK_ij = [dout_last / dG_p]_ij * dG_q,ij ,   dout_last/dG_p = pre_p * suf_p,
//   pre_p = (p == 0 or op_p is '+') ? 1 : out_{p-1},   suf_p = prod_{r > p, op_r is '*'} G_r
// (the product rule: the partner Grams are RE-EVALUATED in registers, never read from memory).
// One launch handles one item and up to 16 of its columns and returns
//   out[c]  = sum_ij W_ij pre suf kappa f'(sq_ij) u_c^2      u_c = (x_ic - x_jc) sc_c
//             -> d/d(lengthscale of column c) = out[c] * (-2 / lengthscale)   for sc_c = 1/lengthscale
//   out[16] = sum_ij W_ij pre suf f(sq_ij)                    -> d/d(kappa of the item)
//   out[17] = sum_i  W_ii / 2                                  -> d/ds = 2 s out[17] (mode 0)
// mode 0: W = weight K^-1 - alpha alpha^T over the lower triangle of a symmetric problem (mirror
//         counted, diagonal halved): the evidence gradient 0.5 tr(W dK)  (gauss_procc.py:631-638 + autograd);
// mode 1: W = an explicit m x n matrix (the incoming gradient of KernelFunction.kernel(a, b): rows are
//         b-points, columns a-points) -- the backward of the operator seam (kernels.py:136-159).
constexpr int GI_MAX = 8;    // items of a composite kernel
constexpr int GC_MAX = 32;   // columns per item

struct GradItem {
  int kind, ncols, sub, pad;
  double arg_scale, kappa, p0;
};
struct GradDesc {
  GradItem it[GI_MAX];
  int cols[GI_MAX][GC_MAX];
  double sc[GI_MAX][GC_MAX];
  int sub_op[GI_MAX];
  int nitems, nsub;
};

// value v (without kappa) and h = dv/dsq of one item
__device__ __forceinline__ void item_value(int kind, double s, double arg_scale, double p0, double& v, double& h) {
  if (kind == STPYB_K_SE) {
    v = exp(arg_scale * s);
    h = arg_scale * v;
  } else if (kind == STPYB_K_MATERN12) {
    const double r = sqrt(s);
    v = exp(-r);
    h = (r > 0.0) ? -0.5 * v / r : 0.0;
  } else if (kind == STPYB_K_MATERN32) {
    const double t = sqrt(s) * 1.7320508075688772;
    const double e = exp(-t);
    v = (1.0 + t) * e;
    h = -1.5 * e;
  } else if (kind == STPYB_K_MATERN52) {
    const double t = sqrt(s) * 2.23606797749979;
    const double e = exp(-t);
    v = (1.0 + t + t * t * 0.3333333333333333) * e;
    h = -0.8333333333333334 * (1.0 + t) * e;
  } else if (kind == STPYB_K_POLY) {
    const double t = s + 1.0;
    v = (p0 == 2.0) ? t * t : (p0 == 3.0) ? t * t * t : (p0 == 1.0) ? t : pow(t, p0);
    h = 0.0;
  } else {
    v = s;
    h = 0.0;
  }
}

__global__ void __launch_bounds__(256) kernel_grad_kernel(GradDesc ds, const double* __restrict__ XR, i64 ldxr, i64 m,
                                                         const double* __restrict__ XC, i64 ldxc, i64 n, int d,
                                                         int mode, const double* __restrict__ Cmat, i64 ldc,
                                                         const double* __restrict__ alpha, double weight,
                                                         int pass_item, int col_off, i64 ntiles, i64 tiles_n,
                                                         double* out) {
  extern __shared__ double sm[];  // xi[GT][dS], xj[GT][dS], red[8]
  const int dS = d | 1;           // odd row stride: the 16 column groups of a warp hit distinct banks
  double* xi = sm;
  double* xj = sm + GT * dS;
  double* red = xj + GT * dS;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const GradItem pit = ds.it[pass_item];
  const int ncp = (pit.ncols - col_off < 16) ? (pit.ncols - col_off) : 16;  // columns of this pass
  double g[16];
#pragma unroll
  for (int u = 0; u < 16; ++u) g[u] = 0.0;
  double gk = 0.0, gd = 0.0;

  for (i64 tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    i64 ti, tj;
    if (mode == 0) {
      ti = (i64)((sqrt(8.0 * (double)tile + 1.0) - 1.0) * 0.5);
      while ((ti + 1) * (ti + 2) / 2 <= tile) ++ti;
      while (ti * (ti + 1) / 2 > tile) --ti;
      tj = tile - ti * (ti + 1) / 2;
    } else {
      ti = tile / tiles_n;
      tj = tile - ti * tiles_n;
    }
    const i64 i0 = ti * GT, j0 = tj * GT;
    __syncthreads();
    for (int idx = threadIdx.x; idx < GT * d; idx += 256) {
      const int r = idx / d, c = idx - r * d;
      xi[r * dS + c] = (i0 + r < m) ? XR[(i0 + r) * ldxr + c] : 0.0;
      xj[r * dS + c] = (j0 + r < n) ? XC[(j0 + r) * ldxc + c] : 0.0;
    }
    __syncthreads();
    for (int a = 0; a < 4; ++a) {
      const i64 i = i0 + ty * 4 + a;
      if (i >= m) continue;
      const double* pi = xi + (ty * 4 + a) * dS;
      for (int b = 0; b < 4; ++b) {
        const i64 j = j0 + tx * 4 + b;
        if (j >= n || (mode == 0 && j > i)) continue;
        const double* pj = xj + (tx * 4 + b) * dS;
        double w;
        if (mode == 0) {
          w = weight * Cmat[i * ldc + j] - alpha[i] * alpha[j];
          if (i == j) {
            w *= 0.5;  // 0.5 tr(W dK): off-diagonal entries stand for themselves and their mirror
            gd += w;
          }
        } else {
          w = Cmat[i * ldc + j];
        }
        // left fold with the derivative bookkeeping for the pass item's sub-kernel
        double outv = 0.0, gsub = 0.0, pre = 1.0, suf = 1.0, v_it = 0.0, h_it = 0.0;
        int cur = 0;
        for (int q = 0; q <= ds.nitems; ++q) {
          const int sub = (q < ds.nitems) ? ds.it[q].sub : -1;
          if (sub != cur) {  // sub-kernel `cur` is complete: fold it
            const int op = ds.sub_op[cur];
            if (cur == pit.sub) pre = (cur > 0 && op == STPYB_OP_MUL) ? outv : 1.0;
            else if (cur > pit.sub && op == STPYB_OP_MUL) suf *= gsub;
            outv = (cur == 0) ? gsub : (op == STPYB_OP_MUL ? outv * gsub : outv + gsub);
            gsub = 0.0;
            cur = sub;
            if (q == ds.nitems) break;
          }
          const GradItem it = ds.it[q];
          double s = 0.0;
          if (it.kind <= STPYB_K_MATERN52) {
            for (int c = 0; c < it.ncols; ++c) {
              const int col = ds.cols[q][c];
              const double u = (pi[col] - pj[col]) * ds.sc[q][c];
              s = fma(u, u, s);
            }
          } else {
            for (int c = 0; c < it.ncols; ++c) {
              const int col = ds.cols[q][c];
              s = fma(pi[col], pj[col], s);
            }
          }
          double v, h;
          item_value(it.kind, s, it.arg_scale, it.p0, v, h);
          gsub += it.kappa * v + (it.kind == STPYB_K_LINEAR ? it.p0 : 0.0);
          if (q == pass_item) {
            v_it = v;
            h_it = it.kappa * h;
          }
        }
        const double cw = w * pre * suf;
        gk = fma(cw, v_it, gk);
        const double ch = cw * h_it;
#pragma unroll
        for (int u = 0; u < 16; ++u) {
          if (u < ncp) {
            const int col = ds.cols[pass_item][col_off + u];
            const double uu = (pi[col] - pj[col]) * ds.sc[pass_item][col_off + u];
            g[u] = fma(ch, uu * uu, g[u]);
          }
        }
      }
    }
  }
#pragma unroll
  for (int u = 0; u < 16; ++u) {
    if (u < ncp) {
      const double t = block_sum<256>(g[u], red);
      if (threadIdx.x == 0) atomicAdd(out + u, t);
    }
  }
  const double t1 = block_sum<256>(gk, red);
  const double t2 = block_sum<256>(gd, red);
  if (threadIdx.x == 0) {
    atomicAdd(out + 16, t1);
    atomicAdd(out + 17, t2);
  }
}

}  // namespace stpyb

using namespace stpyb;

extern "C" int stpyb_potri(const double* L, long long n, long long ld, const double* dinv, double* work,
                           long long ldw, double* Kinv, long long ldki, void* stream) {
  return potri(L, n, ld, dinv, work, ldw, Kinv, ldki, (cudaStream_t)stream);
}

extern "C" int stpyb_kernel_grad(const double* XR, long long m, long long ldxr, const double* XC, long long n,
                                 long long ldxc, int d, int nitems, const int* kinds, const int* ncols,
                                 const int* subs, const int* cols_flat32, const double* sc_flat32,
                                 const double* arg_scales, const double* kappas, const double* p0s, int nsub,
                                 const int* sub_ops, int pass_item, int col_off, int mode, const double* Cmat,
                                 long long ldc, const double* alpha, double weight, double* out18, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  if (m <= 0 || n <= 0) return -2;
  if (d <= 0 || d > STPYB_MAX_DIM) return -7;
  if (nitems <= 0 || nitems > GI_MAX) return -8;
  if (nsub <= 0 || nsub > GI_MAX) return -17;
  if (pass_item < 0 || pass_item >= nitems) return -19;
  if (mode != 0 && mode != 1) return -21;
  if (mode == 0 && (m != n || alpha == nullptr)) return -24;
  GradDesc ds;
  memset(&ds, 0, sizeof(ds));
  ds.nitems = nitems;
  ds.nsub = nsub;
  int prev = 0;
  for (int q = 0; q < nitems; ++q) {
    if (kinds[q] < 0 || kinds[q] >= STPYB_K_COUNT) return -9;
    if (ncols[q] <= 0 || ncols[q] > GC_MAX) return -10;
    if (subs[q] < prev || subs[q] > prev + 1 || subs[q] >= nsub) return -11;  // items sorted by sub-kernel, no gaps
    prev = subs[q];
    ds.it[q].kind = kinds[q];
    ds.it[q].ncols = ncols[q];
    ds.it[q].sub = subs[q];
    ds.it[q].arg_scale = arg_scales[q];
    ds.it[q].kappa = kappas[q];
    ds.it[q].p0 = p0s[q];
    for (int c = 0; c < ncols[q]; ++c) {
      const int col = cols_flat32[q * GC_MAX + c];
      if (col < 0 || col >= d) return -12;
      ds.cols[q][c] = col;
      ds.sc[q][c] = sc_flat32[q * GC_MAX + c];
    }
  }
  if (subs[0] != 0 || prev != nsub - 1) return -11;
  for (int p = 0; p < nsub; ++p) ds.sub_op[p] = sub_ops[p];
  if (col_off < 0 || col_off >= ncols[pass_item]) return -20;
  STPYB_CUDA(cudaMemsetAsync(out18, 0, 18 * sizeof(double), st));
  const long long Tm = (m + GT - 1) / GT, Tn = (n + GT - 1) / GT;
  const long long ntiles = (mode == 0) ? Tm * (Tm + 1) / 2 : Tm * Tn;
  const int dS = d | 1;
  const size_t smem = (size_t)(2 * GT * dS + 8) * sizeof(double);
  static bool configured[64] = {false};
  if (first_use_on_device(configured)) {
    STPYB_CUDA(cudaFuncSetAttribute(kernel_grad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 80 * 1024));
  }
  int dev = 0, sms = 148;
  if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  long long grid = ntiles < 4LL * sms ? ntiles : 4LL * sms;  // persistent: every CTA walks tiles grid-stride
  kernel_grad_kernel<<<(unsigned)grid, 256, smem, st>>>(ds, XR, ldxr, m, XC, ldxc, n, d, mode, Cmat, ldc, alpha, weight,
                                                       pass_item, col_off, ntiles, Tn, out18);
  STPYB_COUNT_LAUNCH();
  STPYB_CUDA(cudaGetLastError());
  return 0;
}
