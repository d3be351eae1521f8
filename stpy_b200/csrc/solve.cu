// Triangular solves and reductions around the Cholesky factor.
//
// Replaces, on the GP path of the reference:
//   alpha = K^-1 y                      gauss_procc.py:376 (lstsq), :635 (solve), estimator.py:37
//   B = K^-1 K*^T and diag(K* B^T)      gauss_procc.py:378, 391-395
//   logdet / quadratic form of the LML  gauss_procc.py:633-637, estimator.py:35-39
//
// Single right-hand sides are HBM-bound (every entry of L is read once): the
// solves walk the factor in 128-wide block columns, applying the inverted
// diagonal block kept by POTRF and then a streaming warp-shuffle GEMV update.
// Many right-hand sides (posterior variance) are stored as ROWS (the layout the
// Gram kernel produces for K*), so  V^T = K* L^-T  is a sequence of NT GEMMs.
#include "gemm_nt.cuh"
#include "stpyb_internal.h"

namespace stpyb {

// x_blk <- Linv * x_blk  (TRANS=0)  or  Linv^T * x_blk (TRANS=1); Linv dense [128][128], b valid rows.
template <int TRANS>
__global__ void __launch_bounds__(1024, 1) blk_gemv_kernel(const double* __restrict__ Linv, double* x, int b) {
  __shared__ __align__(16) double xs[DB];
  __shared__ double part[8][DB];
  const int tid = threadIdx.x;
  if (tid < DB) xs[tid] = (tid < b) ? x[tid] : 0.0;
  __syncthreads();
  if (!TRANS) {
    const int warp = tid >> 5, lane = tid & 31;
    const double2 x0 = *reinterpret_cast<const double2*>(&xs[2 * lane]);
    const double2 x1 = *reinterpret_cast<const double2*>(&xs[64 + 2 * lane]);
#pragma unroll
    for (int rr = 0; rr < 4; ++rr) {
      const int i = warp * 4 + rr;
      const double2 l0 = *reinterpret_cast<const double2*>(Linv + i * DB + 2 * lane);
      const double2 l1 = *reinterpret_cast<const double2*>(Linv + i * DB + 64 + 2 * lane);
      double s = l0.x * x0.x + l0.y * x0.y + l1.x * x1.x + l1.y * x1.y;
      s = warp_sum(s);
      if (lane == 0 && i < b) x[i] = s;
    }
  } else {
    const int rg = tid >> 7, j = tid & 127;
    double s = 0.0;
#pragma unroll 4
    for (int i = rg; i < DB; i += 8) s = fma(Linv[i * DB + j], xs[i], s);
    part[rg][j] = s;
    __syncthreads();
    if (tid < DB && tid < b) {
      double t = 0.0;
#pragma unroll
      for (int r = 0; r < 8; ++r) t += part[r][tid];
      x[tid] = t;
    }
  }
}

// x[r] -= sum_c L[r][c0+c] * x[c0+c]  for rows r in [r0, n); one warp per row, 4 rows in flight.
__global__ void __launch_bounds__(256) trsv_fwd_update_kernel(const double* __restrict__ L, i64 ld, i64 n,
                                                             i64 r0, i64 c0, double* x) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const double2 x0 = *reinterpret_cast<const double2*>(x + c0 + 2 * lane);
  const double2 x1 = *reinterpret_cast<const double2*>(x + c0 + 64 + 2 * lane);
  const i64 base = r0 + ((i64)blockIdx.x * 8 + warp) * 4;
  double s[4];
#pragma unroll
  for (int rr = 0; rr < 4; ++rr) {
    const i64 r = base + rr;
    s[rr] = 0.0;
    if (r < n) {
      const double* row = L + r * ld + c0;
      const double2 l0 = *reinterpret_cast<const double2*>(row + 2 * lane);
      const double2 l1 = *reinterpret_cast<const double2*>(row + 64 + 2 * lane);
      s[rr] = l0.x * x0.x + l0.y * x0.y + l1.x * x1.x + l1.y * x1.y;
    }
  }
#pragma unroll
  for (int rr = 0; rr < 4; ++rr) s[rr] = warp_sum(s[rr]);
  if (lane < 4) {
    const i64 r = base + lane;
    const double mine = (lane == 0) ? s[0] : (lane == 1) ? s[1] : (lane == 2) ? s[2] : s[3];
    if (r < n) x[r] -= mine;
  }
}

// x[c] -= sum_r L[r0+r][c] * x[r0+r]  for columns c in [0, cend): the block row r0.. of L is streamed
// once.  Each thread owns two adjacent columns (16-byte loads) and a quarter of the rows; the four
// row groups of a CTA are combined through shared memory, so every column sees 4x more loads in
// flight than a one-thread-per-column sweep.
__global__ void __launch_bounds__(256) trsv_bwd_update_kernel(const double* __restrict__ L, i64 ld, i64 r0,
                                                             int bw, i64 cend, double* x) {
  __shared__ double xs[DB];
  __shared__ double2 part[4][64];
  if (threadIdx.x < DB) xs[threadIdx.x] = (threadIdx.x < bw) ? x[r0 + threadIdx.x] : 0.0;
  __syncthreads();
  const int cg = threadIdx.x & 63, rg = threadIdx.x >> 6;  // 64 column pairs x 4 row groups
  const i64 c = ((i64)blockIdx.x * 64 + cg) * 2;            // cend is a multiple of 128
  double2 s0 = make_double2(0.0, 0.0), s1 = s0;
  if (c < cend) {
    const double* p = L + r0 * ld + c;
    int r = rg;
    for (; r + 4 < bw; r += 8) {
      const double2 a = *reinterpret_cast<const double2*>(p + (i64)r * ld);
      const double2 b = *reinterpret_cast<const double2*>(p + (i64)(r + 4) * ld);
      s0.x = fma(a.x, xs[r], s0.x);
      s0.y = fma(a.y, xs[r], s0.y);
      s1.x = fma(b.x, xs[r + 4], s1.x);
      s1.y = fma(b.y, xs[r + 4], s1.y);
    }
    if (r < bw) {
      const double2 a = *reinterpret_cast<const double2*>(p + (i64)r * ld);
      s0.x = fma(a.x, xs[r], s0.x);
      s0.y = fma(a.y, xs[r], s0.y);
    }
  }
  part[rg][cg] = make_double2(s0.x + s1.x, s0.y + s1.y);
  __syncthreads();
  if (rg == 0 && c < cend) {
    double2 t = part[0][cg];
#pragma unroll
    for (int q = 1; q < 4; ++q) {
      t.x += part[q][cg].x;
      t.y += part[q][cg].y;
    }
    double2* px = reinterpret_cast<double2*>(x + c);
    double2 v = *px;
    v.x -= t.x;
    v.y -= t.y;
    *px = v;
  }
}

int trsv_lower(const double* L, i64 n, i64 ld, const double* dinv, double* x, int transposed, cudaStream_t st) {
  if (n <= 0) return 0;
  if ((ld & 1) || (((uintptr_t)L) & 15) || (((uintptr_t)x) & 15)) return -3;
  const i64 nblk = (n + DB - 1) / DB;
  if (!transposed) {
    for (i64 k = 0; k < nblk; ++k) {
      const int b = (int)((n - k * DB < DB) ? (n - k * DB) : DB);
      blk_gemv_kernel<0><<<1, 1024, 0, st>>>(dinv + k * (i64)(DB * DB), x + k * DB, b);
      STPYB_COUNT_LAUNCH();
      const i64 r0 = (k + 1) * DB;
      if (r0 < n) {
        const i64 rows = n - r0;
        trsv_fwd_update_kernel<<<(unsigned)((rows + 31) / 32), 256, 0, st>>>(L, ld, n, r0, k * DB, x);
        STPYB_COUNT_LAUNCH();
      }
    }
  } else {
    for (i64 k = nblk - 1; k >= 0; --k) {
      const int b = (int)((n - k * DB < DB) ? (n - k * DB) : DB);
      blk_gemv_kernel<1><<<1, 1024, 0, st>>>(dinv + k * (i64)(DB * DB), x + k * DB, b);
      STPYB_COUNT_LAUNCH();
      const i64 cend = k * DB;
      if (cend > 0) {
        trsv_bwd_update_kernel<<<(unsigned)((cend + 127) / 128), 256, 0, st>>>(L, ld, k * DB, b, cend, x);
        STPYB_COUNT_LAUNCH();
      }
    }
  }
  STPYB_CUDA(cudaGetLastError());
  return 0;
}

// out[0] = ||z||^2, out[1] = 2 sum log L_ii, out[2] = 0.5 out[0] + 0.5 weight out[1]
// (the reference's sign convention, no n/2 log 2pi term: gauss_procc.py:631-638).
__global__ void __launch_bounds__(1024, 1) lml_reduce_kernel(const double* __restrict__ L, i64 n, i64 ld,
                                                            const double* __restrict__ z, double weight,
                                                            double* out) {
  __shared__ double red[32];
  double q = 0.0, ld2 = 0.0;
  for (i64 i = threadIdx.x; i < n; i += 1024) {
    const double zi = z[i];
    q = fma(zi, zi, q);
    ld2 += log(L[i * ld + i]);
  }
  q = block_sum<1024>(q, red);
  ld2 = block_sum<1024>(ld2, red);
  if (threadIdx.x == 0) {
    out[0] = q;
    out[1] = 2.0 * ld2;
    out[2] = 0.5 * q + 0.5 * weight * 2.0 * ld2;
  }
}

// One CTA per row i of V[rows x cols]: s = sum_j V_ij^2; out_i = mode 0: s ; 1: sqrt(kss_i - s) ; 2: kss_i - s
__global__ void __launch_bounds__(256) row_sumsq_kernel(const double* __restrict__ V, i64 cols, i64 ldv,
                                                       const double* __restrict__ kss, int mode, double* out) {
  __shared__ double red[8];
  const double* row = V + (i64)blockIdx.x * ldv;
  double s = 0.0;
  for (i64 j = threadIdx.x; j < cols; j += 256) {
    const double v = row[j];
    s = fma(v, v, s);
  }
  s = block_sum<256>(s, red);
  if (threadIdx.x == 0) {
    double o = s;
    if (mode == 1) o = sqrt(kss[blockIdx.x] - s);
    else if (mode == 2) o = kss[blockIdx.x] - s;
    out[blockIdx.x] = o;
  }
}

// out_i = sum_j M_ij v_j ; one CTA per row (rows are long: K* is n_t x n).
__global__ void __launch_bounds__(256) gemv_rows_kernel(const double* __restrict__ M, i64 cols, i64 ldm,
                                                       const double* __restrict__ v, double* out) {
  __shared__ double red[8];
  const double* row = M + (i64)blockIdx.x * ldm;
  double s = 0.0;
  for (i64 j = threadIdx.x; j < cols; j += 256) s = fma(row[j], v[j], s);
  s = block_sum<256>(s, red);
  if (threadIdx.x == 0) out[blockIdx.x] = s;
}

// Bt <- Bt * L^-T ; Bt is [nt x n] (each ROW is one right-hand side).
int trsm_rt(const double* L, i64 n, i64 ld, const double* dinv, double* Bt, i64 nt, i64 ldbt, cudaStream_t st) {
  if (n <= 0 || nt <= 0) return 0;
  if ((ld & 1) || (ldbt & 1) || (((uintptr_t)L) & 15) || (((uintptr_t)Bt) & 15)) return -3;
  const i64 nblk = (n + DB - 1) / DB;
  for (i64 k = 0; k < nblk; ++k) {
    const int b = (int)((n - k * DB < DB) ? (n - k * DB) : DB);
    double* Bk = Bt + k * DB;
    STPYB_TRY(gemm_nt((int)nt, b, b, Bk, ldbt, dinv + k * (i64)(DB * DB), DB, Bk, ldbt, 1.0, 0.0, TRI_FULL, 1, st));
    const i64 c0 = (k + 1) * DB;
    if (c0 < n) {
      STPYB_TRY(gemm_nt((int)nt, (int)(n - c0), b, Bk, ldbt, L + c0 * ld + k * DB, ld, Bt + c0, ldbt, -1.0, 1.0,
                        TRI_FULL, 0, st));
    }
  }
  return 0;
}

}  // namespace stpyb

using namespace stpyb;

extern "C" int stpyb_trsv(const double* L, long long n, long long ld, const double* dinv, double* x,
                          int transposed, void* stream) {
  return trsv_lower(L, n, ld, dinv, x, transposed, (cudaStream_t)stream);
}

extern "C" int stpyb_potrs_vec(const double* L, long long n, long long ld, const double* dinv, double* x,
                               void* stream) {
  STPYB_TRY(trsv_lower(L, n, ld, dinv, x, 0, (cudaStream_t)stream));
  return trsv_lower(L, n, ld, dinv, x, 1, (cudaStream_t)stream);
}

extern "C" int stpyb_lml(const double* L, long long n, long long ld, const double* z, double weight,
                         double* out3, void* stream) {
  if (n <= 0) return -2;
  lml_reduce_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(L, n, ld, z, weight, out3);
  STPYB_COUNT_LAUNCH();
  STPYB_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int stpyb_trsm_rt(const double* L, long long n, long long ld, const double* dinv, double* Bt,
                             long long nt, long long ldbt, void* stream) {
  return trsm_rt(L, n, ld, dinv, Bt, nt, ldbt, (cudaStream_t)stream);
}

extern "C" int stpyb_row_sumsq(const double* V, long long rows, long long cols, long long ldv,
                               const double* kss_or_null, int mode, double* out, void* stream) {
  if (rows <= 0) return 0;
  if (mode != 0 && kss_or_null == nullptr) return -5;
  row_sumsq_kernel<<<(unsigned)rows, 256, 0, (cudaStream_t)stream>>>(V, cols, ldv, kss_or_null, mode, out);
  STPYB_COUNT_LAUNCH();
  STPYB_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int stpyb_gemv_rows(const double* M, long long rows, long long cols, long long ldm, const double* v,
                               double* out, void* stream) {
  if (rows <= 0) return 0;
  gemv_rows_kernel<<<(unsigned)rows, 256, 0, (cudaStream_t)stream>>>(M, cols, ldm, v, out);
  STPYB_COUNT_LAUNCH();
  STPYB_CUDA(cudaGetLastError());
  return 0;
}

namespace stpyb {
// y[c] -= sum_r A[r][c] * v[r]  for a tall panel A (rows x w): each CTA reduces 256 rows for all
// w (<= 1024) columns (threads run along the contiguous column index) and commits with one
// atomicAdd per column.  Used by the distributed backward solve (column-owned panels).
__global__ void __launch_bounds__(256) gemv_t_sub_kernel(const double* __restrict__ A, i64 rows, int w, i64 ld,
                                                        const double* __restrict__ v, double* y) {
  __shared__ double vs[256];
  const i64 r0 = (i64)blockIdx.x * 256;
  const int nr = (int)((rows - r0 < 256) ? (rows - r0) : 256);
  if (threadIdx.x < nr) vs[threadIdx.x] = v[r0 + threadIdx.x];
  __syncthreads();
  for (int c = threadIdx.x; c < w; c += 256) {
    const double* p = A + r0 * ld + c;
    double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
    int r = 0;
    for (; r + 4 <= nr; r += 4) {
      s0 = fma(p[(i64)(r + 0) * ld], vs[r + 0], s0);
      s1 = fma(p[(i64)(r + 1) * ld], vs[r + 1], s1);
      s2 = fma(p[(i64)(r + 2) * ld], vs[r + 2], s2);
      s3 = fma(p[(i64)(r + 3) * ld], vs[r + 3], s3);
    }
    for (; r < nr; ++r) s0 = fma(p[(i64)r * ld], vs[r], s0);
    atomicAdd(y + c, -((s0 + s1) + (s2 + s3)));
  }
}
}  // namespace stpyb

extern "C" int stpyb_gemv_t_sub(const double* A, long long rows, int w, long long ld, const double* v, double* y,
                                void* stream) {
  if (rows <= 0 || w <= 0) return 0;
  stpyb::gemv_t_sub_kernel<<<(unsigned)((rows + 255) / 256), 256, 0, (cudaStream_t)stream>>>(A, rows, w, ld, v, y);
  STPYB_COUNT_LAUNCH();
  STPYB_CUDA(cudaGetLastError());
  return 0;
}

// One hop of the distributed backward sweep alpha = L^-T z, executed by the owner of a block
// column: seg <- z_g ; seg -= L[below, g]^T alpha[below] ; seg <- L_gg^-T seg.  A single C call so
// that the host enqueues one hop with one FFI crossing (the sweep is latency-bound).
extern "C" int stpyb_dist_alpha_step(const double* Lcol, long long ld, long long below, int w, const double* dinv,
                                     const double* zrow, const double* alpha_below, double* seg, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  if (w <= 0) return -4;
  STPYB_CUDA(cudaMemcpyAsync(seg, zrow, (size_t)w * sizeof(double), cudaMemcpyDeviceToDevice, st));
  if (below > 0) {
    stpyb::gemv_t_sub_kernel<<<(unsigned)((below + 255) / 256), 256, 0, st>>>(Lcol + (long long)w * ld, below, w, ld,
                                                                             alpha_below, seg);
    STPYB_COUNT_LAUNCH();
  }
  return stpyb::trsv_lower(Lcol, w, ld, dinv, seg, 1, st);
}
