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
#include <mutex>
#include <vector>
#include "gemm_nt.cuh"
#include "stpyb_internal.h"

namespace stpyb {

// ---- single-launch triangular solves ---------------------------------------------------------
// One launch per sweep.  CTA number i (a ticket drawn from a global counter, so CTAs start in
// dependency order and a CTA only ever waits on CTAs that already run -- no co-residency
// assumption) owns the 128 unknowns of block i:
//   forward  : x_i = inv(L_ii) (y_i - sum_{k<i} L[i,k] x_k)      streams block ROW i of L
//   backward : x_i = inv(L_ii)^T (z_i - sum_{k>i} L[k,i]^T x_k)   streams block COLUMN i of L
// Blocks finish in order, so one monotone counter `ready` publishes progress.  A CTA streams the
// tiles whose x_k is already final without any synchronisation, keeps its partial sums in
// registers (one reduction at the very end), and issues the loads of the NEXT tile before it
// checks whether that tile's x_k has arrived -- the tile at the frontier is therefore already in
// flight when its right-hand side shows up.  Compared with the earlier chain of 2 launches per
// block (1024 launches per solve at n = 65 536) the serial step shrinks from a launch gap + tail
// to: flag -> 32-byte loads of x_k from L2 -> FMAs -> reduction -> 128 x 128 product with the
// inverted diagonal block -> flag.
constexpr int TS_THREADS = 512;      // 16 warps; a warp owns 8 rows of a tile, a lane 4 adjacent columns:
constexpr int TS_ROWS = DB / (TS_THREADS / 32);  // 8      the whole 128 KB tile is in flight at once

__device__ __forceinline__ int ld_acquire(const int* p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release(int* p, int v) {
  asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ double4 ldcg4(const double* p) {  // 32 bytes straight from L2 (written by another CTA)
  const double2 a = __ldcg(reinterpret_cast<const double2*>(p));
  const double2 b = __ldcg(reinterpret_cast<const double2*>(p) + 1);
  return make_double4(a.x, a.y, b.x, b.y);
}
__device__ __forceinline__ double4 ld4(const double* p) {
  const double2 a = *reinterpret_cast<const double2*>(p);
  const double2 b = *(reinterpret_cast<const double2*>(p) + 1);
  return make_double4(a.x, a.y, b.x, b.y);
}

// ws[0] = ticket counter, ws[1] = number of solved blocks (both zeroed before the launch)
template <int TRANS>
__global__ void __launch_bounds__(TS_THREADS, 1)
trsv_chain_kernel(const double* __restrict__ L, i64 ld, i64 n, const double* __restrict__ dinv, double* x, int* ws,
                  int nblk) {
  __shared__ int s_i, s_avail;
  __shared__ __align__(16) double sv[DB];
  __shared__ double part[TS_THREADS / 32][DB];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) s_i = atomicAdd(ws, 1);
  __syncthreads();
  const int t = s_i;                       // ticket = position in the dependency order
  const int i = TRANS ? nblk - 1 - t : t;  // block owned by this CTA
  const int b = (int)((n - (i64)i * DB < DB) ? (n - (i64)i * DB) : DB);
  int avail = 0;                           // blocks known to be solved (in ticket order)
  // forward: acc[q] = partial dot product of row warp*16+q over this lane's 4 columns, all tiles so far
  // backward: acc4[c] = partial sum for column 4*lane+c over this warp's rows
  double acc[TS_ROWS];
#pragma unroll
  for (int q = 0; q < TS_ROWS; ++q) acc[q] = 0.0;
  double4 accT = make_double4(0.0, 0.0, 0.0, 0.0);

  for (int kt = 0; kt < t; ++kt) {
    const int k = TRANS ? nblk - 1 - kt : kt;  // tile (i, k) forward, (k, i) backward
    // tile loads first (they do not depend on x) ...
    const int r0 = warp * TS_ROWS;
    double4 lv[TS_ROWS];
    if (!TRANS) {
      const double* base = L + ((i64)i * DB + r0) * ld + (i64)k * DB + 4 * lane;
#pragma unroll
      for (int q = 0; q < TS_ROWS; ++q)
        lv[q] = (r0 + q < b) ? ld4(base + (i64)q * ld) : make_double4(0.0, 0.0, 0.0, 0.0);
    } else {
      const int bk = (int)((n - (i64)k * DB < DB) ? (n - (i64)k * DB) : DB);
      const double* base = L + ((i64)k * DB + r0) * ld + (i64)i * DB + 4 * lane;
#pragma unroll
      for (int q = 0; q < TS_ROWS; ++q)
        lv[q] = (r0 + q < bk) ? ld4(base + (i64)q * ld) : make_double4(0.0, 0.0, 0.0, 0.0);
    }
    // ... then make sure x_k is final
    if (kt >= avail) {
      if (tid == 0) {
        int r;
        while ((r = ld_acquire(ws + 1)) <= kt) __nanosleep(20);
        s_avail = r;
      }
      __syncthreads();
      avail = s_avail;
      __syncthreads();
    }
    if (!TRANS) {
      const double4 xv = ldcg4(x + (i64)k * DB + 4 * lane);  // k < i <= nblk - 1: a forward tile's block k is never the partial one
#pragma unroll
      for (int q = 0; q < TS_ROWS; ++q) {
        acc[q] = fma(lv[q].x, xv.x, acc[q]);
        acc[q] = fma(lv[q].y, xv.y, acc[q]);
        acc[q] = fma(lv[q].z, xv.z, acc[q]);
        acc[q] = fma(lv[q].w, xv.w, acc[q]);
      }
    } else {
      // x_k[r0 .. r0+8): the same 64 bytes for every lane of the warp (broadcast loads from L2).  Only the last
      // block can be partial: its missing rows were loaded as zero tile rows, and x is not read past n.
      const i64 xb = (i64)k * DB + r0;
#pragma unroll
      for (int q4 = 0; q4 < TS_ROWS; q4 += 4) {
        double xs[4];
        if (xb + q4 + 3 < n) {
          const double4 xv = ldcg4(x + xb + q4);
          xs[0] = xv.x; xs[1] = xv.y; xs[2] = xv.z; xs[3] = xv.w;
        } else {
#pragma unroll
          for (int u = 0; u < 4; ++u) xs[u] = (xb + q4 + u < n) ? __ldcg(x + xb + q4 + u) : 0.0;
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          accT.x = fma(lv[q4 + u].x, xs[u], accT.x);
          accT.y = fma(lv[q4 + u].y, xs[u], accT.y);
          accT.z = fma(lv[q4 + u].z, xs[u], accT.z);
          accT.w = fma(lv[q4 + u].w, xs[u], accT.w);
        }
      }
    }
  }

  // v = rhs_i - (accumulated sum), in shared memory
  if (!TRANS) {
#pragma unroll
    for (int q = 0; q < TS_ROWS; ++q) {
      const double s = warp_sum(acc[q]);
      if (lane == 0) {
        const int r = warp * TS_ROWS + q;
        sv[r] = (r < b) ? x[(i64)i * DB + r] - s : 0.0;
      }
    }
  } else {
    part[warp][4 * lane + 0] = accT.x;
    part[warp][4 * lane + 1] = accT.y;
    part[warp][4 * lane + 2] = accT.z;
    part[warp][4 * lane + 3] = accT.w;
    __syncthreads();
    if (tid < DB) {
      double s = 0.0;
#pragma unroll
      for (int w = 0; w < TS_THREADS / 32; ++w) s += part[w][tid];
      sv[tid] = (tid < b) ? x[(i64)i * DB + tid] - s : 0.0;
    }
  }
  __syncthreads();
  // x_i = inv(L_ii) v  (TRANS: inv(L_ii)^T v); the inverted block is dense [128][128], zero outside b x b
  const double* Li = dinv + (i64)i * (DB * DB);
  if (!TRANS) {
    // warp owns 16 rows; lane 4 adjacent columns of each
    const double4 vv = *reinterpret_cast<const double4*>(&sv[4 * lane]);
    double o[TS_ROWS];
#pragma unroll
    for (int q = 0; q < TS_ROWS; ++q) {
      const double4 l = ld4(Li + (warp * TS_ROWS + q) * DB + 4 * lane);
      o[q] = l.x * vv.x + l.y * vv.y + l.z * vv.z + l.w * vv.w;
    }
#pragma unroll
    for (int q = 0; q < TS_ROWS; ++q) {
      const double s = warp_sum(o[q]);
      const int r = warp * TS_ROWS + q;
      if (lane == 0 && r < b) x[(i64)i * DB + r] = s;
    }
  } else {
    // x_i[c] = sum_r Li[r][c] v[r]: warp owns 16 rows r, lane 4 columns; combine the warps through shared memory
    double4 o = make_double4(0.0, 0.0, 0.0, 0.0);
#pragma unroll
    for (int q = 0; q < TS_ROWS; ++q) {
      const int r = warp * TS_ROWS + q;
      const double4 l = ld4(Li + r * DB + 4 * lane);
      const double v = sv[r];
      o.x = fma(l.x, v, o.x);
      o.y = fma(l.y, v, o.y);
      o.z = fma(l.z, v, o.z);
      o.w = fma(l.w, v, o.w);
    }
    __syncthreads();  // part[] is reused
    part[warp][4 * lane + 0] = o.x;
    part[warp][4 * lane + 1] = o.y;
    part[warp][4 * lane + 2] = o.z;
    part[warp][4 * lane + 3] = o.w;
    __syncthreads();
    if (tid < b) {
      double s = 0.0;
#pragma unroll
      for (int w = 0; w < TS_THREADS / 32; ++w) s += part[w][tid];
      x[(i64)i * DB + tid] = s;
    }
  }
  // publish: every writer fences, then one thread releases the counter
  __threadfence();
  __syncthreads();
  if (tid == 0) st_release(ws + 1, t + 1);
}

// Per-stream scratch (ticket + progress counter): solves may run concurrently on several streams (the
// hyper-parameter sweep), so the two ints cannot be shared.  Allocated on first use, never freed.
struct TrsvScratch {
  int dev;
  cudaStream_t st;
  int* ws;
};
static std::mutex g_trsv_mu;
static std::vector<TrsvScratch> g_trsv_scratch;

static int* trsv_scratch_for(cudaStream_t st) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return nullptr;
  std::lock_guard<std::mutex> lock(g_trsv_mu);
  for (const TrsvScratch& e : g_trsv_scratch)
    if (e.dev == dev && e.st == st) return e.ws;
  int* ws = nullptr;
  if (cudaMalloc(&ws, 64) != cudaSuccess) return nullptr;
  g_trsv_scratch.push_back({dev, st, ws});
  return ws;
}

int trsv_lower(const double* L, i64 n, i64 ld, const double* dinv, double* x, int transposed, cudaStream_t st) {
  if (n <= 0) return 0;
  // 32-byte row segments are loaded as two 16-byte halves: rows of L and x must be 16-byte aligned
  if ((ld & 1) || (((uintptr_t)L) & 15) || (((uintptr_t)x) & 15) || (((uintptr_t)dinv) & 15)) return -3;
  const i64 nblk = (n + DB - 1) / DB;
  if (nblk > 2147483647LL) return -2;
  int* ws = trsv_scratch_for(st);
  if (ws == nullptr) return STPYB_ERR_CUDA + (int)cudaErrorMemoryAllocation;
  STPYB_CUDA(cudaMemsetAsync(ws, 0, 2 * sizeof(int), st));
  prof_begin(PROF_SOLVE, (double)n * (double)n, st);
  if (!transposed)
    trsv_chain_kernel<0><<<(unsigned)nblk, TS_THREADS, 0, st>>>(L, ld, n, dinv, x, ws, (int)nblk);
  else
    trsv_chain_kernel<1><<<(unsigned)nblk, TS_THREADS, 0, st>>>(L, ld, n, dinv, x, ws, (int)nblk);
  prof_end(st);
  STPYB_COUNT_LAUNCH();
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
