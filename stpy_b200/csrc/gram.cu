// Fused kernel Gram-matrix construction.
//
// Replaces the per-kernel builders of the reference, which materialise 5-6 n^2
// temporaries per call (mm, two broadcast adds, scale, exp, scale):
//   squared_exponential_kernel   stpy/kernels.py:368-398
//   ard_kernel                   stpy/kernels.py:552-583
//   matern_kernel                stpy/kernels.py:811-859   (scipy cdist: direct differences)
//   ard_matern_kernel            stpy/kernels.py:917-970   (torch.cdist: GEMM expansion + clamp)
//   polynomial_kernel            stpy/kernels.py:766-784
//   linear_kernel                stpy/kernels.py:300-320
//   kernel algebra (+, *)        stpy/kernels.py:136-159
//   "+ s^2 I"                    gauss_procc.py:163, 633
// Here: one pass.  The cross term b a^T is the DMMA contraction of gemm_nt.cuh
// (K = padded input dimension), and the squared distance, lengthscale scaling,
// exp / Matern / power map, kappa, the +/* accumulation with the previous
// sub-kernel and the noise diagonal are applied to the accumulator registers
// before one vectorised store.  Orientation follows the reference: K[j,i] =
// k(b_j, a_i), shape (|b|, |a|).
#include "gemm_nt.cuh"
#include "stpyb_internal.h"
#include "../../include/stpyb.h"

namespace stpyb {

struct PrepArgs {
  double scale[STPYB_MAX_DIM];
  int idx[STPYB_MAX_DIM];
  int dg;      // selected columns
  int divide;  // 1: x / scale (matern_kernel), 0: x * scale (ard kernels multiply by 1/gamma)
};

// Xp[i][k] = X[i][idx[k]] (*|/) scale[k], zero padded to dpad; norms[i] = sum_k Xp[i][k]^2
__global__ void __launch_bounds__(256) gram_prep_kernel(const double* __restrict__ X, i64 n, i64 ldx, PrepArgs pa,
                                                       double* __restrict__ Xp, int dpad, double* __restrict__ norms) {
  const i64 i = (i64)blockIdx.x * 256 + threadIdx.x;
  if (i >= n) return;
  const double* row = X + i * ldx;
  double s = 0.0;
  for (int k = 0; k < pa.dg; ++k) {
    double v = row[pa.idx[k]];
    v = pa.divide ? v / pa.scale[k] : v * pa.scale[k];
    Xp[i * dpad + k] = v;
    s += v * v;
  }
  for (int k = pa.dg; k < dpad; ++k) Xp[i * dpad + k] = 0.0;
  if (norms) norms[i] = s;
}

struct KernelMap {
  int kind;
  double arg_scale, kappa, p0;
};

// kernel value (without kappa) from the cross term and the two squared norms; KIND is a
// compile-time constant in the Gram epilogue so the map is branch-free there
template <int KIND>
__device__ __forceinline__ double kernel_map(double dot, double na, double nb, double arg_scale, double p0) {
  if (KIND == STPYB_K_SE) {
    const double sq = (-2.0 * dot + na) + nb;
    return exp(arg_scale * sq);
  } else if (KIND == STPYB_K_POLY) {
    const double t = dot + 1.0;
    if (p0 == 2.0) return t * t;
    if (p0 == 3.0) return t * t * t;
    if (p0 == 1.0) return t;
    return pow(t, p0);
  } else if (KIND == STPYB_K_LINEAR) {
    return dot;
  } else {
    double sq = (-2.0 * dot + na) + nb;
    sq = sq > 0.0 ? sq : 0.0;
    const double r = sqrt(sq);
    if (KIND == STPYB_K_MATERN12) return exp(-r);
    if (KIND == STPYB_K_MATERN32) {
      const double t = r * 1.7320508075688772;
      return (1.0 + t) * exp(-t);
    }
    const double t = r * 2.23606797749979;
    return (1.0 + t + t * t * 0.3333333333333333) * exp(-t);
  }
}

__device__ __forceinline__ double kernel_value(const KernelMap& km, double dot, double na, double nb) {
  double v;
  switch (km.kind) {
    case STPYB_K_SE: v = kernel_map<STPYB_K_SE>(dot, na, nb, km.arg_scale, km.p0); break;
    case STPYB_K_MATERN12: v = kernel_map<STPYB_K_MATERN12>(dot, na, nb, km.arg_scale, km.p0); break;
    case STPYB_K_MATERN32: v = kernel_map<STPYB_K_MATERN32>(dot, na, nb, km.arg_scale, km.p0); break;
    case STPYB_K_MATERN52: v = kernel_map<STPYB_K_MATERN52>(dot, na, nb, km.arg_scale, km.p0); break;
    case STPYB_K_POLY: v = kernel_map<STPYB_K_POLY>(dot, na, nb, km.arg_scale, km.p0); break;
    default: return km.kappa * dot + km.p0;
  }
  return km.kappa * v;
}

template <int KIND>
__device__ __forceinline__ double matern_from_r(double r) {
  if (KIND == STPYB_K_MATERN12) return exp(-r);
  if (KIND == STPYB_K_MATERN32) {
    const double t = r * 1.7320508075688772;
    return (1.0 + t) * exp(-t);
  }
  const double t = r * 2.23606797749979;
  return (1.0 + t + t * t * 0.3333333333333333) * exp(-t);
}

template <int KIND>
struct EpiGram {
  static constexpr bool kPreload = false;
  static constexpr bool kRowBatch = true;
  __device__ __forceinline__ void preload(int, int, int, double&, double&) const {}
  __device__ __forceinline__ void preload_finish(double&, double&) const {}
  KernelMap km;
  const double* na;  // norms of a-points (columns)
  const double* nb;  // norms of b-points (rows)
  const double* Ap;  // prepped points, for the direct-difference refinement
  const double* Bp;
  int dpad;
  int refine;        // Matern only: recompute cancellation-prone distances by direct differences
  int op;            // STPYB_OP_SET / ADD / MUL with the value already in K
  double diag_add;   // added where row == col (after op)
  double* C;
  i64 ldc;
  int vec;

  __device__ __forceinline__ double refined(int row, int col) const {
    const double* pa = Ap + (i64)col * dpad;
    const double* pb = Bp + (i64)row * dpad;
    double s = 0.0;
    for (int k = 0; k < dpad; ++k) {
      const double df = pa[k] - pb[k];
      s = fma(df, df, s);
    }
    return matern_from_r<KIND>(sqrt(s));
  }

  // one call per accumulator row; deliberately NOT inlined: the exp / sqrt expansions of 2*NI
  // elements are a few KB of code, and inlining them MI times overflowed the instruction cache
  // (ncu: 63 % of warp samples stalled on no_instructions before this change)
  __device__ __noinline__ void apply_row(int row, int col_base, int N, double v0, double v1, double v2, double v3,
                                         double v4, double v5, double v6, double v7) const {
    constexpr int NI = 4;
    const double acc[NI][2] = {{v0, v1}, {v2, v3}, {v4, v5}, {v6, v7}};
    const double b2 = nb[row];
    double o[NI][2];
    double a2[NI][2];
#pragma unroll
    for (int j = 0; j < NI; ++j) {
      const int col = col_base + j * 8;
      a2[j][0] = (col < N) ? na[col] : 0.0;
      a2[j][1] = (col + 1 < N) ? na[col + 1] : 0.0;
    }
#pragma unroll
    for (int j = 0; j < NI; ++j) {
#pragma unroll
      for (int e = 0; e < 2; ++e) o[j][e] = kernel_map<KIND>(acc[j][e], a2[j][e], b2, km.arg_scale, km.p0);
    }
    if (KIND >= STPYB_K_MATERN12 && KIND <= STPYB_K_MATERN52) {
      if (refine) {
#pragma unroll
        for (int j = 0; j < NI; ++j) {
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const int col = col_base + j * 8 + e;
            const double sq = (-2.0 * acc[j][e] + a2[j][e]) + b2;
            if (col < N && sq < 1e-3 * (a2[j][e] + b2)) o[j][e] = refined(row, col);
          }
        }
      }
    }
#pragma unroll
    for (int j = 0; j < NI; ++j) {
      const int col = col_base + j * 8;
      if (col >= N) continue;
      const int nc = (col + 1 < N) ? 2 : 1;
      double o0, o1;
      if (KIND == STPYB_K_LINEAR) {
        o0 = km.kappa * o[j][0] + km.p0;
        o1 = km.kappa * o[j][1] + km.p0;
      } else {
        o0 = km.kappa * o[j][0];
        o1 = km.kappa * o[j][1];
      }
      double* p = C + (i64)row * ldc + col;
      if (op != STPYB_OP_SET) {
        double c0, c1 = 0.0;
        if (vec && nc == 2) {
          const double2 c = *reinterpret_cast<const double2*>(p);
          c0 = c.x;
          c1 = c.y;
        } else {
          c0 = p[0];
          if (nc == 2) c1 = p[1];
        }
        if (op == STPYB_OP_ADD) {
          o0 = c0 + o0;
          o1 = c1 + o1;
        } else {
          o0 = c0 * o0;
          o1 = c1 * o1;
        }
      }
      if (row == col) o0 += diag_add;
      if (row == col + 1) o1 += diag_add;
      if (vec && nc == 2) {
        *reinterpret_cast<double2*>(p) = make_double2(o0, o1);
      } else {
        p[0] = o0;
        if (nc == 2) p[1] = o1;
      }
    }
  }
};

// out[i] = k(b_i, a_i): the diagonal of a Gram block (kernel_diag, stpy/kernels.py:112-134,
// and the 1x1 kernel calls of gauss_procc.py:347).
__global__ void __launch_bounds__(256) gram_diag_kernel(KernelMap km, const double* __restrict__ Ap,
                                                       const double* __restrict__ na, const double* __restrict__ Bp,
                                                       const double* __restrict__ nb, i64 n, int dpad, int op,
                                                       double* out) {
  const i64 i = (i64)blockIdx.x * 256 + threadIdx.x;
  if (i >= n) return;
  double dot = 0.0;
  for (int k = 0; k < dpad; ++k) dot = fma(Ap[i * dpad + k], Bp[i * dpad + k], dot);
  double v = kernel_value(km, dot, na[i], nb[i]);
  if (op == STPYB_OP_ADD) v = out[i] + v;
  else if (op == STPYB_OP_MUL) v = out[i] * v;
  out[i] = v;
}


// ---- shared-distance sweep: one cross-term tile, many kernels' epilogues ----
struct MultiMaps {
  int kinds[64];
  double arg_scales[64];
  double kappas[64];
  int nk;
};

struct EpiGramMulti {
  static constexpr bool kPreload = false;
  static constexpr bool kRowBatch = false;
  __device__ __forceinline__ void preload(int, int, int, double&, double&) const {}
  __device__ __forceinline__ void preload_finish(double&, double&) const {}
  MultiMaps mm;
  const double* na;
  const double* Ap;
  int dpad;
  double diag_add;
  double* C;
  i64 ldc, stride;
  int vec;

  __device__ __forceinline__ double sqdist(int row, int col, double dot) const {
    const double a2 = na[col], b2 = na[row];
    double sq = (-2.0 * dot + a2) + b2;
    if (sq < 1e-3 * (a2 + b2)) {  // cancellation: direct differences (exact 0 on the diagonal)
      const double* pa = Ap + (i64)col * dpad;
      const double* pb = Ap + (i64)row * dpad;
      double s = 0.0;
      for (int k = 0; k < dpad; ++k) {
        const double df = pa[k] - pb[k];
        s = fma(df, df, s);
      }
      sq = s;
    }
    return sq;
  }
  __device__ __forceinline__ double map(int q, double sq) const {
    const int kind = mm.kinds[q];
    const double sc = mm.arg_scales[q];
    double v;
    if (kind == STPYB_K_SE) {
      v = exp(sc * sq);
    } else {
      const double r = sqrt(sq) * sc;
      if (kind == STPYB_K_MATERN12) v = exp(-r);
      else if (kind == STPYB_K_MATERN32) { const double t = r * 1.7320508075688772; v = (1.0 + t) * exp(-t); }
      else { const double t = r * 2.23606797749979; v = (1.0 + t + t * t * 0.3333333333333333) * exp(-t); }
    }
    return mm.kappas[q] * v;
  }
  __device__ __forceinline__ void apply(int row, int col, double v0, double v1, int nc) const {
    const double s0 = sqdist(row, col, v0);
    const double s1 = (nc == 2) ? sqdist(row, col + 1, v1) : 0.0;
    double* p = C + (i64)row * ldc + col;
    for (int q = 0; q < mm.nk; ++q) {
      double o0 = map(q, s0);
      double o1 = (nc == 2) ? map(q, s1) : 0.0;
      if (row == col) o0 += diag_add;
      if (row == col + 1) o1 += diag_add;
      double* pq = p + (i64)q * stride;
      if (vec && nc == 2) {
        *reinterpret_cast<double2*>(pq) = make_double2(o0, o1);
      } else {
        pq[0] = o0;
        if (nc == 2) pq[1] = o1;
      }
    }
  }
};

}  // namespace stpyb

using namespace stpyb;

extern "C" int stpyb_gram_prep(const double* X, long long n, long long ldx, const int* cols_host, int dg,
                               const double* scale_host, int scale_len, int divide, double* Xp, int dpad,
                               double* norms_or_null, void* stream) {
  if (n <= 0) return 0;
  if (dg <= 0 || dg > STPYB_MAX_DIM) return -5;
  if (dpad < dg || (dpad & 3)) return -10;
  if (scale_len != 0 && scale_len != 1 && scale_len != dg) return -7;
  PrepArgs pa;
  pa.dg = dg;
  pa.divide = divide;
  for (int k = 0; k < dg; ++k) {
    pa.idx[k] = cols_host ? cols_host[k] : k;
    pa.scale[k] = (scale_len == 0) ? 1.0 : (scale_len == 1 ? scale_host[0] : scale_host[k]);
  }
  gram_prep_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(X, n, ldx, pa, Xp, dpad,
                                                                                 norms_or_null);
  STPYB_COUNT_LAUNCH();
  STPYB_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int stpyb_gram(int kind, const double* Ap, const double* na, long long n, const double* Bp,
                          const double* nb, long long m, int dpad, double arg_scale, double kappa, double p0,
                          int refine, int op, double diag_add, int lower_only, double* K, long long ldk,
                          void* stream) {
  if (kind < 0 || kind >= STPYB_K_COUNT) return -1;
  if (op < 0 || op > STPYB_OP_MUL) return -13;
  if (n <= 0 || m <= 0) return 0;
  GemmArgs g;
  g.A = Bp; g.B = Ap; g.lda = dpad; g.ldb = dpad;
  g.M = (int)m; g.N = (int)n; g.K = dpad;
  g.tri = lower_only ? TRI_LOWER : TRI_FULL; g.kskip = 0;
  const cudaStream_t st = (cudaStream_t)stream;
  prof_begin(PROF_GRAM, (lower_only ? 0.5 : 1.0) * 2.0 * (double)m * (double)n * dpad, st);
  int rc;
#define STPYB_GRAM_CASE(KIND)                                                                         \
  case KIND: {                                                                                        \
    EpiGram<KIND> e;                                                                                  \
    e.km.kind = kind; e.km.arg_scale = arg_scale; e.km.kappa = kappa; e.km.p0 = p0;                   \
    e.na = na; e.nb = nb; e.Ap = Ap; e.Bp = Bp; e.dpad = dpad;                                        \
    e.refine = (refine && kind >= STPYB_K_MATERN12 && kind <= STPYB_K_MATERN52) ? 1 : 0;              \
    e.op = op; e.diag_add = diag_add; e.C = K; e.ldc = ldk;                                           \
    e.vec = ((ldk & 1) == 0 && (((uintptr_t)K) & 15) == 0) ? 1 : 0;                                   \
    rc = launch_gemm_nt<CfgStream, EpiGram<KIND>>(g, e, st);                                          \
  } break;
  switch (kind) {
    STPYB_GRAM_CASE(STPYB_K_SE)
    STPYB_GRAM_CASE(STPYB_K_MATERN12)
    STPYB_GRAM_CASE(STPYB_K_MATERN32)
    STPYB_GRAM_CASE(STPYB_K_MATERN52)
    STPYB_GRAM_CASE(STPYB_K_POLY)
    STPYB_GRAM_CASE(STPYB_K_LINEAR)
    default: rc = -1;
  }
#undef STPYB_GRAM_CASE
  prof_end(st);
  return rc;
}

extern "C" int stpyb_gram_diag(int kind, const double* Ap, const double* na, const double* Bp, const double* nb,
                               long long n, int dpad, double arg_scale, double kappa, double p0, int op,
                               double* out, void* stream) {
  if (kind < 0 || kind >= STPYB_K_COUNT) return -1;
  if (n <= 0) return 0;
  KernelMap km;
  km.kind = kind; km.arg_scale = arg_scale; km.kappa = kappa; km.p0 = p0;
  gram_diag_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(km, Ap, na, Bp, nb, n, dpad, op,
                                                                                 out);
  STPYB_COUNT_LAUNCH();
  STPYB_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int stpyb_gram_multi(int nk, const int* kinds, const double* arg_scales, const double* kappas,
                                const double* Ap, const double* na, long long n, int dpad, double diag_add,
                                double* K, long long ldk, long long stride_k, void* stream) {
  if (nk <= 0 || nk > 64) return -1;
  if (n <= 0) return 0;
  GemmArgs g;
  g.A = Ap; g.B = Ap; g.lda = dpad; g.ldb = dpad;
  g.M = (int)n; g.N = (int)n; g.K = dpad; g.tri = TRI_LOWER; g.kskip = 0;
  EpiGramMulti e;
  e.mm.nk = nk;
  for (int q = 0; q < nk; ++q) {
    if (kinds[q] < STPYB_K_SE || kinds[q] > STPYB_K_MATERN52) return -2;
    e.mm.kinds[q] = kinds[q];
    e.mm.arg_scales[q] = arg_scales[q];
    e.mm.kappas[q] = kappas[q];
  }
  e.na = na; e.Ap = Ap; e.dpad = dpad; e.diag_add = diag_add; e.C = K; e.ldc = ldk; e.stride = stride_k;
  e.vec = ((ldk & 1) == 0 && (stride_k & 1) == 0 && (((uintptr_t)K) & 15) == 0) ? 1 : 0;
  return launch_gemm_nt<CfgStream, EpiGramMulti>(g, e, (cudaStream_t)stream);
}
