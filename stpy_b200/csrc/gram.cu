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
#include <cstdlib>
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
  double kp[6];  // STPYB_K_MATERN_NU: {nu, gam1, gam2, 1/Gamma(1+mu), 1/Gamma(1-mu), 2^(1-nu)/Gamma(nu)}
};

// Modified Bessel function of the second kind K_nu(x), x > 0, real nu >= 0 (scipy.special.kv in
// stpy/kernels.py:858).  Temme's method: with nu = nl + mu, |mu| <= 1/2, K_mu and K_mu+1 come from
// Temme's series (x <= 2; the Gamma-function combinations gam1, gam2, 1/Gamma(1 +- mu) depend on nu only
// and are supplied by the host) or from Steed's evaluation of the second continued fraction (x > 2), and
// the stable upward recurrence K_{m+1} = K_{m-1} + (2 m / x) K_m climbs to nu.
__device__ double bessel_k(const double* kp, double x) {
  const double nu = kp[0];
  const int nl = (int)(nu + 0.5);
  const double mu = nu - nl, mu2 = mu * mu;
  const double xi = 1.0 / x, xi2 = 2.0 * xi;
  const double EPS = 1e-16;
  double rkmu, rk1;
  if (x < 2.0) {
    const double x2 = 0.5 * x;
    const double pimu = 3.141592653589793 * mu;
    const double fact = (fabs(pimu) < EPS) ? 1.0 : pimu / sin(pimu);
    double d = -log(x2);
    double e = mu * d;
    const double fact2 = (fabs(e) < EPS) ? 1.0 : sinh(e) / e;
    double ff = fact * (kp[1] * cosh(e) + kp[2] * fact2 * d);
    double sum = ff;
    e = exp(e);
    double p = 0.5 * e / kp[3];
    double q = 0.5 / (e * kp[4]);
    double c = 1.0;
    d = x2 * x2;
    double sum1 = p;
    for (int i = 1; i < 500; ++i) {
      ff = (i * ff + p + q) / (i * (double)i - mu2);
      c *= d / i;
      p /= (i - mu);
      q /= (i + mu);
      const double del = c * ff;
      sum += del;
      sum1 += c * (p - i * ff);
      if (fabs(del) < fabs(sum) * EPS) break;
    }
    rkmu = sum;
    rk1 = sum1 * xi2;
  } else {
    double b = 2.0 * (1.0 + x);
    double d = 1.0 / b;
    double h = d, delh = d;
    double q1 = 0.0, q2 = 1.0;
    const double a1 = 0.25 - mu2;
    double q = a1, c = a1;
    double a = -a1;
    double s = 1.0 + q * delh;
    for (int i = 2; i < 500; ++i) {
      a -= 2 * (i - 1);
      c = -a * c / i;
      const double qnew = (q1 - b * q2) / a;
      q1 = q2;
      q2 = qnew;
      q += c * qnew;
      b += 2.0;
      d = 1.0 / (b + a * d);
      delh = (b * d - 1.0) * delh;
      h += delh;
      const double dels = q * delh;
      s += dels;
      if (fabs(dels / s) < EPS) break;
    }
    h = a1 * h;
    rkmu = sqrt(3.141592653589793 / (2.0 * x)) * exp(-x) / s;
    rk1 = rkmu * (mu + x + 0.5 - h) * xi;
  }
  for (int i = 1; i <= nl; ++i) {
    const double t = (mu + i) * xi2 * rk1 + rkmu;
    rkmu = rk1;
    rk1 = t;
  }
  return rkmu;
}

// general-nu Matern from the scaled distance r: 2^(1-nu)/Gamma(nu) t^nu K_nu(t), t = sqrt(2 nu) r, with exact zeros
// moved to machine epsilon as the reference does (kernels.py:854)
__device__ __noinline__ double matern_nu_from_r(const double* kp, double r) {
  if (r == 0.0) r = 2.220446049250313e-16;
  const double t = sqrt(2.0 * kp[0]) * r;
  const double kv = bessel_k(kp, t);
  return (kv == 0.0) ? 0.0 : kp[5] * pow(t, kp[0]) * kv;
}

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
  } else if (KIND == STPYB_K_MATERN_NU) {
    return 0.0;  // evaluated through gram_value (needs the Bessel constants)
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
    case STPYB_K_MATERN_NU: {
      double sq = (-2.0 * dot + na) + nb;
      v = matern_nu_from_r(km.kp, sqrt(sq > 0.0 ? sq : 0.0));
    } break;
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

// ---- the Gram kernel -----------------------------------------------------------------------
// The contraction depth here is the (padded) input dimension, 4..64, so this stage is bound by
// the n^2 output stream and the kernel map, not by the tensor pipe: at d = 8 one 8x8 output atom
// costs two DMMA.8x8x4 but 64 exp / sqrt evaluations and 512 bytes of stores.  The generic
// contraction kernel (128x64 tiles, 4 warps, ~210 registers, 2 CTAs per SM) ran this stage at
// 8 resident warps per SM and 0.13-0.18 of the HBM rate -- ncu (profiles/ncu_gram_r02.json):
// warps active 12 %, FP64 pipe 18 %, DRAM 14 %, stalls `wait` and `long_scoreboard`: latency-bound.
// This kernel is built for occupancy instead: 64x64 output tile per CTA, 8 warps, each warp owns
// 8 rows x 64 columns (8 DMMA atoms, 16 accumulator doubles per lane), operand fragments are read
// straight from global memory (the prepped inputs are a few MB and stay in L1/L2; no shared
// memory, no barrier), the map is applied in registers and every lane stores 16-byte pairs that
// form full 32-byte sectors.  Fragment of lane (g = lane/4, t = lane%4): A(row g, k t),
// B(k t, col g), C(row g, cols 2t, 2t+1).
constexpr int GT_M = 64, GT_N = 64;

struct GramArgs {
  const double* Ap;  // prepped a-points (columns of K), [n][dpad]
  const double* na;
  const double* Bp;  // prepped b-points (rows of K), [m][dpad]
  const double* nb;
  int n, m, dpad;
  KernelMap km;
  int refine, op, lower_only, vec;
  double diag_add;
  double* C;
  i64 ldc;
  int tiles_n, tri_rows;
  i64 tri_count;
};

template <int KIND>
__device__ __forceinline__ double gram_refined(const GramArgs& ga, int row, int col) {
  const double* pa = ga.Ap + (i64)col * ga.dpad;
  const double* pb = ga.Bp + (i64)row * ga.dpad;
  double s = 0.0;
  for (int k = 0; k < ga.dpad; ++k) {
    const double df = pa[k] - pb[k];
    s = fma(df, df, s);
  }
  if (KIND == STPYB_K_MATERN_NU) return matern_nu_from_r(ga.km.kp, sqrt(s));
  return matern_from_r<KIND>(sqrt(s));
}

// kernel value without kappa; the general-nu Matern goes through the Bessel routine
template <int KIND>
__device__ __forceinline__ double gram_value(const GramArgs& ga, double dot, double a2, double b2) {
  if (KIND == STPYB_K_MATERN_NU) {
    const double sq = (-2.0 * dot + a2) + b2;
    return matern_nu_from_r(ga.km.kp, sqrt(sq > 0.0 ? sq : 0.0));
  }
  return kernel_map<KIND>(dot, a2, b2, ga.km.arg_scale, ga.km.p0);
}

// Interior tiles (all 64 x 64 outputs in range, aligned rows, plain SET, no diagonal to touch) take
// a lean path: the first version of this kernel spent ~120 of its ~147 instructions per output
// element on bounds / alignment / accumulate-mode / diagonal handling replicated per atom and was
// ISSUE-bound at 0.87 elements per cycle and SM (SASS count, profiles/gram_sass_mix_r02.txt); here
// those decisions are made once per CTA and the atom loop is: 2 x (kernel map) + one 16-byte store.
template <int KIND>
__device__ __forceinline__ void gram_tile_fast(const GramArgs& ga, int m0, int n0, int warp, int g, int t) {
  const int dpad = ga.dpad;
  const int row = m0 + warp * 8 + g;
  const double* pb = ga.Bp + (i64)row * dpad + t;
  const double* pa = ga.Ap + (i64)(n0 + g) * dpad + t;
  const i64 astep = (i64)8 * dpad;
  double acc[8][2];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j][0] = acc[j][1] = 0.0;
  for (int k = 0; k < dpad; k += 4) {
    const double a = __ldg(pb + k);
    double b[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) b[j] = __ldg(pa + j * astep + k);
#pragma unroll
    for (int j = 0; j < 8; ++j) dmma884(acc[j][0], acc[j][1], a, b[j]);
  }
  const double b2 = __ldg(ga.nb + row);
  const double arg_scale = ga.km.arg_scale, p0 = ga.km.p0, kappa = ga.km.kappa;
  const double2* na2 = reinterpret_cast<const double2*>(ga.na + n0 + 2 * t);
  double2* crow = reinterpret_cast<double2*>(ga.C + (i64)row * ga.ldc + n0 + 2 * t);
  const bool refine = ((KIND >= STPYB_K_MATERN12 && KIND <= STPYB_K_MATERN52) || KIND == STPYB_K_MATERN_NU) && ga.refine;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const double2 a2 = __ldg(na2 + 4 * j);  // read-only path: the loads may be hoisted above the stores of earlier atoms
    double o0 = gram_value<KIND>(ga, acc[j][0], a2.x, b2);
    double o1 = gram_value<KIND>(ga, acc[j][1], a2.y, b2);
    if (refine) {
      const double sq0 = (-2.0 * acc[j][0] + a2.x) + b2;
      const double sq1 = (-2.0 * acc[j][1] + a2.y) + b2;
      if (sq0 < 1e-3 * (a2.x + b2)) o0 = gram_refined<KIND>(ga, row, n0 + 8 * j + 2 * t);
      if (sq1 < 1e-3 * (a2.y + b2)) o1 = gram_refined<KIND>(ga, row, n0 + 8 * j + 2 * t + 1);
    }
    if (KIND == STPYB_K_LINEAR) {
      o0 = kappa * o0 + p0;
      o1 = kappa * o1 + p0;
    } else {
      o0 = kappa * o0;
      o1 = kappa * o1;
    }
    crow[4 * j] = make_double2(o0, o1);
  }
}

template <int KIND, int MINB>
__global__ void __launch_bounds__(256, MINB) gram_tile_kernel(GramArgs ga) {
  // tile decode: lower-only launches enumerate tile rows ti with ti+1 tiles (tj <= ti) up to tri_rows,
  // then full rows of tiles_n tiles
  i64 bid = blockIdx.x;
  int ti, tj;
  if (ga.lower_only && bid < ga.tri_count) {
    i64 r = (i64)((sqrt(8.0 * (double)bid + 1.0) - 1.0) * 0.5);
    while ((r + 1) * (r + 2) / 2 <= bid) ++r;
    while (r * (r + 1) / 2 > bid) --r;
    ti = (int)r;
    tj = (int)(bid - r * (r + 1) / 2);
  } else {
    const i64 l = ga.lower_only ? bid - ga.tri_count : bid;
    const int base = ga.lower_only ? ga.tri_rows : 0;
    ti = base + (int)(l / ga.tiles_n);
    tj = (int)(l % ga.tiles_n);
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  const int m0 = ti * GT_M, n0 = tj * GT_N;
  if (m0 + GT_M <= ga.m && n0 + GT_N <= ga.n && ga.vec && ga.op == STPYB_OP_SET &&
      (ga.diag_add == 0.0 || ti != tj)) {
    gram_tile_fast<KIND>(ga, m0, n0, warp, g, t);
    return;
  }
  const int row = m0 + warp * 8 + g;
  const int rowc = row < ga.m ? row : ga.m - 1;  // clamped for loads; stores are guarded
  const int dpad = ga.dpad;

  double acc[8][2];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j][0] = acc[j][1] = 0.0;
  const double* pb = ga.Bp + (i64)rowc * dpad + t;
  for (int k = 0; k < dpad; k += 4) {
    const double a = pb[k];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      int c = n0 + 8 * j + g;
      c = c < ga.n ? c : ga.n - 1;
      dmma884(acc[j][0], acc[j][1], a, ga.Ap[(i64)c * dpad + t + k]);
    }
  }

  if (row >= ga.m) return;
  const double b2 = ga.nb[rowc];
  const double arg_scale = ga.km.arg_scale, p0 = ga.km.p0, kappa = ga.km.kappa;
  double* crow = ga.C + (i64)row * ga.ldc;
#pragma unroll 1
  for (int j = 0; j < 8; ++j) {
    const int col = n0 + 8 * j + 2 * t;
    if (col >= ga.n) continue;
    const int nc = (col + 1 < ga.n) ? 2 : 1;
    const double a20 = ga.na[col];
    const double a21 = (nc == 2) ? ga.na[col + 1] : 0.0;
    // acc[j] with a runtime j: the loop is deliberately not unrolled (edge tiles are rare, code size matters more)
    double d0 = 0.0, d1 = 0.0;
#pragma unroll
    for (int q = 0; q < 8; ++q)
      if (q == j) {
        d0 = acc[q][0];
        d1 = acc[q][1];
      }
    double o0 = gram_value<KIND>(ga, d0, a20, b2);
    double o1 = gram_value<KIND>(ga, d1, a21, b2);
    if ((KIND >= STPYB_K_MATERN12 && KIND <= STPYB_K_MATERN52) || KIND == STPYB_K_MATERN_NU) {
      if (ga.refine) {
        // scipy-cdist semantics: where the expansion cancels, recompute from direct differences
        const double sq0 = (-2.0 * d0 + a20) + b2;
        const double sq1 = (-2.0 * d1 + a21) + b2;
        if (sq0 < 1e-3 * (a20 + b2)) o0 = gram_refined<KIND>(ga, row, col);
        if (nc == 2 && sq1 < 1e-3 * (a21 + b2)) o1 = gram_refined<KIND>(ga, row, col + 1);
      }
    }
    if (KIND == STPYB_K_LINEAR) {
      o0 = kappa * o0 + p0;
      o1 = kappa * o1 + p0;
    } else {
      o0 = kappa * o0;
      o1 = kappa * o1;
    }
    double* p = crow + col;
    if (ga.op != STPYB_OP_SET) {
      const double c0 = p[0];
      const double c1 = (nc == 2) ? p[1] : 0.0;
      if (ga.op == STPYB_OP_ADD) {
        o0 = c0 + o0;
        o1 = c1 + o1;
      } else {
        o0 = c0 * o0;
        o1 = c1 * o1;
      }
    }
    if (row == col) o0 += ga.diag_add;
    if (row == col + 1) o1 += ga.diag_add;
    if (ga.vec && nc == 2) {
      *reinterpret_cast<double2*>(p) = make_double2(o0, o1);
    } else {
      p[0] = o0;
      if (nc == 2) p[1] = o1;
    }
  }
}

template <int KIND>
static int launch_gram(GramArgs ga, cudaStream_t st) {
  const int tiles_m = ceil_div(ga.m, GT_M);
  ga.tiles_n = ceil_div(ga.n, GT_N);
  i64 grid;
  if (ga.lower_only) {
    ga.tri_rows = tiles_m < ga.tiles_n ? tiles_m : ga.tiles_n;
    ga.tri_count = (i64)ga.tri_rows * (ga.tri_rows + 1) / 2;
    grid = ga.tri_count + (i64)(tiles_m - ga.tri_rows) * ga.tiles_n;
  } else {
    ga.tri_rows = 0;
    ga.tri_count = 0;
    grid = (i64)tiles_m * ga.tiles_n;
  }
  if (grid <= 0) return 0;
  if (grid > 2147483647LL) return -2;
  // resident CTAs per SM the register budget is compiled for: 3 (80 registers, 24 warps) by default; the Matern maps
  // spill ~90 bytes there, so STPYB_GRAM_MINB=2 (126 registers, 16 warps, no spill) is kept for comparison
  static int minb = 0;
  if (minb == 0) {
    const char* ev = getenv("STPYB_GRAM_MINB");
    minb = (ev && atoi(ev) == 2) ? 2 : 3;
  }
  if (minb == 2) gram_tile_kernel<KIND, 2><<<(unsigned)grid, 256, 0, st>>>(ga);
  else gram_tile_kernel<KIND, 3><<<(unsigned)grid, 256, 0, st>>>(ga);
  STPYB_COUNT_LAUNCH();
  STPYB_CUDA(cudaGetLastError());
  return 0;
}

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
                          const double* kparams_host_or_null, void* stream) {
  if (kind < 0 || kind >= STPYB_K_COUNT) return -1;
  if (op < 0 || op > STPYB_OP_MUL) return -13;
  if (kind == STPYB_K_MATERN_NU && kparams_host_or_null == nullptr) return -18;
  if (n <= 0 || m <= 0) return 0;
  if (n > 2147483647LL || m > 2147483647LL) return -4;
  if (dpad <= 0 || (dpad & 3)) return -8;
  GramArgs ga;
  ga.Ap = Ap; ga.na = na; ga.Bp = Bp; ga.nb = nb;
  ga.n = (int)n; ga.m = (int)m; ga.dpad = dpad;
  ga.km.kind = kind; ga.km.arg_scale = arg_scale; ga.km.kappa = kappa; ga.km.p0 = p0;
  for (int q = 0; q < 6; ++q) ga.km.kp[q] = kparams_host_or_null ? kparams_host_or_null[q] : 0.0;
  const bool is_matern = (kind >= STPYB_K_MATERN12 && kind <= STPYB_K_MATERN52) || kind == STPYB_K_MATERN_NU;
  ga.refine = (refine && is_matern) ? 1 : 0;
  ga.op = op; ga.lower_only = lower_only ? 1 : 0; ga.diag_add = diag_add; ga.C = K; ga.ldc = ldk;
  ga.vec = ((ldk & 1) == 0 && (((uintptr_t)K) & 15) == 0 && (((uintptr_t)na) & 15) == 0) ? 1 : 0;
  const cudaStream_t st = (cudaStream_t)stream;
  prof_begin(PROF_GRAM, (lower_only ? 0.5 : 1.0) * 2.0 * (double)m * (double)n * dpad, st);
  int rc;
  switch (kind) {
    case STPYB_K_SE: rc = launch_gram<STPYB_K_SE>(ga, st); break;
    case STPYB_K_MATERN12: rc = launch_gram<STPYB_K_MATERN12>(ga, st); break;
    case STPYB_K_MATERN32: rc = launch_gram<STPYB_K_MATERN32>(ga, st); break;
    case STPYB_K_MATERN52: rc = launch_gram<STPYB_K_MATERN52>(ga, st); break;
    case STPYB_K_POLY: rc = launch_gram<STPYB_K_POLY>(ga, st); break;
    case STPYB_K_LINEAR: rc = launch_gram<STPYB_K_LINEAR>(ga, st); break;
    case STPYB_K_MATERN_NU: rc = launch_gram<STPYB_K_MATERN_NU>(ga, st); break;
    default: rc = -1;
  }
  prof_end(st);
  return rc;
}

extern "C" int stpyb_gram_diag(int kind, const double* Ap, const double* na, const double* Bp, const double* nb,
                               long long n, int dpad, double arg_scale, double kappa, double p0, int op,
                               double* out, const double* kparams_host_or_null, void* stream) {
  if (kind < 0 || kind >= STPYB_K_COUNT) return -1;
  if (kind == STPYB_K_MATERN_NU && kparams_host_or_null == nullptr) return -14;
  if (n <= 0) return 0;
  KernelMap km;
  km.kind = kind; km.arg_scale = arg_scale; km.kappa = kappa; km.p0 = p0;
  for (int q = 0; q < 6; ++q) km.kp[q] = kparams_host_or_null ? kparams_host_or_null[q] : 0.0;
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
