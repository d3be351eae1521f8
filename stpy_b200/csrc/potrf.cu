// Blocked Cholesky factorisation (lower, row-major, in place).
//
// Replaces torch.linalg.cholesky / lstsq / lu_factor / slogdet+solve on the GP
// path: stpy/estimator.py:35, stpy/continuous_processes/gauss_procc.py:367-378,
// 633-635.  Right-looking over outer panels (1024 columns on one GPU), left-looking
// inside a panel; per 128-wide block column j of a panel:
//   C0 in-panel update   : the block column receives all previous block columns of
//                          the panel in one DMMA update of depth K = 128 j
//   C1 potrf_diag_kernel : factor the 128x128 diagonal block in shared memory
//                          and form inv(L_jj) (kept for all later solves)
//   C2 panel TRSM        : A21 <- A21 * inv(L_jj)^T  as a DMMA GEMM (in place)
// and once per outer panel
//   C3 trailing update   : A22 <- A22 - L21 * L21^T  as a DMMA SYRK over the
//                          lower tiles, K = outer panel width, issued in two parts
//                          so that the next panel (side stream) overlaps the second.
#include <cstdlib>
#include "gemm_nt_tma.cuh"
#include "stpyb_internal.h"
#include "../../include/stpyb.h"

namespace stpyb {

constexpr int SLD = 132;  // row stride = 4 (mod 16) doubles: every DMMA fragment load (lane -> row g, k t) is bank-conflict-free
constexpr int LEAF = 8;

// One CTA (512 threads) factors a diagonal block of order b <= 128 held in shared memory and
// forms its triangular inverse.  The serial dependency chain of a Cholesky (pivot -> scale ->
// update, 128 times) bounds this kernel, so it is organised to keep that chain short and to put
// everything that is a small matrix product on DMMA:
//   * 16 leaf panels of 8 columns; the 8x8 leaf is factored by ONE thread entirely in registers
//     (rsqrt instead of sqrt + divide, no barriers inside the leaf);
//   * the rows below the leaf are solved one thread per row by forward substitution (36 FMAs);
//   * the rank-8 trailing update is done 8x8 tile by tile with two DMMA.8x8x4 per tile;
//   * the 16 leaf inverses are formed in parallel, then the off-leaf blocks of inv(L) are built by
//     16 warps, one per 8-column block column, each sweeping down its column with DMMA
//     accumulations  T = sum_k L[r][k] X[k][c],  X[r][c] = -inv(leaf_r) T  (no block barrier).
// inv(L) is parked transposed in the strict upper triangle of the same buffer, its diagonal in
// dinv[].  Fragment of lane (g = lane/4, t = lane%4): A(row g, k t), B(k t, col g), C(row g, cols 2t, 2t+1).
__global__ void __launch_bounds__(512, 1)
potrf_diag_kernel(double* __restrict__ A, i64 lda, int b, double* __restrict__ Linv, int* info, int j0,
                  long long* stamps, double* __restrict__ mirror, i64 ldm) {
  extern __shared__ __align__(16) double S[];  // [128][SLD] + dinv[128] + scratch[16][72]
  long long tl = 0, ts = 0, tu = 0, tc = 0;  // per-phase cycle totals (thread 0, only when stamps != nullptr)
  if (stamps && threadIdx.x == 0) stamps[0] = clock64();
  double* dinv = S + DB * SLD;
  double* scratch = dinv + DB;
  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, t = lane & 3;

  for (int idx = tid; idx < DB * DB; idx += 512) {
    const int r = idx >> 7, c = idx & 127;
    double v = 0.0;
    if (r < b && c <= r) v = A[(i64)r * lda + c];
    else if (r >= b && c == r) v = 1.0;  // identity padding keeps tail blocks on the same code path
    S[r * SLD + c] = v;
  }
  __syncthreads();
  if (stamps && threadIdx.x == 0) { stamps[1] = clock64(); tc = stamps[1]; }

  for (int p = 0; p < DB / LEAF; ++p) {
    const int o = p * LEAF;
    // ---- leaf: Cholesky of the 8x8 diagonal block, one thread, registers only
    if (tid == 0) {
      double a[LEAF][LEAF], rs[LEAF];
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
      for (int i = 0; i < LEAF; ++i) {
        dinv[o + i] = rs[i];
#pragma unroll
        for (int j = 0; j <= i; ++j) S[(o + i) * SLD + o + j] = a[i][j];
      }
    }
    __syncthreads();
    if (stamps && threadIdx.x == 0) { const long long tt = clock64(); tl += tt - tc; tc = tt; }
    // ---- panel: rows below the leaf, one thread per row: forward substitution against the leaf
    if (tid < DB && tid >= o + LEAF) {
      double* row = S + tid * SLD + o;
      double w[LEAF];
#pragma unroll
      for (int k = 0; k < LEAF; ++k) w[k] = row[k];
#pragma unroll
      for (int c = 0; c < LEAF; ++c) {
        w[c] *= dinv[o + c];
#pragma unroll
        for (int k = c + 1; k < LEAF; ++k) w[k] = fma(-w[c], S[(o + k) * SLD + o + c], w[k]);
      }
#pragma unroll
      for (int c = 0; c < LEAF; ++c) row[c] = w[c];
    }
    __syncthreads();
    if (stamps && threadIdx.x == 0) { const long long tt = clock64(); ts += tt - tc; tc = tt; }
    // ---- rank-8 update of the trailing block: 8x8 tiles on/below the diagonal, two DMMAs each
    {
      const int t0 = o + LEAF;
      const int nt = (DB - t0) >> 3;
      // static assignment without index decoding: warp pair (rp, rp + 8) shares tile rows rp and
      // nt-1-rp (together nt+1 tiles) and splits them by column parity; two tiles are in flight
      // at a time so their load -> DMMA -> DMMA -> store chains overlap
      const int rp = warp & 7, half = warp >> 3;
#pragma unroll 1
      for (int sel = 0; sel < 2; ++sel) {
        const int I = sel == 0 ? rp : nt - 1 - rp;
        if (I < 0 || I >= nt || (sel == 0 && 2 * rp > nt - 1) || (sel == 1 && I <= rp)) continue;
        const double* pa = S + (t0 + 8 * I + g) * SLD + o + t;
        const double a0 = -pa[0], a1 = -pa[4];
        for (int J = half; J <= I; J += 4) {
          const int J2 = J + 2;
          const bool two = J2 <= I;
          const double* pb = S + (t0 + 8 * J + g) * SLD + o + t;
          double* pc = S + (t0 + 8 * I + g) * SLD + t0 + 8 * J + 2 * t;
          const double* pb2 = two ? pb + 16 * SLD : pb;
          double* pc2 = two ? pc + 16 : pc;
          double c0 = pc[0], c1 = pc[1], e0 = pc2[0], e1 = pc2[1];
          const double b0 = pb[0], b1 = pb[4], f0 = pb2[0], f1 = pb2[4];
          dmma884(c0, c1, a0, b0);
          dmma884(e0, e1, a0, f0);
          dmma884(c0, c1, a1, b1);
          dmma884(e0, e1, a1, f1);
          pc[0] = c0;
          pc[1] = c1;
          if (two) {
            pc2[0] = e0;
            pc2[1] = e1;
          }
        }
      }
    }
    __syncthreads();
    if (stamps && threadIdx.x == 0) { const long long tt = clock64(); tu += tt - tc; tc = tt; }
  }
  if (stamps && threadIdx.x == 0) { stamps[2] = clock64(); stamps[5] = tl; stamps[6] = ts; stamps[7] = tu; }

  // ---- the 16 leaf inverses, one thread each (lane 0 of every warp), transposed into the upper triangle
  if (lane == 0) {
    const int o = warp * LEAF;
    double l[LEAF][LEAF], x[LEAF][LEAF], rs[LEAF];
#pragma unroll
    for (int i = 0; i < LEAF; ++i) {
      rs[i] = dinv[o + i];
#pragma unroll
      for (int j = 0; j < i; ++j) l[i][j] = S[(o + i) * SLD + o + j];
    }
#pragma unroll
    for (int j = 0; j < LEAF; ++j) {
      x[j][j] = rs[j];
#pragma unroll
      for (int i = j + 1; i < LEAF; ++i) {
        double acc = 0.0;
#pragma unroll
        for (int k = j; k < i; ++k) acc = fma(l[i][k], x[k][j], acc);
        x[i][j] = -acc * rs[i];
      }
    }
#pragma unroll
    for (int i = 1; i < LEAF; ++i)
#pragma unroll
      for (int j = 0; j < i; ++j) S[(o + j) * SLD + o + i] = x[i][j];
  }
  __syncthreads();

  // ---- inverse assembly: warp c owns block column c of X = inv(L); sweep the block rows below
  {
    const int c = warp;
    double* sc = scratch + c * 72;  // [8][9]
    for (int rb = c + 1; rb < DB / LEAF; ++rb) {
      double c0 = 0.0, c1 = 0.0, d0 = 0.0, d1 = 0.0;  // two accumulation chains (k-steps 0 / 1)
      const double* la = S + (rb * LEAF + g) * SLD + t;        // L[rb*8 + g][k]
      const double* xb = S + (c * LEAF + g) * SLD + t;         // X[k][c*8 + g] lives at S[c*8+g][k] for k > c*8+g
      {
        // kb == c: X[c][c] is the (lower triangular) leaf inverse, masked against the L entries
        // that share the square
        double bv0 = 0.0, bv1 = 0.0;
        if (t > g) bv0 = xb[c * LEAF];
        else if (t == g) bv0 = dinv[c * LEAF + g];
        if (4 + t > g) bv1 = xb[c * LEAF + 4];
        else if (4 + t == g) bv1 = dinv[c * LEAF + g];
        dmma884(c0, c1, la[c * LEAF], bv0);
        dmma884(d0, d1, la[c * LEAF + 4], bv1);
      }
#pragma unroll 4
      for (int kb = c + 1; kb < rb; ++kb) {
        dmma884(c0, c1, la[kb * LEAF], xb[kb * LEAF]);
        dmma884(d0, d1, la[kb * LEAF + 4], xb[kb * LEAF + 4]);
      }
      c0 += d0;
      c1 += d1;
      // T (C layout: row g, columns 2t, 2t+1) -> scratch, to be re-read as a B fragment
      sc[g * 9 + 2 * t] = c0;
      sc[g * 9 + 2 * t + 1] = c1;
      __syncwarp();
      // X[rb][c] = -inv(leaf_rb) * T ; inv(leaf_rb)[i][a] = S[rb*8+a][rb*8+i] (a < i), dinv on the diagonal
      double o0 = 0.0, o1 = 0.0;
#pragma unroll
      for (int ks = 0; ks < 2; ++ks) {
        const int a = 4 * ks + t;  // A fragment element (row g, k = a)
        double av = 0.0;
        if (a < g) av = S[(rb * LEAF + a) * SLD + rb * LEAF + g];
        else if (a == g) av = dinv[rb * LEAF + g];
        dmma884(o0, o1, av, sc[a * 9 + g]);
      }
      __syncwarp();
      S[(c * LEAF + 2 * t) * SLD + rb * LEAF + g] = -o0;
      S[(c * LEAF + 2 * t + 1) * SLD + rb * LEAF + g] = -o1;
      __syncwarp();
    }
  }
  __syncthreads();
  if (stamps && threadIdx.x == 0) stamps[3] = clock64();

  for (int idx = tid; idx < DB * DB; idx += 512) {
    const int r = idx >> 7, c = idx & 127;
    if (r < b && c <= r) {
      A[(i64)r * lda + c] = S[r * SLD + c];
      if (mirror) mirror[(i64)r * ldm + c] = S[r * SLD + c];
    }
    double li = 0.0;
    if (r < b && c < r) li = S[c * SLD + r];
    else if (r < b && c == r) li = dinv[r];
    Linv[idx] = li;
  }
  __syncthreads();
  if (stamps && threadIdx.x == 0) stamps[4] = clock64();
}

int potrf_diag(double* A, i64 lda, int b, double* Linv, int* info, int j0, cudaStream_t st, double* mirror, i64 ldm) {
  static bool configured[64] = {false};
  const int smem = (DB * SLD + DB + 16 * 72) * (int)sizeof(double);
  if (first_use_on_device(configured)) {
    STPYB_CUDA(cudaFuncSetAttribute(potrf_diag_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  }
  prof_begin(PROF_DIAG, (double)b * b * b / 3.0, st);
  potrf_diag_kernel<<<1, 512, smem, st>>>(A, lda, b, Linv, info, j0, nullptr, mirror, ldm);
  prof_end(st);
  STPYB_COUNT_LAUNCH();
  STPYB_CUDA(cudaGetLastError());
  return 0;
}

// C[M x N] = alpha * A B^T + beta * C, generic entry used by every blocked stage.
int gemm_nt(int M, int N, int K, const double* A, i64 lda, const double* B, i64 ldb, double* C, i64 ldc,
            double alpha, double beta, int tri, int square_cfg, cudaStream_t st, int kskip, double* mirror, i64 ldm) {
  GemmArgs g;
  g.A = A; g.B = B; g.lda = lda; g.ldb = ldb; g.M = M; g.N = N; g.K = K; g.tri = tri; g.kskip = kskip;
  if (beta == 1.0 && (alpha == 1.0 || alpha == -1.0) && !square_cfg) {
    EpiAccum e;
    e.C = C; e.ldc = ldc; e.negate = (alpha < 0.0) ? 1 : 0;
    e.vec = ((ldc & 1) == 0 && (((uintptr_t)C) & 15) == 0) ? 1 : 0;
    // 32-wide K slices, 2-stage ring: measured 33.4 vs 31.6 TFLOP/s for the K = 512 trailing update
    // against the 16-wide / 4-stage tile (profiles/gemm_cfg_sweep_r01.txt)
    static int use_tma = -1;  // STPYB_TMA=0 selects the cp.async (LDGSTS) staging instead of TMA
    if (use_tma < 0) {
      const char* ev = getenv("STPYB_TMA");
      use_tma = ev ? atoi(ev) : 0;
    }
    if (use_tma && K >= 4) {
      const int rc = launch_gemm_nt_tma<CfgStreamK32, EpiAccum>(g, e, st);
      if (rc != -20) return rc;  // -20: tensor-map encoding unavailable -> LDGSTS staging
    }
    return launch_gemm_nt<CfgStreamK32, EpiAccum>(g, e, st);
  }
  EpiAxpby e = make_axpby(C, ldc, alpha, beta);
  e.C2 = mirror;
  e.ldc2 = ldm;
  if (square_cfg) return launch_gemm_nt<CfgSquare, EpiAxpby>(g, e, st);
  return launch_gemm_nt<CfgStream, EpiAxpby>(g, e, st);
}

// Factor the tall panel P (rows x w, top w x w block = diagonal block) in place.  Left-looking
// inside the panel: block column j first receives the contribution of ALL previous block columns
// in one update of depth K = j (instead of j/128 rank-128 updates, each re-reading and re-writing
// the same C tiles), then its diagonal block is factored and the rows below are solved.
int potrf_panel(double* P, i64 rows, int w, i64 ldp, double* dinv, int* info, i64 j0, cudaStream_t st, double* pack,
                i64 ldpack) {
  // pack != nullptr: every final entry of the factored panel (diagonal blocks, solved rows) is stored a second
  // time at pack[r * ldpack + c] -- the contiguous buffer the distributed schedule broadcasts -- by the kernels
  // that produce it, instead of a separate copy pass over the panel afterwards
  for (int j = 0; j < w; j += DB) {
    const int b = (w - j < DB) ? (w - j) : DB;
    double* Pjj = P + (i64)j * ldp + j;
    if (j > 0) {
      const double* Prow = P + (i64)j * ldp;  // rows j.., columns 0..j of the panel
      prof_begin(PROF_PANEL_UPD, 2.0 * ((double)(rows - j) * b - 0.5 * (double)b * b) * j, st);
      STPYB_TRY(gemm_nt((int)(rows - j), b, j, Prow, ldp, Prow, ldp, Pjj, ldp, -1.0, 1.0, TRI_LOWER, 0, st));
      prof_end(st);
    }
    double* Li = dinv + (i64)(j / DB) * (DB * DB);
    STPYB_TRY(potrf_diag(Pjj, ldp, b, Li, info, (int)(j0 + j), st, pack ? pack + (i64)j * ldpack + j : nullptr, ldpack));
    const i64 below = rows - (j + b);
    if (below > 0) {
      // in-place panel TRSM: rows below the diagonal block
      double* P21 = P + (i64)(j + b) * ldp + j;
      prof_begin(PROF_TRSM, (double)below * b * b, st);
      STPYB_TRY(gemm_nt((int)below, b, b, P21, ldp, Li, DB, P21, ldp, 1.0, 0.0, TRI_FULL, 1, st, 0,
                        pack ? pack + (i64)(j + b) * ldpack + j : nullptr, ldpack));
      prof_end(st);
    }
  }
  return 0;
}

// Look-ahead: a high-priority side stream factors panel J+1 while the main stream applies panel J
// to the columns right of it.  One side stream and two events per device, created on first use.
struct LookAhead {
  cudaStream_t side = nullptr;
  cudaEvent_t cols_ready = nullptr, panel_done = nullptr;
  int state = 0;  // 0 untried, 1 ready, -1 unavailable
};

static LookAhead* lookahead_for_current_device() {
  static LookAhead tab[64];
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
  LookAhead& la = tab[dev];
  if (la.state == 0) {
    int lo = 0, hi = 0;
    la.state = -1;
    if (cudaDeviceGetStreamPriorityRange(&lo, &hi) == cudaSuccess &&
        cudaStreamCreateWithPriority(&la.side, cudaStreamNonBlocking, hi) == cudaSuccess &&
        cudaEventCreateWithFlags(&la.cols_ready, cudaEventDisableTiming) == cudaSuccess &&
        cudaEventCreateWithFlags(&la.panel_done, cudaEventDisableTiming) == cudaSuccess)
      la.state = 1;
  }
  return la.state == 1 ? &la : nullptr;
}

static i64 g_lookahead_min_n = -2;  // -2: not read yet; -1: overlap disabled
static i64 lookahead_min_n() {
  if (g_lookahead_min_n == -2) {
    const char* ev = getenv("STPYB_LOOKAHEAD_MIN_N");  // 0 = always, negative = never
    g_lookahead_min_n = ev ? atoll(ev) : 4096;
    if (g_lookahead_min_n < 0) g_lookahead_min_n = -1;
  }
  return g_lookahead_min_n;
}

static int syrk_lower(const double* P, double* C, i64 m, i64 ncols, int k, i64 lda, cudaStream_t st) {
  // C (m x ncols, the first ncols columns of a lower-triangular trailing matrix) -= P P^T, lower tiles only
  prof_begin(PROF_SYRK, (2.0 * (double)m * (double)ncols - (double)ncols * (double)ncols) * k, st);
  const int rc = gemm_nt((int)m, (int)ncols, k, P, lda, P, lda, C, lda, -1.0, 1.0, TRI_LOWER, 0, st);
  prof_end(st);
  return rc;
}

int potrf_lower(double* A, i64 n, i64 lda, double* dinv, int* info, int outer, cudaStream_t st) {
  if (n <= 0) return 0;
  if ((lda & 1) || (((uintptr_t)A) & 15) || (((uintptr_t)dinv) & 15)) return -3;
  // STPYB_POTRF_NO_LOOKAHEAD in outer_block: this call keeps everything on the caller's stream (the per-device
  // side stream and its two events are shared, so callers that already overlap several factorisations
  // stream against stream -- the hyper-parameter sweep -- opt out per call instead of flipping a global)
  const bool no_la = (outer & STPYB_POTRF_NO_LOOKAHEAD) != 0;
  outer &= ~STPYB_POTRF_NO_LOOKAHEAD;
  if (outer < DB) outer = DB;
  outer = (outer / DB) * DB;
  STPYB_CUDA(cudaMemsetAsync(info, 0, sizeof(int), st));
  const i64 min_n = lookahead_min_n();
  LookAhead* la = (!no_la && min_n >= 0 && n >= min_n && n > 2 * (i64)outer) ? lookahead_for_current_device() : nullptr;
  if (la == nullptr) {
    for (i64 J = 0; J < n; J += outer) {
      const int jb = (int)((n - J < outer) ? (n - J) : outer);
      STPYB_TRY(potrf_panel(A + J * lda + J, n - J, jb, lda, dinv + (J / DB) * (i64)(DB * DB), info, J, st));
      const i64 trail = n - (J + jb);
      if (trail > 0) STPYB_TRY(syrk_lower(A + (J + jb) * lda + J, A + (J + jb) * lda + (J + jb), trail, trail, jb, lda, st));
    }
    return 0;
  }
  // Panel J is factored when iteration J starts.  Its update is split in two: the columns of panel J+1
  // first; then panel J+1 is factored on the side stream while the main stream updates the rest.
  STPYB_TRY(potrf_panel(A, n, (int)((n < outer) ? n : outer), lda, dinv, info, 0, st));
  for (i64 J = 0; J < n; J += outer) {
    const int jb = (int)((n - J < outer) ? (n - J) : outer);
    const i64 J1 = J + jb, trail = n - J1;
    if (trail <= 0) break;
    const int nb = (int)((trail < outer) ? trail : outer);
    const i64 rest = trail - nb;
    STPYB_TRY(syrk_lower(A + J1 * lda + J, A + J1 * lda + J1, trail, nb, jb, lda, st));
    double* dinv1 = dinv + (J1 / DB) * (i64)(DB * DB);
    if (rest > 0) {
      STPYB_CUDA(cudaEventRecord(la->cols_ready, st));
      STPYB_CUDA(cudaStreamWaitEvent(la->side, la->cols_ready, 0));
      STPYB_TRY(potrf_panel(A + J1 * lda + J1, trail, nb, lda, dinv1, info, J1, la->side));
      STPYB_CUDA(cudaEventRecord(la->panel_done, la->side));
      const i64 J2 = J1 + nb;
      STPYB_TRY(syrk_lower(A + J2 * lda + J, A + J2 * lda + J2, rest, rest, jb, lda, st));
      STPYB_CUDA(cudaStreamWaitEvent(st, la->panel_done, 0));
    } else {
      STPYB_TRY(potrf_panel(A + J1 * lda + J1, trail, nb, lda, dinv1, info, J1, st));
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

extern "C" int stpyb_set_lookahead_min_n(long long min_n, long long* old_or_null) {
  const i64 old = lookahead_min_n();
  if (old_or_null) *old_or_null = old;
  g_lookahead_min_n = (min_n < 0) ? -1 : min_n;
  return 0;
}

extern "C" int stpyb_potrf_diag_profile(double* A, long long lda, int b, double* Linv, int* info_dev,
                                        long long* stamps8_dev, void* stream) {
  const int smem = (DB * SLD + DB + 16 * 72) * (int)sizeof(double);
  STPYB_CUDA(cudaFuncSetAttribute(potrf_diag_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  potrf_diag_kernel<<<1, 512, smem, (cudaStream_t)stream>>>(A, lda, b, Linv, info_dev, 0, stamps8_dev, nullptr, 0);
  STPYB_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int stpyb_potrf_panel(double* P, long long rows, int w, long long ldp, double* dinv, int* info_dev,
                                 long long j0, double* pack_or_null, long long ldpack, void* stream) {
  if (rows < w || w <= 0) return -2;
  if ((ldp & 1) || (((uintptr_t)P) & 15)) return -4;
  if (pack_or_null && ldpack < w) return -10;
  return potrf_panel(P, rows, w, ldp, dinv, info_dev, j0, (cudaStream_t)stream, pack_or_null, ldpack);
}

extern "C" int stpyb_gemm_nt(int M, int N, int K, const double* A, long long lda, const double* B,
                             long long ldb, double* C, long long ldc, double alpha, double beta, int lower,
                             void* stream) {
  return gemm_nt(M, N, K, A, lda, B, ldb, C, ldc, alpha, beta, lower ? TRI_LOWER : TRI_FULL, 0,
                 (cudaStream_t)stream);
}

// A batch of independent C_i = alpha A_i B_i^T + beta C_i updates (same K and leading dimensions)
// issued with ONE call: fork from `main_stream` onto up to `nside` side streams (round robin) and
// join back, so that the last partial wave of one launch overlaps the first of the next and the
// host crosses the FFI once per factorisation step.  Used by the multi-GPU trailing update, where
// a rank owns several block columns per step.
extern "C" int stpyb_gemm_nt_batch(int count, const int* M, const int* N, int K, const double* const* A,
                                   long long lda, const double* const* B, long long ldb, double* const* C,
                                   long long ldc, double alpha, double beta, int lower, void* main_stream,
                                   void* const* side_streams, int nside) {
  if (count <= 0) return 0;
  cudaStream_t mainst = (cudaStream_t)main_stream;
  const int tri = lower ? TRI_LOWER : TRI_FULL;
  if (count == 1 || nside <= 0) {
    for (int i = 0; i < count; ++i)
      STPYB_TRY(gemm_nt(M[i], N[i], K, A[i], lda, B[i], ldb, C[i], ldc, alpha, beta, tri, 0, mainst));
    return 0;
  }
  // fork / join events belong to a device: one set per device, created on first use there
  struct BatchEvents { cudaEvent_t fork = nullptr, join[8] = {nullptr}; };
  static BatchEvents evtab[64];
  int dev = 0;
  STPYB_CUDA(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64) return -14;
  cudaEvent_t& fork_ev = evtab[dev].fork;
  cudaEvent_t* join_ev = evtab[dev].join;
  if (nside > 8) nside = 8;
  if (!fork_ev) {
    STPYB_CUDA(cudaEventCreateWithFlags(&fork_ev, cudaEventDisableTiming));
    for (int s = 0; s < 8; ++s) STPYB_CUDA(cudaEventCreateWithFlags(&join_ev[s], cudaEventDisableTiming));
  }
  const int used = count < nside ? count : nside;
  STPYB_CUDA(cudaEventRecord(fork_ev, mainst));
  for (int s = 0; s < used; ++s) STPYB_CUDA(cudaStreamWaitEvent((cudaStream_t)side_streams[s], fork_ev, 0));
  for (int i = 0; i < count; ++i)
    STPYB_TRY(gemm_nt(M[i], N[i], K, A[i], lda, B[i], ldb, C[i], ldc, alpha, beta, tri, 0,
                      (cudaStream_t)side_streams[i % used]));
  for (int s = 0; s < used; ++s) {
    STPYB_CUDA(cudaEventRecord(join_ev[s], (cudaStream_t)side_streams[s]));
    STPYB_CUDA(cudaStreamWaitEvent(mainst, join_ev[s], 0));
  }
  return 0;
}
