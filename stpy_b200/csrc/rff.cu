// Random-Fourier-feature embedding and the streamed normal equations of the
// feature-space Bayesian linear regression.
//
// Replaces RFFEmbedding.embed (stpy/embeddings/embedding.py:225-241: W.mm(x^T),
// cos / sin, cat, transpose — four n*m temporaries) with one DMMA projection
// whose accumulator goes through a trigonometric register epilogue, and
// KernelizedFeatures.precompute / theta_mean
// (stpy/continuous_processes/kernelized_features.py:228, 237, 256: Q = embed(x),
// Q.T @ Q, Q.T @ y) with a chunked  embed^T -> SYRK  stream that never stores the
// n x m feature matrix.
#include "gemm_nt.cuh"
#include "stpyb_internal.h"
#include "../../include/stpyb.h"

namespace stpyb {

// FEAT_ROW = 1: tile rows index features (transposed output, m x n); 0: tile columns do (n x m).
template <int FEAT_ROW>
struct EpiRff {
  static constexpr bool kPreload = false;
  static constexpr bool kRowBatch = true;
  __device__ __forceinline__ void preload(int, int, int, double&, double&) const {}
  __device__ __forceinline__ void preload_finish(double&, double&) const {}
  const double* bias;
  const double* featw;
  int mode;  // 0: split cos|sin, 1: cos(. + bias)
  int m;
  double scale;
  double* C;
  i64 ldc;
  int vec;
  __device__ __forceinline__ double one(int f, double v) const {
    double r;
    if (mode == 0) {
      r = (f < (m >> 1)) ? cos(v) : sin(v);
    } else {
      r = cos(bias ? v + bias[f] : v);
    }
    r *= scale;
    if (featw) r *= featw[f];
    return r;
  }
  // not inlined: the sin / cos expansions are large, one copy serves all accumulator rows
  __device__ __noinline__ void apply_row(int row, int col_base, int N, double v0, double v1, double v2, double v3,
                                         double v4, double v5, double v6, double v7) const {
    constexpr int NI = 4;
    const double acc[NI][2] = {{v0, v1}, {v2, v3}, {v4, v5}, {v6, v7}};
    double o[NI][2];
#pragma unroll
    for (int j = 0; j < NI; ++j) {
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int col = col_base + j * 8 + e;
        o[j][e] = (col < N) ? one(FEAT_ROW ? row : col, acc[j][e]) : 0.0;
      }
    }
#pragma unroll
    for (int j = 0; j < NI; ++j) {
      const int col = col_base + j * 8;
      if (col >= N) continue;
      double* p = C + (i64)row * ldc + col;
      if (vec && col + 1 < N) {
        *reinterpret_cast<double2*>(p) = make_double2(o[j][0], o[j][1]);
      } else {
        p[0] = o[j][0];
        if (col + 1 < N) p[1] = o[j][1];
      }
    }
  }
};

int rff_embed(const double* Xp, i64 n, const double* Wp, int m, int dpad, const double* bias, const double* featw,
              int mode, double scale, int transposed, double* Phi, i64 ldphi, cudaStream_t st) {
  if (n <= 0 || m <= 0) return 0;
  if (mode != 0 && mode != 1) return -8;
  GemmArgs g;
  g.lda = dpad; g.ldb = dpad; g.K = dpad; g.tri = TRI_FULL; g.kskip = 0;
  const int vec = ((ldphi & 1) == 0 && (((uintptr_t)Phi) & 15) == 0) ? 1 : 0;
  if (transposed) {
    g.A = Wp; g.B = Xp; g.M = m; g.N = (int)n;
    EpiRff<1> e{bias, featw, mode, m, scale, Phi, ldphi, vec};
    return launch_gemm_nt<CfgStream, EpiRff<1>>(g, e, st);
  }
  g.A = Xp; g.B = Wp; g.M = (int)n; g.N = m;
  EpiRff<0> e{bias, featw, mode, m, scale, Phi, ldphi, vec};
  return launch_gemm_nt<CfgStream, EpiRff<0>>(g, e, st);
}

}  // namespace stpyb

using namespace stpyb;

extern "C" int stpyb_rff_embed(const double* Xp, long long n, const double* Wp, int m, int dpad,
                               const double* bias_or_null, const double* featw_or_null, int mode, double scale,
                               int transposed_out, double* Phi, long long ldphi, void* stream) {
  return rff_embed(Xp, n, Wp, m, dpad, bias_or_null, featw_or_null, mode, scale, transposed_out, Phi, ldphi,
                   (cudaStream_t)stream);
}

namespace stpyb {

// out[i] += sum_j M[i][j] v[j]; one CTA per row.  Accumulates Phi^T y next to the SYRK.
__global__ void __launch_bounds__(256) gemv_rows_acc_kernel(const double* __restrict__ M, i64 cols, i64 ldm,
                                                           const double* __restrict__ v, double* out) {
  __shared__ double red[8];
  const double* row = M + (i64)blockIdx.x * ldm;
  double s = 0.0;
  for (i64 j = threadIdx.x; j < cols; j += 256) s = fma(row[j], v[j], s);
  s = block_sum<256>(s, red);
  if (threadIdx.x == 0) out[blockIdx.x] += s;
}

// Side stream + events for the embed / SYRK pipeline, one set per device, created on first use.
struct EmbedPipe {
  cudaStream_t side = nullptr;
  cudaEvent_t start = nullptr, filled[2] = {nullptr, nullptr}, consumed[2] = {nullptr, nullptr};
  int state = 0;  // 0 untried, 1 ready, -1 unavailable
};

static EmbedPipe* embed_pipe_for_current_device() {
  static EmbedPipe tab[64];
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
  EmbedPipe& p = tab[dev];
  if (p.state == 0) {
    int lo = 0, hi = 0;
    p.state = -1;
    bool ok = cudaDeviceGetStreamPriorityRange(&lo, &hi) == cudaSuccess &&
              cudaStreamCreateWithPriority(&p.side, cudaStreamNonBlocking, hi) == cudaSuccess &&
              cudaEventCreateWithFlags(&p.start, cudaEventDisableTiming) == cudaSuccess;
    for (int b = 0; ok && b < 2; ++b)
      ok = cudaEventCreateWithFlags(&p.filled[b], cudaEventDisableTiming) == cudaSuccess &&
           cudaEventCreateWithFlags(&p.consumed[b], cudaEventDisableTiming) == cudaSuccess;
    if (ok) p.state = 1;
  }
  return p.state == 1 ? &p : nullptr;
}

}  // namespace stpyb

// V (lower tiles of the leading m x m block) += Phi^T Phi and V[m][0..m) += Phi^T y, Phi never stored in
// full: the rows are embedded (transposed) half a chunk at a time into the two halves of `scratch` on a side
// stream while the main stream contracts the previous half.  V[m][m] is not touched.
extern "C" int stpyb_rff_normal_eq(const double* Xp, const double* y, long long n, const double* Wp, int m,
                                   int dpad, const double* bias_or_null, const double* featw_or_null, int mode,
                                   double scale, long long chunk, double* scratch, long long ldscratch, double* V,
                                   long long ldv, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  if (n <= 0) return 0;
  if (chunk <= 0 || ldscratch < chunk || (ldscratch & 1)) return -11;
  if (ldv < m + 1 || (ldv & 1)) return -15;
  double* vy = V + (long long)m * ldv;  // row m of V: Phi^T y
  long long half = (chunk / 2 / 64) * 64;
  EmbedPipe* pipe = (half >= 1024 && n > half) ? embed_pipe_for_current_device() : nullptr;
  if (pipe == nullptr) {
    for (long long r0 = 0; r0 < n; r0 += chunk) {
      const long long len = (n - r0 < chunk) ? (n - r0) : chunk;
      STPYB_TRY(rff_embed(Xp + r0 * dpad, len, Wp, m, dpad, bias_or_null, featw_or_null, mode, scale, 1, scratch,
                          ldscratch, st));
      gemv_rows_acc_kernel<<<(unsigned)m, 256, 0, st>>>(scratch, len, ldscratch, y + r0, vy);
      STPYB_COUNT_LAUNCH();
      STPYB_CUDA(cudaGetLastError());
      STPYB_TRY(gemm_nt(m, m, (int)len, scratch, ldscratch, scratch, ldscratch, V, ldv, 1.0, 1.0, TRI_LOWER, 0, st));
    }
    return 0;
  }
  STPYB_CUDA(cudaEventRecord(pipe->start, st));
  STPYB_CUDA(cudaStreamWaitEvent(pipe->side, pipe->start, 0));
  long long it = 0;
  for (long long r0 = 0; r0 < n; r0 += half, ++it) {
    const long long len = (n - r0 < half) ? (n - r0) : half;
    const int b = (int)(it & 1);
    double* buf = scratch + (long long)b * half;
    if (it >= 2) STPYB_CUDA(cudaStreamWaitEvent(pipe->side, pipe->consumed[b], 0));
    STPYB_TRY(rff_embed(Xp + r0 * dpad, len, Wp, m, dpad, bias_or_null, featw_or_null, mode, scale, 1, buf,
                        ldscratch, pipe->side));
    STPYB_CUDA(cudaEventRecord(pipe->filled[b], pipe->side));
    STPYB_CUDA(cudaStreamWaitEvent(st, pipe->filled[b], 0));
    gemv_rows_acc_kernel<<<(unsigned)m, 256, 0, st>>>(buf, len, ldscratch, y + r0, vy);
    STPYB_COUNT_LAUNCH();
    STPYB_CUDA(cudaGetLastError());
    STPYB_TRY(gemm_nt(m, m, (int)len, buf, ldscratch, buf, ldscratch, V, ldv, 1.0, 1.0, TRI_LOWER, 0, st));
    STPYB_CUDA(cudaEventRecord(pipe->consumed[b], st));
  }
  return 0;
}
