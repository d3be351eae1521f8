// Peer-memory plumbing and the fused "solve + publish" hop of the distributed backward sweep.
//
// One process per GPU: each rank cudaMalloc's a symmetric buffer, exports it with
// cudaIpcGetMemHandle, and maps its peers' buffers (cudaIpcOpenMemHandle).  Over NVLink /
// NVSwitch a kernel on one GPU can then store straight into every peer's HBM.  The backward
// sweep alpha = L^-T z is a chain of n/nbw tiny dependent steps whose owner changes every hop;
// with NCCL each hop pays a collective launch + rendezvous (~100 us).  Here the owner's last
// kernel of the hop writes the new alpha segment into ALL ranks' buffers and raises a flag
// there (st.global on peer pointers + __threadfence_system), and the next owner's stream simply
// waits on its local flag -- compute and "broadcast" are one kernel, no collective call.
#include "common.cuh"
#include "stpyb_internal.h"
#include "../../include/stpyb.h"

namespace stpyb {

struct PeerTable {
  double* alpha[STPYB_MAX_PEERS];
  int* flags[STPYB_MAX_PEERS];
  int world;
};

// seg (w doubles, already final in the LOCAL alpha buffer at offset off) -> every peer, then flag[g] = epoch
__global__ void __launch_bounds__(512) alpha_publish_kernel(PeerTable pt, int self, long long off, int w, int g,
                                                           int epoch) {
  const double* src = pt.alpha[self] + off;
  for (int p = 0; p < pt.world; ++p) {
    if (p == self) continue;
    double* dst = pt.alpha[p] + off;
    for (int i = threadIdx.x; i < w; i += blockDim.x) dst[i] = src[i];
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x < pt.world) {
    volatile int* f = pt.flags[threadIdx.x] + g;
    *f = epoch;
  }
}

// spin until flags[lo..hi) of the LOCAL flag array carry `epoch`; bounded: after `limit` cycles the
// kernel gives up and records the failure in *err (a lost peer must not hang the GPU)
__global__ void alpha_wait_kernel(const int* flags, int lo, int hi, int epoch, long long limit, int* err) {
  const long long t0 = clock64();
  for (int g = lo + (int)threadIdx.x; g < hi; g += blockDim.x) {
    const volatile int* f = flags + g;
    while (*f != epoch) {
      if (clock64() - t0 > limit) {
        atomicExch(err, 1 + g);
        return;
      }
      __nanosleep(64);
    }
  }
  __threadfence_system();
}

}  // namespace stpyb

using namespace stpyb;

extern "C" int stpyb_p2p_alloc(long long bytes, void** dev_ptr_out, void* ipc_handle_64) {
  void* p = nullptr;
  STPYB_CUDA(cudaMalloc(&p, (size_t)bytes));
  STPYB_CUDA(cudaMemset(p, 0, (size_t)bytes));
  cudaIpcMemHandle_t h;
  STPYB_CUDA(cudaIpcGetMemHandle(&h, p));
  static_assert(sizeof(h) == 64, "ipc handle size");
  memcpy(ipc_handle_64, &h, 64);
  *dev_ptr_out = p;
  return 0;
}

extern "C" int stpyb_p2p_open(const void* ipc_handle_64, void** dev_ptr_out) {
  cudaIpcMemHandle_t h;
  memcpy(&h, ipc_handle_64, 64);
  void* p = nullptr;
  STPYB_CUDA(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
  *dev_ptr_out = p;
  return 0;
}

extern "C" int stpyb_p2p_close(void* dev_ptr) {
  STPYB_CUDA(cudaIpcCloseMemHandle(dev_ptr));
  return 0;
}

extern "C" int stpyb_p2p_free(void* dev_ptr) {
  STPYB_CUDA(cudaFree(dev_ptr));
  return 0;
}

extern "C" int stpyb_p2p_alpha_publish(void* const* peer_alpha, void* const* peer_flags, int world, int self,
                                       long long off, int w, int g, int epoch, void* stream) {
  if (world < 1 || world > STPYB_MAX_PEERS) return -3;
  PeerTable pt;
  pt.world = world;
  for (int p = 0; p < world; ++p) {
    pt.alpha[p] = (double*)peer_alpha[p];
    pt.flags[p] = (int*)peer_flags[p];
  }
  alpha_publish_kernel<<<1, 512, 0, (cudaStream_t)stream>>>(pt, self, off, w, g, epoch);
  STPYB_COUNT_LAUNCH();
  STPYB_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int stpyb_p2p_wait_flags(const int* local_flags, int lo, int hi, int epoch, long long limit_cycles,
                                    int* err_dev, void* stream) {
  if (hi <= lo) return 0;
  alpha_wait_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(local_flags, lo, hi, epoch, limit_cycles, err_dev);
  STPYB_COUNT_LAUNCH();
  STPYB_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int stpyb_memcpy_d2d(void* dst, const void* src, long long bytes, void* stream) {
  STPYB_CUDA(cudaMemcpyAsync(dst, src, (size_t)bytes, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
  return 0;
}

namespace stpyb {

// Right-looking step of the distributed backward sweep on EVERY rank: once alpha_g (bw values at
// seg) is visible (flag[g] == epoch; pass flags == nullptr on the rank that just produced it), fold
// it into the pending right-hand sides of all local block columns left of g:
//   zrow[c] -= sum_r L[r][c] * seg[r]     r < bw (rows of block g), c < ncols.
// Each thread owns two adjacent columns and a quarter of the rows (cf. trsv_bwd_update_kernel).
__global__ void __launch_bounds__(256) dist_strip_kernel(const double* __restrict__ Lstrip, i64 ld, int bw, i64 ncols,
                                                        const double* seg, double* zrow, const int* flags, int g,
                                                        int epoch, long long limit, int* err) {
  __shared__ double xs[1024];
  __shared__ double2 part[8][32];
  __shared__ int give_up;
  if (threadIdx.x == 0) {
    give_up = 0;
    if (flags) {
      const volatile int* f = flags + g;
      const long long t0 = clock64();
      while (*f != epoch) {
        if (clock64() - t0 > limit) {
          atomicExch(err, 1 + g);
          give_up = 1;
          break;
        }
        __nanosleep(32);
      }
      __threadfence_system();
    }
  }
  __syncthreads();
  if (give_up) return;
  for (int i = threadIdx.x; i < 1024; i += 256) xs[i] = (i < bw) ? ((const volatile double*)seg)[i] : 0.0;
  __syncthreads();
  // a CTA owns 64 columns (32 lanes x 2) and splits the bw rows over its 8 warps; every lane keeps 8 independent
  // 16-byte loads in flight (the first version had 2 and was bound by the L2 round trip: ~64 dependent trips per hop)
  const int cg = threadIdx.x & 31, rg = threadIdx.x >> 5;
  const i64 c = ((i64)blockIdx.x * 32 + cg) * 2;  // ncols is a multiple of 128
  double2 s = make_double2(0.0, 0.0);
  if (c < ncols) {
    const double* p = Lstrip + c;
    for (int r0 = rg; r0 < bw; r0 += 64) {
      double2 a[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int r = r0 + 8 * u;
        a[u] = (r < bw) ? *reinterpret_cast<const double2*>(p + (i64)r * ld) : make_double2(0.0, 0.0);
      }
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int r = r0 + 8 * u;
        const double xv = (r < bw) ? xs[r] : 0.0;
        s.x = fma(a[u].x, xv, s.x);
        s.y = fma(a[u].y, xv, s.y);
      }
    }
  }
  part[rg][cg] = s;
  __syncthreads();
  if (rg == 0 && c < ncols) {
    double2 t = part[0][cg];
#pragma unroll
    for (int q = 1; q < 8; ++q) {
      t.x += part[q][cg].x;
      t.y += part[q][cg].y;
    }
    double2* pz = reinterpret_cast<double2*>(zrow + c);
    double2 v = *pz;
    v.x -= t.x;
    v.y -= t.y;
    *pz = v;
  }
}

// Owner of block column g (width w <= 512): alpha_g = L_gg^-T z_g with the inverted 128-blocks, in
// ONE CTA, then the segment is stored into every rank's symmetric alpha buffer and flag[g] raised
// (the "broadcast" is the tail of the solve kernel).
__global__ void __launch_bounds__(1024, 1) dist_solve_publish_kernel(const double* __restrict__ Lgg, i64 ld, int w,
                                                                    const double* __restrict__ dinv,
                                                                    const double* zrow, PeerTable pt, int self,
                                                                    long long off, int g, int epoch) {
  __shared__ double xs[1024];
  __shared__ double part[8][DB];
  const int tid = threadIdx.x;
  xs[tid] = (tid < w) ? zrow[tid] : 0.0;
  __syncthreads();
  const int nq = (w + DB - 1) / DB;
  for (int q = nq - 1; q >= 0; --q) {
    const int bq = (w - q * DB < DB) ? (w - q * DB) : DB;
    {  // xs_q <- Linv_q^T xs_q
      const double* Li = dinv + (i64)q * (DB * DB);
      const int rg = tid >> 7, j = tid & 127;
      double s = 0.0;
#pragma unroll 4
      for (int i = rg; i < DB; i += 8) s = fma(Li[i * DB + j], xs[q * DB + i], s);
      part[rg][j] = s;
      __syncthreads();
      if (tid < DB) {
        double t = 0.0;
#pragma unroll
        for (int r = 0; r < 8; ++r) t += part[r][tid];
        xs[q * DB + tid] = (tid < bq) ? t : 0.0;
      }
      __syncthreads();
    }
    if (q > 0) {
      // xs[c] -= sum_r L[q*128 + r][c] xs_q[r] for c < q*128: 8 row groups x 128 columns per pass, 16 independent
      // loads per thread (one thread per column with 2 loads in flight cost ~40 us per 512-block)
      const int rg = tid >> 7, c = tid & 127;
      for (int cb = 0; cb < q * DB; cb += DB) {
        const double* p = Lgg + (i64)(q * DB) * ld + cb + c;
        double v[16];
#pragma unroll
        for (int u = 0; u < 16; ++u) {
          const int r = rg + 8 * u;
          v[u] = (r < bq) ? p[(i64)r * ld] : 0.0;
        }
        double s0 = 0.0, s1 = 0.0;
#pragma unroll
        for (int u = 0; u < 16; u += 2) {
          s0 = fma(v[u], xs[q * DB + rg + 8 * u], s0);
          s1 = fma(v[u + 1], xs[q * DB + rg + 8 * (u + 1)], s1);
        }
        part[rg][c] = s0 + s1;
        __syncthreads();
        if (tid < DB) {
          double t = 0.0;
#pragma unroll
          for (int r = 0; r < 8; ++r) t += part[r][tid];
          xs[cb + tid] -= t;
        }
        __syncthreads();
      }
    }
  }
  // publish: local copy first, then every peer, then the flags
  for (int p = 0; p < pt.world; ++p) {
    double* dst = pt.alpha[p] + off;
    if (tid < w) dst[tid] = xs[tid];
  }
  __threadfence_system();
  __syncthreads();
  if (tid < pt.world) {
    volatile int* f = pt.flags[tid] + g;
    *f = epoch;
  }
}

}  // namespace stpyb

extern "C" int stpyb_dist_strip(const double* Lstrip, long long ld, int bw, long long ncols, const double* seg,
                                double* zrow, const int* flags_or_null, int g, int epoch, long long limit_cycles,
                                int* err_dev, void* stream) {
  if (ncols <= 0 || bw <= 0) {
    // nothing to fold in, but the stream must still observe the flag before later readers of seg
    if (flags_or_null) return stpyb_p2p_wait_flags(flags_or_null, g, g + 1, epoch, limit_cycles, err_dev, stream);
    return 0;
  }
  if (bw > 1024 || (ld & 1) || (ncols & 127)) return -3;
  stpyb::dist_strip_kernel<<<(unsigned)((ncols + 63) / 64), 256, 0, (cudaStream_t)stream>>>(
      Lstrip, ld, bw, ncols, seg, zrow, flags_or_null, g, epoch, limit_cycles, err_dev);
  STPYB_COUNT_LAUNCH();
  STPYB_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int stpyb_dist_solve_publish(const double* Lgg, long long ld, int w, const double* dinv, const double* zrow,
                                        void* const* peer_alpha, void* const* peer_flags, int world, int self,
                                        long long off, int g, int epoch, void* stream) {
  if (w <= 0 || w > 1024) return -3;
  if (world < 1 || world > STPYB_MAX_PEERS) return -8;
  stpyb::PeerTable pt;
  pt.world = world;
  for (int p = 0; p < world; ++p) {
    pt.alpha[p] = (double*)peer_alpha[p];
    pt.flags[p] = (int*)peer_flags[p];
  }
  stpyb::dist_solve_publish_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(Lgg, ld, w, dinv, zrow, pt, self, off, g,
                                                                       epoch);
  STPYB_COUNT_LAUNCH();
  STPYB_CUDA(cudaGetLastError());
  return 0;
}
