// Prints how TMA swizzle modes place a {16 doubles x 32 rows} box in shared memory.
// nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tma_swizzle_probe tma_swizzle_probe.cu
#include <cstdio>
#include <cuda.h>
#include <cuda_runtime.h>
typedef CUresult (*Enc)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                        const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                        CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
__global__ void k(const __grid_constant__ CUtensorMap map, double* out) {
  extern __shared__ __align__(1024) double sm[];
  __shared__ __align__(8) unsigned long long bar;
  unsigned b = (unsigned)__cvta_generic_to_shared(&bar), d = (unsigned)__cvta_generic_to_shared(sm);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(b));
    asm volatile("fence.mbarrier_init.release.cluster;");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(16 * 32 * 8));
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(d),
                 "l"(&map), "r"(b), "r"(0), "r"(0)
                 : "memory");
  }
  __syncthreads();
  unsigned ok = 0;
  while (!ok) {
    asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0, 1, 0, p; }" : "=r"(ok) : "r"(b));
  }
  for (int i = threadIdx.x; i < 512; i += blockDim.x) out[i] = sm[i];
}
int main() {
  void* fn; cudaDriverEntryPointQueryResult q;
  cudaFree(0);
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
  Enc enc = (Enc)fn;
  double h[64 * 64];
  for (int r = 0; r < 64; ++r) for (int c = 0; c < 64; ++c) h[r * 64 + c] = r * 100 + c;
  double *dm, *dout; cudaMalloc(&dm, sizeof(h)); cudaMalloc(&dout, 512 * 8);
  cudaMemcpy(dm, h, sizeof(h), cudaMemcpyHostToDevice);
  const char* names[] = {"NONE", "32B", "64B", "128B", "128B_ATOM_32B", "128B_ATOM_32B_FLIP_8B", "128B_ATOM_64B"};
  for (int sw : {0, 3, 4, 6}) {
    alignas(64) CUtensorMap map;
    cuuint64_t dims[2] = {64, 64}; cuuint64_t str[1] = {64 * 8}; cuuint32_t box[2] = {16, 32}; cuuint32_t es[2] = {1, 1};
    CUresult r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, dm, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     (CUtensorMapSwizzle)sw, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("swizzle %s: encode rc=%d\n", names[sw], (int)r);
    if (r != CUDA_SUCCESS) continue;
    cudaMemset(dout, 0, 512 * 8);
    k<<<1, 128, 16 * 32 * 8>>>(map, dout);
    cudaError_t e = cudaDeviceSynchronize();
    double o[512]; cudaMemcpy(o, dout, sizeof(o), cudaMemcpyDeviceToHost);
    printf(" err=%s; smem rows (logical row r: physical 32-byte chunk order of its 4 chunks)\n", cudaGetErrorString(e));
    for (int r2 = 0; r2 < 16; ++r2) {
      printf("  smem row %2d:", r2);
      for (int c = 0; c < 16; c += 2) printf(" %5.0f", o[r2 * 16 + c]);
      printf("\n");
    }
  }
  return 0;
}
