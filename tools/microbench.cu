// Microbenchmarks that size the FP64 pipes of sm_100a: DMMA.8x8x4 issue rate per SM
// sub-partition as a function of resident warps and independent accumulators, and the plain
// DFMA rate.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o microbench microbench.cu
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

template <int ILP>
__global__ void dmma_rate(double* out, int iters, long long* clocks) {
  double c[ILP][2];
  double a = threadIdx.x * 1e-3, b = 1.0 + threadIdx.x * 1e-4;
#pragma unroll
  for (int i = 0; i < ILP; ++i) c[i][0] = c[i][1] = 0.0;
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i) dmma(c[i][0], c[i][1], a, b);
  }
  long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += c[i][0] + c[i][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) clocks[blockIdx.x] = t1 - t0;
}

// distinct A/B fragments per accumulator row/column, like a real 64x32 warp tile (8 A x 4 B)
__global__ void dmma_tile(double* out, int iters, long long* clocks) {
  double c[8][4][2];
  double a[8], b[4];
  for (int i = 0; i < 8; ++i) a[i] = threadIdx.x * 1e-3 + i;
  for (int j = 0; j < 4; ++j) b[j] = 1.0 + threadIdx.x * 1e-4 + j;
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) c[i][j][0] = c[i][j][1] = 0.0;
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) dmma(c[i][j][0], c[i][j][1], a[i], b[j]);
  }
  long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) s += c[i][j][0] + c[i][j][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) clocks[blockIdx.x] = t1 - t0;
}

template <int ILP>
__global__ void dfma_rate(double* out, int iters, long long* clocks) {
  double c[ILP];
  double a = 1.0 + threadIdx.x * 1e-9, b = 1e-9;
#pragma unroll
  for (int i = 0; i < ILP; ++i) c[i] = i;
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i) c[i] = fma(c[i], a, b);
  }
  long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += c[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) clocks[blockIdx.x] = t1 - t0;
}

template <class F>
void run(const char* name, F launch, int threads, int per_iter_ops_per_warp, int iters, double flop_per_op) {
  double* out;
  long long* clk;
  cudaMalloc(&out, 148 * 4 * 1024 * sizeof(double));
  cudaMalloc(&clk, 148 * 4 * sizeof(long long));
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  for (int blocks_per_sm = 1; blocks_per_sm <= 2; ++blocks_per_sm) {
    int grid = 148 * blocks_per_sm;
    launch(grid, threads, out, 10, clk);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    launch(grid, threads, out, iters, clk);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    long long h[148 * 2];
    cudaMemcpy(h, clk, grid * sizeof(long long), cudaMemcpyDeviceToHost);
    long long mx = 0;
    for (int i = 0; i < grid; ++i) mx = h[i] > mx ? h[i] : mx;
    int warps = threads / 32 * blocks_per_sm;
    double ops_per_smsp = (double)iters * per_iter_ops_per_warp * warps / 4.0;
    double total_flop = (double)iters * per_iter_ops_per_warp * (threads / 32) * grid * flop_per_op;
    printf("%-28s thr=%4d cta/sm=%d warps/smsp=%4.1f  clk/op/smsp=%7.2f  %8.2f TFLOP/s (%.3f ms, err=%s)\n", name,
           threads, blocks_per_sm, warps / 4.0, (double)mx / ops_per_smsp, total_flop / (ms * 1e-3) / 1e12, ms,
           cudaGetErrorString(cudaGetLastError()));
  }
  cudaFree(out);
  cudaFree(clk);
}

int main() {
  const int iters = 2000;
  for (int threads : {128, 256, 512, 1024}) {
    run("dmma ILP=2", [](int g, int t, double* o, int it, long long* c) { dmma_rate<2><<<g, t>>>(o, it, c); }, threads, 2, iters, 512);
    run("dmma ILP=8", [](int g, int t, double* o, int it, long long* c) { dmma_rate<8><<<g, t>>>(o, it, c); }, threads, 8, iters, 512);
    run("dmma ILP=32", [](int g, int t, double* o, int it, long long* c) { dmma_rate<32><<<g, t>>>(o, it, c); }, threads, 32, iters / 4, 512);
    if (threads <= 256)
      run("dmma tile 8x4 (64x32)", [](int g, int t, double* o, int it, long long* c) { dmma_tile<<<g, t>>>(o, it, c); }, threads, 32, iters / 4, 512);
  }
  for (int threads : {128, 256, 512, 1024}) {
    run("dfma ILP=8", [](int g, int t, double* o, int it, long long* c) { dfma_rate<8><<<g, t>>>(o, it, c); }, threads, 8, iters, 64);
    run("dfma ILP=32", [](int g, int t, double* o, int it, long long* c) { dfma_rate<32><<<g, t>>>(o, it, c); }, threads, 32, iters, 64);
  }
  return 0;
}
