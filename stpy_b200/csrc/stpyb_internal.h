// Internal (non-ABI) declarations shared by the .cu translation units.
#pragma once
#include <cuda_runtime.h>

namespace stpyb {

typedef long long i64;

constexpr int DB = 128;  // diagonal block order used by POTRF and every blocked solve

int gemm_nt(int M, int N, int K, const double* A, i64 lda, const double* B, i64 ldb, double* C, i64 ldc,
            double alpha, double beta, int tri, int square_cfg, cudaStream_t st, int kskip = 0,
            double* mirror = nullptr, i64 ldm = 0);
int potrf_diag(double* A, i64 lda, int b, double* Linv, int* info, int j0, cudaStream_t st, double* mirror = nullptr,
               i64 ldm = 0);
int potrf_panel(double* P, i64 rows, int w, i64 ldp, double* dinv, int* info, i64 j0, cudaStream_t st,
                double* pack = nullptr, i64 ldpack = 0);
int potrf_lower(double* A, i64 n, i64 lda, double* dinv, int* info, int outer, cudaStream_t st);

int trsm_rt(const double* L, i64 n, i64 ld, const double* dinv, double* Bt, i64 nt, i64 ldbt, cudaStream_t st);
int trsv_lower(const double* L, i64 n, i64 ld, const double* dinv, double* x, int transposed, cudaStream_t st);

}  // namespace stpyb
