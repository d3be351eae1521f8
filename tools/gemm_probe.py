"""Throughput of stpyb_gemm_nt for shapes that isolate main loop / epilogue / memory effects."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from stpy_b200 import _lib as L

L.load()
dev = torch.device("cuda")

def run(M, N, K, alpha, beta, lower, reps=5, tag=""):
    A, lda = L.empty_matrix(M, K); A.normal_()
    B, ldb = L.empty_matrix(N, K); B.normal_()
    C, ldc = L.empty_matrix(M, N); C.zero_()
    def call():
        L.call("stpyb_gemm_nt", M, N, K, L.ptr(A), lda, L.ptr(B), ldb, L.ptr(C), ldc, alpha, beta, lower, L.stream_ptr())
    call(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); call(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    fl = 2.0 * M * N * K * (0.5 if lower else 1.0)
    print("%-34s M=%6d N=%6d K=%5d a=%+.0f b=%.0f lower=%d  %8.3f ms  %6.2f TFLOP/s" % (tag, M, N, K, alpha, beta, lower, best, fl / best / 1e9))

run(8192, 8192, 4096, 1.0, 0.0, 0, tag="long K, store only")
run(8192, 8192, 16384, 1.0, 0.0, 0, tag="very long K")
run(4736, 4736, 8192, 1.0, 0.0, 0, tag="exactly 37x74 tiles (~9 waves)")
run(8192, 8192, 256, 1.0, 0.0, 0, tag="K=256 store only (L2 operands)")
run(8192, 8192, 256, -1.0, 1.0, 0, tag="K=256 accumulate")
run(8192, 8192, 512, -1.0, 1.0, 0, tag="K=512 accumulate")
run(8192, 8192, 1024, -1.0, 1.0, 0, tag="K=1024 accumulate")
run(32768, 32768, 256, -1.0, 1.0, 1, tag="SYRK-like lower 32k K=256")
run(32768, 32768, 512, -1.0, 1.0, 1, tag="SYRK-like lower 32k K=512")
run(32768, 32768, 256, 1.0, 0.0, 1, tag="lower 32k K=256 store only")
run(65536, 128, 128, 1.0, 0.0, 0, tag="panel-TRSM shape (stream cfg)")
