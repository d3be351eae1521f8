/* stpyb.h — C ABI of libstpyb.so, the sm_100a implementation of stpy's
 * Gaussian-process hot path (Gram -> Cholesky -> triangular solves -> fit /
 * mean_std / log-marginal-likelihood, plus random Fourier features).
 *
 * stpy (the reference) is pure Python: it has no FFI layer.  The seam this
 * library sits behind is therefore the set of torch / scipy calls made by the
 * reference's hot functions; each entry point names the reference lines whose
 * arithmetic it replaces (paths relative to the stpy repository root).  The
 * host side that calls these (stpy_b200/*.py, via ctypes) mirrors the
 * reference's classes one to one; INTEGRATION.md shows the binding a stpy
 * maintainer would add.
 *
 * Conventions
 *  - every matrix is row-major float64 in DEVICE memory, leading dimension in
 *    elements; matrices handed to the dense kernels need an even leading
 *    dimension and a 16-byte aligned base (torch allocations are 512-byte aligned);
 *  - `stream` is a cudaStream_t passed as void*; all calls are asynchronous on
 *    it and never synchronise;
 *  - return value: 0 ok; <0 invalid argument (minus its 1-based position);
 *    1000+e CUDA runtime error e.  No C++ exception crosses the boundary (the library itself never calls
 *    NCCL: collectives are issued by the host side through torch.distributed);
 *  - Cholesky failure is reported LAPACK-style through `info_dev` (device int):
 *    0 ok, i>0 = the leading minor of order i is not positive definite (what
 *    torch.linalg.cholesky raises on, estimator.py:35).
 *  - ownership: the caller owns every buffer; the library keeps no state apart
 *    from opaque handles created by the *_create calls.
 */
#ifndef STPYB_H
#define STPYB_H

#ifdef __cplusplus
extern "C" {
#endif

#define STPYB_MAX_DIM 64 /* max selected input columns per sub-kernel */
#define STPYB_DB 128     /* order of the inverted diagonal blocks kept by potrf */
#define STPYB_MAX_PEERS 16 /* ranks of one NVLink domain addressed by the peer-memory kernels */

/* kernel kinds (stpy/kernels.py: get_kernel_internal, lines 167-261) */
enum {
  STPYB_K_SE = 0,       /* squared_exponential :368-398, ard :552-583 (inputs pre-scaled) */
  STPYB_K_MATERN12 = 1, /* matern :811-859 / ard_matern :917-970, nu = 0.5 */
  STPYB_K_MATERN32 = 2, /* nu = 1.5 */
  STPYB_K_MATERN52 = 3, /* nu = 2.5 */
  STPYB_K_POLY = 4,     /* polynomial :766-784, p0 = degree */
  STPYB_K_LINEAR = 5,   /* linear :300-320, p0 = offset */
  STPYB_K_MATERN_NU = 6,/* matern :852-859, general nu through the modified Bessel function K_nu; kparams */
  STPYB_K_COUNT = 7
};
/* how a sub-kernel's Gram combines with what is already in K (kernels.py:146-157) */
enum { STPYB_OP_SET = 0, STPYB_OP_ADD = 1, STPYB_OP_MUL = 2 };

int stpyb_version(void);

/* Instrumentation used by bench.py.  stpyb_profile(1) resets the counters and makes the
 * factorisation bracket each of its launches with CUDA events on the launching stream;
 * stpyb_profile_read (after a synchronise) returns, per category {0 diagonal block,
 * 1 panel TRSM, 2 in-panel update, 3 trailing SYRK, 4 Gram, 5 other, 6 single-RHS triangular
 * solves (HBM-bound: algorithmic bytes = 4 x the flops reported)}, out[3*c+0..2] =
 * {milliseconds, algorithmic flops, timed launches}, and the number of kernels launched
 * by the library since the reset. */
int stpyb_profile(int enable);
int stpyb_profile_read(double* out21, long long* launches);
/* Debug: factor one diagonal block (order b <= 128) and record clock64() stamps of the kernel's
 * phases: {start, loaded, factored, inverse assembled, stored, sum(leaf), sum(panel solve), sum(update)}. */
int stpyb_potrf_diag_profile(double* A, long long lda, int b, double* Linv, int* info_dev,
                             long long* stamps8_dev, void* stream);

/* ---- Gram construction ------------------------------------------------- */

/* Column-select (`group`), lengthscale-scale and zero-pad the inputs to dpad
 * (multiple of 4) columns, and form the row norms used by the distance
 * expansion.  Replaces a[:, group], mm(a, diag(1/gamma)), sum(a**2, dim=1):
 * kernels.py:387-391, 572-577, 840-843.  cols_host / scale_host are HOST
 * arrays (dg ints / scale_len in {0,1,dg} doubles).  divide=1 computes
 * x/scale (matern_kernel), 0 computes x*scale. */
int stpyb_gram_prep(const double* X, long long n, long long ldx, const int* cols_host, int dg,
                    const double* scale_host, int scale_len, int divide, double* Xp, int dpad,
                    double* norms_or_null, void* stream);

/* K[j,i] (op)= kappa * f(b_j, a_i) (+ diag_add on j==i), K is (m x n): the
 * reference's orientation (|b|,|a|), kernels.py:393-398.  Ap/Bp/na/nb come
 * from stpyb_gram_prep.  arg_scale is the factor on the squared distance for
 * SE (-0.5/gamma^2, or -0.5 for pre-scaled ARD inputs).  refine=1 (Matern
 * kinds) recomputes cancellation-prone distances by direct differences, the
 * behaviour of scipy's cdist in matern_kernel; refine=0 reproduces
 * torch.cdist's clamped expansion used by ard_matern_kernel.  lower_only=1
 * (symmetric case, a is b) writes only tiles on or below the diagonal.
 * kparams_host_or_null: 6 HOST doubles for STPYB_K_MATERN_NU, NULL otherwise:
 * {nu, gam1, gam2, 1/Gamma(1+mu), 1/Gamma(1-mu), 2^(1-nu)/Gamma(nu)} with mu = nu - round(nu) and
 * gam1 = (1/Gamma(1-mu) - 1/Gamma(1+mu)) / (2 mu), gam2 = (1/Gamma(1-mu) + 1/Gamma(1+mu)) / 2 (the
 * Gamma-function combinations of the Temme series for K_nu). */
int stpyb_gram(int kind, const double* Ap, const double* na, long long n, const double* Bp,
               const double* nb, long long m, int dpad, double arg_scale, double kappa, double p0,
               int refine, int op, double diag_add, int lower_only, double* K, long long ldk,
               const double* kparams_host_or_null, void* stream);

/* out[i] (op)= kappa * f(b_i, a_i): kernel_diag (kernels.py:112-134) and the
 * n_t 1x1 kernel calls of gauss_procc.py:347. */
int stpyb_gram_diag(int kind, const double* Ap, const double* na, const double* Bp, const double* nb,
                    long long n, int dpad, double arg_scale, double kappa, double p0, int op,
                    double* out, const double* kparams_host_or_null, void* stream);

/* One shared distance tile -> nk kernels' Gram matrices (model-selection
 * sweep over isotropic SE / Matern kernels on one dataset; the shape of
 * categorical_mixture.py:48-65).  kinds/arg_scales/kappas are HOST arrays of
 * length nk (nk <= 64); kernel q is written to K + q*stride_k (lower tiles
 * only), with diag_add on the diagonal.  Inputs are the UNSCALED prepped
 * points; arg_scale is -0.5/gamma^2 for SE and 1/gamma for Matern. */
int stpyb_gram_multi(int nk, const int* kinds, const double* arg_scales, const double* kappas,
                     const double* Ap, const double* na, long long n, int dpad, double diag_add,
                     double* K, long long ldk, long long stride_k, void* stream);

/* ---- Cholesky and solves ------------------------------------------------ */

/* In-place lower Cholesky of the n x n matrix (only the lower triangle is
 * read or written).  dinv receives ceil(n/128) dense 128x128 blocks holding
 * inv(L_kk) of every diagonal block; the solves below consume them.
 * Replaces torch.linalg.cholesky (estimator.py:35) and the lstsq / LU
 * factorisations of gauss_procc.py:367-378, 633-635.  outer_block (a multiple
 * of 128; 1024 is the measured optimum): K-depth of the trailing SYRK.
 * OR-ing STPYB_POTRF_NO_LOOKAHEAD into outer_block keeps this call on `stream` alone. */
#define STPYB_POTRF_NO_LOOKAHEAD 0x40000000
int stpyb_potrf(double* K_inout, long long n, long long ld, double* dinv, int* info_dev,
                int outer_block, void* stream);
/* Factorisations of order >= min_n (default 4096, or the environment variable
 * STPYB_LOOKAHEAD_MIN_N) overlap the next panel, on a high-priority side stream, with the
 * trailing update of the current one; smaller ones run on the caller's stream only (so that
 * many of them can share the device from different streams: stpy_b200/sweep.py).
 * Negative min_n disables the overlap.  Returns the previous value through old_or_null. */
int stpyb_set_lookahead_min_n(long long min_n, long long* old_or_null);

/* x <- L^-1 x (transposed=0) or L^-T x (transposed=1), single right-hand side. */
int stpyb_trsv(const double* L, long long n, long long ld, const double* dinv, double* x,
               int transposed, void* stream);
/* x <- (L L^T)^-1 x : alpha = K^-1 y (gauss_procc.py:376, estimator.py:37). */
int stpyb_potrs_vec(const double* L, long long n, long long ld, const double* dinv, double* x,
                    void* stream);
/* Bt <- Bt L^-T for Bt (nt x n), each ROW one right-hand side: V^T = K* L^-T,
 * the solve behind gauss_procc.py:378 (B = lstsq(K, K*^T)^T). */
int stpyb_trsm_rt(const double* L, long long n, long long ld, const double* dinv, double* Bt,
                  long long nt, long long ldbt, void* stream);
/* out3 = { ||z||^2, logdet K = 2 sum log L_ii, 0.5*||z||^2 + 0.5*weight*logdet }
 * with z = L^-1 y: the value returned by _log_marginal_squared
 * (gauss_procc.py:631-638) and Estimator.log_marginal (estimator.py:32-40). */
int stpyb_lml(const double* L, long long n, long long ld, const double* z, double weight,
              double* out3, void* stream);
/* out_i = sum_j V_ij^2 (mode 0), sqrt(kss_i - sum) (mode 1: posterior std,
 * gauss_procc.py:391-395), kss_i - sum (mode 2: variance). */
int stpyb_row_sumsq(const double* V, long long rows, long long cols, long long ldv,
                    const double* kss_or_null, int mode, double* out, void* stream);
/* out = M v, M (rows x cols): posterior mean K* alpha (gauss_procc.py:381). */
int stpyb_gemv_rows(const double* M, long long rows, long long cols, long long ldm, const double* v,
                    double* out, void* stream);
/* C = alpha A B^T + beta C, A (M x K), B (N x K), row-major; lower=1 skips
 * tiles above the diagonal.  The DMMA contraction every blocked stage uses;
 * exported for the full posterior covariance (gauss_procc.py:396-399), the
 * distributed schedule and tests. */
int stpyb_gemm_nt(int M, int N, int K, const double* A, long long lda, const double* B, long long ldb,
                  double* C, long long ldc, double alpha, double beta, int lower, void* stream);

/* ---- LML gradient --------------------------------------------------------- */

/* Kinv (n x n, lower triangle valid) = (L L^T)^-1 via U = L^-T and U U^T.
 * work is an n x ldw scratch matrix.  Feeds the analytic gradient that
 * replaces autograd through solve/slogdet (gauss_procc.py:631-638 + backward). */
int stpyb_potri(const double* L, long long n, long long ld, const double* dinv, double* work,
                long long ldw, double* Kinv, long long ldki, void* stream);
/* Derivative pass over a COMPOSITE kernel: sum_ij W_ij dK_ij/dtheta for the lengthscales and the
 * amplitude of one item, with K the reference's left fold of sub-kernel Grams by + and *
 * (kernels.py:146-157) and each sub-kernel a sum of items (additive groups, kernels.py:700-729).
 * Replaces autograd's backward through exp / mm / cdist of every kernel builder
 * (kernels.py:390-398, 572-583, 944-962, 780-784) and, in mode 0, through slogdet / solve
 * (gauss_procc.py:631-638).  All descriptor arrays are HOST arrays:
 *   item q < nitems (<= 8, sorted by sub-kernel): kinds[q], ncols[q] (<= 32), subs[q],
 *   cols_flat32[q*32 + c] (input column), sc_flat32[q*32 + c] (factor on that column, 1/lengthscale),
 *   arg_scales[q] (SE: factor on the squared scaled distance), kappas[q], p0s[q] (degree / offset);
 *   sub_ops[p] for p < nsub: STPYB_OP_SET for p = 0, then ADD or MUL.
 * XR (m x d rows = the b-points of K) and XC (n x d, the a-points) are the RAW inputs.
 * mode 0: m == n, XR == XC, W = weight * Kinv - alpha alpha^T taken over the lower triangle of
 *         Kinv = Cmat (mirror included, diagonal halved: the value 0.5 tr(W dK));
 * mode 1: W = Cmat, an explicit m x n matrix (the backward of KernelFunction.kernel).
 * out18 (device): [0..15] = sum W dout/dG kappa f'(sq) u_c^2 for columns col_off..col_off+15 of item
 * pass_item (times -2/lengthscale = the lengthscale derivative), [16] = sum W dout/dG f (the derivative
 * w.r.t. the item's kappa), [17] = 0.5 trace(W) (mode 0; times 2s = d/ds). */
int stpyb_kernel_grad(const double* XR, long long m, long long ldxr, const double* XC, long long n,
                      long long ldxc, int d, int nitems, const int* kinds, const int* ncols,
                      const int* subs, const int* cols_flat32, const double* sc_flat32,
                      const double* arg_scales, const double* kappas, const double* p0s, int nsub,
                      const int* sub_ops, int pass_item, int col_off, int mode, const double* Cmat,
                      long long ldc, const double* alpha, double weight, double* out18, void* stream);

/* ---- stack of Gram matrices (multiple-kernel learning) ----------------------------------- */

/* out = sum_q weights[q] * K_q + diag_add * I over a stack of k (<= 64) n x n Gram matrices (matrix q at
 * stack + q*stride, e.g. from stpyb_gram_multi); weights_host is a HOST array.  lower=1 reads and writes only
 * j <= i.  Replaces torch.sum(torch.stack([alpha*K ...])) + eye*lam*s^2 (mkl_estimator.py:90). */
int stpyb_stack_combine(const double* stack, int k, const double* weights_host, long long n, long long ld,
                        long long stride, double diag_add, int lower, double* out, long long ldo, void* stream);
/* out_k[q] = beta^T K_q beta for every matrix of the stack (lower=1: symmetric matrices stored as their lower
 * triangle).  With beta = K(alpha)^-1 y this is minus the gradient of y^T K(alpha)^-1 y in alpha_q, the weight
 * objective of mkl_estimator.py:60-64. */
int stpyb_stack_quadform(const double* stack, int k, long long n, long long ld, long long stride, int lower,
                         const double* beta, double* out_k, void* stream);

/* ---- symmetric eigendecomposition (Nystrom features) ------------------------------------ */

/* One-sided Jacobi eigensolver for a symmetric (Gram) matrix, replacing torch.linalg.eigh in
 * nystrom_fea.py:116-136, 188-196.  Work matrices H, W are np x np (np even, >= n; row-major, ld).
 *   stpyb_jacobi_init        H = A zero-padded, W = I
 *   stpyb_jacobi_sweep       np-1 rounds of np/2 disjoint row-pair rotations keeping H = W A; *rotated_dev =
 *                            number of pairs with |H_p . H_q| > tol |H_p| |H_q| (0: converged)
 *   stpyb_jacobi_eigenvalues lam_i = H_i . W_i; the rows of W are the eigenvectors (unsorted). */
int stpyb_jacobi_init(const double* A, long long lda, double* H, double* W, long long n, long long np, long long ld,
                      void* stream);
int stpyb_jacobi_sweep(double* H, double* W, long long np, long long ld, double tol, int* rotated_dev, void* stream);
int stpyb_jacobi_eigenvalues(const double* H, const double* W, long long np, long long ld, double* lam, void* stream);

/* ---- random Fourier features ---------------------------------------------- */

/* Phi[i,f] = scale * featw_f * trig(x_i . w_f + bias_f)   (n x m), or its
 * transpose (m x n) when transposed_out=1.  mode 0 (reference default,
 * embedding.py:234-239): f < m/2 -> cos, f >= m/2 -> sin, no bias;
 * mode 1 (biased, :232): cos with bias.  featw (sqrt quadrature weights,
 * embedding.py:450-466) and bias may be null.  Xp (n x dpad) / Wp (m x dpad)
 * come from stpyb_gram_prep with unit scale. */
int stpyb_rff_embed(const double* Xp, long long n, const double* Wp, int m, int dpad,
                    const double* bias_or_null, const double* featw_or_null, int mode, double scale,
                    int transposed_out, double* Phi, long long ldphi, void* stream);

/* Normal equations of the feature-space regression without materialising Phi
 * (n x m): V[0:m,0:m] (lower) += Phi^T Phi, V[m,0:m] += (Phi^T y)^T (V[m,m] is
 * not touched), streaming row chunks of at most `chunk` points through
 * `scratch` (m x ldscratch, ldscratch >= chunk): the two halves of scratch are
 * filled alternately by the embedding on a side stream while the caller's
 * stream contracts the other half.  V is (m+1) x ldv and must be zeroed (or
 * hold a previous partial sum) on entry.  Replaces
 * kernelized_features.py:228, 237 (Q = embed(x); Q.T @ Q) and the Q.T @ y of :256. */
int stpyb_rff_normal_eq(const double* Xp, const double* y, long long n, const double* Wp, int m, int dpad,
                        const double* bias_or_null, const double* featw_or_null, int mode, double scale,
                        long long chunk, double* scratch, long long ldscratch, double* V, long long ldv,
                        void* stream);

/* ---- building block of the multi-GPU factorisation --------------------------- */

/* Factor one tall panel in place: P is rows x w (row-major, ldp), whose top
 * w x w block is the (already updated) diagonal block.  On return the top
 * block holds L_jj, the rows below hold L_ij = A_ij L_jj^-T, and dinv receives
 * the ceil(w/128) inverted 128x128 diagonal sub-blocks.  info_dev is set (if
 * still 0) to j0 + the 1-based index of the first non-positive pivot.
 * stpy_b200/distributed.py drives this per block column of the
 * block-column-cyclic layout and broadcasts the result with NCCL.  With
 * pack_or_null != NULL the kernels that produce the factored panel also store
 * it at pack[r * ldpack + c] (the contiguous broadcast buffer), so no separate
 * copy pass over the panel sits on the critical chain. */
int stpyb_potrf_panel(double* P, long long rows, int w, long long ldp, double* dinv, int* info_dev,
                      long long j0, double* pack_or_null, long long ldpack, void* stream);


/* y[c] -= sum_r A[r][c] v[r] for a tall panel A (rows x w): the transposed GEMV of the
 * distributed backward solve alpha = L^-T z over column-owned panels. */
int stpyb_gemv_t_sub(const double* A, long long rows, int w, long long ld, const double* v, double* y,
                     void* stream);

/* `count` independent updates C_i = alpha A_i B_i^T + beta C_i (common K and leading dimensions)
 * in one call: forked from main_stream onto up to nside side streams and joined back.  M, N and
 * the pointer tables are HOST arrays.  The multi-GPU trailing update of one factorisation step. */
int stpyb_gemm_nt_batch(int count, const int* M, const int* N, int K, const double* const* A, long long lda,
                        const double* const* B, long long ldb, double* const* C, long long ldc, double alpha,
                        double beta, int lower, void* main_stream, void* const* side_streams, int nside);

/* One hop of the distributed backward sweep (owner of a block column of width w): seg <- z_g,
 * seg -= L[below, g]^T alpha_below, seg <- L_gg^-T seg.  Lcol points at the diagonal block of the
 * column inside the rank's slab (the `below` rows follow it), dinv at its inverted sub-blocks. */
int stpyb_dist_alpha_step(const double* Lcol, long long ld, long long below, int w, const double* dinv,
                          const double* zrow, const double* alpha_below, double* seg, void* stream);

/* ---- peer memory over NVLink (one process per GPU) -------------------------------------------
 * stpyb_p2p_alloc: cudaMalloc a zeroed symmetric buffer and export its 64-byte IPC handle;
 * stpyb_p2p_open / _close: map / unmap a peer's buffer; stpyb_p2p_free: release the own one. */
int stpyb_p2p_alloc(long long bytes, void** dev_ptr_out, void* ipc_handle_64);
int stpyb_p2p_open(const void* ipc_handle_64, void** dev_ptr_out);
int stpyb_p2p_close(void* dev_ptr);
int stpyb_p2p_free(void* dev_ptr);
/* Fused "publish" of a backward-sweep hop: the w doubles at offset `off` of the own alpha buffer
 * are stored into every peer's buffer (st.global on mapped peer pointers), followed by a
 * system-scope fence and flag[g] = epoch on every rank.  peer_alpha / peer_flags are HOST tables
 * of `world` device pointers (entry `self` is the local buffer). */
int stpyb_p2p_alpha_publish(void* const* peer_alpha, void* const* peer_flags, int world, int self,
                            long long off, int w, int g, int epoch, void* stream);
/* In-stream wait until local_flags[lo..hi) == epoch; gives up after limit_cycles and writes
 * 1 + (first missing flag) to *err_dev instead of hanging. */
/* Right-looking backward sweep over peer memory.  stpyb_dist_strip (every rank, every hop): wait
 * in-kernel for flag[g] (unless flags is null), then zrow[c] -= sum_r Lstrip[r][c] * seg[r] for the
 * bw rows of block g and the rank's ncols local columns left of it.  stpyb_dist_solve_publish
 * (owner of block g): alpha_g = L_gg^-T z_g in one CTA, stored into EVERY rank's symmetric buffer
 * at `off` followed by flag[g] = epoch -- the broadcast is the tail of the solve kernel. */
int stpyb_dist_strip(const double* Lstrip, long long ld, int bw, long long ncols, const double* seg,
                     double* zrow, const int* flags_or_null, int g, int epoch, long long limit_cycles,
                     int* err_dev, void* stream);
int stpyb_dist_solve_publish(const double* Lgg, long long ld, int w, const double* dinv, const double* zrow,
                             void* const* peer_alpha, void* const* peer_flags, int world, int self,
                             long long off, int g, int epoch, void* stream);
/* Asynchronous device-to-device copy on `stream` (moves results out of a symmetric buffer). */
int stpyb_memcpy_d2d(void* dst, const void* src, long long bytes, void* stream);
int stpyb_p2p_wait_flags(const int* local_flags, int lo, int hi, int epoch, long long limit_cycles,
                         int* err_dev, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* STPYB_H */
