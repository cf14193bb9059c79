/* tn_b200.h -- C ABI of the B200-native Gauss-Newton / ALS sweep kernels.
 *
 * The reference (niccogc/TensorNetworksFork) has no FFI/plugin boundary: its hot
 * path is Python on torch.einsum / torch.linalg.  These entry
 * points are what a binding for that path replaces, one per reference function.
 * Every pointer is a DEVICE pointer to fp64 data unless stated; sizes are in
 * elements; `stream` is a cudaStream_t passed as void*.  Every function enqueues
 * on `stream`, returns 0 on success or a negative TN_E* code, and never throws;
 * tn_last_error() gives the message for the calling thread.
 *
 * Shared vocabulary
 *   S           samples in the shard / batch
 *   rows        S * V "virtual rows": a sample whose output Hessian H_s is written as
 *               sum_t lam_t u_t u_t^T contributes V rows (V = 1 when C = 1)
 *   factor      a per-row vector F[row / div, 0..m) read with row stride ld; the local
 *               Jacobian of a core is the Kronecker product of three factors
 *               (left environment, site input, right environment): network.py:101-113
 *   map_kind    feature map applied to the site input while it is read:
 *               TN_MAP_IDENTITY  phi[p] = x[p]                    (poly-mode, models/tensor_train.py:223)
 *               TN_MAP_SINCOS    phi = [cos(pi x/2), sin(pi x/2)] (models/tnml.py:11-16)
 *               TN_MAP_POLY      phi[p] = x^p, p = 0..m-1         (models/tnml.py:18-23)
 */
#ifndef TN_B200_H
#define TN_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TN_OK 0
#define TN_EINVAL (-1)   /* bad argument                                   */
#define TN_ECUDA (-2)    /* CUDA runtime error (see tn_last_error)         */
#define TN_ENOTSPD (-3)  /* reserved: factorisation status is in *info     */
#define TN_EUNSUPPORTED (-4)

#define TN_MAP_IDENTITY 0
#define TN_MAP_SINCOS 1
#define TN_MAP_POLY 2

/* One Kronecker factor of the local Jacobian. */
typedef struct tn_factor {
    const double *ptr; /* row r lives at ptr + (r / div) * ld                         */
    int64_t ld;        /* row stride in elements                                      */
    int32_t m;         /* entries per row AFTER the feature map                       */
    int32_t div;       /* row divisor: 1, or V when the factor is shared by V rows    */
    int32_t map_kind;  /* TN_MAP_*; for SINCOS / POLY one raw value is read per row   */
    int32_t _pad;
} tn_factor;

int tn_version(void);
const char *tn_last_error(void);
/* Number of SMs of the current device (grid sizing on the host side). */
int tn_sm_count(void);
/* Number of kernels this library has launched in the process so far (bench bookkeeping). */
int64_t tn_launch_count(void);

/* ---- environments: TensorNetwork.compute_stacks / left|right_update_stacks / forward
 *      (tensor/network.py:55-71, 152-172, 115-137).
 * out[row, b] = sum_{a,p} env_in[row / env_div, a] * phi[row / cdiv, p] * core[a, p, b]
 * env_in == NULL means r_in == 1 and env_in == 1 (chain end).  `core` is [r_in, f, r_out]
 * contiguous, already oriented for the sweep direction by the caller.
 * If dot != NULL nothing is stored to `out`; instead
 *   yhat[row] = sum_b out[row, b] * dot[(row / dot_div) * dot_ld + b]   (the prediction).       */
int tn_env_update(const double *env_in, int64_t env_ld, int env_div, const double *x, int64_t x_ld, int map_kind, int f,
                  int cdiv, const double *core, double *out, int64_t out_ld, const double *dot, int64_t dot_ld,
                  int dot_div, double *yhat, int64_t rows, int r_in, int r_out, void *stream);

/* ---- class-leg rows: fold the output Hessian into the left factor
 *      (what the 3-operand einsum of network.py:207-212 does implicitly).
 * F[(s,t), a] = sum_c U[s,t,c] * env[s,c,a]        t < V
 * G[s, a]     = sum_c g[s,c]   * env[s,c,a]        (right-hand-side factor), either may be NULL  */
int tn_class_rows(const double *env, const double *U, const double *g, double *F, double *G, int64_t S, int C,
                  int V, int r, void *stream);

/* ---- Gram build: TensorNetwork.get_A_b (tensor/network.py:174-217), never materialising J.
 * With pair(F)[(i<=j)] = F[i]*F[j]:
 *   M[qa, qb, qc] (+)= sum_rows w[row] * pair(fa)[qa] * pair(fb)[qb] * pair(fc)[qc]
 * M has na*nb*nc entries, n = m(m+1)/2: the unique entries of A = J^T H J under the
 * Kronecker symmetry.  `w` may be NULL (all ones).  `work` holds ksplit partial copies of M
 * (ksplit >= 1 chosen by tn_gram_ksplit); the result is reduced into M deterministically.
 * mode: 0 = fp64 (FP64 tensor pipe, DMMA), 1 = TF32 tcgen05, 2 = 3xTF32 tcgen05, 3 = FP16 operands on tcgen05 (kind::f16,
 * fp32 accumulate; factors scaled per factor to [0.5, 1) first: a preconditioner-grade Gram for the exact refinement of
 * tn_cg, twice the samples per shared-memory transaction of mode 1).  Modes 1-3 accumulate in fp32 (TMEM) and flush to fp64
 * every `flush_rows` rows.  accumulate != 0 adds to M instead of overwriting.                     */
int tn_gram_ksplit(int64_t rows, int ma, int mb, int mc, int mode);
int tn_gram_kr3(int mode, const tn_factor *fa, const tn_factor *fb, const tn_factor *fc, const double *w,
                int64_t rows, double *M, double *work, int ksplit, int accumulate, void *stream);

/* Tensor-core modes: rows accumulated in fp32 (TMEM) between two fp64 flushes, for later calls of the calling thread; returns
 * the previous value; 0 = default (2048), negative = query only.                                    */
int tn_gram_tc_flush_rows(int rows);

/* b[ia, ib, ic] (+)= sum_rows w[row] * fa[ia] * fb[ib] * fc[ic]    (network.py:215)            */
int tn_rhs_ksplit(int64_t rows, int ma, int mb, int mc);
int tn_rhs_kr3(const tn_factor *fa, const tn_factor *fb, const tn_factor *fc, const double *w, int64_t rows,
               double *b, double *work, int ksplit, int accumulate, void *stream);

/* Gram / right-hand side of a Jacobian WITHOUT Kronecker structure: column i of J is
 * f1[t1[i]] * f2[t2[i]] * f3[t3[i]] (device int tables of length P).  Used for the cum-sum train, whose
 * per-feature coupling J[s,a,p,b] = x[p] cumL[a,p] R[b,p] (reference tensor/layers.py:408-477, SURVEY.md
 * Appendix C) is not a Kronecker product.  rhs_only == 0: out (P x P, dense, row stride P) (+)= J^T diag(w) J;
 * rhs_only != 0: out (P) (+)= J^T w.  fp64.                                                            */
int tn_generic_ksplit(int64_t rows, int P, int rhs_only);
int tn_gram_generic(const tn_factor *f1, const tn_factor *f2, const tn_factor *f3, const int *t1, const int *t2,
                    const int *t3, int P, const double *w, int64_t rows, double *out, int rhs_only, double *work,
                    int ksplit, int accumulate, void *stream);

/* ---- local solve: TensorNetwork.solve_system (tensor/network.py:293-327).
 * sigma_out[0] = mean_i |A_ii| computed from M (1 if 0).  role_of_pos[t] says which role
 * (0=a,1=b,2=c) parameter position t = 0,1,2 plays; m_pos[t] is its size.                       */
int tn_gram_sigma(const double *M, const int *m_pos, const int *role_of_pos, double *sigma_out, void *stream);
/* A[i, j] = M[pair...] / sigma + (i == j) * ridge      (P x P, row stride lda)                  */
int tn_gram_expand(const double *M, const int *m_pos, const int *role_of_pos, const double *sigma, double ridge,
                   double *A, int64_t lda, void *stream);
/* rhs[i] = -( b[i] / sigma + ridge * theta[i] )                                                 */
int tn_rhs_prepare(const double *b, const double *theta, const double *sigma, double ridge, double *rhs, int64_t P,
                   void *stream);
/* In-place blocked Cholesky of the lower triangle of A, then the two triangular solves on rhs.
 * work: tn_cholesky_work_elems(P) doubles (the inverted 64 x 64 diagonal blocks and, for P > 8192, the
 * inverted 512 x 512 diagonal blocks that turn the substitutions' block solves into matrix-vector
 * products).  info[0] = 0, or k>0 if the leading minor of order k is not positive definite (caller
 * raises LinAlgError as torch.linalg.cholesky does).                                              */
int64_t tn_cholesky_work_elems(int64_t P);
int tn_cholesky_solve(double *A, int64_t lda, int64_t P, double *rhs, double *work, int *info, void *stream);

/* Mixed-precision variant of tn_cholesky_solve for large systems whose matrix was itself accumulated on the tensor
 * cores (gram modes 1, 2).  A must hold the FULL symmetric matrix.  The P^3/3 trailing updates of the factorisation run
 * as 3xTF32 tcgen05 GEMMs (fp32 accumulate), panels and diagonal blocks in fp64; the resulting factor (accurate to ~1e-5)
 * preconditions an fp64 conjugate-gradient refinement whose operator is read from the untouched strict upper triangle of A
 * plus the saved diagonal, until ||A x - b|| <= rtol ||b|| or max_iter iterations.
 * stats[0] = relative residual reached, stats[1] = refinement iterations used (device doubles; may be NULL).
 * info as above; when info[0] != 0, or the residual stays above rtol, rhs does not hold a usable solution and the caller
 * re-expands A and calls tn_cholesky_solve.  work: tn_cholesky_mixed_work_elems(P) doubles.
 * Unlike tn_cholesky_solve this call SYNCHRONISES the stream once per refinement iteration (the host reads the stop flag),
 * so it cannot be captured into a CUDA graph.                                                                      */
int64_t tn_cholesky_mixed_work_elems(int64_t P);
int tn_cholesky_solve_mixed(double *A, int64_t lda, int64_t P, double *rhs, double *work, int *info, double rtol,
                            int max_iter, double *stats, void *stream);

/* theta <- theta + lr * step with the optional adaptive shrink / max-norm projection
 * (TensorNode.update_node, tensor/node.py:178-203).  max_norm <= 0 disables the projection.     */
int tn_update_node(double *theta, const double *step, int64_t P, double lr, int adaptive_step, double max_norm,
                   double *scratch, void *stream);

/* ---- QR re-gauge: node_orthonormalize_left/right (tensor/network.py:625-707).
 * a (m x n, row-major, m >= n) <- Q (reduced, LAPACK sign convention); r (n x n) <- R.          */
int tn_qr(double *a, int m, int n, double *r, void *stream);

/* ---- matrix-free local operator: the matvec of lanczos_swipe / scipy_swipe
 *      (tensor/network.py:770-790, 896-918):  out = J^T diag(w) J v  on virtual rows.           */
int64_t tn_matvec_work_elems(int64_t rows, int ma, int mb, int mc);
int tn_matvec_kr3(const tn_factor *fa, const tn_factor *fb, const tn_factor *fc, const double *w, int64_t rows,
                  const double *v, double *out, double *work, void *stream);

/* ---- on-device Krylov drivers: the local solves of lanczos_swipe / scipy_swipe (tensor/network.py:770-832, 896-932; the
 *      reference runs SciPy's cg / minres on the host in float32, or an eager torch loop for Lanczos), and the exact
 *      refinement of the tensor-core Gram modes.  The operator is
 *          Op v = A0 v / sigma[0] + ridge * v          (sigma == NULL: 1)
 *      with A0 either built in -- J^T diag(w) J of this rank's rows from three Kronecker factors (apply == NULL) -- or
 *      enqueued by the caller's `apply` callback; `allreduce`, when set, sums the P-vector A0 v over the ranks of a
 *      sample-sharded run (replaces the missing collective of the reference) before the scaling.
 *      All recurrence scalars stay on the device.  Unlike the other entry points these drivers SYNCHRONISE the stream every
 *      `poll_every` iterations to read the convergence flag (poll_every == 0: never -- a fixed sequence of launches whose
 *      kernels turn into no-ops once converged, capturable in a CUDA graph).
 *      stats (5 device doubles, may be NULL): [0] relative residual reached (cg: |b - Op x| / |b|; minres: its estimate
 *      phibar / |b|; lanczos: the last beta), [1] iterations, [2] 1 if stopped by the tolerance, [3] operator applications,
 *      [4] (cg only) the value of the stopping criterion.                                                                */
typedef int (*tn_apply_fn)(void *ctx, const double *v, double *out, void *stream);
typedef int (*tn_allreduce_fn)(void *ctx, double *buf, int64_t n, void *stream);
typedef struct tn_operator {
    const tn_factor *fa, *fb, *fc; /* built-in operator (apply == NULL); only fb may carry a feature map      */
    const double *w;               /* row weights or NULL                                                     */
    int64_t rows;
    tn_apply_fn apply;             /* or: enqueue out = A0 v on `stream`                                      */
    void *apply_ctx;
    tn_allreduce_fn allreduce;     /* optional: enqueue the in-place sum of buf[0..n) over the ranks          */
    void *allreduce_ctx;
    const double *sigma;           /* device scalar or NULL                                                   */
    double ridge;
    int64_t P;
    int32_t apply_is_global;       /* != 0: `apply` already sums over the ranks; `allreduce` then only serves the
                                      collective convergence decision (ranks must leave the iteration together) */
    int32_t _pad;
} tn_operator;

/* Conjugate gradients on Op x = b (scipy.sparse.linalg.cg as called at network.py:921-925: stop at |r| <= rtol |b| or
 * max_iter), optionally preconditioned by the Cholesky factor L (lower triangle, row stride lda, with the `Lwork` and `Linfo`
 * of tn_cholesky_factor; Linfo[0] != 0 turns the call into a no-op).  use_x0 == 0: x starts at 0, or at (L L^T)^-1 b when
 * L is given.  With L the stopping test is |L^-T L^-1 r| <= rtol |x|: z = (L L^T)^-1 r estimates the ERROR of x when L L^T is
 * close to Op, so the forward error is bounded whatever the condition number (a residual test lets it float with it).
 * With L the factor of the TF32 / 3xTF32 Gram and the built-in fp64 operator this is the refinement that
 * makes the tensor-core Gram modes solve the fp64 system of solve_system (network.py:293-327).                          */
int64_t tn_cg_work_elems(const tn_operator *op);
int tn_cg(const tn_operator *op, const double *L, int64_t lda, const double *Lwork, const int *Linfo, const double *b,
          double *x, int use_x0, int max_iter, double rtol, int poll_every, double *work, double *stats, void *stream);
/* MINRES without preconditioner or shift (scipy.sparse.linalg.minres as called at network.py:921-925). */
int64_t tn_minres_work_elems(const tn_operator *op);
int tn_minres(const tn_operator *op, const double *b, double *x, int use_x0, int max_iter, double rtol, int poll_every,
              double *work, double *stats, void *stream);
/* Lanczos-Galerkin solve of lanczos_swipe (network.py:793-824): r0 = b - Op x0 (x0 may be NULL = 0), max_iter Lanczos
 * vectors without re-orthogonalisation (stop when |w_j| < tol), x = x0 + V T^-1 (|r0| e1).                              */
int64_t tn_lanczos_work_elems(const tn_operator *op, int max_iter);
int tn_lanczos(const tn_operator *op, const double *b, const double *x0, double *x, int max_iter, double tol,
               int poll_every, double *work, double *stats, void *stream);

/* Factorisation alone (the preconditioner of tn_cg): lower triangle of A <- L, work (tn_cholesky_work_elems(P)) <- the
 * inverted 64 x 64 diagonal blocks.  tensor_core == 1: trailing updates as 3xTF32 tcgen05 GEMMs (factor accurate to ~1e-5);
 * tensor_core == 2: one TF32 pass per K step (~1e-3: a factor that only preconditions tn_cg).
 * tn_cholesky_apply: x <- L^-T L^-1 x.                                                                               */
int tn_cholesky_factor(double *A, int64_t lda, int64_t P, int tensor_core, double *work, int *info, void *stream);
int tn_cholesky_apply(const double *L, int64_t lda, int64_t P, double *x, const double *work, const int *info, void *stream);

/* out[0] (+)= sum_rows w |fa|^2 |fb|^2 |fc|^2 = trace(J^T diag(w) J) in fp64, out[1] (+)= the same with |w|: the exact
 * sigma = mean |A_ii| of solve_system (network.py:298) for the tensor-core Gram modes when no weight is negative.       */
int tn_gram_trace(const tn_factor *fa, const tn_factor *fb, const tn_factor *fc, const double *w, int64_t rows, double *out,
                  int accumulate, void *stream);

/* ---- per-sample small matrix products of the patch/pixel ("conv-TT") layer
 *      (TensorConvolutionTrainLayer, tensor/layers.py:791-890; closed forms in SURVEY.md Appendix C):
 * out[s, i, j] (+)= sum_k A[s*sA + i*iA + k*kA] * B[s*sB + k*kB + j*jB],  out [S, I, J] contiguous.
 * A stride sA or sB of 0 shares that operand between all samples.                                   */
int tn_bmm(const double *A, int64_t sA, int64_t iA, int64_t kA, const double *B, int64_t sB, int64_t kB, int64_t jB,
           double *out, int64_t S, int I, int K, int J, int accumulate, void *stream);

/* Row-reduced outer product (right-hand side / J^T pass of a conv-TT patch core, get_b + the second einsum of the matvec,
 * tensor/network.py:258-291, 790):  out[i, j] (+)= sum_row w[row] * G[(row / gdiv) * ldg + i] * W[row * ldw + j],
 * out (ra x m) contiguous, ra <= 128, w may be NULL.  Row ranges are combined with fp64 atomics.                     */
int tn_outer_rows(const double *G, int64_t ldg, int gdiv, int ra, const double *W, int64_t ldw, int m, const double *w,
                  int64_t rows, double *out, int accumulate, void *stream);

/* Row-wise products with a shared matrix (J v pass of a conv-TT patch core, tensor/network.py:789; also the per-sample
 * contractions of a conv-TT column with its shared cores):
 * z[row * ldz + i] = sum_j W[row * ldw + j] * V[i * ldv + j],  i < ra (any ra), ldz >= ra.                            */
int tn_rows_dot(const double *W, int64_t ldw, int m, const double *V, int64_t ldv, int ra, int64_t rows, double *z,
                int64_t ldz, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* TN_B200_H */
