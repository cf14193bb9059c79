// Householder QR of a small tall matrix in one CTA (the QR re-gauge of a core).
//
// Replaces torch.linalg.qr(mode='reduced') in node_orthonormalize_left/right (reference
// tensor/network.py:644,686).  Matrices are (r_l*f) x r_r, at most a few hundred rows by a few
// dozen columns, so this is latency work: one CTA, the matrix in shared memory when it fits.
// Sign convention follows LAPACK dgeqr2/dorg2r (beta = -sign(alpha)*norm), which is what torch
// returns, so Q and R match the reference entry by entry up to rounding.
#include "common.cuh"

namespace tn {

constexpr int QR_THREADS = 256;

__device__ double qr_block_sum(double v, double* red) {
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    double t = 0.0;
    for (int k = 0; k < QR_THREADS / 32; ++k) t += red[k];
    return t;
}

// a: m x n row-major (global).  q_g: scratch m x n (global) used only when the matrix does not fit
// in shared memory.  On exit a = Q, r = R (n x n, zero below the diagonal).
__global__ void __launch_bounds__(QR_THREADS)
qr_kernel(double* __restrict__ a_g, int m, int n, double* __restrict__ r_g, double* __restrict__ q_g, int use_smem) {
    extern __shared__ double sm[];
    __shared__ double red[QR_THREADS / 32];
    __shared__ double s_tau[128];
    double* a = use_smem ? sm : a_g;
    double* q = use_smem ? sm + (size_t)m * n : q_g;
    const int tid = threadIdx.x, lane = tid & 31, wp = tid >> 5;
    if (use_smem)
        for (int idx = tid; idx < m * n; idx += QR_THREADS) a[idx] = a_g[idx];
    __syncthreads();

    for (int k = 0; k < n; ++k) {
        double part = 0.0;
        for (int i = k + 1 + tid; i < m; i += QR_THREADS) part += a[i * n + k] * a[i * n + k];
        const double xn2 = qr_block_sum(part, red);
        const double alpha = a[k * n + k];
        double tau = 0.0, beta = alpha, scal = 0.0;
        if (xn2 != 0.0) {
            beta = -copysign(sqrt(alpha * alpha + xn2), alpha);
            tau = (beta - alpha) / beta;
            scal = 1.0 / (alpha - beta);
        }
        __syncthreads();  // everyone has read alpha
        for (int i = k + 1 + tid; i < m; i += QR_THREADS) a[i * n + k] *= scal;
        if (tid == 0) {
            a[k * n + k] = beta;
            s_tau[k] = tau;
        }
        __syncthreads();
        // apply H_k = I - tau v v^T (v_k = 1) to columns k+1..n-1, one warp per column
        for (int j = k + 1 + wp; j < n; j += QR_THREADS / 32) {
            double wj = (lane == 0) ? a[k * n + j] : 0.0;
            for (int i = k + 1 + lane; i < m; i += 32) wj = fma(a[i * n + k], a[i * n + j], wj);
            for (int o = 16; o > 0; o >>= 1) wj += __shfl_xor_sync(0xffffffffu, wj, o);
            wj *= tau;
            if (lane == 0) a[k * n + j] -= wj;
            for (int i = k + 1 + lane; i < m; i += 32) a[i * n + j] = fma(-a[i * n + k], wj, a[i * n + j]);
        }
        __syncthreads();
    }
    // R
    for (int idx = tid; idx < n * n; idx += QR_THREADS) {
        const int i = idx / n, j = idx - i * n;
        r_g[idx] = (j >= i) ? a[i * n + j] : 0.0;
    }
    // Q = H_0 H_1 ... H_{n-1} [I; 0]
    for (int idx = tid; idx < m * n; idx += QR_THREADS) {
        const int i = idx / n, j = idx - i * n;
        q[idx] = (i == j) ? 1.0 : 0.0;
    }
    __syncthreads();
    for (int k = n - 1; k >= 0; --k) {
        const double tau = s_tau[k];
        for (int j = k + wp; j < n; j += QR_THREADS / 32) {
            double wj = (lane == 0) ? q[k * n + j] : 0.0;
            for (int i = k + 1 + lane; i < m; i += 32) wj = fma(a[i * n + k], q[i * n + j], wj);
            for (int o = 16; o > 0; o >>= 1) wj += __shfl_xor_sync(0xffffffffu, wj, o);
            wj *= tau;
            if (lane == 0) q[k * n + j] -= wj;
            for (int i = k + 1 + lane; i < m; i += 32) q[i * n + j] = fma(-a[i * n + k], wj, q[i * n + j]);
        }
        __syncthreads();
    }
    for (int idx = tid; idx < m * n; idx += QR_THREADS) a_g[idx] = q[idx];
}

}  // namespace tn

extern "C" int tn_qr(double* a, int m, int n, double* r, void* stream) {
    using namespace tn;
    TN_CHECK_ARG(a && r && m >= n && n >= 1 && n <= 128, "tn_qr: need m >= n, 1 <= n <= 128 (got %d x %d)", m, n);
    const size_t bytes = 2 * (size_t)m * n * sizeof(double);
    const int use_smem = bytes <= 200 * 1024;
    cudaStream_t st = as_stream(stream);
    AsyncScratch qscratch;
    if (!use_smem) TN_CUDA(qscratch.alloc((size_t)m * n * sizeof(double), st));
    double* q_g = static_cast<double*>(qscratch.ptr);
    if (use_smem) TN_SMEM(qr_kernel, bytes);
    qr_kernel<<<1, QR_THREADS, use_smem ? bytes : 0, st>>>(a, m, n, r, q_g, use_smem);
    TN_LAUNCH_CHECK();
    return TN_OK;
}
