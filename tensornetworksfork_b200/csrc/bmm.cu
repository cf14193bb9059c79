// Per-sample small matrix products for the patch/pixel "conv-TT" layer (reference tensor/layers.py:791-890,
// SURVEY.md Appendix C): the site input of a patch core is Y_s = X_s C_k, the Jacobian of a pixel core is
// K_s X_s, the right environment enters as Y_s R_s -- one tiny product per sample, with per-sample operands on
// both sides, which the environment kernel (shared core) cannot express.
//
//   out[s, i, j] (+)= sum_k A[s*sA + i*iA + k*kA] * B[s*sB + k*kB + j*jB]          out: [S, I, J] contiguous
//
// HBM-bound streaming work (a few flops per byte): one CTA per sample group, operands staged in shared memory
// with coalesced loads, each thread owns output elements (i, j) of one sample.
#include "common.cuh"

namespace tn {

__global__ void __launch_bounds__(256)
bmm_kernel(const double* __restrict__ A, int64_t sA, int64_t iA, int64_t kA, const double* __restrict__ B, int64_t sB,
           int64_t kB, int64_t jB, double* __restrict__ out, int64_t S, int I, int K, int J, int accumulate, int spb) {
    extern __shared__ double sm[];
    const int nA = I * K, nB = K * J, nO = I * J;
    double* a_s = sm;                       // [spb][I*K]
    double* b_s = sm + (size_t)spb * nA;    // [spb][K*J]
    for (int64_t s0 = (int64_t)blockIdx.x * spb; s0 < S; s0 += (int64_t)gridDim.x * spb) {
        const int ns = (int)((S - s0 < spb) ? S - s0 : spb);
        __syncthreads();
        // stage A: iterate in the order that is contiguous in memory (kA == 1: k fastest, else i fastest)
        for (int idx = threadIdx.x; idx < ns * nA; idx += blockDim.x) {
            const int ls = idx / nA, e = idx - ls * nA;
            int i, k;
            if (kA <= iA) { i = e / K; k = e - i * K; } else { k = e / I; i = e - k * I; }
            a_s[ls * nA + i * K + k] = A[(s0 + ls) * sA + (int64_t)i * iA + (int64_t)k * kA];
        }
        for (int idx = threadIdx.x; idx < ns * nB; idx += blockDim.x) {
            const int ls = idx / nB, e = idx - ls * nB;
            int k, j;
            if (jB <= kB) { k = e / J; j = e - k * J; } else { j = e / K; k = e - j * K; }
            b_s[ls * nB + k * J + j] = B[(s0 + ls) * sB + (int64_t)k * kB + (int64_t)j * jB];
        }
        __syncthreads();
        for (int idx = threadIdx.x; idx < ns * nO; idx += blockDim.x) {
            const int ls = idx / nO, e = idx - ls * nO;
            const int i = e / J, j = e - i * J;
            const double* ar = a_s + ls * nA + i * K;
            const double* bc = b_s + ls * nB + j;
            double acc = 0.0;
            for (int k = 0; k < K; ++k) acc = fma(ar[k], bc[k * J], acc);
            double* o = out + (s0 + ls) * (int64_t)nO + e;
            *o = accumulate ? *o + acc : acc;
        }
    }
}

}  // namespace tn

extern "C" int tn_bmm(const double* A, int64_t sA, int64_t iA, int64_t kA, const double* B, int64_t sB, int64_t kB, int64_t jB,
                      double* out, int64_t S, int I, int K, int J, int accumulate, void* stream) {
    using namespace tn;
    TN_CHECK_ARG(A && B && out && S >= 0 && I >= 1 && K >= 1 && J >= 1, "tn_bmm: bad arguments");
    if (S == 0) return TN_OK;
    const size_t per = ((size_t)I * K + (size_t)K * J) * sizeof(double);
    TN_CHECK_ARG(per <= 200 * 1024, "tn_bmm: operands of one sample (%zu bytes) do not fit shared memory", per);
    // samples per CTA: enough output elements to keep 256 threads busy, within 64 KB of staging
    int spb = (int)((256 + (size_t)I * J - 1) / ((size_t)I * J));
    if (spb < 1) spb = 1;
    while (spb > 1 && spb * per > 64 * 1024) --spb;
    if (spb > S) spb = (int)S;
    const size_t smem = spb * per;
    static size_t configured = 0;
    if (smem > 48 * 1024 && smem > configured) {
        TN_CUDA(cudaFuncSetAttribute(bmm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = smem;
    }
    int64_t blocks = ceil_div64(S, spb);
    if (blocks > 16LL * sm_count()) blocks = 16LL * sm_count();
    bmm_kernel<<<(unsigned)blocks, 256, smem, as_stream(stream)>>>(A, sA, iA, kA, B, sB, kB, jB, out, S, I, K, J, accumulate, spb);
    TN_LAUNCH_CHECK();
    return TN_OK;
}
