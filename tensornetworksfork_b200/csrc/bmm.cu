// Per-sample small matrix products for the patch/pixel "conv-TT" layer (reference tensor/layers.py:791-890,
// SURVEY.md Appendix C): the site input of a patch core is Y_s = X_s C_k, the Jacobian of a pixel core is
// K_s X_s, the right environment enters as Y_s R_s -- one tiny product per sample, with per-sample operands on
// both sides, which the environment kernel (shared core) cannot express.
//
//   out[s, i, j] (+)= sum_k A[s*sA + i*iA + k*kA] * B[s*sB + k*kB + j*jB]          out: [S, I, J] contiguous
//
// HBM-bound streaming work (a few flops per byte): one CTA per sample group, operands staged in shared memory
// with coalesced loads, each thread owns output elements (i, j) of one sample.
#include <stdlib.h>
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


// ---- row-reduced outer product: out[i, j] (+)= sum_row w[row] * G[row*ldg + i] * W[row*ldw + j]      (ra x m, K = rows)
// The right-hand side / J^T pass of a patch core whose folded site input W is thousands of columns wide (conv-TT): a
// tall-skinny GEMM G^T W that streams W once (HBM-bound).  CTA tile: all ra rows x 64 columns, 16 sample rows per stage;
// row ranges are split across blockIdx.y and combined with fp64 atomics.
constexpr int OR_TN = 64, OR_KC = 16, OR_MAXRA = 128;
__global__ void __launch_bounds__(256)
outer_rows_kernel(const double* __restrict__ G, int64_t ldg, int gdiv, int ra, const double* __restrict__ W, int64_t ldw, int m,
                  const double* __restrict__ w, int64_t rows, int64_t rows_per_split, double* __restrict__ out) {
    __shared__ double sG[OR_KC][OR_MAXRA + 1];
    __shared__ double sW[OR_KC][OR_TN + 1];
    const int tid = threadIdx.x;
    const int tj = tid & 15, ti = tid >> 4;            // 16 x 16 threads: columns tj + 16*q, rows ti + 16*p
    const int j0 = blockIdx.x * OR_TN;
    const int64_t k_begin = (int64_t)blockIdx.y * rows_per_split;
    const int64_t k_end = (k_begin + rows_per_split < rows) ? k_begin + rows_per_split : rows;
    double acc[OR_MAXRA / 16][4];
#pragma unroll
    for (int p = 0; p < OR_MAXRA / 16; ++p)
#pragma unroll
        for (int q = 0; q < 4; ++q) acc[p][q] = 0.0;
    const int np = (ra + 15) / 16;
    for (int64_t k0 = k_begin; k0 < k_end; k0 += OR_KC) {
        __syncthreads();
        for (int idx = tid; idx < OR_KC * ra; idx += 256) {
            const int kk = idx / ra, i = idx - kk * ra;
            const int64_t row = k0 + kk;
            double v = 0.0;
            if (row < k_end) v = G[(gdiv == 1 ? row : row / gdiv) * ldg + i] * (w ? w[row] : 1.0);
            sG[kk][i] = v;
        }
        for (int idx = tid; idx < OR_KC * OR_TN; idx += 256) {
            const int kk = idx >> 6, j = idx & 63;
            const int64_t row = k0 + kk;
            sW[kk][j] = (row < k_end && j0 + j < m) ? W[row * ldw + j0 + j] : 0.0;
        }
        __syncthreads();
#pragma unroll 4
        for (int kk = 0; kk < OR_KC; ++kk) {
            double b[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) b[q] = sW[kk][tj + 16 * q];
#pragma unroll
            for (int p = 0; p < OR_MAXRA / 16; ++p) {
                if (p < np) {
                    const double a = sG[kk][ti + 16 * p];      // entries past ra are never stored below
#pragma unroll
                    for (int q = 0; q < 4; ++q) acc[p][q] = fma(a, b[q], acc[p][q]);
                }
            }
        }
    }
#pragma unroll
    for (int p = 0; p < OR_MAXRA / 16; ++p) {
        const int i = ti + 16 * p;
        if (p < np && i < ra) {
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int j = j0 + tj + 16 * q;
                if (j < m) atomicAdd(out + (int64_t)i * m + j, acc[p][q]);
            }
        }
    }
}

// ---- row-wise products with a shared matrix: z[row, i] = sum_j W[row*ldw + j] * V[i*ldv + j]          (rows x ra, K = m)
// The J v pass of a conv-TT patch core (first einsum of the matvec, tensor/network.py:789): W is streamed once.
// CTA tile: 64 rows x all ra outputs, 16 columns of W / V per stage.
constexpr int RD_TR = 64, RD_KC = 16;
__global__ void __launch_bounds__(256)
rows_dot_kernel(const double* __restrict__ W, int64_t ldw, int m, const double* __restrict__ V, int64_t ldv, int ra, int64_t rows,
                double* __restrict__ z, int64_t ldz) {
    __shared__ double sW[RD_TR][RD_KC + 1];
    __shared__ double sV[OR_MAXRA][RD_KC + 1];
    const int tid = threadIdx.x;
    const int tr = tid & 15, ti = tid >> 4;            // rows tr + 16*q (4 per thread), outputs ti + 16*p
    const int64_t r0 = (int64_t)blockIdx.x * RD_TR;
    double acc[OR_MAXRA / 16][4];
#pragma unroll
    for (int p = 0; p < OR_MAXRA / 16; ++p)
#pragma unroll
        for (int q = 0; q < 4; ++q) acc[p][q] = 0.0;
    const int np = (ra + 15) / 16;
    for (int j0 = 0; j0 < m; j0 += RD_KC) {
        __syncthreads();
        for (int idx = tid; idx < RD_TR * RD_KC; idx += 256) {
            const int rr = idx >> 4, jj = idx & 15;
            const int64_t row = r0 + rr;
            sW[rr][jj] = (row < rows && j0 + jj < m) ? W[row * ldw + j0 + jj] : 0.0;
        }
        for (int idx = tid; idx < ra * RD_KC; idx += 256) {
            const int i = idx >> 4, jj = idx & 15;
            sV[i][jj] = (j0 + jj < m) ? V[(int64_t)i * ldv + j0 + jj] : 0.0;
        }
        __syncthreads();
#pragma unroll 4
        for (int jj = 0; jj < RD_KC; ++jj) {
            double a[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) a[q] = sW[tr + 16 * q][jj];
#pragma unroll
            for (int p = 0; p < OR_MAXRA / 16; ++p) {
                if (p < np) {
                    const double b = sV[ti + 16 * p][jj];
#pragma unroll
                    for (int q = 0; q < 4; ++q) acc[p][q] = fma(a[q], b, acc[p][q]);
                }
            }
        }
    }
#pragma unroll
    for (int p = 0; p < OR_MAXRA / 16; ++p) {
        const int i = ti + 16 * p;
        if (p < np && i < ra) {
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int64_t row = r0 + tr + 16 * q;
                if (row < rows) z[row * ldz + i] = acc[p][q];
            }
        }
    }
}

// ---- FP64 tensor-core (DMMA m8n8k4) versions of the two passes.  Both are tall-skinny GEMMs with ~ra/4 flop per byte of W,
//      i.e. FP64-pipe bound at ra = 38; the FMA kernels above stay as the fallback for unaligned operands.
constexpr int TD_LD = 20;     // padded row of a staged 16-wide K slab (doubles): 40 words = 8 mod 32, so the 16 lanes of a
                              // 64-bit shared-memory wavefront (4 rows x 4 k) hit 32 distinct banks

// z[row, i] = sum_j W[row, j] V[i, j]:  M = rows (128 per CTA, 16 per warp), N = ra (NT tiles of 8), K = m (16 per stage).
template <int NT>
__global__ void __launch_bounds__(256)
rows_dot_dmma_kernel(const double* __restrict__ W, int64_t ldw, int m, const double* __restrict__ V, int64_t ldv, int ra_total,
                     int64_t rows, double* __restrict__ z, int64_t ldz) {
    __shared__ double sW[128 * TD_LD];
    __shared__ double sV[NT * 8 * TD_LD];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int fr = lane >> 2, fk = lane & 3;
    const int64_t r0 = (int64_t)blockIdx.x * 128;
    const int i0 = blockIdx.y * (NT * 8);                  // this CTA's block of outputs (ra_total may exceed NT*8)
    V += (int64_t)i0 * ldv;
    z += i0;
    const int ra = min(NT * 8, ra_total - i0);
    double acc[2][NT][2];
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < NT; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
    // the next 16-column slab is fetched into registers while the current one is multiplied
    constexpr int NVR = (NT * 8 * 16 + 255) / 256;
    double wreg[8], vreg[NVR];
    auto fetch = [&](int j0) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int idx = tid + u * 256;
            const int rr = idx >> 4, jj = idx & 15;
            const int64_t row = r0 + rr;
            wreg[u] = (row < rows && j0 + jj < m) ? W[row * ldw + j0 + jj] : 0.0;
        }
#pragma unroll
        for (int u = 0; u < NVR; ++u) {
            const int idx = tid + u * 256;
            const int i = idx >> 4, jj = idx & 15;
            vreg[u] = (idx < NT * 8 * 16 && i < ra && j0 + jj < m) ? V[(int64_t)i * ldv + j0 + jj] : 0.0;
        }
    };
    fetch(0);
    for (int j0 = 0; j0 < m; j0 += 16) {
        __syncthreads();
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int idx = tid + u * 256;
            sW[(idx >> 4) * TD_LD + (idx & 15)] = wreg[u];
        }
#pragma unroll
        for (int u = 0; u < NVR; ++u) {
            const int idx = tid + u * 256;
            if (idx < NT * 8 * 16) sV[(idx >> 4) * TD_LD + (idx & 15)] = vreg[u];
        }
        __syncthreads();
        if (j0 + 16 < m) fetch(j0 + 16);
        const double* a_s = sW + (warp * 16 + fr) * TD_LD + fk;
        const double* b_s = sV + fr * TD_LD + fk;
#pragma unroll
        for (int k4 = 0; k4 < 16; k4 += 4) {
            const double a0 = a_s[k4], a1 = a_s[8 * TD_LD + k4];
#pragma unroll
            for (int j = 0; j < NT; ++j) {
                const double b = b_s[j * 8 * TD_LD + k4];
                dmma884(acc[0][j][0], acc[0][j][1], a0, b);
                dmma884(acc[1][j][0], acc[1][j][1], a1, b);
            }
        }
    }
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const int64_t row = r0 + warp * 16 + i * 8 + fr;
        if (row >= rows) continue;
#pragma unroll
        for (int j = 0; j < NT; ++j) {
            const int col = j * 8 + fk * 2;
            if (col < ra) z[row * ldz + col] = acc[i][j][0];
            if (col + 1 < ra) z[row * ldz + col + 1] = acc[i][j][1];
        }
    }
}

// out[i, j] += sum_row w[row] G[row, i] W[row, j]:  M = ra (MT tiles of 8), N = 64 columns per CTA (one 8-column tile per warp),
// K = rows (16 per stage); row ranges split over blockIdx.y, combined with fp64 atomics.
template <int MT>
__global__ void __launch_bounds__(256)
outer_rows_dmma_kernel(const double* __restrict__ G, int64_t ldg, int gdiv, int ra, const double* __restrict__ W, int64_t ldw, int m,
                       const double* __restrict__ w, int64_t rows, int64_t rows_per_split, double* __restrict__ out) {
    constexpr int GL = MT * 8 + 4;      // [k][i] slab; row stride = +-8 words mod 32: conflict-free fragment reads
    constexpr int WL = 64 + 4;          // [k][j] slab, same
    __shared__ double sG[16 * GL];
    __shared__ double sW[16 * WL];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int fr = lane >> 2, fk = lane & 3;
    const int j0 = blockIdx.x * 64;
    const int64_t k_begin = (int64_t)blockIdx.y * rows_per_split;
    const int64_t k_end = (k_begin + rows_per_split < rows) ? k_begin + rows_per_split : rows;
    double acc[MT][2];
#pragma unroll
    for (int i = 0; i < MT; ++i) acc[i][0] = acc[i][1] = 0.0;
    constexpr int NGR = (16 * MT * 8 + 255) / 256;
    double greg[NGR], wreg[4];
    auto fetch = [&](int64_t k0) {
#pragma unroll
        for (int u = 0; u < NGR; ++u) {
            const int idx = tid + u * 256;
            const int kk = idx / (MT * 8), i = idx - kk * (MT * 8);
            const int64_t row = k0 + kk;
            double v = 0.0;
            if (idx < 16 * MT * 8 && row < k_end && i < ra) v = G[(gdiv == 1 ? row : row / gdiv) * ldg + i] * (w ? w[row] : 1.0);
            greg[u] = v;
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int idx = tid + u * 256;
            const int kk = idx >> 6, j = idx & 63;
            const int64_t row = k0 + kk;
            wreg[u] = (row < k_end && j0 + j < m) ? W[row * ldw + j0 + j] : 0.0;
        }
    };
    if (k_begin < k_end) fetch(k_begin);
    for (int64_t k0 = k_begin; k0 < k_end; k0 += 16) {
        __syncthreads();
#pragma unroll
        for (int u = 0; u < NGR; ++u) {
            const int idx = tid + u * 256;
            if (idx < 16 * MT * 8) sG[(idx / (MT * 8)) * GL + (idx % (MT * 8))] = greg[u];
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int idx = tid + u * 256;
            sW[(idx >> 6) * WL + (idx & 63)] = wreg[u];
        }
        __syncthreads();
        if (k0 + 16 < k_end) fetch(k0 + 16);
#pragma unroll
        for (int k4 = 0; k4 < 16; k4 += 4) {
            const double b = sW[(k4 + fk) * WL + warp * 8 + fr];          // B[k][n] = W[row k][column n]
#pragma unroll
            for (int i = 0; i < MT; ++i) {
                const double a = sG[(k4 + fk) * GL + i * 8 + fr];         // A[i][k] = G[row k][i]
                dmma884(acc[i][0], acc[i][1], a, b);
            }
        }
    }
#pragma unroll
    for (int i = 0; i < MT; ++i) {
        const int row = i * 8 + fr;
        if (row >= ra) continue;
        const int col = j0 + warp * 8 + fk * 2;
        if (col < m) atomicAdd(out + (int64_t)row * m + col, acc[i][0]);
        if (col + 1 < m) atomicAdd(out + (int64_t)row * m + col + 1, acc[i][1]);
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
    if (smem > 48 * 1024) TN_SMEM(bmm_kernel, smem);
    int64_t blocks = ceil_div64(S, spb);
    if (blocks > 16LL * sm_count()) blocks = 16LL * sm_count();
    bmm_kernel<<<(unsigned)blocks, 256, smem, as_stream(stream)>>>(A, sA, iA, kA, B, sB, kB, jB, out, S, I, K, J, accumulate, spb);
    TN_LAUNCH_CHECK();
    return TN_OK;
}

extern "C" int tn_outer_rows(const double* G, int64_t ldg, int gdiv, int ra, const double* W, int64_t ldw, int m, const double* w,
                             int64_t rows, double* out, int accumulate, void* stream) {
    using namespace tn;
    TN_CHECK_ARG(G && W && out && ra >= 1 && ra <= OR_MAXRA && m >= 1 && rows >= 0 && gdiv >= 1, "tn_outer_rows: bad arguments (ra <= %d)", OR_MAXRA);
    cudaStream_t st = as_stream(stream);
    if (!accumulate) TN_CUDA(cudaMemsetAsync(out, 0, (size_t)ra * m * sizeof(double), st));
    if (rows == 0) return TN_OK;
    const int64_t gx = ceil_div64(m, OR_TN);
    int64_t splits = ceil_div64(4LL * sm_count(), gx);
    const int64_t max_splits = ceil_div64(rows, 4 * OR_KC);
    if (splits > max_splits) splits = max_splits;
    if (splits < 1) splits = 1;
    if (splits > 65535) splits = 65535;
    const int64_t rps = ceil_div64(ceil_div64(rows, splits), OR_KC) * OR_KC;
    splits = ceil_div64(rows, rps);
    dim3 grid((unsigned)gx, (unsigned)splits);
    const int mt = (ra + 7) / 8;
    if (!getenv("TN_CONV_NO_DMMA")) {
#define TN_OR_DMMA(MTV) outer_rows_dmma_kernel<MTV><<<grid, 256, 0, st>>>(G, ldg, gdiv, ra, W, ldw, m, w, rows, rps, out)
        if (mt <= 1) TN_OR_DMMA(1);
        else if (mt <= 2) TN_OR_DMMA(2);
        else if (mt <= 4) TN_OR_DMMA(4);
        else if (mt <= 5) TN_OR_DMMA(5);
        else if (mt <= 8) TN_OR_DMMA(8);
        else if (mt <= 12) TN_OR_DMMA(12);
        else TN_OR_DMMA(16);
#undef TN_OR_DMMA
    } else {
        outer_rows_kernel<<<grid, 256, 0, st>>>(G, ldg, gdiv, ra, W, ldw, m, w, rows, rps, out);
    }
    TN_LAUNCH_CHECK();
    return TN_OK;
}

extern "C" int tn_rows_dot(const double* W, int64_t ldw, int m, const double* V, int64_t ldv, int ra, int64_t rows, double* z,
                           int64_t ldz, void* stream) {
    using namespace tn;
    TN_CHECK_ARG(W && V && z && ra >= 1 && m >= 1 && rows >= 0 && ldz >= ra, "tn_rows_dot: bad arguments");
    if (rows == 0) return TN_OK;
    cudaStream_t st = as_stream(stream);
    if (!getenv("TN_CONV_NO_DMMA")) {
        const int64_t gx = ceil_div64(rows, 128);
        TN_CHECK_ARG(gx <= 0x7fffffff, "tn_rows_dot: too many rows");
        const int nt = (ra + 7) / 8;
#define TN_RD_DMMA(NTV) rows_dot_dmma_kernel<NTV><<<dim3((unsigned)gx, (unsigned)ceil_div64(ra, NTV * 8)), 256, 0, st>>>(W, ldw, m, V, ldv, ra, rows, z, ldz)
        if (nt <= 1) TN_RD_DMMA(1);
        else if (nt <= 2) TN_RD_DMMA(2);
        else if (nt <= 4) TN_RD_DMMA(4);
        else if (nt <= 5) TN_RD_DMMA(5);
        else if (nt <= 8) TN_RD_DMMA(8);
        else if (nt <= 12) TN_RD_DMMA(12);
        else TN_RD_DMMA(16);
#undef TN_RD_DMMA
        TN_LAUNCH_CHECK();
        return TN_OK;
    }
    const int64_t grid = ceil_div64(rows, RD_TR);
    TN_CHECK_ARG(grid <= 0x7fffffff, "tn_rows_dot: too many rows");
    for (int i0 = 0; i0 < ra; i0 += OR_MAXRA) {
        const int rb = (ra - i0 < OR_MAXRA) ? ra - i0 : OR_MAXRA;
        rows_dot_kernel<<<(unsigned)grid, 256, 0, st>>>(W, ldw, m, V + (int64_t)i0 * ldv, ldv, rb, rows, z + i0, ldz);
        TN_LAUNCH_CHECK();
    }
    return TN_OK;
}
