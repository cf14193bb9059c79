// Regularised local solve: blocked Cholesky + triangular solves, fp64.
//
// Replaces torch.linalg.cholesky + torch.cholesky_solve inside TensorNetwork.solve_system
// (reference tensor/network.py:311-320).  Two-level right-looking factorisation of the lower
// triangle, in place: 64-wide diagonal blocks are factored and inverted by one CTA, the panel
// below is a small GEMM with the inverted block, trailing updates are SYRK tiles (64x64 inside an
// outer panel, 128x128 with K = outer panel width for the rest, so the C-tile read-modify-write is
// amortised over a long K).  P ranges from 4 to 41 876 (14 GB) on the named configurations.
#include <stdlib.h>
#include <cooperative_groups.h>
#include "common.cuh"

namespace tn {

constexpr int CH_NB = 64;

// ---- diagonal block: factor (lower) and invert ------------------------------------------------
// Latency work on one CTA.  Factorisation: warp 0 alone, left-looking, two rows per lane, __syncwarp only
// (no block barriers in the 64-step dependency chain).  Inversion: all 8 warps, 4 lanes per column of
// L^{-1}; lane l keeps the entries x[q], q = l (mod 4), of its column in registers, so a row step is a
// 16-term dot product, two shuffles and a divide.
__device__ void potrf_diag_body(double* __restrict__ A, int64_t lda, int64_t j, int nb, double* __restrict__ Linv,
                                int* __restrict__ info, double* dsm) {
    double (*a)[CH_NB + 1] = reinterpret_cast<double (*)[CH_NB + 1]>(dsm);
    __shared__ int bad;
    __shared__ double s_piv;
    __shared__ double s_rdiag[CH_NB];
    if (*info != 0) return;
    const int tid = threadIdx.x;
    if (tid == 0) bad = 0;
    for (int idx = tid; idx < CH_NB * CH_NB; idx += 256) {
        const int r = idx >> 6, c = idx & 63;
        a[r][c] = (r < nb && c <= r) ? A[(j + r) * lda + j + c] : ((r == c) ? 1.0 : 0.0);   // identity padding
    }
    __syncthreads();
    // ---- factorisation, left-looking: thread (r, l) owns a quarter of row r's dot product
    {
        const int r = tid >> 2, l = tid & 3;
        for (int c = 0; c < nb; ++c) {
            double s0 = 0.0, s1 = 0.0;
            int t = l;
            for (; t + 4 < c; t += 8) {
                s0 = fma(a[r][t], a[c][t], s0);
                s1 = fma(a[r][t + 4], a[c][t + 4], s1);
            }
            if (t < c) s0 = fma(a[r][t], a[c][t], s0);
            double sd = s0 + s1;
            sd += __shfl_xor_sync(0xffffffffu, sd, 1);
            sd += __shfl_xor_sync(0xffffffffu, sd, 2);
            const double v = a[r][c] - sd;
            if (r == c && l == 0) s_piv = v;
            __syncthreads();
            const double d = s_piv;
            if (!(d > 0.0)) {   // also catches NaN; uniform across the CTA
                if (tid == 0) {
                    bad = 1;
                    *info = (int)(j + c + 1);
                }
                break;
            }
            const double root = sqrt(d);
            if (l == 0 && r >= c) a[r][c] = (r == c) ? root : v / root;
            __syncthreads();
        }
    }
    __syncthreads();
    if (bad) return;
    for (int idx = tid; idx < nb * nb; idx += 256) {
        const int r = idx / nb, c = idx % nb;
        if (c <= r) A[(j + r) * lda + j + c] = a[r][c];
    }
    if (tid < CH_NB) s_rdiag[tid] = 1.0 / a[tid][tid];
    __syncthreads();
    // ---- inverse: column t of X = L^{-1}, 4 lanes per column; lane l keeps x[q], q = l (mod 4), in registers
    {
        const int t = tid >> 2, l = tid & 3;
        double x[16];          // x[m] = X[4m + l][t]
#pragma unroll
        for (int m = 0; m < 16; ++m) x[m] = (4 * m + l == t) ? s_rdiag[t] : 0.0;
        for (int r = 1; r < CH_NB; ++r) {
            double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
#pragma unroll
            for (int m = 0; m < 16; m += 4) {       // entries with q >= r are still zero in x, so no bound is needed... except q == r
                s0 = fma(a[r][4 * m + l], x[m], s0);
                s1 = fma(a[r][4 * (m + 1) + l], x[m + 1], s1);
                s2 = fma(a[r][4 * (m + 2) + l], x[m + 2], s2);
                s3 = fma(a[r][4 * (m + 3) + l], x[m + 3], s3);
            }
            double sdot = (s0 + s1) + (s2 + s3);
            sdot += __shfl_xor_sync(0xffffffffu, sdot, 1);
            sdot += __shfl_xor_sync(0xffffffffu, sdot, 2);
            if (r > t) {
                const double xr = -sdot * s_rdiag[r];
#pragma unroll
                for (int m = 0; m < 16; ++m)
                    if (4 * m + l == r) x[m] = xr;
            }
        }
#pragma unroll
        for (int m = 0; m < 16; ++m) Linv[(4 * m + l) * CH_NB + t] = x[m];
    }
}

__global__ void __launch_bounds__(256)
potrf_diag_kernel(double* __restrict__ A, int64_t lda, int64_t j, int nb, double* __restrict__ Linv, int* __restrict__ info) {
    extern __shared__ double dsm[];
    potrf_diag_body(A, lda, j, nb, Linv, info, dsm);
}

// ---- panel: X = B * Linv^T for the rows below the diagonal block -------------------------------
__device__ void trsm_panel_body(double* __restrict__ A, int64_t lda, int64_t j, int nb, const double* __restrict__ Linv, int64_t P,
                                int64_t rb, double* dsm) {
    double (*b)[CH_NB + 1] = reinterpret_cast<double (*)[CH_NB + 1]>(dsm);
    double (*li)[CH_NB + 1] = reinterpret_cast<double (*)[CH_NB + 1]>(dsm + CH_NB * (CH_NB + 1));
    const int tid = threadIdx.x;
    const int64_t r0 = j + nb + rb * CH_NB;
    for (int idx = tid; idx < CH_NB * CH_NB; idx += 256) {
        const int r = idx >> 6, c = idx & 63;
        const int64_t row = r0 + r;
        b[r][c] = (row < P && c < nb) ? A[row * lda + j + c] : 0.0;
        li[r][c] = Linv[idx];
    }
    __syncthreads();
    const int tx = tid & 15, ty = tid >> 4;
    double acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int q = 0; q < 4; ++q) acc[i][q] = 0.0;
    for (int t = 0; t < nb; ++t) {
        double x[4], y[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) x[i] = b[ty * 4 + i][t];
#pragma unroll
        for (int q = 0; q < 4; ++q) y[q] = li[tx + 16 * q][t];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int q = 0; q < 4; ++q) acc[i][q] = fma(x[i], y[q], acc[i][q]);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int64_t row = r0 + ty * 4 + i;
        if (row >= P) continue;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int c = tx + 16 * q;
            if (c < nb) A[row * lda + j + c] = acc[i][q];
        }
    }
}

__global__ void __launch_bounds__(256)
trsm_panel_kernel(double* __restrict__ A, int64_t lda, int64_t j, int nb, const double* __restrict__ Linv, int64_t P,
                  const int* __restrict__ info) {
    extern __shared__ double dsm[];
    if (*info != 0) return;
    trsm_panel_body(A, lda, j, nb, Linv, P, blockIdx.x, dsm);
}

// ---- trailing update: C[i][q] -= sum_t A[i][k0+t] * A[q][k0+t] on the lower tiles of
//      columns [c0, c1), rows [c0, P).  BT x BT tile per CTA, 256 threads, (BT/16)^2 per thread.
template <int BT>
__device__ void syrk_update_body(double* __restrict__ A, int64_t lda, int64_t P, int64_t c0, int64_t c1, int64_t k0, int kb,
                                 int bi, int bj) {
    constexpr int TM = BT / 16;
    constexpr int KC = 16;
    __shared__ double sa[KC][BT + 2];
    __shared__ double sb[KC][BT + 2];
    if (bj > bi) return;  // tile strictly above the diagonal
    const int64_t r0 = c0 + (int64_t)bi * BT, q0 = c0 + (int64_t)bj * BT;
    if (r0 >= P || q0 >= c1) return;
    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;
    double acc[TM][TM];
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int q = 0; q < TM; ++q) acc[i][q] = 0.0;

    const int lt = tid & 15;   // k lane while loading
    const int lr = tid >> 4;   // row lane while loading
    for (int kk = 0; kk < kb; kk += KC) {
        __syncthreads();
#pragma unroll
        for (int rr = 0; rr < BT; rr += 16) {
            const int r = rr + lr;
            const int64_t ra = r0 + r, rb = q0 + r;
            const bool kin = (kk + lt) < kb;
            sa[lt][r] = (kin && ra < P) ? A[ra * lda + k0 + kk + lt] : 0.0;
            sb[lt][r] = (kin && rb < P && rb < c1) ? A[rb * lda + k0 + kk + lt] : 0.0;
        }
        __syncthreads();
#pragma unroll
        for (int t = 0; t < KC; ++t) {
            double x[TM], y[TM];
#pragma unroll
            for (int i = 0; i < TM; ++i) x[i] = sa[t][ty * TM + i];
#pragma unroll
            for (int q = 0; q < TM; ++q) y[q] = sb[t][tx + 16 * q];
#pragma unroll
            for (int i = 0; i < TM; ++i)
#pragma unroll
                for (int q = 0; q < TM; ++q) acc[i][q] = fma(x[i], y[q], acc[i][q]);
        }
    }
#pragma unroll
    for (int i = 0; i < TM; ++i) {
        const int64_t row = r0 + ty * TM + i;
        if (row >= P) continue;
#pragma unroll
        for (int q = 0; q < TM; ++q) {
            const int64_t col = q0 + tx + 16 * q;
            if (col < c1 && col <= row) A[row * lda + col] -= acc[i][q];
        }
    }
}

template <int BT>
__global__ void __launch_bounds__(256)
syrk_update_kernel(double* __restrict__ A, int64_t lda, int64_t P, int64_t c0, int64_t c1, int64_t k0, int kb,
                   const int* __restrict__ info) {
    if (*info != 0) return;
    syrk_update_body<BT>(A, lda, P, c0, c1, k0, kb, blockIdx.y, blockIdx.x);
}

// ---- whole factorisation in ONE cooperative launch for small systems (P <= CF_MAXP): the three phases of every
//      64-column step are separated by grid-wide barriers instead of kernel boundaries (54 launches -> 1 at P = 1152).
constexpr int CF_MAXP = 4096;
__global__ void __launch_bounds__(256)
cholesky_fused_kernel(double* __restrict__ A, int64_t lda, int P, double* __restrict__ Linv_all, int* __restrict__ info) {
    cooperative_groups::grid_group grid = cooperative_groups::this_grid();
    extern __shared__ double dsm[];
    const int nblk = (P + CH_NB - 1) / CH_NB;
    for (int b = 0; b < nblk; ++b) {
        const int64_t j = (int64_t)b * CH_NB;
        const int nb = min(CH_NB, P - (int)j);
        if (blockIdx.x == 0) potrf_diag_body(A, lda, j, nb, Linv_all + (size_t)b * CH_NB * CH_NB, info, dsm);
        __threadfence();
        grid.sync();
        if (*reinterpret_cast<volatile int*>(info) != 0) return;      // uniform: every CTA reads the same flag after the barrier
        const int below = P - (int)j - nb;
        if (below <= 0) break;
        const int nrb = (below + CH_NB - 1) / CH_NB;
        for (int rb = blockIdx.x; rb < nrb; rb += gridDim.x) {
            __syncthreads();
            trsm_panel_body(A, lda, j, nb, Linv_all + (size_t)b * CH_NB * CH_NB, P, rb, dsm);
        }
        __threadfence();
        grid.sync();
        const int ntile = nrb * (nrb + 1) / 2;
        for (int t = blockIdx.x; t < ntile; t += gridDim.x) {
            int bi = (int)((sqrt(8.0 * t + 1.0) - 1.0) * 0.5);
            while ((bi + 1) * (bi + 2) / 2 <= t) ++bi;
            while (bi * (bi + 1) / 2 > t) --bi;
            const int bj = t - bi * (bi + 1) / 2;
            __syncthreads();
            syrk_update_body<64>(A, lda, P, j + nb, P, j, nb, bi, bj);
        }
        __threadfence();
        grid.sync();
    }
}

// ---- large trailing update on the FP64 tensor pipe (DMMA m8n8k4), 128x128 tile, K streamed by cp.async.
//      Same contract as syrk_update_kernel; requires kb % 16 == 0 and lda % 2 == 0, k0 % 2 == 0.
constexpr int DS_BT = 128;   // tile
constexpr int DS_KC = 16;    // k per stage
constexpr int DS_LD = DS_KC + 2;   // padded smem row (doubles): 16-byte aligned rows, conflict-free fragment reads
constexpr int DS_STAGES = 3;

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gmem_src) : "memory");
}

constexpr int DS_THREADS = 512;   // 16 warps as 4 x 4: warp tile 32 x 32 (four warps per scheduler hide the DMMA / LDS latency)
__global__ void __launch_bounds__(DS_THREADS, 1)
syrk_update_dmma_kernel(double* __restrict__ A, int64_t lda, int64_t P, int64_t c0, int64_t c1, int64_t k0, int kb,
                        const int* __restrict__ info) {
    extern __shared__ double dsm[];
    if (*info != 0) return;
    const int bi = blockIdx.y, bj = blockIdx.x;
    if (bj > bi) return;
    const int64_t r0 = c0 + (int64_t)bi * DS_BT, q0 = c0 + (int64_t)bj * DS_BT;
    if (r0 >= P || q0 >= c1) return;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wm = warp >> 2, wn = warp & 3;
    double* sA = dsm;                                  // [stage][128][DS_LD]
    double* sB = dsm + DS_STAGES * DS_BT * DS_LD;

    double acc[4][4][2];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

    // each thread copies 2 + 2 sixteen-byte pieces per stage: piece = (row, half-pair index 0..7)
    auto issue = [&](int stage, int kk) {
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const int piece = tid + u * DS_THREADS;    // 0..1023
            const int row = piece >> 3, part = piece & 7;
            int64_t ra = r0 + row, rb = q0 + row;
            if (ra >= P) ra = P - 1;                   // clamped rows are never stored
            if (rb >= P) rb = P - 1;
            cp_async16(sA + ((size_t)stage * DS_BT + row) * DS_LD + part * 2, A + ra * lda + k0 + kk + part * 2);
            cp_async16(sB + ((size_t)stage * DS_BT + row) * DS_LD + part * 2, A + rb * lda + k0 + kk + part * 2);
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    const int nk = kb / DS_KC;
    for (int s = 0; s < DS_STAGES - 1; ++s) {
        if (s < nk) issue(s, s * DS_KC);
        else asm volatile("cp.async.commit_group;" ::: "memory");
    }
    const int fr = lane >> 2, fk = lane & 3;
    for (int it = 0; it < nk; ++it) {
        asm volatile("cp.async.wait_group %0;" ::"n"(DS_STAGES - 2) : "memory");
        __syncthreads();
        const int nxt = it + DS_STAGES - 1;
        if (nxt < nk) issue(nxt % DS_STAGES, nxt * DS_KC);
        else asm volatile("cp.async.commit_group;" ::: "memory");
        const double* a_s = sA + ((size_t)(it % DS_STAGES) * DS_BT + wm * 32 + fr) * DS_LD + fk;
        const double* b_s = sB + ((size_t)(it % DS_STAGES) * DS_BT + wn * 32 + fr) * DS_LD + fk;
#pragma unroll
        for (int k4 = 0; k4 < DS_KC; k4 += 4) {
            double af[4], bf[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) af[i] = a_s[(size_t)i * 8 * DS_LD + k4];
#pragma unroll
            for (int j = 0; j < 4; ++j) bf[j] = b_s[(size_t)j * 8 * DS_LD + k4];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) dmma884(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
        }
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int64_t row = r0 + wm * 32 + i * 8 + fr;
        if (row >= P) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int64_t col = q0 + wn * 32 + j * 8 + fk * 2;
            double* cptr = A + row * lda + col;
            if (col + 1 < c1 && col + 1 <= row) {
                double2 v = *reinterpret_cast<double2*>(cptr);
                v.x -= acc[i][j][0];
                v.y -= acc[i][j][1];
                *reinterpret_cast<double2*>(cptr) = v;
            } else {
                if (col < c1 && col <= row) cptr[0] -= acc[i][j][0];
                if (col + 1 < c1 && col + 1 <= row) cptr[1] -= acc[i][j][1];
            }
        }
    }
}

// ---- triangular solves -------------------------------------------------------------------------
// y_j = Linv_j * rhs_j (forward) or Linv_j^T * rhs_j (backward), in place, one CTA of 64 threads.
__global__ void trsv_diag_kernel(double* __restrict__ rhs, int64_t j, int nb, const double* __restrict__ Linv, int transpose,
                                 const int* __restrict__ info) {
    __shared__ double x[CH_NB];
    if (*info != 0) return;
    const int t = threadIdx.x;
    x[t] = (t < nb) ? rhs[j + t] : 0.0;
    __syncthreads();
    if (t < nb) {
        double s = 0.0;
        if (!transpose) {
            for (int q = 0; q <= t; ++q) s = fma(Linv[t * CH_NB + q], x[q], s);
        } else {
            for (int q = t; q < nb; ++q) s = fma(Linv[q * CH_NB + t], x[q], s);
        }
        rhs[j + t] = s;
    }
}

// forward: rhs[i] -= sum_t A[i][j+t] * y[t] for i >= j+nb.  One warp per row.
__global__ void __launch_bounds__(256)
trsv_fwd_update_kernel(const double* __restrict__ A, int64_t lda, int64_t P, int64_t j, int nb, double* __restrict__ rhs,
                       const int* __restrict__ info) {
    __shared__ double y[CH_NB];
    if (*info != 0) return;
    if (threadIdx.x < CH_NB) y[threadIdx.x] = (threadIdx.x < nb) ? rhs[j + threadIdx.x] : 0.0;
    __syncthreads();
    const int lane = threadIdx.x & 31, wp = threadIdx.x >> 5;
    for (int64_t i = j + nb + (int64_t)blockIdx.x * 8 + wp; i < P; i += (int64_t)gridDim.x * 8) {
        const double* row = A + i * lda + j;
        double s = 0.0;
        if (lane < nb) s = row[lane] * y[lane];
        if (lane + 32 < nb) s = fma(row[lane + 32], y[lane + 32], s);
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (lane == 0) rhs[i] -= s;
    }
}

// backward: rhs[c] -= sum_t A[j+t][c] * x[t] for c < j.  One thread per column (coalesced rows).
__global__ void __launch_bounds__(256)
trsv_bwd_update_kernel(const double* __restrict__ A, int64_t lda, int64_t j, int nb, double* __restrict__ rhs,
                       const int* __restrict__ info) {
    __shared__ double x[CH_NB];
    if (*info != 0) return;
    if (threadIdx.x < CH_NB) x[threadIdx.x] = (threadIdx.x < nb) ? rhs[j + threadIdx.x] : 0.0;
    __syncthreads();
    for (int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; c < j; c += (int64_t)gridDim.x * blockDim.x) {
        double s = 0.0;
        for (int t = 0; t < nb; ++t) s = fma(A[(j + t) * lda + c], x[t], s);
        rhs[c] -= s;
    }
}

// Both substitutions in one launch for small systems (P <= TS_MAXP): one CTA, rhs in shared memory,
// blocks of 64 processed in order; the 64x64 block solves use the inverted diagonal blocks.
constexpr int TS_MAXP = 8192;
__global__ void __launch_bounds__(1024)
trsv_small_kernel(const double* __restrict__ A, int64_t lda, int P, double* __restrict__ rhs, const double* __restrict__ Linv_all,
                  const int* __restrict__ info) {
    extern __shared__ double x[];      // [P] right-hand side / solution, then [64] block result
    double* yb = x + ((P + 63) / 64) * 64;
    if (*info != 0) return;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nblk = (P + CH_NB - 1) / CH_NB;
    for (int i = tid; i < nblk * CH_NB; i += 1024) x[i] = (i < P) ? rhs[i] : 0.0;
    __syncthreads();
    // forward: L y = rhs
    for (int b = 0; b < nblk; ++b) {
        const int j = b * CH_NB;
        const int nb = min(CH_NB, P - j);
        const double* Li = Linv_all + (size_t)b * CH_NB * CH_NB;
        // y_b = Linv_b * x_b : 64 rows, 16 threads per row
        {
            const int r = tid >> 4, l = tid & 15;
            double sdot = 0.0;
            for (int q = l; q <= r; q += 16) sdot = fma(Li[r * CH_NB + q], x[j + q], sdot);
            sdot += __shfl_xor_sync(0xffffffffu, sdot, 8);
            sdot += __shfl_xor_sync(0xffffffffu, sdot, 4);
            sdot += __shfl_xor_sync(0xffffffffu, sdot, 2);
            sdot += __shfl_xor_sync(0xffffffffu, sdot, 1);
            if (l == 0) yb[r] = sdot;
        }
        __syncthreads();
        if (tid < CH_NB) x[j + tid] = (tid < nb) ? yb[tid] : 0.0;
        // x[i] -= sum_t L[i][j+t] y[t], i > block : one warp per row
        for (int i = j + CH_NB + warp; i < P; i += 32) {
            const double* row = A + (int64_t)i * lda + j;
            double sdot = row[lane] * yb[lane] + row[lane + 32] * yb[lane + 32];
            for (int o = 16; o > 0; o >>= 1) sdot += __shfl_xor_sync(0xffffffffu, sdot, o);
            if (lane == 0) x[i] -= sdot;
        }
        __syncthreads();
    }
    // backward: L^T z = y
    for (int b = nblk - 1; b >= 0; --b) {
        const int j = b * CH_NB;
        const int nb = min(CH_NB, P - j);
        const double* Li = Linv_all + (size_t)b * CH_NB * CH_NB;
        {
            const int r = tid >> 4, l = tid & 15;
            double sdot = 0.0;
            for (int q = r + l; q < CH_NB; q += 16) sdot = fma(Li[q * CH_NB + r], x[j + q], sdot);   // Linv^T
            sdot += __shfl_xor_sync(0xffffffffu, sdot, 8);
            sdot += __shfl_xor_sync(0xffffffffu, sdot, 4);
            sdot += __shfl_xor_sync(0xffffffffu, sdot, 2);
            sdot += __shfl_xor_sync(0xffffffffu, sdot, 1);
            if (l == 0) yb[r] = sdot;
        }
        __syncthreads();
        if (tid < CH_NB) x[j + tid] = (tid < nb) ? yb[tid] : 0.0;
        // x[c] -= sum_t L[j+t][c] z[t], c < j : one thread per column (coalesced rows)
        for (int c = tid; c < j; c += 1024) {
            double sdot = 0.0;
            for (int t = 0; t < nb; ++t) sdot = fma(A[(int64_t)(j + t) * lda + c], yb[t], sdot);
            x[c] -= sdot;
        }
        __syncthreads();
    }
    for (int i = tid; i < P; i += 1024) rhs[i] = x[i];
}

}  // namespace tn

// work = [inverted 64 x 64 diagonal blocks][for P > 8192: inverted 512 x 512 diagonal blocks (the substitutions' block solves as GEMVs)]
static int64_t chol_linv_elems(int64_t P) { return tn::ceil_div64(P, tn::CH_NB) * tn::CH_NB * tn::CH_NB; }
static int64_t chol_x512_elems(int64_t P) { return (P > 8192) ? tn::ceil_div64(P, 512) * 512 * 512 : 0; }
extern "C" int64_t tn_cholesky_work_elems(int64_t P) { return chol_linv_elems(P) + chol_x512_elems(P); }

namespace tn {

// ---- substitution for large systems, 512-wide super-blocks: one CTA solves the 512 x 512 triangle (using the inverted
//      64 x 64 diagonal blocks), then a grid-wide kernel applies the off-diagonal panel to the rest of the right-hand side.
constexpr int TB_W = 512;

// forward (transpose == 0): solve L[w,w] y = rhs[w] for the window w = [j0, j0+nbw); backward: L[w,w]^T z = rhs[w].
__global__ void __launch_bounds__(1024)
trsv_block_kernel(const double* __restrict__ A, int64_t lda, int64_t j0, int nbw, double* __restrict__ rhs,
                  const double* __restrict__ Linv_all, int transpose, const int* __restrict__ stop) {
    __shared__ double x[TB_W];
    __shared__ double yb[CH_NB];
    if (*stop != 0) return;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nblk = (nbw + CH_NB - 1) / CH_NB;
    for (int i = tid; i < TB_W; i += 1024) x[i] = (i < nbw) ? rhs[j0 + i] : 0.0;
    __syncthreads();
    const int r = tid >> 4, l = tid & 15;
    if (!transpose) {
        for (int b = 0; b < nblk; ++b) {
            const int j = b * CH_NB;
            const double* Li = Linv_all + (size_t)((j0 + j) / CH_NB) * CH_NB * CH_NB;
            double sdot = 0.0;
            for (int q = l; q <= r; q += 16) sdot = fma(Li[r * CH_NB + q], x[j + q], sdot);
            sdot += __shfl_xor_sync(0xffffffffu, sdot, 8);
            sdot += __shfl_xor_sync(0xffffffffu, sdot, 4);
            sdot += __shfl_xor_sync(0xffffffffu, sdot, 2);
            sdot += __shfl_xor_sync(0xffffffffu, sdot, 1);
            if (l == 0) yb[r] = sdot;
            __syncthreads();
            if (tid < CH_NB) x[j + tid] = (j + tid < nbw) ? yb[tid] : 0.0;
            {   // rows below the block, four per warp pass: all loads are issued before the reductions
                const double y0 = yb[lane], y1 = yb[lane + 32];
                for (int i0 = j + CH_NB + warp; i0 < nbw; i0 += 128) {
                    double sd[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const int i = i0 + 32 * u;
                        const double* row = A + (j0 + (i < nbw ? i : i0)) * lda + j0 + j;
                        sd[u] = row[lane] * y0 + row[lane + 32] * y1;
                    }
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1)
#pragma unroll
                        for (int u = 0; u < 4; ++u) sd[u] += __shfl_xor_sync(0xffffffffu, sd[u], o);
                    if (lane == 0) {
#pragma unroll
                        for (int u = 0; u < 4; ++u)
                            if (i0 + 32 * u < nbw) x[i0 + 32 * u] -= sd[u];
                    }
                }
            }
            __syncthreads();
        }
    } else {
        for (int b = nblk - 1; b >= 0; --b) {
            const int j = b * CH_NB;
            const int nb = min(CH_NB, nbw - j);
            const double* Li = Linv_all + (size_t)((j0 + j) / CH_NB) * CH_NB * CH_NB;
            double sdot = 0.0;
            for (int q = r + l; q < CH_NB; q += 16) sdot = fma(Li[q * CH_NB + r], x[j + q], sdot);   // Linv^T
            sdot += __shfl_xor_sync(0xffffffffu, sdot, 8);
            sdot += __shfl_xor_sync(0xffffffffu, sdot, 4);
            sdot += __shfl_xor_sync(0xffffffffu, sdot, 2);
            sdot += __shfl_xor_sync(0xffffffffu, sdot, 1);
            if (l == 0) yb[r] = sdot;
            __syncthreads();
            if (tid < CH_NB) x[j + tid] = (tid < nb) ? yb[tid] : 0.0;
            {   // columns left of the block: 2 threads per column (even / odd rows of the block), 4 independent sums each
                const int c = tid >> 1, h = tid & 1;
                double sd = 0.0;
                if (c < j) {
                    const double* col = A + (j0 + j + h) * lda + j0 + c;
                    double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
                    int t = h;
                    for (; t + 6 < nb; t += 8) {
                        s0 = fma(col[(int64_t)(t - h) * lda], yb[t], s0);
                        s1 = fma(col[(int64_t)(t - h + 2) * lda], yb[t + 2], s1);
                        s2 = fma(col[(int64_t)(t - h + 4) * lda], yb[t + 4], s2);
                        s3 = fma(col[(int64_t)(t - h + 6) * lda], yb[t + 6], s3);
                    }
                    for (; t < nb; t += 2) s0 = fma(col[(int64_t)(t - h) * lda], yb[t], s0);
                    sd = (s0 + s1) + (s2 + s3);
                }
                sd += __shfl_xor_sync(0xffffffffu, sd, 1);
                if (h == 0 && c < j) x[c] -= sd;
            }
            __syncthreads();
        }
    }
    for (int i = tid; i < nbw; i += 1024) rhs[j0 + i] = x[i];
}

// forward panel: rhs[i] -= sum_t A[i][j0+t] * y[t], i >= j0+nbw.  One warp per row, the row piece (<= 512 doubles) coalesced.
__global__ void __launch_bounds__(256)
trsv_fwd_panel_kernel(const double* __restrict__ A, int64_t lda, int64_t P, int64_t j0, int nbw, double* __restrict__ rhs,
                      const int* __restrict__ stop, const double* __restrict__ ysrc = nullptr) {
    __shared__ double y[TB_W];
    if (*stop != 0) return;
    // ysrc: the window's solution comes from the block-inverse product (trsv_gemv512_kernel); CTA 0 also stores it into rhs
    for (int i = threadIdx.x; i < TB_W; i += 256) {
        const double v = (i < nbw) ? (ysrc ? ysrc[i] : rhs[j0 + i]) : 0.0;
        y[i] = v;
        if (ysrc && blockIdx.x == 0 && i < nbw) rhs[j0 + i] = v;
    }
    __syncthreads();
    const int lane = threadIdx.x & 31, wp = threadIdx.x >> 5;
    for (int64_t i = j0 + nbw + (int64_t)blockIdx.x * 8 + wp; i < P; i += (int64_t)gridDim.x * 8) {
        const double* row = A + i * lda + j0;
        double s0 = 0.0, s1 = 0.0;
#pragma unroll 4
        for (int t = lane * 2; t < nbw; t += 64) {        // nbw is even (multiple of 64) on this path
            const double2 a = *reinterpret_cast<const double2*>(row + t);
            s0 = fma(a.x, y[t], s0);
            s1 = fma(a.y, y[t + 1], s1);
        }
        double sd = s0 + s1;
        for (int o = 16; o > 0; o >>= 1) sd += __shfl_xor_sync(0xffffffffu, sd, o);
        if (lane == 0) rhs[i] -= sd;
    }
}

// backward panel: rhs[c] -= sum_t A[j0+t][c] * x[t], c < j0.  Thread per column (coalesced rows); blockIdx.y splits the
// nbw rows into groups of 64 whose partial sums are added atomically.
__global__ void __launch_bounds__(256)
trsv_bwd_panel_kernel(const double* __restrict__ A, int64_t lda, int64_t j0, int nbw, double* __restrict__ rhs,
                      const int* __restrict__ stop, const double* __restrict__ ysrc = nullptr) {
    __shared__ double x[CH_NB];
    if (*stop != 0) return;
    const int t0 = blockIdx.y * CH_NB;
    if (t0 >= nbw) return;
    const int nt = min(CH_NB, nbw - t0);
    if (threadIdx.x < CH_NB) {
        const double v = (threadIdx.x < nt) ? (ysrc ? ysrc[t0 + threadIdx.x] : rhs[j0 + t0 + threadIdx.x]) : 0.0;
        x[threadIdx.x] = v;
        if (ysrc && blockIdx.x == 0 && threadIdx.x < nt) rhs[j0 + t0 + threadIdx.x] = v;     // the window's solution, stored once
    }
    __syncthreads();
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= j0) return;
    const double* col = A + (j0 + t0) * lda + c;
    double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
    int t = 0;
    for (; t + 3 < nt; t += 4) {
        s0 = fma(col[(int64_t)t * lda], x[t], s0);
        s1 = fma(col[(int64_t)(t + 1) * lda], x[t + 1], s1);
        s2 = fma(col[(int64_t)(t + 2) * lda], x[t + 2], s2);
        s3 = fma(col[(int64_t)(t + 3) * lda], x[t + 3], s3);
    }
    for (; t < nt; ++t) s0 = fma(col[(int64_t)t * lda], x[t], s0);
    atomicAdd(rhs + c, -((s0 + s1) + (s2 + s3)));
}

// ---- inverses of the 512 x 512 diagonal blocks (P > 8192).  The block solve of a substitution step was one CTA walking eight
//      64-column steps of the triangle (trsv_block_kernel: 60 us, 655 x 2 launches per application on the critical path of every
//      conjugate-gradient iteration of the refinement).  With X = L_BB^-1 stored, it is a 512 x 512 triangular matrix-vector product
//      spread over 64 CTAs.  X is built once per factorisation from the inverted 64 x 64 diagonal blocks: for sub-block column j,
//      X_jj = Linv_j and X_ij = -Linv_i * sum_{k=j..i-1} L_ik X_kj (i > j).  One CTA per (block, j); 64 x 64 x 64 products through
//      two padded shared tiles, 4 x 4 outputs per thread.
constexpr int TB_NS = TB_W / CH_NB;     // 8 sub-blocks of 64

__device__ __forceinline__ void mm64_acc(double (*a)[CH_NB + 1], double (*b)[CH_NB + 1], int ty, int tx, double acc[4][4]) {
    // acc[i][q] += sum_t a[ty*4+i][t] * b[t][tx + 16 q]
    for (int t = 0; t < CH_NB; ++t) {
        double x[4], y[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) x[i] = a[ty * 4 + i][t];
#pragma unroll
        for (int q = 0; q < 4; ++q) y[q] = b[t][tx + 16 * q];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int q = 0; q < 4; ++q) acc[i][q] = fma(x[i], y[q], acc[i][q]);
    }
}

__global__ void __launch_bounds__(256)
blkinv512_kernel(const double* __restrict__ A, int64_t lda, int64_t P, const double* __restrict__ Linv_all, double* __restrict__ X512,
                 const int* __restrict__ info) {
    extern __shared__ double dsm[];
    double (*ta)[CH_NB + 1] = reinterpret_cast<double (*)[CH_NB + 1]>(dsm);
    double (*tb)[CH_NB + 1] = reinterpret_cast<double (*)[CH_NB + 1]>(dsm + CH_NB * (CH_NB + 1));
    if (*info != 0) return;
    const int64_t B = blockIdx.x / TB_NS;
    const int j = blockIdx.x % TB_NS;
    const int64_t j0 = B * TB_W;
    double* X = X512 + B * (int64_t)TB_W * TB_W;
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    if (j0 + (int64_t)j * CH_NB >= P) return;          // sub-block column past the end of the matrix
    // X_jj = Linv_j
    {
        const double* Lj = Linv_all + ((j0 / CH_NB) + j) * (int64_t)CH_NB * CH_NB;
        for (int idx = tid; idx < CH_NB * CH_NB; idx += 256) {
            const int r = idx >> 6, c = idx & 63;
            X[(int64_t)(j * CH_NB + r) * TB_W + j * CH_NB + c] = (c <= r) ? Lj[idx] : 0.0;
        }
    }
    for (int i = j + 1; i < TB_NS; ++i) {
        if (j0 + (int64_t)i * CH_NB >= P) break;
        __syncthreads();                                 // X_kj of the previous rows (written by this CTA) visible; tiles free
        double acc[4][4];
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int b = 0; b < 4; ++b) acc[a][b] = 0.0;
        for (int k = j; k < i; ++k) {
            for (int idx = tid; idx < CH_NB * CH_NB; idx += 256) {
                const int r = idx >> 6, c = idx & 63;
                const int64_t gr = j0 + (int64_t)i * CH_NB + r, gc = j0 + (int64_t)k * CH_NB + c;
                ta[r][c] = (gr < P && gc < P) ? A[gr * lda + gc] : 0.0;                                   // L_ik
                tb[r][c] = X[(int64_t)(k * CH_NB + r) * TB_W + j * CH_NB + c];                             // X_kj
            }
            __syncthreads();
            mm64_acc(ta, tb, ty, tx, acc);
            __syncthreads();
        }
        // X_ij = -Linv_i * S
        {
            const double* Li = Linv_all + ((j0 / CH_NB) + i) * (int64_t)CH_NB * CH_NB;
            for (int idx = tid; idx < CH_NB * CH_NB; idx += 256) {
                const int r = idx >> 6, c = idx & 63;
                ta[r][c] = (c <= r) ? Li[idx] : 0.0;
            }
#pragma unroll
            for (int a = 0; a < 4; ++a)
#pragma unroll
                for (int b = 0; b < 4; ++b) tb[ty * 4 + a][tx + 16 * b] = acc[a][b];
            __syncthreads();
            double out[4][4];
#pragma unroll
            for (int a = 0; a < 4; ++a)
#pragma unroll
                for (int b = 0; b < 4; ++b) out[a][b] = 0.0;
            mm64_acc(ta, tb, ty, tx, out);
#pragma unroll
            for (int a = 0; a < 4; ++a)
#pragma unroll
                for (int b = 0; b < 4; ++b) X[(int64_t)(i * CH_NB + ty * 4 + a) * TB_W + j * CH_NB + tx + 16 * b] = -out[a][b];
        }
    }
}

// y = X x (forward, lower triangle: row r uses columns 0..r) or y = X^T x (backward: column c uses rows c..nbw-1) for the window
// [j0, j0 + nbw).  Forward: one warp per row, the row read coalesced.  Backward: 64 columns per CTA, 4 row phases per column
// (coalesced 512-byte row segments), reduced through shared memory.  y goes to ytmp; the panel kernel that follows (or the copy
// kernel for the last window) stores it into the right-hand side.
__global__ void __launch_bounds__(256)
trsv_gemv512_kernel(const double* __restrict__ X, int nbw, const double* __restrict__ x, double* __restrict__ ytmp, int transpose,
                    const int* __restrict__ stop) {
    __shared__ double sx[TB_W];
    __shared__ double red[4][CH_NB];
    if (*stop != 0) return;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int i = tid; i < TB_W; i += 256) sx[i] = (i < nbw) ? x[i] : 0.0;
    __syncthreads();
    if (!transpose) {
        const int r = blockIdx.x * 8 + warp;
        if (r < nbw) {
            const double* row = X + (int64_t)r * TB_W;
            double s0 = 0.0, s1 = 0.0;
            for (int c = lane * 2; c <= r; c += 64) {
                const double2 v = *reinterpret_cast<const double2*>(row + c);
                s0 = fma(v.x, sx[c], s0);
                if (c + 1 <= r) s1 = fma(v.y, sx[c + 1], s1);
            }
            double sd = s0 + s1;
            for (int o = 16; o > 0; o >>= 1) sd += __shfl_xor_sync(0xffffffffu, sd, o);
            if (lane == 0) ytmp[r] = sd;
        }
    } else {
        const int c = blockIdx.x * CH_NB + (tid & 63), ph = tid >> 6;
        double s = 0.0;
        if (c < nbw)
            for (int r = c + ph; r < nbw; r += 4) s = fma(X[(int64_t)r * TB_W + c], sx[r], s);
        red[ph][tid & 63] = s;
        __syncthreads();
        if (tid < CH_NB && c < nbw) ytmp[c] = (red[0][tid] + red[1][tid]) + (red[2][tid] + red[3][tid]);
    }
}

__global__ void trsv_copy_kernel(const double* __restrict__ ytmp, double* __restrict__ dst, int n, const int* __restrict__ stop) {
    if (*stop != 0) return;
    for (int i = threadIdx.x; i < n; i += blockDim.x) dst[i] = ytmp[i];
}

static int cholesky_configure() {
    constexpr size_t kBlkSmem = 2 * CH_NB * (CH_NB + 1) * sizeof(double);   // trsm needs two blocks, potrf one
    constexpr size_t kDmmaSmem = (size_t)2 * DS_STAGES * DS_BT * DS_LD * sizeof(double);
    TN_SMEM(syrk_update_dmma_kernel, kDmmaSmem);
    TN_SMEM(potrf_diag_kernel, kBlkSmem);
    TN_SMEM(trsm_panel_kernel, kBlkSmem);
    return TN_OK;
}

// Blocked right-looking factorisation of the lower triangle, in place.  X != nullptr: the outer trailing updates with at least
// tc_min_n rows run on the tensor cores in 3xTF32 (syrk_tc.cu); X must hold syrk_tc_work_floats(P, NBO) floats.
int cholesky_factorize(double* A, int64_t lda, int64_t P, double* work, int* info, cudaStream_t st, float* X, int64_t NBO) {
    constexpr size_t kBlkSmem = 2 * CH_NB * (CH_NB + 1) * sizeof(double);
    constexpr size_t kDmmaSmem = (size_t)2 * DS_STAGES * DS_BT * DS_LD * sizeof(double);
    int rc = cholesky_configure();
    if (rc != TN_OK) return rc;
    int coop_ok = 0;
    {
        int dev = 0;
        TN_CUDA(cudaGetDevice(&dev));
        TN_CUDA(cudaDeviceGetAttribute(&coop_ok, cudaDevAttrCooperativeLaunch, dev));
        if (coop_ok) TN_SMEM(cholesky_fused_kernel, kBlkSmem);
    }
    if (coop_ok && P <= CF_MAXP && P > CH_NB && !X && !getenv("TN_CHOL_NO_FUSED")) {
        int Pi = (int)P;
        void* args[] = {(void*)&A, (void*)&lda, (void*)&Pi, (void*)&work, (void*)&info};
        int grid = sm_count();
        const int need = (int)(ceil_div64(P, CH_NB) * (ceil_div64(P, CH_NB) + 1) / 2);
        if (grid > need) grid = need;
        TN_CUDA(cudaLaunchCooperativeKernel((void*)cholesky_fused_kernel, dim3(grid), dim3(256), args, kBlkSmem, st));
        count_launch();
        return TN_OK;
    }
    const int64_t tc_min_n = 1024;
    // tensor-core path, two levels: inside an outer panel the 64-wide steps update only their own 256-wide sub-panel (DMMA); each
    // finished sub-panel then updates the rest of the outer panel on the tensor cores (K = 256), and the outer panel the rest
    // of the matrix (K = NBO)
    const int64_t NBI = (X && NBO % 256 == 0 && NBO > 256 && !getenv("TN_CHOL_NO_MID")) ? 256 : NBO;
    for (int64_t J = 0; J < P; J += NBO) {
        const int64_t Jend = (J + NBO < P) ? J + NBO : P;
        for (int64_t I = J; I < Jend; I += NBI) {
            const int64_t Iend = (I + NBI < Jend) ? I + NBI : Jend;
            for (int64_t j = I; j < Iend; j += CH_NB) {
                const int nb = (int)((Iend - j < CH_NB) ? Iend - j : CH_NB);
                double* Linv = work + (j / CH_NB) * CH_NB * CH_NB;
                potrf_diag_kernel<<<1, 256, kBlkSmem, st>>>(A, lda, j, nb, Linv, info);
                count_launch();
                const int64_t below = P - (j + nb);
                if (below > 0) {
                    trsm_panel_kernel<<<(unsigned)ceil_div64(below, CH_NB), 256, kBlkSmem, st>>>(A, lda, j, nb, Linv, P, info);
                    count_launch();
                    if (j + nb < Iend) {
                        const int64_t c0 = j + nb;
                        if (nb % DS_KC == 0 && lda % 2 == 0 && Iend - c0 >= DS_BT && P - c0 > 2048 && !getenv("TN_CHOL_INNER_FMA")) {
                            dim3 grid((unsigned)ceil_div64(Iend - c0, DS_BT), (unsigned)ceil_div64(P - c0, DS_BT));
                            syrk_update_dmma_kernel<<<grid, DS_THREADS, kDmmaSmem, st>>>(A, lda, P, c0, Iend, j, nb, info);
                        } else {
                            dim3 grid((unsigned)ceil_div64(Iend - c0, 64), (unsigned)ceil_div64(P - c0, 64));
                            syrk_update_kernel<64><<<grid, 256, 0, st>>>(A, lda, P, c0, Iend, j, nb, info);
                        }
                        count_launch();
                    }
                }
            }
            if (Iend < Jend) {
                // sub-panel [I, Iend) -> columns [Iend, Jend) of the outer panel, rows [Iend, P)
                const int64_t width = Jend - Iend;
                const int kbi = (int)(Iend - I);
                if (X && P - Iend >= tc_min_n && (width % 256 == 0 || Jend == P)) {
                    rc = syrk_tc_update(A, lda, P, Iend, I, kbi, X, info, st, (Jend == P) ? 0 : width);
                    if (rc != TN_OK) return rc;
                } else if (kbi % DS_KC == 0 && lda % 2 == 0 && P - Iend > 512) {
                    dim3 grid((unsigned)ceil_div64(width, DS_BT), (unsigned)ceil_div64(P - Iend, DS_BT));
                    syrk_update_dmma_kernel<<<grid, DS_THREADS, kDmmaSmem, st>>>(A, lda, P, Iend, Jend, I, kbi, info);
                    count_launch();
                } else {
                    dim3 grid((unsigned)ceil_div64(width, 64), (unsigned)ceil_div64(P - Iend, 64));
                    syrk_update_kernel<64><<<grid, 256, 0, st>>>(A, lda, P, Iend, Jend, I, kbi, info);
                    count_launch();
                }
            }
        }
        if (Jend < P) {
            const int64_t n = P - Jend;
            const int kbo = (int)(Jend - J);
            if (X && n >= tc_min_n) {
                rc = syrk_tc_update(A, lda, P, Jend, J, kbo, X, info, st);
                if (rc != TN_OK) return rc;
            } else if (n > 512 && kbo % DS_KC == 0 && lda % 2 == 0) {
                dim3 grid((unsigned)ceil_div64(n, DS_BT), (unsigned)ceil_div64(n, DS_BT));
                syrk_update_dmma_kernel<<<grid, DS_THREADS, kDmmaSmem, st>>>(A, lda, P, Jend, P, J, kbo, info);
            } else if (n > 512) {
                dim3 grid((unsigned)ceil_div64(n, 128), (unsigned)ceil_div64(n, 128));
                syrk_update_kernel<128><<<grid, 256, 0, st>>>(A, lda, P, Jend, P, J, kbo, info);
            } else {
                dim3 grid((unsigned)ceil_div64(n, 64), (unsigned)ceil_div64(n, 64));
                syrk_update_kernel<64><<<grid, 256, 0, st>>>(A, lda, P, Jend, P, J, (int)(Jend - J), info);
            }
        }
        TN_LAUNCH_CHECK();
    }
    if (chol_x512_elems(P) > 0) {      // inverses of the 512 x 512 diagonal blocks for the substitutions
        TN_SMEM(blkinv512_kernel, kBlkSmem);
        blkinv512_kernel<<<(unsigned)(ceil_div64(P, TB_W) * TB_NS), 256, kBlkSmem, st>>>(A, lda, P, work, work + chol_linv_elems(P), info);
        TN_LAUNCH_CHECK();
    }
    return TN_OK;
}

// rhs <- L^{-T} L^{-1} rhs with the factor in the lower triangle of A and the inverted diagonal blocks in work.
// Every kernel returns at once when *stop != 0.
int cholesky_substitute(const double* A, int64_t lda, int64_t P, double* rhs, const double* work, const int* stop,
                        cudaStream_t st) {
    if (P <= TS_MAXP) {
        const size_t smem = ((size_t)ceil_div64(P, 64) * 64 + 64) * sizeof(double);
        TN_SMEM(trsv_small_kernel, smem);
        trsv_small_kernel<<<1, 1024, smem, st>>>(A, lda, (int)P, rhs, work, stop);
        TN_LAUNCH_CHECK();
        return TN_OK;
    }
    const int sms = sm_count();
    if (lda % 2 != 0) {      // unaligned rows: the 64-wide kernels
        for (int64_t j = 0; j < P; j += CH_NB) {
            const int nb = (int)((P - j < CH_NB) ? P - j : CH_NB);
            const double* Linv = work + (j / CH_NB) * CH_NB * CH_NB;
            trsv_diag_kernel<<<1, CH_NB, 0, st>>>(rhs, j, nb, Linv, 0, stop);
            count_launch();
            const int64_t below = P - (j + nb);
            if (below > 0) {
                int64_t blocks = ceil_div64(below, 8);
                if (blocks > 4LL * sms) blocks = 4LL * sms;
                trsv_fwd_update_kernel<<<(unsigned)blocks, 256, 0, st>>>(A, lda, P, j, nb, rhs, stop);
                count_launch();
            }
        }
        for (int64_t j = ((P - 1) / CH_NB) * CH_NB; j >= 0; j -= CH_NB) {
            const int nb = (int)((P - j < CH_NB) ? P - j : CH_NB);
            const double* Linv = work + (j / CH_NB) * CH_NB * CH_NB;
            trsv_diag_kernel<<<1, CH_NB, 0, st>>>(rhs, j, nb, Linv, 1, stop);
            count_launch();
            if (j > 0) {
                int64_t blocks = ceil_div64(j, 256);
                if (blocks > 4LL * sms) blocks = 4LL * sms;
                trsv_bwd_update_kernel<<<(unsigned)blocks, 256, 0, st>>>(A, lda, j, nb, rhs, stop);
                count_launch();
            }
        }
        TN_LAUNCH_CHECK();
        return TN_OK;
    }
    // block solves as products with the inverted 512 x 512 diagonal blocks (TN_TRSV_NO_BLKINV=1: the serial block kernel)
    const bool blkinv = !getenv("TN_TRSV_NO_BLKINV");
    const double* X512 = work + chol_linv_elems(P);
    AsyncScratch yscratch;
    if (blkinv) TN_CUDA(yscratch.alloc(TB_W * sizeof(double), st));
    double* ytmp = static_cast<double*>(yscratch.ptr);
    for (int64_t j0 = 0; j0 < P; j0 += TB_W) {
        const int nbw = (int)((P - j0 < TB_W) ? P - j0 : TB_W);
        if (blkinv) trsv_gemv512_kernel<<<(unsigned)ceil_div64(nbw, 8), 256, 0, st>>>(X512 + (j0 / TB_W) * (int64_t)TB_W * TB_W, nbw, rhs + j0, ytmp, 0, stop);
        else trsv_block_kernel<<<1, 1024, 0, st>>>(A, lda, j0, nbw, rhs, work, 0, stop);
        count_launch();
        const int64_t below = P - (j0 + nbw);
        if (below > 0) {
            int64_t blocks = ceil_div64(below, 8);
            if (blocks > 8LL * sms) blocks = 8LL * sms;
            trsv_fwd_panel_kernel<<<(unsigned)blocks, 256, 0, st>>>(A, lda, P, j0, nbw, rhs, stop, ytmp);
            count_launch();
        } else if (blkinv) {
            trsv_copy_kernel<<<1, 256, 0, st>>>(ytmp, rhs + j0, nbw, stop);
            count_launch();
        }
    }
    for (int64_t j0 = ((P - 1) / TB_W) * TB_W; j0 >= 0; j0 -= TB_W) {
        const int nbw = (int)((P - j0 < TB_W) ? P - j0 : TB_W);
        if (blkinv) trsv_gemv512_kernel<<<(unsigned)ceil_div64(nbw, CH_NB), 256, 0, st>>>(X512 + (j0 / TB_W) * (int64_t)TB_W * TB_W, nbw, rhs + j0, ytmp, 1, stop);
        else trsv_block_kernel<<<1, 1024, 0, st>>>(A, lda, j0, nbw, rhs, work, 1, stop);
        count_launch();
        if (j0 > 0) {
            dim3 grid((unsigned)ceil_div64(j0, 256), (unsigned)ceil_div64(nbw, CH_NB));
            trsv_bwd_panel_kernel<<<grid, 256, 0, st>>>(A, lda, j0, nbw, rhs, stop, ytmp);
            count_launch();
        } else if (blkinv) {
            trsv_copy_kernel<<<1, 256, 0, st>>>(ytmp, rhs + j0, nbw, stop);
            count_launch();
        }
    }
    TN_LAUNCH_CHECK();
    return TN_OK;
}

int64_t cholesky_default_nbo(int64_t P) {
    // outer panel width: wider panels amortise the read-modify-write of the trailing matrix (measured at P = 41 876:
    // 256 -> 1385 ms, 512 -> 1210 ms, 768 -> 1140 ms, 1024 -> 1125 ms)
    int64_t NBO = (P > 16384) ? 768 : ((P > 8192) ? 512 : ((P > 4096) ? 256 : ((P > 1024) ? 128 : CH_NB)));
    if (const char* e = getenv("TN_CHOL_NBO")) {
        const int v = atoi(e);
        if (v >= CH_NB && v % CH_NB == 0) NBO = v;
    }
    return NBO;
}

// ---- fp64 refinement around the tensor-core factor ------------------------------------------------------------------
// The factorisation overwrites the lower triangle only, so the strict upper triangle of A still holds the system; with the
// saved diagonal it gives the fp64 operator of the conjugate-gradient refinement without a second copy of A.
__global__ void diag_save_kernel(const double* __restrict__ A, int64_t lda, int64_t P, double* __restrict__ d) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < P; i += (int64_t)gridDim.x * blockDim.x) d[i] = A[i * lda + i];
}

__global__ void symv_diag_kernel(const double* __restrict__ d, const double* __restrict__ x, double* __restrict__ y, int64_t P,
                                 const int* __restrict__ stop) {
    if (*stop != 0) return;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < P; i += (int64_t)gridDim.x * blockDim.x) y[i] = d[i] * x[i];
}

// y += U x + U^T x, U = strict upper triangle of A.  64 x 256 tile per CTA, thread = column.
__global__ void __launch_bounds__(256)
symv_upper_kernel(const double* __restrict__ A, int64_t lda, int64_t P, const double* __restrict__ x, double* __restrict__ y,
                  const int* __restrict__ stop) {
    __shared__ double xi[64];
    __shared__ double rowpart[64][8];
    if (*stop != 0) return;
    const int64_t i0 = (int64_t)blockIdx.y * 64, j0 = (int64_t)blockIdx.x * 256;
    if (j0 + 255 <= i0) return;                 // no element with col > row in this tile
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid < 64) xi[tid] = (i0 + tid < P) ? x[i0 + tid] : 0.0;
    __syncthreads();
    const int64_t col = j0 + tid;
    const double xj = (col < P) ? x[col] : 0.0;
    const int64_t nrow = (P - i0 < 64) ? P - i0 : 64;
    double colacc = 0.0;
    const double* a = A + i0 * lda + col;
#pragma unroll 8
    for (int r = 0; r < 64; ++r) {
        const double v = (r < nrow && col < P && col > i0 + r) ? a[(int64_t)r * lda] : 0.0;
        colacc = fma(v, xi[r], colacc);
        double rp = v * xj;
        for (int o = 16; o > 0; o >>= 1) rp += __shfl_xor_sync(0xffffffffu, rp, o);
        if (lane == 0) rowpart[r][warp] = rp;
    }
    if (col < P) atomicAdd(y + col, colacc);
    __syncthreads();
    if (tid < nrow) {
        double s = 0.0;
#pragma unroll
        for (int w = 0; w < 8; ++w) s += rowpart[tid][w];
        atomicAdd(y + i0 + tid, s);
    }
}

// out[0] += sum a[i] * b[i]
__global__ void __launch_bounds__(256)
dot_kernel(const double* __restrict__ a, const double* __restrict__ b, int64_t n, double* __restrict__ out, const int* __restrict__ stop) {
    __shared__ double red[8];
    if (*stop != 0) return;
    double s = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) s = fma(a[i], b[i], s);
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int k = 0; k < 8; ++k) t += red[k];
        atomicAdd(out, t);
    }
}

// r = b - q, rnorm2 += r.r, bnorm2 += b.b
__global__ void __launch_bounds__(256)
pcg_residual_kernel(const double* __restrict__ b, const double* __restrict__ q, double* __restrict__ r, int64_t n,
                    double* __restrict__ scal, const int* __restrict__ stop) {
    __shared__ double red[2][8];
    if (*stop != 0) return;
    double s = 0.0, sb = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const double bi = b[i], v = bi - q[i];
        r[i] = v;
        s = fma(v, v, s);
        sb = fma(bi, bi, sb);
    }
    for (int o = 16; o > 0; o >>= 1) {
        s += __shfl_xor_sync(0xffffffffu, s, o);
        sb += __shfl_xor_sync(0xffffffffu, sb, o);
    }
    if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = s; red[1][threadIdx.x >> 5] = sb; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0, tb = 0.0;
        for (int k = 0; k < 8; ++k) { t += red[0][k]; tb += red[1][k]; }
        atomicAdd(scal + 1, t);
        atomicAdd(scal + 0, tb);
    }
}

// scal: [0] |b|^2  [1] |r|^2  [2],[3] r.z (alternating by iteration parity)  [4] p.Ap  [5] iterations done
//       [6] relative residual  [7] accumulator of the next |r|^2
__global__ void pcg_latch_kernel(const int* __restrict__ info, int* __restrict__ stop) { *stop = (*info != 0) ? 1 : 0; }

// Before iteration `it`: record the residual, stop when it is small enough (or NaN), clear the accumulators of the iteration.
__global__ void pcg_check_kernel(double* __restrict__ scal, int* __restrict__ stop, double rtol, int it) {
    if (*stop != 0) return;
    const double bn = scal[0], rn = scal[1];
    const double rel = (bn > 0.0) ? sqrt(rn / bn) : 0.0;
    scal[6] = rel;
    scal[5] = (double)it;
    if (!(rel > rtol)) *stop = 1;
    scal[4] = 0.0;
    scal[7] = 0.0;
    scal[2 + ((it + 1) & 1)] = 0.0;                 // the r.z slot this iteration accumulates into
}

__global__ void pcg_commit_kernel(double* __restrict__ scal, const int* __restrict__ stop) {
    if (*stop == 0) scal[1] = scal[7];
}

__global__ void __launch_bounds__(256)
pcg_direction_kernel(const double* __restrict__ z, double* __restrict__ p, int64_t n, const double* __restrict__ scal, int it,
                     const int* __restrict__ stop) {
    if (*stop != 0) return;
    const double beta = (it == 0) ? 0.0 : scal[2 + ((it + 1) & 1)] / scal[2 + (it & 1)];
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        p[i] = (it == 0) ? z[i] : fma(beta, p[i], z[i]);
}

// x += alpha p, r -= alpha q, |r|^2 (scal[1] must have been zeroed)
__global__ void __launch_bounds__(256)
pcg_step_kernel(double* __restrict__ x, double* __restrict__ r, const double* __restrict__ p, const double* __restrict__ q, int64_t n,
                double* __restrict__ scal, double* __restrict__ rn_out, int it, const int* __restrict__ stop) {
    __shared__ double red[8];
    if (*stop != 0) return;
    const double alpha = scal[2 + ((it + 1) & 1)] / scal[4];
    double s = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        x[i] = fma(alpha, p[i], x[i]);
        const double v = fma(-alpha, q[i], r[i]);
        r[i] = v;
        s = fma(v, v, s);
    }
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int k = 0; k < 8; ++k) t += red[k];
        atomicAdd(rn_out, t);
    }
}

__global__ void pcg_finish_kernel(const double* __restrict__ x, double* __restrict__ rhs, int64_t n, const int* __restrict__ info,
                                  const double* __restrict__ scal, double* __restrict__ stats) {
    if (blockIdx.x == 0 && threadIdx.x == 0 && stats) {
        stats[0] = scal[6];
        stats[1] = scal[5];
    }
    if (*info != 0) return;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) rhs[i] = x[i];
}

}  // namespace tn

extern "C" int tn_cholesky_solve(double* A, int64_t lda, int64_t P, double* rhs, double* work, int* info, void* stream) {
    using namespace tn;
    TN_CHECK_ARG(A && work && info && P >= 1 && lda >= P, "tn_cholesky_solve: bad arguments");
    cudaStream_t st = as_stream(stream);
    TN_CUDA(cudaMemsetAsync(info, 0, sizeof(int), st));
    int rc = cholesky_factorize(A, lda, P, work, info, st, nullptr, cholesky_default_nbo(P));
    if (rc != TN_OK) return rc;
    if (rhs) rc = cholesky_substitute(A, lda, P, rhs, work, info, st);
    return rc;
}

extern "C" int64_t tn_cholesky_mixed_work_elems(int64_t P) {
    const int64_t Pp = tn::ceil_div64(P, 64) * 64;
    return tn_cholesky_work_elems(P) + 7 * Pp + 16;
}

extern "C" int tn_cholesky_solve_mixed(double* A, int64_t lda, int64_t P, double* rhs, double* work, int* info, double rtol,
                                       int max_iter, double* stats, void* stream) {
    using namespace tn;
    TN_CHECK_ARG(A && rhs && work && info && P >= 1 && lda >= P && max_iter >= 0 && rtol > 0.0, "tn_cholesky_solve_mixed: bad arguments");
    cudaStream_t st = as_stream(stream);
    const int64_t Pp = ceil_div64(P, 64) * 64;
    double* vec = work + tn_cholesky_work_elems(P);
    double *d = vec, *b = vec + Pp, *x = vec + 2 * Pp, *r = vec + 3 * Pp, *z = vec + 4 * Pp, *p = vec + 5 * Pp, *q = vec + 6 * Pp;
    double* scal = vec + 7 * Pp;                                   // 8 doubles + the stop flag
    int* stop = reinterpret_cast<int*>(scal + 8);
    const int sms = sm_count();
    int64_t vb = ceil_div64(P, 256);
    if (vb > 2LL * sms) vb = 2LL * sms;
    const unsigned vblocks = (unsigned)vb;

    TN_CUDA(cudaMemsetAsync(info, 0, sizeof(int), st));
    TN_CUDA(cudaMemsetAsync(scal, 0, 16 * sizeof(double), st));
    diag_save_kernel<<<vblocks, 256, 0, st>>>(A, lda, P, d);
    TN_LAUNCH_CHECK();
    TN_CUDA(cudaMemcpyAsync(b, rhs, (size_t)P * sizeof(double), cudaMemcpyDeviceToDevice, st));

    int64_t NBO = (P > 8192) ? 1024 : cholesky_default_nbo(P);
    if (const char* e = getenv("TN_CHOL_TC_NBO")) {
        const int v = atoi(e);
        if (v >= CH_NB && v % CH_NB == 0) NBO = v;
    }
    AsyncScratch xscratch;
    TN_CUDA(xscratch.alloc((size_t)syrk_tc_work_floats(P, (int)NBO) * sizeof(float), st));
    float* X = static_cast<float*>(xscratch.ptr);
    int rc = cholesky_factorize(A, lda, P, work, info, st, X, NBO);
    if (rc != TN_OK) return rc;

    auto symv = [&](const double* v, double* out) -> int {
        symv_diag_kernel<<<vblocks, 256, 0, st>>>(d, v, out, P, stop);
        count_launch();
        dim3 grid((unsigned)ceil_div64(P, 256), (unsigned)ceil_div64(P, 64));
        symv_upper_kernel<<<grid, 256, 0, st>>>(A, lda, P, v, out, stop);
        TN_LAUNCH_CHECK();
        return TN_OK;
    };
    // x0 = (L L^T)^-1 b with the approximate factor; a failed factorisation (info != 0) switches everything below off
    pcg_latch_kernel<<<1, 1, 0, st>>>(info, stop);
    TN_LAUNCH_CHECK();
    TN_CUDA(cudaMemcpyAsync(x, b, (size_t)P * sizeof(double), cudaMemcpyDeviceToDevice, st));
    rc = cholesky_substitute(A, lda, P, x, work, stop, st);
    if (rc != TN_OK) return rc;
    rc = symv(x, q);
    if (rc != TN_OK) return rc;
    pcg_residual_kernel<<<vblocks, 256, 0, st>>>(b, q, r, P, scal, stop);
    TN_LAUNCH_CHECK();
    for (int it = 0; it < max_iter; ++it) {
        pcg_check_kernel<<<1, 1, 0, st>>>(scal, stop, rtol, it);
        count_launch();
        if (it >= 1) {      // an iteration is ~10 ms of queued work; do not queue hundreds of no-op launches after convergence
            int h_stop = 0;
            TN_CUDA(cudaMemcpyAsync(&h_stop, stop, sizeof(int), cudaMemcpyDeviceToHost, st));
            TN_CUDA(cudaStreamSynchronize(st));
            if (h_stop) break;
        }
        TN_CUDA(cudaMemcpyAsync(z, r, (size_t)P * sizeof(double), cudaMemcpyDeviceToDevice, st));
        rc = cholesky_substitute(A, lda, P, z, work, stop, st);
        if (rc != TN_OK) return rc;
        dot_kernel<<<vblocks, 256, 0, st>>>(r, z, P, scal + 2 + ((it + 1) & 1), stop);
        count_launch();
        pcg_direction_kernel<<<vblocks, 256, 0, st>>>(z, p, P, scal, it, stop);
        count_launch();
        rc = symv(p, q);
        if (rc != TN_OK) return rc;
        dot_kernel<<<vblocks, 256, 0, st>>>(p, q, P, scal + 4, stop);
        count_launch();
        pcg_step_kernel<<<vblocks, 256, 0, st>>>(x, r, p, q, P, scal, scal + 7, it, stop);
        TN_LAUNCH_CHECK();
        pcg_commit_kernel<<<1, 1, 0, st>>>(scal, stop);
        count_launch();
    }
    pcg_check_kernel<<<1, 1, 0, st>>>(scal, stop, rtol, max_iter);
    count_launch();
    pcg_finish_kernel<<<vblocks, 256, 0, st>>>(x, rhs, P, info, scal, stats);
    TN_LAUNCH_CHECK();
    return TN_OK;
}

// Factorisation only (the preconditioner of tn_cg): lower triangle of A <- L, work <- inverted diagonal blocks.
extern "C" int tn_cholesky_factor(double* A, int64_t lda, int64_t P, int tensor_core, double* work, int* info, void* stream) {
    using namespace tn;
    TN_CHECK_ARG(A && work && info && P >= 1 && lda >= P, "tn_cholesky_factor: bad arguments");
    cudaStream_t st = as_stream(stream);
    TN_CUDA(cudaMemsetAsync(info, 0, sizeof(int), st));
    if (!tensor_core) return cholesky_factorize(A, lda, P, work, info, st, nullptr, cholesky_default_nbo(P));
    int64_t NBO = (P > 8192) ? 1024 : cholesky_default_nbo(P);
    if (const char* e = getenv("TN_CHOL_TC_NBO")) {
        const int v = atoi(e);
        if (v >= CH_NB && v % CH_NB == 0) NBO = v;
    }
    AsyncScratch xscratch;
    TN_CUDA(xscratch.alloc((size_t)syrk_tc_work_floats(P, (int)NBO) * sizeof(float), st));
    float* X = static_cast<float*>(xscratch.ptr);
    syrk_tc_set_passes(tensor_core == 2 ? 1 : 3);        // 2: one TF32 pass per K step (preconditioner-grade factor)
    const int rc = cholesky_factorize(A, lda, P, work, info, st, X, NBO);
    syrk_tc_set_passes(3);
    return rc;
}

// x <- L^-T L^-1 x with the factor and work of tn_cholesky_factor / tn_cholesky_solve (skipped when info[0] != 0).
extern "C" int tn_cholesky_apply(const double* L, int64_t lda, int64_t P, double* x, const double* work, const int* info, void* stream) {
    using namespace tn;
    TN_CHECK_ARG(L && x && work && info && P >= 1 && lda >= P, "tn_cholesky_apply: bad arguments");
    return cholesky_substitute(L, lda, P, x, work, info, as_stream(stream));
}
