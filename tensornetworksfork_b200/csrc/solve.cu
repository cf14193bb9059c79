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

extern "C" int64_t tn_cholesky_work_elems(int64_t P) {
    return tn::ceil_div64(P, tn::CH_NB) * tn::CH_NB * tn::CH_NB;
}

extern "C" int tn_cholesky_solve(double* A, int64_t lda, int64_t P, double* rhs, double* work, int* info, void* stream) {
    using namespace tn;
    TN_CHECK_ARG(A && work && info && P >= 1 && lda >= P, "tn_cholesky_solve: bad arguments");
    cudaStream_t st = as_stream(stream);
    constexpr size_t kBlkSmem = 2 * CH_NB * (CH_NB + 1) * sizeof(double);   // trsm needs two blocks, potrf one
    constexpr size_t kDmmaSmem = (size_t)2 * DS_STAGES * DS_BT * DS_LD * sizeof(double);
    static bool configured = false;
    if (!configured) {
        TN_CUDA(cudaFuncSetAttribute(syrk_update_dmma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kDmmaSmem));
        TN_CUDA(cudaFuncSetAttribute(potrf_diag_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kBlkSmem));
        TN_CUDA(cudaFuncSetAttribute(trsm_panel_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kBlkSmem));
        configured = true;
    }
    TN_CUDA(cudaMemsetAsync(info, 0, sizeof(int), st));
    static int coop_ok = -1;
    if (coop_ok < 0) {
        int dev = 0, v = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&v, cudaDevAttrCooperativeLaunch, dev);
        coop_ok = v;
        if (v) TN_CUDA(cudaFuncSetAttribute(cholesky_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kBlkSmem));
    }
    bool factored = false;
    if (coop_ok && P <= CF_MAXP && P > CH_NB && !getenv("TN_CHOL_NO_FUSED")) {
        int Pi = (int)P;
        void* args[] = {(void*)&A, (void*)&lda, (void*)&Pi, (void*)&work, (void*)&info};
        int grid = sm_count();
        const int need = (int)(ceil_div64(P, CH_NB) * (ceil_div64(P, CH_NB) + 1) / 2);
        if (grid > need) grid = need;
        TN_CUDA(cudaLaunchCooperativeKernel((void*)cholesky_fused_kernel, dim3(grid), dim3(256), args, kBlkSmem, st));
        count_launch();
        factored = true;
    }
    // outer panel width: wider panels amortise the read-modify-write of the trailing matrix (measured at P = 41 876:
    // 256 -> 1385 ms, 512 -> 1210 ms, 768 -> 1140 ms, 1024 -> 1125 ms)
    int64_t NBO = (P > 16384) ? 768 : ((P > 8192) ? 512 : ((P > 4096) ? 256 : ((P > 1024) ? 128 : CH_NB)));
    if (const char* e = getenv("TN_CHOL_NBO")) {
        const int v = atoi(e);
        if (v >= CH_NB && v % CH_NB == 0) NBO = v;
    }
    for (int64_t J = 0; J < P && !factored; J += NBO) {
        const int64_t Jend = (J + NBO < P) ? J + NBO : P;
        for (int64_t j = J; j < Jend; j += CH_NB) {
            const int nb = (int)((Jend - j < CH_NB) ? Jend - j : CH_NB);
            double* Linv = work + (j / CH_NB) * CH_NB * CH_NB;
            potrf_diag_kernel<<<1, 256, kBlkSmem, st>>>(A, lda, j, nb, Linv, info);
            count_launch();
            const int64_t below = P - (j + nb);
            if (below > 0) {
                trsm_panel_kernel<<<(unsigned)ceil_div64(below, CH_NB), 256, kBlkSmem, st>>>(A, lda, j, nb, Linv, P, info);
                count_launch();
                if (j + nb < Jend) {
                    const int64_t c0 = j + nb;
                    if (nb % DS_KC == 0 && lda % 2 == 0 && Jend - c0 >= DS_BT && P - c0 > 2048 && !getenv("TN_CHOL_INNER_FMA")) {
                        dim3 grid((unsigned)ceil_div64(Jend - c0, DS_BT), (unsigned)ceil_div64(P - c0, DS_BT));
                        syrk_update_dmma_kernel<<<grid, DS_THREADS, kDmmaSmem, st>>>(A, lda, P, c0, Jend, j, nb, info);
                    } else {
                        dim3 grid((unsigned)ceil_div64(Jend - c0, 64), (unsigned)ceil_div64(P - c0, 64));
                        syrk_update_kernel<64><<<grid, 256, 0, st>>>(A, lda, P, c0, Jend, j, nb, info);
                    }
                    count_launch();
                }
            }
        }
        if (Jend < P) {
            const int64_t n = P - Jend;
            const int kbo = (int)(Jend - J);
            if (n > 512 && kbo % DS_KC == 0 && lda % 2 == 0) {
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
    if (rhs && P <= TS_MAXP) {
        const size_t smem = ((size_t)ceil_div64(P, 64) * 64 + 64) * sizeof(double);
        static size_t ts_configured = 0;
        if (smem > ts_configured) {
            TN_CUDA(cudaFuncSetAttribute(trsv_small_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            ts_configured = smem;
        }
        trsv_small_kernel<<<1, 1024, smem, st>>>(A, lda, (int)P, rhs, work, info);
        TN_LAUNCH_CHECK();
    } else if (rhs) {
        const int sms = sm_count();
        for (int64_t j = 0; j < P; j += CH_NB) {
            const int nb = (int)((P - j < CH_NB) ? P - j : CH_NB);
            const double* Linv = work + (j / CH_NB) * CH_NB * CH_NB;
            trsv_diag_kernel<<<1, CH_NB, 0, st>>>(rhs, j, nb, Linv, 0, info);
            count_launch();
            const int64_t below = P - (j + nb);
            if (below > 0) {
                int64_t blocks = ceil_div64(below, 8);
                if (blocks > 4LL * sms) blocks = 4LL * sms;
                trsv_fwd_update_kernel<<<(unsigned)blocks, 256, 0, st>>>(A, lda, P, j, nb, rhs, info);
                count_launch();
            }
        }
        for (int64_t j = ((P - 1) / CH_NB) * CH_NB; j >= 0; j -= CH_NB) {
            const int nb = (int)((P - j < CH_NB) ? P - j : CH_NB);
            const double* Linv = work + (j / CH_NB) * CH_NB * CH_NB;
            trsv_diag_kernel<<<1, CH_NB, 0, st>>>(rhs, j, nb, Linv, 1, info);
            count_launch();
            if (j > 0) {
                int64_t blocks = ceil_div64(j, 256);
                if (blocks > 4LL * sms) blocks = 4LL * sms;
                trsv_bwd_update_kernel<<<(unsigned)blocks, 256, 0, st>>>(A, lda, j, nb, rhs, info);
                count_launch();
            }
        }
        TN_LAUNCH_CHECK();
    }
    return TN_OK;
}
