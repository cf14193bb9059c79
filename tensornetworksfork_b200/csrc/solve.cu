// Regularised local solve: blocked Cholesky + triangular solves, fp64.
//
// Replaces torch.linalg.cholesky + torch.cholesky_solve inside TensorNetwork.solve_system
// (reference tensor/network.py:311-320).  Two-level right-looking factorisation of the lower
// triangle, in place: 64-wide diagonal blocks are factored and inverted by one CTA, the panel
// below is a small GEMM with the inverted block, trailing updates are SYRK tiles (64x64 inside an
// outer panel, 128x128 with K = outer panel width for the rest, so the C-tile read-modify-write is
// amortised over a long K).  P ranges from 4 to 41 876 (14 GB) on the named configurations.
#include "common.cuh"

namespace tn {

constexpr int CH_NB = 64;

// ---- diagonal block: factor (lower) and invert ------------------------------------------------
__global__ void __launch_bounds__(256)
potrf_diag_kernel(double* __restrict__ A, int64_t lda, int64_t j, int nb, double* __restrict__ Linv, int* __restrict__ info) {
    extern __shared__ double dsm[];
    double (*a)[CH_NB + 1] = reinterpret_cast<double (*)[CH_NB + 1]>(dsm);
    double (*li)[CH_NB + 1] = reinterpret_cast<double (*)[CH_NB + 1]>(dsm + CH_NB * (CH_NB + 1));
    if (*info != 0) return;
    const int tid = threadIdx.x;
    for (int idx = tid; idx < CH_NB * CH_NB; idx += 256) {
        const int r = idx >> 6, c = idx & 63;
        a[r][c] = (r < nb && c <= r) ? A[(j + r) * lda + j + c] : 0.0;
        li[r][c] = 0.0;
    }
    for (int c = 0; c < nb; ++c) {
        __syncthreads();
        const double d = a[c][c];
        if (!(d > 0.0)) {  // also catches NaN
            if (tid == 0) *info = (int)(j + c + 1);
            return;
        }
        const double sd = sqrt(d);
        __syncthreads();
        for (int r = c + tid; r < nb; r += 256) a[r][c] = (r == c) ? sd : a[r][c] / sd;
        __syncthreads();
        const int rem = nb - c - 1;
        for (int idx = tid; idx < rem * rem; idx += 256) {
            const int rr = c + 1 + idx / rem, cc = c + 1 + idx % rem;
            if (cc <= rr) a[rr][cc] = fma(-a[rr][c], a[cc][c], a[rr][cc]);
        }
    }
    __syncthreads();
    for (int idx = tid; idx < nb * nb; idx += 256) {
        const int r = idx / nb, c = idx % nb;
        if (c <= r) A[(j + r) * lda + j + c] = a[r][c];
    }
    if (tid < nb) {  // column tid of L^{-1} by forward substitution
        const int t = tid;
        li[t][t] = 1.0 / a[t][t];
        for (int r = t + 1; r < nb; ++r) {
            double s = 0.0;
            for (int q = t; q < r; ++q) s = fma(a[r][q], li[q][t], s);
            li[r][t] = -s / a[r][r];
        }
    }
    __syncthreads();
    for (int idx = tid; idx < CH_NB * CH_NB; idx += 256) Linv[idx] = li[idx >> 6][idx & 63];
}

// ---- panel: X = B * Linv^T for the rows below the diagonal block -------------------------------
__global__ void __launch_bounds__(256)
trsm_panel_kernel(double* __restrict__ A, int64_t lda, int64_t j, int nb, const double* __restrict__ Linv, int64_t P,
                  const int* __restrict__ info) {
    extern __shared__ double dsm[];
    double (*b)[CH_NB + 1] = reinterpret_cast<double (*)[CH_NB + 1]>(dsm);
    double (*li)[CH_NB + 1] = reinterpret_cast<double (*)[CH_NB + 1]>(dsm + CH_NB * (CH_NB + 1));
    if (*info != 0) return;
    const int tid = threadIdx.x;
    const int64_t r0 = j + nb + (int64_t)blockIdx.x * CH_NB;
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

// ---- trailing update: C[i][q] -= sum_t A[i][k0+t] * A[q][k0+t] on the lower tiles of
//      columns [c0, c1), rows [c0, P).  BT x BT tile per CTA, 256 threads, (BT/16)^2 per thread.
template <int BT>
__global__ void __launch_bounds__(256)
syrk_update_kernel(double* __restrict__ A, int64_t lda, int64_t P, int64_t c0, int64_t c1, int64_t k0, int kb,
                   const int* __restrict__ info) {
    constexpr int TM = BT / 16;
    constexpr int KC = 16;
    __shared__ double sa[KC][BT + 2];
    __shared__ double sb[KC][BT + 2];
    if (*info != 0) return;
    const int bi = blockIdx.y, bj = blockIdx.x;
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

// ---- large trailing update on the FP64 tensor pipe (DMMA m8n8k4), 128x128 tile, K streamed by cp.async.
//      Same contract as syrk_update_kernel; requires kb % 16 == 0 and lda % 2 == 0, k0 % 2 == 0.
constexpr int DS_BT = 128;   // tile
constexpr int DS_KC = 16;    // k per stage
constexpr int DS_LD = DS_KC + 2;   // padded smem row (doubles): 16-byte aligned rows, conflict-free fragment reads
constexpr int DS_STAGES = 3;

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void dmma884(double& d0, double& d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

__global__ void __launch_bounds__(256, 1)
syrk_update_dmma_kernel(double* __restrict__ A, int64_t lda, int64_t P, int64_t c0, int64_t c1, int64_t k0, int kb,
                        const int* __restrict__ info) {
    extern __shared__ double dsm[];
    if (*info != 0) return;
    const int bi = blockIdx.y, bj = blockIdx.x;
    if (bj > bi) return;
    const int64_t r0 = c0 + (int64_t)bi * DS_BT, q0 = c0 + (int64_t)bj * DS_BT;
    if (r0 >= P || q0 >= c1) return;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wm = warp >> 1, wn = warp & 1;          // 4 x 2 warps: warp tile 32 (rows) x 64 (cols)
    double* sA = dsm;                                  // [stage][128][DS_LD]
    double* sB = dsm + DS_STAGES * DS_BT * DS_LD;

    double acc[4][8][2];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

    // each thread copies 4 + 4 sixteen-byte pieces per stage: piece = (row, half-pair index 0..7)
    auto issue = [&](int stage, int kk) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int piece = tid + u * 256;           // 0..1023
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
        const double* b_s = sB + ((size_t)(it % DS_STAGES) * DS_BT + wn * 64 + fr) * DS_LD + fk;
#pragma unroll
        for (int k4 = 0; k4 < DS_KC; k4 += 4) {
            double af[4], bf[8];
#pragma unroll
            for (int i = 0; i < 4; ++i) af[i] = a_s[(size_t)i * 8 * DS_LD + k4];
#pragma unroll
            for (int j = 0; j < 8; ++j) bf[j] = b_s[(size_t)j * 8 * DS_LD + k4];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) dmma884(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
        }
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int64_t row = r0 + wm * 32 + i * 8 + fr;
        if (row >= P) continue;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int64_t col = q0 + wn * 64 + j * 8 + fk * 2;
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

}  // namespace tn

extern "C" int64_t tn_cholesky_work_elems(int64_t P) {
    return tn::ceil_div64(P, tn::CH_NB) * tn::CH_NB * tn::CH_NB;
}

extern "C" int tn_cholesky_solve(double* A, int64_t lda, int64_t P, double* rhs, double* work, int* info, void* stream) {
    using namespace tn;
    TN_CHECK_ARG(A && work && info && P >= 1 && lda >= P, "tn_cholesky_solve: bad arguments");
    cudaStream_t st = as_stream(stream);
    constexpr size_t kBlkSmem = 2 * CH_NB * (CH_NB + 1) * sizeof(double);
    constexpr size_t kDmmaSmem = (size_t)2 * DS_STAGES * DS_BT * DS_LD * sizeof(double);
    static bool configured = false;
    if (!configured) {
        TN_CUDA(cudaFuncSetAttribute(syrk_update_dmma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kDmmaSmem));
        TN_CUDA(cudaFuncSetAttribute(potrf_diag_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kBlkSmem));
        TN_CUDA(cudaFuncSetAttribute(trsm_panel_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kBlkSmem));
        configured = true;
    }
    TN_CUDA(cudaMemsetAsync(info, 0, sizeof(int), st));
    const int64_t NBO = (P > 4096) ? 256 : ((P > 1024) ? 128 : CH_NB);
    for (int64_t J = 0; J < P; J += NBO) {
        const int64_t Jend = (J + NBO < P) ? J + NBO : P;
        for (int64_t j = J; j < Jend; j += CH_NB) {
            const int nb = (int)((Jend - j < CH_NB) ? Jend - j : CH_NB);
            double* Linv = work + (j / CH_NB) * CH_NB * CH_NB;
            potrf_diag_kernel<<<1, 256, kBlkSmem, st>>>(A, lda, j, nb, Linv, info);
            const int64_t below = P - (j + nb);
            if (below > 0) {
                trsm_panel_kernel<<<(unsigned)ceil_div64(below, CH_NB), 256, kBlkSmem, st>>>(A, lda, j, nb, Linv, P, info);
                if (j + nb < Jend) {
                    const int64_t c0 = j + nb;
                    dim3 grid((unsigned)ceil_div64(Jend - c0, 64), (unsigned)ceil_div64(P - c0, 64));
                    syrk_update_kernel<64><<<grid, 256, 0, st>>>(A, lda, P, c0, Jend, j, nb, info);
                }
            }
        }
        if (Jend < P) {
            const int64_t n = P - Jend;
            const int kbo = (int)(Jend - J);
            if (n > 512 && kbo % DS_KC == 0 && lda % 2 == 0) {
                dim3 grid((unsigned)ceil_div64(n, DS_BT), (unsigned)ceil_div64(n, DS_BT));
                syrk_update_dmma_kernel<<<grid, 256, kDmmaSmem, st>>>(A, lda, P, Jend, P, J, kbo, info);
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
    if (rhs) {
        const int sms = sm_count();
        for (int64_t j = 0; j < P; j += CH_NB) {
            const int nb = (int)((P - j < CH_NB) ? P - j : CH_NB);
            const double* Linv = work + (j / CH_NB) * CH_NB * CH_NB;
            trsv_diag_kernel<<<1, CH_NB, 0, st>>>(rhs, j, nb, Linv, 0, info);
            const int64_t below = P - (j + nb);
            if (below > 0) {
                int64_t blocks = ceil_div64(below, 8);
                if (blocks > 4LL * sms) blocks = 4LL * sms;
                trsv_fwd_update_kernel<<<(unsigned)blocks, 256, 0, st>>>(A, lda, P, j, nb, rhs, info);
            }
        }
        for (int64_t j = ((P - 1) / CH_NB) * CH_NB; j >= 0; j -= CH_NB) {
            const int nb = (int)((P - j < CH_NB) ? P - j : CH_NB);
            const double* Linv = work + (j / CH_NB) * CH_NB * CH_NB;
            trsv_diag_kernel<<<1, CH_NB, 0, st>>>(rhs, j, nb, Linv, 1, info);
            if (j > 0) {
                int64_t blocks = ceil_div64(j, 256);
                if (blocks > 4LL * sms) blocks = 4LL * sms;
                trsv_bwd_update_kernel<<<(unsigned)blocks, 256, 0, st>>>(A, lda, j, nb, rhs, info);
            }
        }
        TN_LAUNCH_CHECK();
    }
    return TN_OK;
}
