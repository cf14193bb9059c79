// Matrix-free local operator in ONE launch for small cores: out = J^T diag(w) J v with J[row,(a,p,c)] = fa[a] fb[p] fc[c].
//
// Replaces the two einsums of the reference's matvec closure (tensor/network.py:786-790, 912-918) for the shapes of the long
// chains (configs 3 and 4a: r_l * f <= 128 rows of (a,p), r_r <= 64).  The two-launch path (environment pass with prediction
// epilogue, then the right-hand-side pass) reads the three factors twice and runs its second pass on CUDA cores at a few per
// cent of HBM bandwidth; here a persistent CTA keeps v (as the B operand, [(a,p)][c]) and its slice of the result resident,
// walks over 128-row tiles prefetched with cp.async, and does both contractions on the FP64 tensor pipe (DMMA m8n8k4):
//   pass B   t[row]       = w[row] * sum_c ( sum_(a,p) fa[a] phi[p] v[(a,p), c] ) fc[c]        rows x K=(a,p) x N=c
//   pass C   acc[(a,p),c] += sum_rows ( t fa[a] phi[p] )[row] * fc[row, c]                      M=(a,p) x N=c x K=rows
// The factors are read from HBM once per operator application; algorithmic bytes = 8 * rows * (ma/diva + xraw/divb + mc/divc + 1).
#include "common.cuh"

namespace tn {

constexpr int MF_TR = 128;        // rows per tile
constexpr int MF_THREADS = 256;   // 8 warps x 16 rows

struct MfFactor {
    const double* ptr;
    int64_t ld;
    int m;
    int div;
    int map_kind;
};

__device__ __forceinline__ void mf_cp_async8(void* smem_dst, const void* gmem_src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gmem_src) : "memory");
}

template <int NT>
__global__ void __launch_bounds__(MF_THREADS, 1)
matvec_fused_kernel(MfFactor fa, MfFactor fb, MfFactor fc, const double* __restrict__ w, int64_t rows, const double* __restrict__ v,
                    double* __restrict__ out, int64_t ntiles, const int* __restrict__ stop) {
    if (stop && *stop != 0) return;
    constexpr int LDG = NT * 8 + 8;                  // row stride of v as the B operand: == 8 (mod 16) doubles
    extern __shared__ double sm[];
    const int ma = fa.m, mb = fb.m, mc = fc.m;
    const int K = ma * mb;                           // (a, p) pairs
    const int Kp = (K + 3) & ~3;
    const int sta = (ma + 5) | 1;                    // + zero columns read by the padded tail of K
    const int stc = (NT * 8) | 1;
    const int phs = mb | 1;
    const int xraw = (fb.map_kind == TN_MAP_IDENTITY) ? mb : 1;
    double* s_v = sm;                                            // [Kp][LDG]
    double* s_a = s_v + (size_t)Kp * LDG;                        // [2][MF_TR][sta]
    double* s_c = s_a + 2 * (size_t)MF_TR * sta;                 // [2][MF_TR][stc]
    double* s_x = s_c + 2 * (size_t)MF_TR * stc;                 // [2][MF_TR][xraw]
    double* s_w = s_x + 2 * (size_t)MF_TR * xraw;                // [2][MF_TR]
    double* s_phi = s_w + 2 * MF_TR;                             // [MF_TR][phs]
    double* s_t = s_phi + (size_t)MF_TR * phs;                   // [MF_TR]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    for (int idx = tid; idx < Kp * (NT * 8); idx += MF_THREADS) {
        const int k = idx / (NT * 8), n = idx - k * (NT * 8);
        s_v[k * LDG + n] = (k < K && n < mc) ? v[(int64_t)k * mc + n] : 0.0;
    }
    for (int idx = tid; idx < 2 * MF_TR * sta; idx += MF_THREADS) s_a[idx] = 0.0;      // the pad columns stay zero
    for (int idx = tid; idx < 2 * MF_TR * stc; idx += MF_THREADS) s_c[idx] = 0.0;

    auto prefetch = [&](int64_t tile, int buf) {
        if (tile < ntiles) {
            const int64_t row0 = tile * MF_TR;
            double* da = s_a + (size_t)buf * MF_TR * sta;
            for (int idx = tid; idx < MF_TR * ma; idx += MF_THREADS) {
                const int r = idx / ma, a = idx - r * ma;
                int64_t row = row0 + r;
                if (row >= rows) row = rows - 1;                 // clamped rows get weight zero
                mf_cp_async8(da + r * sta + a, fa.ptr + (fa.div == 1 ? row : row / fa.div) * fa.ld + a);
            }
            double* dc = s_c + (size_t)buf * MF_TR * stc;
            for (int idx = tid; idx < MF_TR * mc; idx += MF_THREADS) {
                const int r = idx / mc, c = idx - r * mc;
                int64_t row = row0 + r;
                if (row >= rows) row = rows - 1;
                mf_cp_async8(dc + r * stc + c, fc.ptr + (fc.div == 1 ? row : row / fc.div) * fc.ld + c);
            }
            double* dx = s_x + (size_t)buf * MF_TR * xraw;
            for (int idx = tid; idx < MF_TR * xraw; idx += MF_THREADS) {
                const int r = idx / xraw, q = idx - r * xraw;
                int64_t row = row0 + r;
                if (row >= rows) row = rows - 1;
                mf_cp_async8(dx + idx, fb.ptr + (fb.div == 1 ? row : row / fb.div) * fb.ld + q);
            }
            if (w) {
                double* dw = s_w + (size_t)buf * MF_TR;
                for (int r = tid; r < MF_TR; r += MF_THREADS) {
                    int64_t row = row0 + r;
                    if (row >= rows) row = rows - 1;
                    mf_cp_async8(dw + r, w + row);
                }
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };

    const int fr = lane >> 2, fk = lane & 3;
    const int da4 = 4 / mb, dp4 = 4 - da4 * mb;           // (a, p) advance of this lane's k by 4
    const int MT = (K + 7) / 8;                           // 8-row tiles of the (a, p) dimension: warp w owns mt = w and w + 8
    // pass C accumulators: [slot of this warp's M tile][n tile][2], live for the whole kernel
    double cacc[2][NT][2];
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < NT; ++j) cacc[i][j][0] = cacc[i][j][1] = 0.0;
    int ca[2], cp[2];
    bool cok[2];
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const int mt = warp + 8 * i;
        const int m_idx = mt * 8 + fr;
        cok[i] = (mt < MT);                               // warp-uniform
        const int mi = (m_idx < K) ? m_idx : 0;
        ca[i] = mi / mb;
        cp[i] = mi - ca[i] * mb;
        if (m_idx >= K) { ca[i] = ma; cp[i] = 0; }        // reads a zero pad column of s_a
    }

    int64_t tile = blockIdx.x;
    int buf = 0;
    __syncthreads();                 // zero fills done before the first async copies land
    prefetch(tile, 0);
    for (; tile < ntiles; tile += gridDim.x, buf ^= 1) {
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncthreads();             // tile data visible; everyone is done with the other buffer, s_phi and s_t
        prefetch(tile + gridDim.x, buf ^ 1);
        const int64_t row0 = tile * MF_TR;
        const double* xin = s_x + (size_t)buf * MF_TR * xraw;
        if (fb.map_kind == TN_MAP_SINCOS) {
            for (int r = tid; r < MF_TR; r += MF_THREADS) {
                double c, sn;
                sincos((0.5 * 3.14159265358979323846) * xin[r], &sn, &c);
                s_phi[r * phs] = c;
                s_phi[r * phs + 1] = sn;
            }
        } else {
            for (int idx = tid; idx < MF_TR * mb; idx += MF_THREADS) {
                const int r = idx / mb, p = idx - r * mb;
                s_phi[r * phs + p] = map_eval(fb.map_kind, xin + r * xraw, p);
            }
        }
        __syncthreads();
        const double* A_ = s_a + (size_t)buf * MF_TR * sta;
        const double* C_ = s_c + (size_t)buf * MF_TR * stc;
        {   // ---- pass B: this warp's 16 rows
            double acc[2][NT][2];
#pragma unroll
            for (int i = 0; i < 2; ++i)
#pragma unroll
                for (int j = 0; j < NT; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
            int ka = fk / mb, kp = fk - ka * mb;
            const double* in0 = A_ + (warp * 16 + fr) * sta;
            const double* in1 = in0 + 8 * sta;
            const double* ph0 = s_phi + (warp * 16 + fr) * phs;
            const double* ph1 = ph0 + 8 * phs;
#pragma unroll 4
            for (int kk = 0; kk < Kp; kk += 4) {
                const double a0 = in0[ka] * ph0[kp];
                const double a1 = in1[ka] * ph1[kp];
                kp += dp4; ka += da4;
                if (kp >= mb) { kp -= mb; ++ka; }
                const double* gp = s_v + (kk + fk) * LDG + fr;
#pragma unroll
                for (int j = 0; j < NT; ++j) {
                    const double b = gp[j * 8];
                    dmma884(acc[0][j][0], acc[0][j][1], a0, b);
                    dmma884(acc[1][j][0], acc[1][j][1], a1, b);
                }
            }
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                const int r = warp * 16 + i * 8 + fr;
                double yd = 0.0;
#pragma unroll
                for (int j = 0; j < NT; ++j)
#pragma unroll
                    for (int e = 0; e < 2; ++e) yd = fma(acc[i][j][e], C_[r * stc + j * 8 + 2 * fk + e], yd);     // pad columns are zero
                yd += __shfl_xor_sync(0xffffffffu, yd, 1);
                yd += __shfl_xor_sync(0xffffffffu, yd, 2);
                if (fk == 0) {
                    const double wr = (row0 + r < rows) ? (w ? s_w[buf * MF_TR + r] : 1.0) : 0.0;
                    s_t[r] = yd * wr;
                }
            }
        }
        __syncthreads();
        // ---- pass C: K = the tile's 128 rows; A[m = (a,p)][k = row] = t fa[a] phi[p], B[k = row][n = c] = fc[row, c]
#pragma unroll 2
        for (int k4 = 0; k4 < MF_TR; k4 += 4) {
            const int r = k4 + fk;
            const double tv = s_t[r];
            double bfr[NT];
            const double* cr = C_ + r * stc + fr;
#pragma unroll
            for (int j = 0; j < NT; ++j) bfr[j] = cr[j * 8];
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                if (cok[i]) {
                    const double a = tv * A_[r * sta + ca[i]] * s_phi[r * phs + cp[i]];
#pragma unroll
                    for (int j = 0; j < NT; ++j) dmma884(cacc[i][j][0], cacc[i][j][1], a, bfr[j]);
                }
            }
        }
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    // ---- add this CTA's partial of out[(a,p), c]
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const int m_idx = (warp + 8 * i) * 8 + fr;
        if (cok[i] && m_idx < K) {
#pragma unroll
            for (int j = 0; j < NT; ++j)
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const int c = j * 8 + 2 * fk + e;
                    if (c < mc) atomicAdd(out + (int64_t)m_idx * mc + c, cacc[i][j][e]);
                }
        }
    }
}

template <int NT>
static int launch_matvec_fused(const MfFactor& a, const MfFactor& b, const MfFactor& c, const double* w, int64_t rows, const double* v,
                               double* out, const int* stop, cudaStream_t st) {
    const int K = a.m * b.m, Kp = (K + 3) & ~3, xraw = (b.map_kind == TN_MAP_IDENTITY) ? b.m : 1;
    const size_t smem = ((size_t)Kp * (NT * 8 + 8) + 2 * (size_t)MF_TR * ((a.m + 5) | 1) + 2 * (size_t)MF_TR * ((NT * 8) | 1) +
                         2 * (size_t)MF_TR * xraw + 2 * MF_TR + (size_t)MF_TR * (b.m | 1) + MF_TR) * sizeof(double);
    if (smem > 220 * 1024) return 1;
    TN_SMEM(matvec_fused_kernel<NT>, smem);
    const int64_t ntiles = ceil_div64(rows, MF_TR);
    int64_t grid = sm_count();
    if (grid > ntiles) grid = ntiles;
    matvec_fused_kernel<NT><<<(unsigned)grid, MF_THREADS, smem, st>>>(a, b, c, w, rows, v, out, ntiles, stop);
    TN_LAUNCH_CHECK();
    return TN_OK;
}

// Returns TN_OK when the fused kernel was launched (out holds J^T diag(w) J v), 1 when the shape is outside its range (the caller
// uses the two-pass path), < 0 on error.  `stop` (device flag, may be null) turns the launch into a no-op when set.
int matvec_fused(const tn_factor* fa, const tn_factor* fb, const tn_factor* fc, const double* w, int64_t rows, const double* v, double* out,
                 const int* stop, cudaStream_t st) {
    const int K = fa->m * fb->m;
    if (getenv("TN_NO_FUSED_MATVEC")) return 1;
    if (K > 128 || K < 4 || fc->m > 64 || rows < 4 * MF_TR) return 1;
    if (fa->map_kind != TN_MAP_IDENTITY || fc->map_kind != TN_MAP_IDENTITY) return 1;
    const MfFactor a{fa->ptr, fa->ld, fa->m, fa->div < 1 ? 1 : fa->div, fa->map_kind};
    const MfFactor b{fb->ptr, fb->ld, fb->m, fb->div < 1 ? 1 : fb->div, fb->map_kind};
    const MfFactor c{fc->ptr, fc->ld, fc->m, fc->div < 1 ? 1 : fc->div, fc->map_kind};
    TN_CUDA(cudaMemsetAsync(out, 0, (size_t)K * fc->m * sizeof(double), st));
    const int nt = (fc->m + 7) / 8;
    switch (nt) {
        case 1: return launch_matvec_fused<1>(a, b, c, w, rows, v, out, stop, st);
        case 2: return launch_matvec_fused<2>(a, b, c, w, rows, v, out, stop, st);
        case 3: return launch_matvec_fused<3>(a, b, c, w, rows, v, out, stop, st);
        case 4: return launch_matvec_fused<4>(a, b, c, w, rows, v, out, stop, st);
        case 5: return launch_matvec_fused<5>(a, b, c, w, rows, v, out, stop, st);
        case 6: return launch_matvec_fused<6>(a, b, c, w, rows, v, out, stop, st);
        default: return launch_matvec_fused<8>(a, b, c, w, rows, v, out, stop, st);
    }
}

}  // namespace tn
