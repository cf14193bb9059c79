// Fused feature-map + environment step, prediction epilogue and class-leg row folding.
//
// Replaces the chained pairwise einsums of TensorNetwork.compute_stacks /
// left_update_stacks / right_update_stacks / forward (reference tensor/network.py:55-71,
// 115-137, 152-172): out[row, b] = sum_{a,p} env[row, a] * phi(x[row / cdiv], p) * core[a, p, b].
//
// Layout: environments are sample-major [rows, r] fp64, so a tile of consecutive rows is one
// contiguous byte range; the site input is read once per sample and mapped in registers.
// One CTA owns 64..128 rows; the core is streamed through shared memory in feature slabs that
// every CTA re-reads from L2 (the core is at most a few hundred KB).
#include <stdlib.h>
#include "common.cuh"

namespace tn {

constexpr int ENV_THREADS = 256;
constexpr int ENV_RT = 4;        // rows per thread

// TXN column lanes x (256/TXN) row groups; each thread owns ENV_RT rows x CN columns (tx + TXN*j).
// The contraction is ordered  out = sum_p phi_p * (sum_a env_a * core[a,p,:])  so the inner loop is pure FMA.
template <int TXN, int CN>
__global__ void __launch_bounds__(ENV_THREADS, (CN <= 3) ? 3 : 2)
env_kernel(const double* __restrict__ env_in, int64_t env_ld, int env_div, const double* __restrict__ x, int64_t x_ld,
           int map_kind, int f, int cdiv, const double* __restrict__ core, double* __restrict__ out,
           int64_t out_ld, const double* __restrict__ dot, int64_t dot_ld, int dot_div,
           double* __restrict__ yhat, const double* __restrict__ yscale, int64_t rows, int r_in, int r_out, int p_chunk) {
    constexpr int TR = (ENV_THREADS / TXN) * ENV_RT;   // rows per CTA
    constexpr int TC = TXN * CN;                        // columns per pass
    extern __shared__ double sm[];
    const int in_st = r_in | 1;
    const int phi_st = f | 1;
    double* s_in = sm;
    double* s_phi = s_in + TR * in_st;
    double* s_g = s_phi + TR * phi_st;                  // [p_chunk][r_in][TC]

    const int tid = threadIdx.x;
    const int tx = tid % TXN;
    const int ty = tid / TXN;
    const int64_t row0 = (int64_t)blockIdx.x * TR;

    for (int idx = tid; idx < TR * r_in; idx += ENV_THREADS) {
        const int r = idx / r_in, a = idx - r * r_in;
        const int64_t row = row0 + r;
        double v = 0.0;
        if (row < rows) v = env_in ? env_in[(env_div == 1 ? row : row / env_div) * env_ld + a] : 1.0;
        s_in[r * in_st + a] = v;
    }
    if (map_kind == TN_MAP_SINCOS) {
        for (int r = tid; r < TR; r += ENV_THREADS) {
            const int64_t row = row0 + r;
            double c = 0.0, sn = 0.0;
            if (row < rows) sincos((0.5 * 3.14159265358979323846) * x[(cdiv == 1 ? row : row / cdiv) * x_ld], &sn, &c);
            s_phi[r * phi_st] = c;
            s_phi[r * phi_st + 1] = sn;
        }
    } else {
        for (int idx = tid; idx < TR * f; idx += ENV_THREADS) {
            const int r = idx / f, p = idx - r * f;
            const int64_t row = row0 + r;
            double v = 0.0;
            if (row < rows) v = map_eval(map_kind, x + (cdiv == 1 ? row : row / cdiv) * x_ld, p);
            s_phi[r * phi_st + p] = v;
        }
    }

    double ydot[ENV_RT];
#pragma unroll
    for (int i = 0; i < ENV_RT; ++i) ydot[i] = 0.0;
    for (int c0 = 0; c0 < r_out; c0 += TC) {
        double acc[ENV_RT][CN];
#pragma unroll
        for (int i = 0; i < ENV_RT; ++i)
#pragma unroll
            for (int j = 0; j < CN; ++j) acc[i][j] = 0.0;

        for (int p0 = 0; p0 < f; p0 += p_chunk) {
            const int pc = min(p_chunk, f - p0);
            __syncthreads();
            // stage core[:, p0:p0+pc, c0:c0+TC] as [p][a][col]
            for (int idx = tid; idx < pc * r_in * TC; idx += ENV_THREADS) {
                const int col = idx % TC;
                const int pa = idx / TC;
                const int pp = pa / r_in, a = pa - pp * r_in;
                const int b = c0 + col;
                s_g[idx] = (b < r_out) ? core[((int64_t)a * f + p0 + pp) * r_out + b] : 0.0;
            }
            __syncthreads();
            for (int pp = 0; pp < pc; ++pp) {
                double t[ENV_RT][CN];
#pragma unroll
                for (int i = 0; i < ENV_RT; ++i)
#pragma unroll
                    for (int j = 0; j < CN; ++j) t[i][j] = 0.0;
                const double* gp = s_g + (size_t)pp * r_in * TC + tx;
                const double* ip = s_in + (ty * ENV_RT) * in_st;
#pragma unroll 2
                for (int a = 0; a < r_in; ++a) {
                    double e[ENV_RT], g[CN];
#pragma unroll
                    for (int i = 0; i < ENV_RT; ++i) e[i] = ip[i * in_st + a];
#pragma unroll
                    for (int j = 0; j < CN; ++j) g[j] = gp[a * TC + TXN * j];
#pragma unroll
                    for (int i = 0; i < ENV_RT; ++i)
#pragma unroll
                        for (int j = 0; j < CN; ++j) t[i][j] = fma(e[i], g[j], t[i][j]);
                }
#pragma unroll
                for (int i = 0; i < ENV_RT; ++i) {
                    const double ph = s_phi[(ty * ENV_RT + i) * phi_st + p0 + pp];
#pragma unroll
                    for (int j = 0; j < CN; ++j) acc[i][j] = fma(ph, t[i][j], acc[i][j]);
                }
            }
        }
#pragma unroll
        for (int i = 0; i < ENV_RT; ++i) {
            const int64_t row = row0 + ty * ENV_RT + i;
            if (row >= rows) continue;
#pragma unroll
            for (int j = 0; j < CN; ++j) {
                const int b = c0 + tx + TXN * j;
                if (b >= r_out) continue;
                if (dot)
                    ydot[i] = fma(acc[i][j], dot[(dot_div == 1 ? row : row / dot_div) * dot_ld + b], ydot[i]);
                else
                    out[row * out_ld + b] = acc[i][j];
            }
        }
    }
    if (dot) {
#pragma unroll
        for (int i = 0; i < ENV_RT; ++i) {
            double v = ydot[i];
#pragma unroll
            for (int o = TXN / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            const int64_t row = row0 + ty * ENV_RT + i;
            if (tx == 0 && row < rows) yhat[row] = yscale ? v * yscale[row] : v;
        }
    }
}

// ---- FP64 tensor-core variant (DMMA m8n8k4).  out = Z * G with Z[row, k=(a,p)] = env[row,a] * phi[row,p] formed in
//      registers as the A fragment, G = core viewed as (r_in*f) x r_out read as the B fragment from shared memory.
//      One CTA = 128 rows, 8 warps x 16 rows; NT 8-column tiles cover r_out.  Same contract as env_kernel.
constexpr int ED_TR = 128;
constexpr int ED_KC = 64;    // k per staged core slab

template <int NT>
__global__ void __launch_bounds__(ENV_THREADS, 2)
env_dmma_kernel(const double* __restrict__ env_in, int64_t env_ld, int env_div, const double* __restrict__ x, int64_t x_ld,
                int map_kind, int f, int cdiv, const double* __restrict__ core, double* __restrict__ out,
                int64_t out_ld, const double* __restrict__ dot, int64_t dot_ld, int dot_div,
                double* __restrict__ yhat, const double* __restrict__ yscale, int64_t rows, int r_in, int r_out) {
    constexpr int LDG = NT * 8 + 4;                  // core slab row stride == 4 (mod 8) doubles: the 16 lanes of a half warp (fr 0..3 x fk 0..3) read fk * LDG + fr from 16 distinct 8-byte banks (NT * 8 + 8 gave 2- and 4-way conflicts: profiles/r2_ncu_env_warp.txt)
    extern __shared__ double sm[];
    const int in_st = (r_in + 5) | 1;                // + zero columns read by the padded tail of K
    const int phi_st = f | 1;
    double* s_in = sm;                               // [ED_TR][in_st]
    double* s_phi = s_in + ED_TR * in_st;            // [ED_TR][phi_st]
    double* s_g = s_phi + ED_TR * phi_st;            // [ED_KC][LDG]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t row0 = (int64_t)blockIdx.x * ED_TR;
    const int K = r_in * f;

    for (int idx = tid; idx < ED_TR * in_st; idx += ENV_THREADS) {
        const int r = idx / in_st, a = idx - r * in_st;
        const int64_t row = row0 + r;
        double v = 0.0;
        if (row < rows && a < r_in) v = env_in ? env_in[(env_div == 1 ? row : row / env_div) * env_ld + a] : 1.0;
        s_in[idx] = v;
    }
    if (map_kind == TN_MAP_SINCOS) {
        for (int r = tid; r < ED_TR; r += ENV_THREADS) {
            const int64_t row = row0 + r;
            double c = 0.0, sn = 0.0;
            if (row < rows) sincos((0.5 * 3.14159265358979323846) * x[(cdiv == 1 ? row : row / cdiv) * x_ld], &sn, &c);
            s_phi[r * phi_st] = c;
            s_phi[r * phi_st + 1] = sn;
        }
    } else {
        for (int idx = tid; idx < ED_TR * f; idx += ENV_THREADS) {
            const int r = idx / f, p = idx - r * f;
            const int64_t row = row0 + r;
            s_phi[r * phi_st + p] = (row < rows) ? map_eval(map_kind, x + (cdiv == 1 ? row : row / cdiv) * x_ld, p) : 0.0;
        }
    }

    double acc[2][NT][2];
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < NT; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

    const int fr = lane >> 2, fk = lane & 3;
    const int da = 4 / f, dp = 4 - da * f;            // (a, p) advance of this lane's k by 4
    int ka = fk / f, kp = fk - ka * f;                // k = fk at the start
    const double* in0 = s_in + (warp * 16 + fr) * in_st;
    const double* in1 = in0 + 8 * in_st;
    const double* ph0 = s_phi + (warp * 16 + fr) * phi_st;
    const double* ph1 = ph0 + 8 * phi_st;

    for (int k0 = 0; k0 < K; k0 += ED_KC) {
        __syncthreads();
        for (int idx = tid; idx < ED_KC * (NT * 8); idx += ENV_THREADS) {
            const int kk = idx / (NT * 8), n = idx - kk * (NT * 8);
            const int k = k0 + kk;
            s_g[kk * LDG + n] = (k < K && n < r_out) ? core[(int64_t)k * r_out + n] : 0.0;
        }
        __syncthreads();
        const int kend = min(ED_KC, ((K - k0) + 3) & ~3);   // the padded tail reads the zero columns of s_in
#pragma unroll 4
        for (int kk = 0; kk < kend; kk += 4) {
            // A fragments: rows fr and fr+8 of this warp's 16, column k = k0 + kk + fk  ->  env[a] * phi[p]
            const double a0 = in0[ka] * ph0[kp];
            const double a1 = in1[ka] * ph1[kp];
            kp += dp; ka += da;
            if (kp >= f) { kp -= f; ++ka; }
            const double* gp = s_g + (kk + fk) * LDG + fr;
#pragma unroll
            for (int j = 0; j < NT; ++j) {
                const double b = gp[j * 8];
                dmma884(acc[0][j][0], acc[0][j][1], a0, b);
                dmma884(acc[1][j][0], acc[1][j][1], a1, b);
            }
        }
    }
    // epilogue: this lane owns rows (warp*16 + fr) and (+8), columns j*8 + 2*fk + {0,1}
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const int64_t row = row0 + warp * 16 + i * 8 + fr;
        double yd = 0.0;
        if (row < rows) {
#pragma unroll
            for (int j = 0; j < NT; ++j) {
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const int b = j * 8 + 2 * fk + e;
                    if (b < r_out) {
                        if (dot) yd = fma(acc[i][j][e], dot[(dot_div == 1 ? row : row / dot_div) * dot_ld + b], yd);
                        else out[row * out_ld + b] = acc[i][j][e];
                    }
                }
            }
        }
        if (dot) {
            yd += __shfl_xor_sync(0xffffffffu, yd, 1);
            yd += __shfl_xor_sync(0xffffffffu, yd, 2);
            if (fk == 0 && row < rows) yhat[row] = yscale ? yd * yscale[row] : yd;
        }
    }
}

// ---- persistent, software-pipelined variant for small contractions (the whole core fits in shared memory):
//      the HBM-bound case (f = 2: configs 3 and 4).  A CTA keeps the core resident, walks over row tiles and
//      prefetches the next tile's environment rows and raw inputs with cp.async while it multiplies the current one.
__device__ __forceinline__ void cp_async8(void* smem_dst, const void* gmem_src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gmem_src) : "memory");
}

template <int NT>
__global__ void __launch_bounds__(ENV_THREADS, 2)
env_dmma_persist_kernel(const double* __restrict__ env_in, int64_t env_ld, int env_div, const double* __restrict__ x, int64_t x_ld,
                        int map_kind, int f, int cdiv, const double* __restrict__ core, double* __restrict__ out,
                        int64_t out_ld, const double* __restrict__ dot, int64_t dot_ld, int dot_div,
                        double* __restrict__ yhat, const double* __restrict__ yscale, int64_t rows, int r_in, int r_out, int64_t ntiles) {
    constexpr int LDG = NT * 8 + 4;
    extern __shared__ double sm[];
    const int in_st = (r_in + 5) | 1;
    const int phi_st = f | 1;
    const int K = r_in * f;
    const int Kp = (K + 3) & ~3;
    const int xraw = (map_kind == TN_MAP_IDENTITY) ? f : 1;      // raw input values per row
    double* s_g = sm;                                            // [Kp][LDG]  resident
    double* s_in = s_g + (size_t)Kp * LDG;                       // [2][ED_TR][in_st]
    double* s_x = s_in + 2 * (size_t)ED_TR * in_st;              // [2][ED_TR][xraw]
    double* s_phi = s_x + 2 * (size_t)ED_TR * xraw;              // [ED_TR][phi_st]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    for (int idx = tid; idx < Kp * (NT * 8); idx += ENV_THREADS) {
        const int k = idx / (NT * 8), n = idx - k * (NT * 8);
        s_g[k * LDG + n] = (k < K && n < r_out) ? core[(int64_t)k * r_out + n] : 0.0;
    }
    for (int idx = tid; idx < 2 * ED_TR * in_st; idx += ENV_THREADS) s_in[idx] = (env_in == nullptr && (idx % in_st) == 0) ? 1.0 : 0.0;

    auto prefetch = [&](int64_t tile, int buf) {
        if (tile < ntiles) {
            const int64_t row0 = tile * ED_TR;
            if (env_in) {
                double* dst = s_in + (size_t)buf * ED_TR * in_st;
                for (int idx = tid; idx < ED_TR * r_in; idx += ENV_THREADS) {
                    const int r = idx / r_in, a = idx - r * r_in;
                    int64_t row = row0 + r;
                    if (row >= rows) row = rows - 1;                 // clamped rows are never stored
                    cp_async8(dst + r * in_st + a, env_in + (env_div == 1 ? row : row / env_div) * env_ld + a);
                }
            }
            double* dx = s_x + (size_t)buf * ED_TR * xraw;
            for (int idx = tid; idx < ED_TR * xraw; idx += ENV_THREADS) {
                const int r = idx / xraw, q = idx - r * xraw;
                int64_t row = row0 + r;
                if (row >= rows) row = rows - 1;
                cp_async8(dx + idx, x + (cdiv == 1 ? row : row / cdiv) * x_ld + q);
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };

    const int fr = lane >> 2, fk = lane & 3;
    const int da = 4 / f, dp = 4 - da * f;
    int64_t tile = blockIdx.x;
    int buf = 0;
    __syncthreads();                 // zero / one fill of s_in done before the first async copies land on it
    prefetch(tile, 0);
    for (; tile < ntiles; tile += gridDim.x, buf ^= 1) {
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncthreads();             // tile data visible to all; everyone is done with the other buffer and with s_phi
        prefetch(tile + gridDim.x, buf ^ 1);
        const int64_t row0 = tile * ED_TR;
        const double* xin = s_x + (size_t)buf * ED_TR * xraw;
        if (map_kind == TN_MAP_SINCOS) {
            for (int r = tid; r < ED_TR; r += ENV_THREADS) {
                double c, sn;
                sincos((0.5 * 3.14159265358979323846) * xin[r], &sn, &c);
                s_phi[r * phi_st] = c;
                s_phi[r * phi_st + 1] = sn;
            }
        } else {
            for (int idx = tid; idx < ED_TR * f; idx += ENV_THREADS) {
                const int r = idx / f, p = idx - r * f;
                s_phi[r * phi_st + p] = map_eval(map_kind, xin + r * xraw, p);
            }
        }
        __syncthreads();
        double acc[2][NT][2];
#pragma unroll
        for (int i = 0; i < 2; ++i)
#pragma unroll
            for (int j = 0; j < NT; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
        int ka = fk / f, kp = fk - ka * f;
        const double* in0 = s_in + (size_t)buf * ED_TR * in_st + (warp * 16 + fr) * in_st;
        const double* in1 = in0 + 8 * in_st;
        const double* ph0 = s_phi + (warp * 16 + fr) * phi_st;
        const double* ph1 = ph0 + 8 * phi_st;
#pragma unroll 4
        for (int kk = 0; kk < Kp; kk += 4) {
            const double a0 = in0[ka] * ph0[kp];
            const double a1 = in1[ka] * ph1[kp];
            kp += dp; ka += da;
            if (kp >= f) { kp -= f; ++ka; }
            const double* gp = s_g + (kk + fk) * LDG + fr;
#pragma unroll
            for (int j = 0; j < NT; ++j) {
                const double b = gp[j * 8];
                dmma884(acc[0][j][0], acc[0][j][1], a0, b);
                dmma884(acc[1][j][0], acc[1][j][1], a1, b);
            }
        }
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            const int64_t row = row0 + warp * 16 + i * 8 + fr;
            double yd = 0.0;
            if (row < rows) {
#pragma unroll
                for (int j = 0; j < NT; ++j) {
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        const int b = j * 8 + 2 * fk + e;
                        if (b < r_out) {
                            if (dot) yd = fma(acc[i][j][e], dot[(dot_div == 1 ? row : row / dot_div) * dot_ld + b], yd);
                            else out[row * out_ld + b] = acc[i][j][e];
                        }
                    }
                }
            }
            if (dot) {
                yd += __shfl_xor_sync(0xffffffffu, yd, 1);
                yd += __shfl_xor_sync(0xffffffffu, yd, 2);
                if (fk == 0 && row < rows) yhat[row] = yscale ? yd * yscale[row] : yd;
            }
        }
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
}

// ---- warp-per-tile variant: every warp owns 16-row tiles and its own double-buffered pipeline, no block barrier in the loop.
//      The environment rows of a tile are one contiguous byte range (sample-major layout), so one elected lane fetches them with
//      a bulk copy (cp.async.bulk, the 1-D TMA path: UBLKCP) that completes on the warp's mbarrier; the raw site inputs (one or
//      f strided values per row) come by cp.async; the result tile goes back through shared memory as one bulk store.  The core
//      stays resident as the B operand.  With ~20 such independent pipelines per SM the FP64 tensor pipe and the HBM stream
//      overlap instead of alternating between the three block barriers per 128-row tile of the kernel above.
//      Preconditions (checked by the launcher): env rows dense (env_ld == r_in, env_div == 1, 16-byte aligned), r_in * 16 * 8 bytes
//      a multiple of 16 (always), the whole core in shared memory.
constexpr int EW_TR = 16;          // rows per warp tile
constexpr int EW_WARPS = 8;

__device__ __forceinline__ void ew_mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(bar)), "r"(count));
}
__device__ __forceinline__ bool ew_mbar_try(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"((uint32_t)__cvta_generic_to_shared(bar)), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void ew_mbar_wait(uint64_t* bar, uint32_t parity) {
    for (uint32_t it = 0; it < (1u << 26); ++it)
        if (ew_mbar_try(bar, parity)) return;
    __trap();      // a protocol bug traps instead of hanging the GPU
}

template <int NT>
__global__ void __launch_bounds__(EW_WARPS * 32, 2)
env_warp_kernel(const double* __restrict__ env_in, int64_t env_ld, const double* __restrict__ x, int64_t x_ld, int map_kind, int f, int cdiv,
                const double* __restrict__ core, double* __restrict__ out, int64_t out_ld, const double* __restrict__ dot, int64_t dot_ld,
                int dot_div, double* __restrict__ yhat, const double* __restrict__ yscale, int64_t rows, int r_in, int r_out,
                int64_t ntiles) {
    constexpr int LDG = NT * 8 + 4;
    extern __shared__ __align__(128) double sm[];
    const int phi_st = f | 1;
    const int K = r_in * f;
    const int Kp = (K + 3) & ~3;
    const int xraw = (map_kind == TN_MAP_IDENTITY) ? f : 1;
    const int in_st = r_in + ((12 - (r_in & 7)) & 7);         // row stride == 4 (mod 8) doubles: the A fragments of a half warp (4 rows x 2..4 k) hit distinct banks
    const int in_elems = EW_TR * in_st + 8;                   // padded tile (pad words stay zero: the padded tail of K reads them)
    const int out_st = NT * 8;                                // staging row stride when the result is not stored densely
    // per-warp region (doubles): [2][in_elems] [2][EW_TR * xraw] [EW_TR * phi_st] [EW_TR * out_st] + 2 barriers, rounded to 16 B
    const int per_warp = (2 * in_elems + 2 * EW_TR * xraw + EW_TR * phi_st + EW_TR * out_st + 2 + 1) & ~1;
    double* s_g = sm;                                         // [Kp][LDG] resident core
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    double* wbase = s_g + (((size_t)Kp * LDG + 1) & ~(size_t)1) + (size_t)warp * per_warp;
    double* s_in = wbase;
    double* s_x = s_in + 2 * in_elems;
    double* s_phi = s_x + 2 * EW_TR * xraw;
    double* s_out = s_phi + EW_TR * phi_st;
    uint64_t* bars = reinterpret_cast<uint64_t*>(s_out + EW_TR * out_st);

    for (int idx = tid; idx < Kp * (NT * 8); idx += EW_WARPS * 32) {
        const int k = idx / (NT * 8), n = idx - k * (NT * 8);
        s_g[k * LDG + n] = (k < K && n < r_out) ? core[(int64_t)k * r_out + n] : 0.0;
    }
    for (int i = lane; i < 2 * in_elems; i += 32) s_in[i] = (env_in == nullptr && (i % in_elems) % in_st == 0 && (i % in_elems) < EW_TR * in_st) ? 1.0 : 0.0;
    if (lane == 0) {
        ew_mbar_init(&bars[0], 1);
        ew_mbar_init(&bars[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();                   // core resident, barriers initialised (the only block barrier of the kernel)

    const int64_t nwarps = (int64_t)gridDim.x * EW_WARPS;
    const bool dense_out = (dot == nullptr) && (out_ld == r_out);
    auto prefetch = [&](int64_t tile, int buf) {
        if (tile < ntiles) {
            const int64_t row0 = tile * EW_TR;
            const int nrow = (int)((rows - row0 < EW_TR) ? rows - row0 : EW_TR);
            if (env_in) {
                // one bulk copy per row (lane = row) into the padded tile; lane 0 announces the tile's bytes first
                const uint32_t bar = (uint32_t)__cvta_generic_to_shared(&bars[buf]);
                if (lane == 0) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"((uint32_t)(nrow * r_in) * 8u) : "memory");
                __syncwarp();
                if (lane < nrow)
                    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                                 ::"r"((uint32_t)__cvta_generic_to_shared(s_in + (size_t)buf * in_elems + (size_t)lane * in_st)),
                                   "l"(env_in + (row0 + lane) * env_ld), "r"((uint32_t)r_in * 8u), "r"(bar) : "memory");
            }
            double* dx = s_x + (size_t)buf * EW_TR * xraw;
            for (int idx = lane; idx < EW_TR * xraw; idx += 32) {
                const int r = idx / xraw, q = idx - r * xraw;
                int64_t row = row0 + r;
                if (row >= rows) row = rows - 1;                      // clamped rows are never stored
                cp_async8(dx + idx, x + (cdiv == 1 ? row : row / cdiv) * x_ld + q);
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };

    const int fr = lane >> 2, fk = lane & 3;
    const int da = 4 / f, dp = 4 - da * f;
    int64_t tile = (int64_t)blockIdx.x * EW_WARPS + warp;
    int buf = 0;
    uint32_t phase[2] = {0u, 0u};
    prefetch(tile, 0);
    for (; tile < ntiles; tile += nwarps, buf ^= 1) {
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        if (env_in) {
            ew_mbar_wait(&bars[buf], phase[buf]);
            phase[buf] ^= 1u;
        }
        __syncwarp();
        if (dense_out) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");      // the previous bulk store has left s_out
        prefetch(tile + nwarps, buf ^ 1);
        const int64_t row0 = tile * EW_TR;
        const double* xin = s_x + (size_t)buf * EW_TR * xraw;
        if (map_kind == TN_MAP_SINCOS) {
            if (lane < EW_TR) {
                double c, sn;
                sincos((0.5 * 3.14159265358979323846) * xin[lane], &sn, &c);
                s_phi[lane * phi_st] = c;
                s_phi[lane * phi_st + 1] = sn;
            }
        } else {
            for (int idx = lane; idx < EW_TR * f; idx += 32) {
                const int r = idx / f, p = idx - r * f;
                s_phi[r * phi_st + p] = map_eval(map_kind, xin + r * xraw, p);
            }
        }
        __syncwarp();
        double acc[2][NT][2];
#pragma unroll
        for (int i = 0; i < 2; ++i)
#pragma unroll
            for (int j = 0; j < NT; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
        int ka = fk / f, kp = fk - ka * f;
        const double* in0 = s_in + (size_t)buf * in_elems + fr * in_st;      // a tail k reads past the row: multiplied by a zero row of the core
        const double* in1 = in0 + 8 * in_st;
        const double* ph0 = s_phi + fr * phi_st;
        const double* ph1 = ph0 + 8 * phi_st;
#pragma unroll 4
        for (int kk = 0; kk < Kp; kk += 4) {
            const int kpc = (kp < f) ? kp : 0;
            const double a0 = in0[ka] * ph0[kpc];
            const double a1 = in1[ka] * ph1[kpc];
            kp += dp; ka += da;
            if (kp >= f) { kp -= f; ++ka; }
            const double* gp = s_g + (kk + fk) * LDG + fr;
#pragma unroll
            for (int j = 0; j < NT; ++j) {
                const double b = gp[j * 8];
                dmma884(acc[0][j][0], acc[0][j][1], a0, b);
                dmma884(acc[1][j][0], acc[1][j][1], a1, b);
            }
        }
        // epilogue: this lane owns rows fr and fr + 8 of the tile, columns j*8 + 2*fk + {0,1}
        if (dot) {
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                const int64_t row = row0 + i * 8 + fr;
                double yd = 0.0;
                if (row < rows) {
#pragma unroll
                    for (int j = 0; j < NT; ++j)
#pragma unroll
                        for (int e = 0; e < 2; ++e) {
                            const int b = j * 8 + 2 * fk + e;
                            if (b < r_out) yd = fma(acc[i][j][e], dot[(dot_div == 1 ? row : row / dot_div) * dot_ld + b], yd);
                        }
                }
                yd += __shfl_xor_sync(0xffffffffu, yd, 1);
                yd += __shfl_xor_sync(0xffffffffu, yd, 2);
                if (fk == 0 && row < rows) yhat[row] = yscale ? yd * yscale[row] : yd;
            }
        } else if (dense_out) {
            // stage the 16 x r_out tile densely and hand it to the bulk-copy engine as one store
#pragma unroll
            for (int i = 0; i < 2; ++i)
#pragma unroll
                for (int j = 0; j < NT; ++j) {
                    const int b = j * 8 + 2 * fk;                  // r_out is even on this path: the pair (b, b + 1) is inside or outside together
                    if (b < r_out) *reinterpret_cast<double2*>(s_out + (i * 8 + fr) * r_out + b) = make_double2(acc[i][j][0], acc[i][j][1]);
                }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            if (lane == 0) {
                const int nrow = (int)((rows - row0 < EW_TR) ? rows - row0 : EW_TR);
                asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                             ::"l"(out + row0 * r_out), "r"((uint32_t)__cvta_generic_to_shared(s_out)), "r"((uint32_t)(nrow * r_out) * 8u) : "memory");
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            }
        } else {
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                const int64_t row = row0 + i * 8 + fr;
                if (row < rows) {
#pragma unroll
                    for (int j = 0; j < NT; ++j)
#pragma unroll
                        for (int e = 0; e < 2; ++e) {
                            const int b = j * 8 + 2 * fk + e;
                            if (b < r_out) out[row * out_ld + b] = acc[i][j][e];
                        }
                }
            }
        }
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    if (dense_out) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

template <int NT>
static int launch_env_dmma(const double* env_in, int64_t env_ld, int env_div, const double* x, int64_t x_ld, int map_kind, int f,
                           int cdiv, const double* core, double* out, int64_t out_ld, const double* dot, int64_t dot_ld,
                           int dot_div, double* yhat, const double* yscale, int64_t rows, int r_in, int r_out, cudaStream_t st) {
    {   // warp-per-tile kernel with bulk-copy (TMA) staging: dense environment rows, the whole core resident
        const int K = r_in * f, Kp = (K + 3) & ~3, xraw = (map_kind == TN_MAP_IDENTITY) ? f : 1;
        const int in_elems = EW_TR * (r_in + ((12 - (r_in & 7)) & 7)) + 8;
        const int per_warp = (2 * in_elems + 2 * EW_TR * xraw + EW_TR * (f | 1) + EW_TR * NT * 8 + 2 + 1) & ~1;
        const size_t wsmem = ((((size_t)Kp * (NT * 8 + 4) + 1) & ~(size_t)1) + (size_t)EW_WARPS * per_warp) * sizeof(double);
        const bool dense_in = (env_in == nullptr) || (env_div == 1 && env_ld % 2 == 0 && (reinterpret_cast<uintptr_t>(env_in) & 15) == 0 && (r_in % 2 == 0));   // every row is a 16-byte aligned bulk copy
        const bool out_ok = dot != nullptr || out_ld != r_out || ((reinterpret_cast<uintptr_t>(out) & 15) == 0 && (r_out % 2 == 0));
        const int64_t wtiles = ceil_div64(rows, EW_TR);
        if (dense_in && out_ok && wsmem <= 220 * 1024 && wtiles >= 8LL * EW_WARPS * sm_count() && !getenv("TN_ENV_NO_WARP")) {
            TN_SMEM(env_warp_kernel<NT>, wsmem);
            int64_t grid = (wsmem <= 110 * 1024 ? 2LL : 1LL) * sm_count();      // two CTAs (16 warp pipelines) per SM when they fit
            env_warp_kernel<NT><<<(unsigned)grid, EW_WARPS * 32, wsmem, st>>>(env_in, env_ld, x, x_ld, map_kind, f, cdiv, core, out, out_ld, dot, dot_ld,
                                                                         dot_div, yhat, yscale, rows, r_in, r_out, wtiles);
            TN_LAUNCH_CHECK();
            return TN_OK;
        }
    }
    {   // persistent pipelined kernel when the whole core slab stays resident
        const int K = r_in * f, Kp = (K + 3) & ~3, xraw = (map_kind == TN_MAP_IDENTITY) ? f : 1;
        const size_t psmem = ((size_t)Kp * (NT * 8 + 4) + 2 * (size_t)ED_TR * ((r_in + 5) | 1) + 2 * (size_t)ED_TR * xraw +
                              (size_t)ED_TR * (f | 1)) * sizeof(double);
        const int64_t ntiles = ceil_div64(rows, ED_TR);
        if ((size_t)Kp * (NT * 8 + 4) * sizeof(double) <= 48 * 1024 && psmem <= 113 * 1024 && ntiles >= 4LL * sm_count() &&
            !getenv("TN_ENV_NO_PERSIST")) {
            TN_SMEM(env_dmma_persist_kernel<NT>, psmem);
            int64_t grid = 2LL * sm_count();
            if (grid > ntiles) grid = ntiles;
            env_dmma_persist_kernel<NT><<<(unsigned)grid, ENV_THREADS, psmem, st>>>(env_in, env_ld, env_div, x, x_ld, map_kind, f, cdiv, core,
                                                                               out, out_ld, dot, dot_ld, dot_div, yhat, yscale, rows, r_in, r_out,
                                                                               ntiles);
            TN_LAUNCH_CHECK();
            return TN_OK;
        }
    }
    const size_t smem = ((size_t)ED_TR * ((r_in + 5) | 1) + (size_t)ED_TR * (f | 1) + (size_t)ED_KC * (NT * 8 + 4)) * sizeof(double);
    if (smem > 113 * 1024) return 1;   // would not leave room for two CTAs per SM: let the caller use the FMA kernel
    TN_SMEM(env_dmma_kernel<NT>, smem);
    const int64_t grid = ceil_div64(rows, ED_TR);
    TN_CHECK_ARG(grid <= 0x7fffffff, "tn_env_update: too many rows");
    env_dmma_kernel<NT><<<(unsigned)grid, ENV_THREADS, smem, st>>>(env_in, env_ld, env_div, x, x_ld, map_kind, f, cdiv, core, out, out_ld,
                                                                 dot, dot_ld, dot_div, yhat, yscale, rows, r_in, r_out);
    TN_LAUNCH_CHECK();
    return TN_OK;
}

__global__ void class_rows_kernel(const double* __restrict__ env, const double* __restrict__ U,
                                  const double* __restrict__ g, double* __restrict__ F, double* __restrict__ G,
                                  int64_t S, int C, int V, int r) {
    const int64_t nF = F ? S * V * r : 0;
    const int64_t nG = G ? S * r : 0;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < nF + nG;
         idx += (int64_t)gridDim.x * blockDim.x) {
        if (idx < nF) {
            const int a = (int)(idx % r);
            const int64_t st = idx / r;
            const int64_t s = st / V;
            const double* u = U + st * C;
            const double* e = env + s * C * r + a;
            double acc = 0.0;
            for (int c = 0; c < C; ++c) acc = fma(u[c], e[(int64_t)c * r], acc);
            F[idx] = acc;
        } else {
            const int64_t k = idx - nF;
            const int a = (int)(k % r);
            const int64_t s = k / r;
            const double* e = env + s * C * r + a;
            double acc = 0.0;
            for (int c = 0; c < C; ++c) acc = fma(g[s * C + c], e[(int64_t)c * r], acc);
            G[k] = acc;
        }
    }
}

}  // namespace tn

extern "C" int tn_env_update(const double* env_in, int64_t env_ld, int env_div, const double* x, int64_t x_ld, int map_kind, int f,
                             int cdiv, const double* core, double* out, int64_t out_ld, const double* dot,
                             int64_t dot_ld, int dot_div, double* yhat, int64_t rows, int r_in, int r_out,
                             void* stream) {
    return tn::env_update_scaled(env_in, env_ld, env_div, x, x_ld, map_kind, f, cdiv, core, out, out_ld, dot, dot_ld, dot_div, yhat, nullptr,
                                 rows, r_in, r_out, tn::as_stream(stream));
}

int tn::env_update_scaled(const double* env_in, int64_t env_ld, int env_div, const double* x, int64_t x_ld, int map_kind, int f, int cdiv,
                          const double* core, double* out, int64_t out_ld, const double* dot, int64_t dot_ld, int dot_div, double* yhat,
                          const double* yscale, int64_t rows, int r_in, int r_out, cudaStream_t stream) {
    using namespace tn;
    TN_CHECK_ARG(rows >= 0 && r_in >= 1 && r_out >= 1 && f >= 1 && cdiv >= 1, "tn_env_update: bad sizes");
    if (rows == 0) return TN_OK;       // an empty shard / batch: nothing to read or write (the pointers may be null)
    TN_CHECK_ARG(env_in != nullptr || r_in == 1, "tn_env_update: env_in == NULL requires r_in == 1");
    TN_CHECK_ARG(x && core, "tn_env_update: null input");
    TN_CHECK_ARG(dot ? (yhat != nullptr && dot_div >= 1) : (out != nullptr), "tn_env_update: missing output");
    TN_CHECK_ARG(map_kind >= 0 && map_kind <= 2, "tn_env_update: unknown map_kind %d", map_kind);
    TN_CHECK_ARG(map_kind != TN_MAP_SINCOS || f == 2, "tn_env_update: sin-cos map has f == 2");
    if (rows == 0) return TN_OK;
    // FP64 tensor-core path for every shape it covers (r_out <= 104, K = r_in*f >= 8); the FMA kernel below is the
    // general one (very wide outputs, tiny contractions, shared-memory overflow).
    if (r_out <= 104 && (int64_t)r_in * f >= 8 && !getenv("TN_ENV_NO_DMMA")) {
        const int nt = (r_out + 7) / 8;
        const int edv = env_div < 1 ? 1 : env_div;
        cudaStream_t st = stream;
        int rc = 1;
#define TN_ENV_DMMA(NTV) rc = launch_env_dmma<NTV>(env_in, env_ld, edv, x, x_ld, map_kind, f, cdiv, core, out, out_ld, dot, dot_ld, dot_div, yhat, yscale, rows, r_in, r_out, st)
        if (nt <= 1) TN_ENV_DMMA(1);
        else if (nt <= 2) TN_ENV_DMMA(2);
        else if (nt <= 3) TN_ENV_DMMA(3);
        else if (nt <= 4) TN_ENV_DMMA(4);
        else if (nt <= 5) TN_ENV_DMMA(5);
        else if (nt <= 6) TN_ENV_DMMA(6);
        else if (nt <= 8) TN_ENV_DMMA(8);
        else TN_ENV_DMMA(13);
#undef TN_ENV_DMMA
        if (rc <= 0) return rc;      // launched (0) or failed (<0); rc == 1 means "does not fit", fall through
    }
    // column tiling: 8 lanes x ceil(r_out/8) columns per thread for narrow outputs (no padded work), 16 x 4 for wide ones
    int TXN = 8, CN = (r_out + 7) / 8;
    if (CN > 5) { TXN = 16; CN = 4; }
    const int TR = (ENV_THREADS / TXN) * ENV_RT, TC = TXN * CN;
    int p_chunk = (int)(48 * 1024 / sizeof(double)) / (r_in * TC);     // <= 48 KB of core slab per stage
    if (p_chunk < 1) p_chunk = 1;
    if (p_chunk > f) p_chunk = f;
    const size_t smem = ((size_t)TR * (r_in | 1) + (size_t)TR * (f | 1) + (size_t)p_chunk * r_in * TC) * sizeof(double);
    TN_CHECK_ARG(smem <= 227 * 1024, "tn_env_update: r_in=%d f=%d needs %zu B of shared memory", r_in, f, smem);
    using Kern = void (*)(const double*, int64_t, int, const double*, int64_t, int, int, int, const double*, double*, int64_t,
                          const double*, int64_t, int, double*, const double*, int64_t, int, int, int);
    static const Kern kerns[6] = {env_kernel<8, 1>, env_kernel<8, 2>, env_kernel<8, 3>, env_kernel<8, 4>, env_kernel<8, 5>, env_kernel<16, 4>};
    const int ki = (TXN == 16) ? 5 : CN - 1;
    TN_SMEM(kerns[ki], smem);
    const int64_t grid = ceil_div64(rows, TR);
    TN_CHECK_ARG(grid <= 0x7fffffff, "tn_env_update: too many rows");
    kerns[ki]<<<(unsigned)grid, ENV_THREADS, smem, stream>>>(env_in, env_ld, env_div < 1 ? 1 : env_div, x, x_ld, map_kind, f,
                                                                       cdiv, core, out, out_ld, dot, dot_ld, dot_div, yhat, yscale, rows,
                                                                       r_in, r_out, p_chunk);
    TN_LAUNCH_CHECK();
    return TN_OK;
}

extern "C" int tn_class_rows(const double* env, const double* U, const double* g, double* F, double* G, int64_t S,
                             int C, int V, int r, void* stream) {
    using namespace tn;
    TN_CHECK_ARG(env && S >= 0 && C >= 1 && r >= 1, "tn_class_rows: bad arguments");
    TN_CHECK_ARG((F == nullptr) || (U != nullptr && V >= 1), "tn_class_rows: F needs U");
    TN_CHECK_ARG((G == nullptr) || (g != nullptr), "tn_class_rows: G needs g");
    const int64_t n = (F ? S * V * r : 0) + (G ? S * r : 0);
    if (n == 0) return TN_OK;
    const int threads = 256;
    int64_t blocks = ceil_div64(n, threads);
    const int64_t cap = (int64_t)sm_count() * 16;
    if (blocks > cap) blocks = cap;
    class_rows_kernel<<<(unsigned)blocks, threads, 0, as_stream(stream)>>>(env, U, g, F, G, S, C, V, r);
    TN_LAUNCH_CHECK();
    return TN_OK;
}
