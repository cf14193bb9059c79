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
           double* __restrict__ yhat, int64_t rows, int r_in, int r_out, int p_chunk) {
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
            if (tx == 0 && row < rows) yhat[row] = v;
        }
    }
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
    using namespace tn;
    TN_CHECK_ARG(rows >= 0 && r_in >= 1 && r_out >= 1 && f >= 1 && cdiv >= 1, "tn_env_update: bad sizes");
    TN_CHECK_ARG(env_in != nullptr || r_in == 1, "tn_env_update: env_in == NULL requires r_in == 1");
    TN_CHECK_ARG(x && core, "tn_env_update: null input");
    TN_CHECK_ARG(dot ? (yhat != nullptr && dot_div >= 1) : (out != nullptr), "tn_env_update: missing output");
    TN_CHECK_ARG(map_kind >= 0 && map_kind <= 2, "tn_env_update: unknown map_kind %d", map_kind);
    TN_CHECK_ARG(map_kind != TN_MAP_SINCOS || f == 2, "tn_env_update: sin-cos map has f == 2");
    if (rows == 0) return TN_OK;
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
                          const double*, int64_t, int, double*, int64_t, int, int, int);
    static const Kern kerns[6] = {env_kernel<8, 1>, env_kernel<8, 2>, env_kernel<8, 3>, env_kernel<8, 4>, env_kernel<8, 5>, env_kernel<16, 4>};
    static size_t configured[6] = {0, 0, 0, 0, 0, 0};
    const int ki = (TXN == 16) ? 5 : CN - 1;
    if (smem > configured[ki]) {
        TN_CUDA(cudaFuncSetAttribute(kerns[ki], cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured[ki] = smem;
    }
    const int64_t grid = ceil_div64(rows, TR);
    TN_CHECK_ARG(grid <= 0x7fffffff, "tn_env_update: too many rows");
    kerns[ki]<<<(unsigned)grid, ENV_THREADS, smem, as_stream(stream)>>>(env_in, env_ld, env_div < 1 ? 1 : env_div, x, x_ld, map_kind, f,
                                                                       cdiv, core, out, out_ld, dot, dot_ld, dot_div, yhat, rows,
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
