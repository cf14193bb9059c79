// Fused feature-map + environment step, prediction epilogue and class-leg row folding.
//
// Replaces the chained pairwise einsums of TensorNetwork.compute_stacks /
// left_update_stacks / right_update_stacks / forward (reference tensor/network.py:55-71,
// 115-137, 152-172): out[row, b] = sum_{a,p} env[row, a] * phi(x[row / cdiv], p) * core[a, p, b].
//
// Layout: environments are sample-major [rows, r] fp64, so a tile of consecutive rows is one
// contiguous byte range; the site input is read once per sample and mapped in registers.
// One CTA owns ENV_TR rows; the core is streamed through shared memory in (a, p) slabs that
// every CTA re-reads from L2 (the core is at most a few hundred KB).
#include "common.cuh"

namespace tn {

constexpr int ENV_TR = 64;       // rows per CTA
constexpr int ENV_TC = 64;       // output columns per pass
constexpr int ENV_THREADS = 256; // 16 (row groups of 4) x 16 (column lanes)

__global__ void __launch_bounds__(ENV_THREADS)
env_kernel(const double* __restrict__ env_in, int64_t env_ld, int env_div, const double* __restrict__ x, int64_t x_ld,
           int map_kind, int f, int cdiv, const double* __restrict__ core, double* __restrict__ out,
           int64_t out_ld, const double* __restrict__ dot, int64_t dot_ld, int dot_div,
           double* __restrict__ yhat, int64_t rows, int r_in, int r_out, int a_chunk) {
    extern __shared__ double sm[];
    const int in_st = r_in | 1;
    const int phi_st = f | 1;
    double* s_in = sm;
    double* s_phi = s_in + ENV_TR * in_st;
    double* s_g = s_phi + ENV_TR * phi_st;

    const int tid = threadIdx.x;
    const int tx = tid & 15;
    const int ty = tid >> 4;
    const int64_t row0 = (int64_t)blockIdx.x * ENV_TR;

    for (int idx = tid; idx < ENV_TR * r_in; idx += ENV_THREADS) {
        const int r = idx / r_in, a = idx - r * r_in;
        const int64_t row = row0 + r;
        double v = 0.0;
        if (row < rows) v = env_in ? env_in[(row / env_div) * env_ld + a] : 1.0;
        s_in[r * in_st + a] = v;
    }
    for (int idx = tid; idx < ENV_TR * f; idx += ENV_THREADS) {
        const int r = idx / f, p = idx - r * f;
        const int64_t row = row0 + r;
        double v = 0.0;
        if (row < rows) v = map_eval(map_kind, x + (row / cdiv) * x_ld, p);
        s_phi[r * phi_st + p] = v;
    }

    double ydot[4] = {0.0, 0.0, 0.0, 0.0};
    for (int c0 = 0; c0 < r_out; c0 += ENV_TC) {
        double acc[4][4];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[i][j] = 0.0;

        for (int a0 = 0; a0 < r_in; a0 += a_chunk) {
            const int ac = min(a_chunk, r_in - a0);
            __syncthreads();
            for (int idx = tid; idx < ac * f * ENV_TC; idx += ENV_THREADS) {
                const int col = idx & (ENV_TC - 1);
                const int ap = idx >> 6;  // (a, p) flattened, ENV_TC == 64
                const int b = c0 + col;
                s_g[idx] = (b < r_out) ? core[((int64_t)a0 * f + ap) * r_out + b] : 0.0;
            }
            __syncthreads();
            for (int a = 0; a < ac; ++a) {
                double ein[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) ein[i] = s_in[(ty * 4 + i) * in_st + a0 + a];
                const double* gp = s_g + (a * f) * ENV_TC + tx;
                for (int p = 0; p < f; ++p) {
                    double z[4], g[4];
#pragma unroll
                    for (int i = 0; i < 4; ++i) z[i] = ein[i] * s_phi[(ty * 4 + i) * phi_st + p];
#pragma unroll
                    for (int j = 0; j < 4; ++j) g[j] = gp[p * ENV_TC + 16 * j];
#pragma unroll
                    for (int i = 0; i < 4; ++i)
#pragma unroll
                        for (int j = 0; j < 4; ++j) acc[i][j] = fma(z[i], g[j], acc[i][j]);
                }
            }
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int64_t row = row0 + ty * 4 + i;
            if (row >= rows) continue;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int b = c0 + tx + 16 * j;
                if (b >= r_out) continue;
                if (dot)
                    ydot[i] = fma(acc[i][j], dot[(row / dot_div) * dot_ld + b], ydot[i]);
                else
                    out[row * out_ld + b] = acc[i][j];
            }
        }
    }
    if (dot) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            double v = ydot[i];
            v += __shfl_xor_sync(0xffffffffu, v, 8);
            v += __shfl_xor_sync(0xffffffffu, v, 4);
            v += __shfl_xor_sync(0xffffffffu, v, 2);
            v += __shfl_xor_sync(0xffffffffu, v, 1);
            const int64_t row = row0 + ty * 4 + i;
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
    int a_chunk = 8192 / (f * ENV_TC);
    if (a_chunk < 1) a_chunk = 1;
    if (a_chunk > r_in) a_chunk = r_in;
    const size_t smem = ((size_t)ENV_TR * (r_in | 1) + (size_t)ENV_TR * (f | 1) + (size_t)a_chunk * f * ENV_TC) * sizeof(double);
    TN_CHECK_ARG(smem <= 227 * 1024, "tn_env_update: r_in=%d f=%d needs %zu B of shared memory", r_in, f, smem);
    static size_t configured = 0;
    if (smem > configured) {
        TN_CUDA(cudaFuncSetAttribute(env_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = smem;
    }
    const int64_t grid = ceil_div64(rows, ENV_TR);
    TN_CHECK_ARG(grid <= 0x7fffffff, "tn_env_update: too many rows");
    env_kernel<<<(unsigned)grid, ENV_THREADS, smem, as_stream(stream)>>>(env_in, env_ld, env_div < 1 ? 1 : env_div, x, x_ld, map_kind, f, cdiv, core,
                                                                         out, out_ld, dot, dot_ld, dot_div, yhat, rows,
                                                                         r_in, r_out, a_chunk);
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
