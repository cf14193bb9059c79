// Shared helpers for the tn_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include "../../include/tn_b200.h"

namespace tn {

void set_error(const char* fmt, ...);

#define TN_CHECK_ARG(cond, ...)            \
    do {                                   \
        if (!(cond)) {                     \
            tn::set_error(__VA_ARGS__);    \
            return TN_EINVAL;              \
        }                                  \
    } while (0)

#define TN_CUDA(call)                                                                   \
    do {                                                                                \
        cudaError_t e__ = (call);                                                       \
        if (e__ != cudaSuccess) {                                                       \
            tn::set_error("%s:%d %s: %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__)); \
            return TN_ECUDA;                                                            \
        }                                                                               \
    } while (0)

#define TN_LAUNCH_CHECK()          \
    do {                           \
        tn::count_launch();        \
        TN_CUDA(cudaGetLastError()); \
    } while (0)

static inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

__host__ __device__ __forceinline__ int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }
__host__ __device__ __forceinline__ int npairs(int m) { return m * (m + 1) / 2; }
// index of the unordered pair (i <= j) among m, row-major over the upper triangle
__host__ __device__ __forceinline__ int pair_index(int i, int j, int m) { return i * m - (i * (i - 1)) / 2 + (j - i); }
// inverse of pair_index
__host__ __device__ inline void pair_decode(int q, int m, int& i, int& j) {
    int ii = 0, rem = q;
    while (rem >= m - ii) {
        rem -= m - ii;
        ++ii;
    }
    i = ii;
    j = ii + rem;
}

// Evaluate one site-input entry under a feature map.  raw = x[row*ld] for SINCOS/POLY.
__device__ __forceinline__ double map_eval(int map_kind, const double* __restrict__ xrow, int p) {
    if (map_kind == TN_MAP_IDENTITY) return xrow[p];
    const double t = xrow[0];
    if (map_kind == TN_MAP_SINCOS) {
        const double a = (0.5 * 3.14159265358979323846) * t;  // same operation order as models/tnml.py:14
        return p == 0 ? cos(a) : sin(a);
    }
    double v = 1.0;  // TN_MAP_POLY
    for (int d = 0; d < p; ++d) v *= t;
    return v;
}

// The two halves of map_eval for software-pipelined kernels: the address to fetch now, the function to apply to the fetched value later.
__device__ __forceinline__ const double* map_raw_ptr(int map_kind, const double* __restrict__ xrow, int p) {
    return (map_kind == TN_MAP_IDENTITY) ? xrow + p : xrow;
}
__device__ __forceinline__ double map_apply(int map_kind, double raw, int p) {
    if (map_kind == TN_MAP_IDENTITY) return raw;
    if (map_kind == TN_MAP_SINCOS) {
        const double a = (0.5 * 3.14159265358979323846) * raw;
        return p == 0 ? cos(a) : sin(a);
    }
    double v = 1.0;
    for (int d = 0; d < p; ++d) v *= raw;
    return v;
}

// Stream-ordered scratch that is returned on EVERY path out of a launcher (error returns included).
struct AsyncScratch {
    void* ptr = nullptr;
    cudaStream_t st = nullptr;
    AsyncScratch() = default;
    AsyncScratch(const AsyncScratch&) = delete;
    AsyncScratch& operator=(const AsyncScratch&) = delete;
    cudaError_t alloc(size_t bytes, cudaStream_t s) {
        st = s;
        return cudaMallocAsync(&ptr, bytes, s);
    }
    ~AsyncScratch() {
        if (ptr) cudaFreeAsync(ptr, st);
    }
};

int sm_count();   // of the calling thread's current device (cached per device)
// Raise a kernel's dynamic shared-memory limit to at least `bytes` on the CURRENT device.  The attribute is per device and per
// function; what has been set is remembered per (device, function) under a mutex, so a process that drives several GPUs (or
// several host threads) configures each of them.
cudaError_t ensure_dyn_smem(const void* func, size_t bytes);
#define TN_SMEM(kernel, bytes) TN_CUDA(tn::ensure_dyn_smem(reinterpret_cast<const void*>(kernel), (size_t)(bytes)))
// tensor-core trailing update of the blocked Cholesky (syrk_tc.cu)
int64_t syrk_tc_work_floats(int64_t n, int kb);
void syrk_tc_set_passes(int passes);   // 3 (default): 3xTF32; 1: one TF32 pass (calling thread's later syrk_tc_update calls)
int syrk_tc_update(double* A, int64_t lda, int64_t P, int64_t c0, int64_t k0, int kb, float* X, const int* info, cudaStream_t st,
                   int64_t col_limit = 0);
void count_launch(int n = 1);   // bookkeeping for tn_launch_count()
// blocked Cholesky internals shared with the Krylov drivers (solve.cu)
int cholesky_factorize(double* A, int64_t lda, int64_t P, double* work, int* info, cudaStream_t st, float* X, int64_t NBO);
int cholesky_substitute(const double* A, int64_t lda, int64_t P, double* rhs, const double* work, const int* stop, cudaStream_t st);
int64_t cholesky_default_nbo(int64_t P);
// one-launch matrix-free operator for small cores (matvec_fused.cu): TN_OK = launched, 1 = shape outside its range
int matvec_fused(const tn_factor* fa, const tn_factor* fb, const tn_factor* fc, const double* w, int64_t rows, const double* v, double* out,
                 const int* stop, cudaStream_t st);
// environment step with an optional per-row scale of the prediction epilogue (env.cu): yhat[row] *= yscale[row]
int env_update_scaled(const double* env_in, int64_t env_ld, int env_div, const double* x, int64_t x_ld, int map_kind, int f, int cdiv,
                      const double* core, double* out, int64_t out_ld, const double* dot, int64_t dot_ld, int dot_div, double* yhat,
                      const double* yscale, int64_t rows, int r_in, int r_out, cudaStream_t st);

// FP64 tensor-core MMA, D(8x8) += A(8x4, row) * B(4x8, col): a = A[lane/4][lane%4], b = B[lane%4][lane/4],
// d0/d1 = D[lane/4][2*(lane%4) + 0/1].
__device__ __forceinline__ void dmma884(double& d0, double& d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

}  // namespace tn
