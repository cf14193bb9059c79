// Matrix-free local operator for the Krylov solvers: out = J^T diag(w) J v.
//
// Replaces the per-batch get_J + two einsums of lanczos_swipe / scipy_swipe (reference
// tensor/network.py:770-790, 896-918).  J[row,(ia,ib,ic)] = fa[ia] fb[ib] fc[ic] is never
// formed: pass 1 contracts v (viewed as a core [ma, mb, mc]) with the three factors -- the same
// kernel as an environment step with the prediction epilogue -- and pass 2 is the right-hand-side
// accumulation with weights w * (J v).
#include "common.cuh"

extern "C" int tn_rhs_ksplit(int64_t rows, int ma, int mb, int mc);

namespace tn {
__global__ void mul_weights_kernel(double* __restrict__ t, const double* __restrict__ w, int64_t n) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) t[i] *= w[i];
}
}  // namespace tn

// work: rows doubles for J v, followed by ksplit * P doubles of right-hand-side partials.
extern "C" int64_t tn_matvec_work_elems(int64_t rows, int ma, int mb, int mc) {
    const int ks = tn_rhs_ksplit(rows, ma, mb, mc);
    return rows + (int64_t)ks * ma * mb * mc;
}

extern "C" int tn_matvec_kr3(const tn_factor* fa, const tn_factor* fb, const tn_factor* fc, const double* w, int64_t rows,
                             const double* v, double* out, double* work, void* stream) {
    using namespace tn;
    TN_CHECK_ARG(fa && fb && fc && v && out && work, "tn_matvec_kr3: null argument");
    TN_CHECK_ARG(fa->map_kind == TN_MAP_IDENTITY && fc->map_kind == TN_MAP_IDENTITY, "tn_matvec_kr3: only the middle factor may carry a feature map");
    int rc = matvec_fused(fa, fb, fc, w, rows, v, out, nullptr, as_stream(stream));      // small cores: both passes in one launch
    if (rc <= 0) return rc;
    double* t = work;
    rc = tn_env_update(fa->ptr, fa->ld, fa->div, fb->ptr, fb->ld, fb->map_kind, fb->m, fb->div, v, nullptr, 0, fc->ptr,
                           fc->ld, fc->div, t, rows, fa->m, fc->m, stream);
    if (rc != TN_OK) return rc;
    if (w && rows > 0) {
        int64_t blocks = ceil_div64(rows, 256);
        if (blocks > 8LL * sm_count()) blocks = 8LL * sm_count();
        mul_weights_kernel<<<(unsigned)blocks, 256, 0, as_stream(stream)>>>(t, w, rows);
        TN_LAUNCH_CHECK();
    }
    const int ks = tn_rhs_ksplit(rows, fa->m, fb->m, fc->m);
    return tn_rhs_kr3(fa, fb, fc, t, rows, out, work + rows, ks, 0, stream);
}
