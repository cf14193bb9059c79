// tcgen05 (TF32 / 3xTF32) Gram kernel -- placeholder until the tensor-core path lands.
#include "common.cuh"

int tn_gram_kr3_tc(int mode, const tn_factor* fa, const tn_factor* fb, const tn_factor* fc, const double* w, int64_t rows,
                   double* M, int accumulate, void* stream) {
    (void)mode; (void)fa; (void)fb; (void)fc; (void)w; (void)rows; (void)M; (void)accumulate; (void)stream;
    tn::set_error("tn_gram_kr3: tensor-core modes are not built yet");
    return TN_EUNSUPPORTED;
}
