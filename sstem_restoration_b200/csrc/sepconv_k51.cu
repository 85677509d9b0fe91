// placeholder until the tuned kernels land: route to generic
#include "common.cuh"
namespace sstem {
int launch_sepconv_fwd_k51(const float* in, const float* v, const float* h, float* out,
                           int64_t B, int64_t C, int64_t H, int64_t W, cudaStream_t s) {
    return launch_sepconv_fwd_generic(in, v, h, out, B, C, H, W, 51, false, s);
}
int launch_sepconv_bwd_taps_k51(const float* g, const float* in, const float* v, const float* h,
                                float* gv, float* gh, int64_t B, int64_t C, int64_t H, int64_t W, cudaStream_t s) {
    return launch_sepconv_bwd_taps_generic(g, in, v, h, gv, gh, B, C, H, W, 51, s);
}
}
