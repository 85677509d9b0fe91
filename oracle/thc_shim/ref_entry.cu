/*
 * ORACLE -- TEST INFRASTRUCTURE ONLY.
 * Plain-pointer entry points around the reference's own host launchers
 * (libs/sepconv/src/SeparableConvolution_kernel.cu:54-73 forward, :152-206
 * backward), which are compiled verbatim next to this file by oracle/Makefile.
 * Plays the role of libs/sepconv/src/SeparableConvolution_cuda.c:13-52 without
 * the THC global state.  Used by tests (P0: CPU oracle vs the real reference
 * kernels on a B200) and to generate tests/golden/sepconv_ref_*.npz.
 */
#include "THC.h"

extern "C" {
void SeparableConvolution_kernel_forward(THCState*, THCudaTensor*, THCudaTensor*, THCudaTensor*, THCudaTensor*);
void SeparableConvolution_kernel_backward(THCState*, THCudaTensor*, THCudaTensor*, THCudaTensor*, THCudaTensor*,
                                          THCudaTensor*, THCudaTensor*, THCudaTensor*);
}

static THCudaTensor make4(const float* p, long a, long b, long c, long d) {
    THCudaTensor t;
    t.data = const_cast<float*>(p);
    t.size[0] = a; t.size[1] = b; t.size[2] = c; t.size[3] = d;
    t.stride[3] = 1; t.stride[2] = d; t.stride[1] = c * d; t.stride[0] = b * c * d;
    return t;
}

extern "C" int ref_sepconv_forward(const float* in, const float* v, const float* h, float* out,
                                   long B, long C, long H, long W, void* stream) {
    THCState st; st.stream = (cudaStream_t)stream; st.last_error = 0;
    THCudaTensor ti = make4(in, B, C, H + 50, W + 50), tv = make4(v, B, 51, H, W),
                 th = make4(h, B, 51, H, W), to = make4(out, B, C, H, W);
    SeparableConvolution_kernel_forward(&st, &ti, &tv, &th, &to);
    return st.last_error;
}

/* The reference backward reads channels 0,1,2 literally (kernel.cu:100-108): C must be 3. */
extern "C" int ref_sepconv_backward(const float* g, const float* in, const float* v, const float* h,
                                    float* gi, float* gv, float* gh,
                                    long B, long C, long H, long W, void* stream) {
    if (C != 3) return -1;
    THCState st; st.stream = (cudaStream_t)stream; st.last_error = 0;
    THCudaTensor tg = make4(g, B, C, H, W), ti = make4(in, B, C, H + 50, W + 50),
                 tv = make4(v, B, 51, H, W), th = make4(h, B, 51, H, W),
                 tgi = make4(gi, B, C, H + 50, W + 50), tgv = make4(gv, B, 51, H, W),
                 tgh = make4(gh, B, 51, H, W);
    SeparableConvolution_kernel_backward(&st, &tg, &ti, &tv, &th, &tgi, &tgv, &tgh);
    return st.last_error;
}
