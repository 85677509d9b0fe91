// Generic-tap-count sepconv kernels (any taps in 1..64, any channel count).
// They are the fallback for taps != 51, the STRICT_ORDER verification mode, and
// the first-correct implementation the tuned 51-tap kernels are checked against.
// One thread per output element, operands read through L1/L2 -- not tuned.
//
// Arithmetic follows libs/sepconv/src/SeparableConvolution_kernel.cu:38-51
// (forward) and :97-111 / :134-149 (tap gradients) of the reference; the code
// is written from the formula, not translated.
#include "common.cuh"

namespace sstem {

template <bool STRICT>
__global__ void __launch_bounds__(256)
sepconv_fwd_generic_kernel(const float* __restrict__ in, const float* __restrict__ v,
                           const float* __restrict__ h, float* __restrict__ out,
                           int64_t total, int C, int H, int W, int K) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int x = (int)(idx % W);
    const int y = (int)((idx / W) % H);
    const int c = (int)((idx / ((int64_t)W * H)) % C);
    const int64_t b = idx / ((int64_t)W * H * C);
    const int IW = W + K - 1, IH = H + K - 1;
    const int64_t plane = (int64_t)H * W;
    const float* pin = in + ((b * C + c) * IH + y) * (int64_t)IW + x;
    const float* pv = v + b * K * plane + (int64_t)y * W + x;
    const float* ph = h + b * K * plane + (int64_t)y * W + x;
    float acc = 0.f;
    if (STRICT) {
        // reference chain: acc = fma(in*v, h, acc), fy outer, fx inner
        for (int fy = 0; fy < K; ++fy) {
            const float vv = __ldg(pv + fy * plane);
            const float* row = pin + (int64_t)fy * IW;
            for (int fx = 0; fx < K; ++fx) {
                const float t = __fmul_rn(__ldg(row + fx), vv);
                acc = __fmaf_rn(t, __ldg(ph + fx * plane), acc);
            }
        }
    } else {
        for (int fy = 0; fy < K; ++fy) {
            const float* row = pin + (int64_t)fy * IW;
            float r = 0.f;
            for (int fx = 0; fx < K; ++fx) r = fmaf(__ldg(row + fx), __ldg(ph + fx * plane), r);
            acc = fmaf(r, __ldg(pv + fy * plane), acc);
        }
    }
    out[idx] = acc;
}

// one thread per (b, f, y, x): writes gv[b,f,y,x] and gh[b,f,y,x]
__global__ void __launch_bounds__(256)
sepconv_bwd_taps_generic_kernel(const float* __restrict__ g, const float* __restrict__ in,
                                const float* __restrict__ v, const float* __restrict__ h,
                                float* __restrict__ gv, float* __restrict__ gh,
                                int64_t total, int C, int H, int W, int K) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int x = (int)(idx % W);
    const int y = (int)((idx / W) % H);
    const int f = (int)((idx / ((int64_t)W * H)) % K);
    const int64_t b = idx / ((int64_t)W * H * K);
    const int IW = W + K - 1, IH = H + K - 1;
    const int64_t plane = (int64_t)H * W;
    const float* pv = v + b * K * plane + (int64_t)y * W + x;
    const float* ph = h + b * K * plane + (int64_t)y * W + x;
    float av = 0.f, ah = 0.f;
    for (int c = 0; c < C; ++c) {
        const float gc = __ldg(g + (b * C + c) * plane + (int64_t)y * W + x);
        const float* pin = in + ((b * C + c) * IH + y) * (int64_t)IW + x;
        float rv = 0.f, rh = 0.f;
        if (gv) {
            const float* row = pin + (int64_t)f * IW;           // fy = f, sum over fx
            for (int fx = 0; fx < K; ++fx) rv = fmaf(__ldg(row + fx), __ldg(ph + fx * plane), rv);
        }
        if (gh) {
            const float* col = pin + f;                          // fx = f, sum over fy
            for (int fy = 0; fy < K; ++fy) rh = fmaf(__ldg(col + (int64_t)fy * IW), __ldg(pv + fy * plane), rh);
        }
        av = fmaf(gc, rv, av);
        ah = fmaf(gc, rh, ah);
    }
    if (gv) gv[idx] = av;
    if (gh) gh[idx] = ah;
}

// one thread per grad_input element (gather form of the adjoint)
__global__ void __launch_bounds__(256)
sepconv_bwd_input_generic_kernel(const float* __restrict__ g, const float* __restrict__ v,
                                 const float* __restrict__ h, float* __restrict__ gi,
                                 int64_t total, int C, int H, int W, int K) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int IW = W + K - 1, IH = H + K - 1;
    const int X = (int)(idx % IW);
    const int Y = (int)((idx / IW) % IH);
    const int c = (int)((idx / ((int64_t)IW * IH)) % C);
    const int64_t b = idx / ((int64_t)IW * IH * C);
    const int64_t plane = (int64_t)H * W;
    const float* pg = g + (b * C + c) * plane;
    const float* pv = v + b * K * plane;
    const float* ph = h + b * K * plane;
    const int fy_lo = max(0, Y - (H - 1)), fy_hi = min(K - 1, Y);
    const int fx_lo = max(0, X - (W - 1)), fx_hi = min(K - 1, X);
    float acc = 0.f;
    for (int fy = fy_lo; fy <= fy_hi; ++fy) {
        const int y = Y - fy;
        float r = 0.f;
        for (int fx = fx_lo; fx <= fx_hi; ++fx) {
            const int64_t o = (int64_t)y * W + (X - fx);
            const float a = __ldg(pg + o) * __ldg(pv + fy * plane + o);
            r = fmaf(a, __ldg(ph + fx * plane + o), r);
        }
        acc += r;
    }
    gi[idx] = acc;
}

static inline unsigned blocks_for(int64_t total, int threads) { return (unsigned)((total + threads - 1) / threads); }

int launch_sepconv_fwd_generic(const float* in, const float* v, const float* h, float* out,
                               int64_t B, int64_t C, int64_t H, int64_t W, int K, bool strict,
                               cudaStream_t s) {
    const int64_t total = B * C * H * W;
    if (strict)
        sepconv_fwd_generic_kernel<true><<<blocks_for(total, 256), 256, 0, s>>>(in, v, h, out, total, (int)C, (int)H, (int)W, K);
    else
        sepconv_fwd_generic_kernel<false><<<blocks_for(total, 256), 256, 0, s>>>(in, v, h, out, total, (int)C, (int)H, (int)W, K);
    count_launch();
    return finish_launch();
}

int launch_sepconv_bwd_taps_generic(const float* g, const float* in, const float* v, const float* h,
                                    float* gv, float* gh,
                                    int64_t B, int64_t C, int64_t H, int64_t W, int K, cudaStream_t s) {
    const int64_t total = B * K * H * W;
    sepconv_bwd_taps_generic_kernel<<<blocks_for(total, 256), 256, 0, s>>>(g, in, v, h, gv, gh, total, (int)C, (int)H, (int)W, K);
    count_launch();
    return finish_launch();
}

int launch_sepconv_bwd_input_generic(const float* g, const float* v, const float* h, float* gi,
                                     int64_t B, int64_t C, int64_t H, int64_t W, int K, cudaStream_t s) {
    const int64_t total = B * C * (H + K - 1) * (W + K - 1);
    sepconv_bwd_input_generic_kernel<<<blocks_for(total, 256), 256, 0, s>>>(g, v, h, gi, total, (int)C, (int)H, (int)W, K);
    count_launch();
    return finish_launch();
}

}  // namespace sstem
