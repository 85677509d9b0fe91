// Stack pre/post-processing either side of the interpolation path, on the device, so only uint8
// sections cross PCIe -- SURVEY.md section 8f, N4.
//
//  * sections_to_input_kernel: sff_scripts_interp/inference.py:69-83 --
//        img1 = np.repeat(section[k-1][None], 3, 0); img2 likewise for k+1
//        inputs = np.concatenate([img1, img2], 0)[None].astype(np.float32) / 255.0
//        inputs = F.pad(inputs, (PAD, PAD, PAD, PAD))                  (zeros)
//    uint8 [B,H,W] x 2  ->  float32 [B,6,H+2P,W+2P]: 2 B read, 24 B written per pixel (HBM-bound
//    streaming write; the host path uploads those 24 B per pixel over PCIe instead).
//  * prediction_to_u8_kernel: inference.py:84-88 --
//        pred = F.pad(pred, (-PAD, -PAD, -PAD, -PAD)); (pred * 255).astype(np.uint8)
//    float32 [B,1,H+2P,W+2P] -> uint8 [B,H,W].
// Both are bit-equal to the numpy expressions (IEEE division by 255, float32 product, C truncation).
#include "common.cuh"

namespace sstem {
namespace {

// one thread = 4 consecutive columns of one padded output row, all six channel planes
__global__ void __launch_bounds__(256)
sections_to_input_kernel(const uint8_t* __restrict__ sec_a, const uint8_t* __restrict__ sec_b,
                         float* __restrict__ out, int H, int W, int pad) {
    const int OW = W + 2 * pad, OH = H + 2 * pad;
    const int x0 = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
    const int y = blockIdx.y;
    const int64_t b = blockIdx.z;
    if (x0 >= OW) return;
    const int sy = y - pad;
    float va[4], vb[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const int sx = x0 + q - pad;
        const bool in = (unsigned)sy < (unsigned)H && (unsigned)sx < (unsigned)W;
        const int64_t o = (b * H + (in ? sy : 0)) * (int64_t)W + (in ? sx : 0);
        va[q] = in ? __fdiv_rn((float)__ldg(sec_a + o), 255.0f) : 0.f;
        vb[q] = (in && sec_b) ? __fdiv_rn((float)__ldg(sec_b + o), 255.0f) : 0.f;
    }
    const int64_t plane = (int64_t)OH * OW;
    const int nch = sec_b ? 6 : 3;                      // one section only: [B,3,..] (the correction module's input_sff)
    float* row = out + b * nch * plane + (int64_t)y * OW + x0;
    const bool vec = (x0 + 4 <= OW) && ((reinterpret_cast<uintptr_t>(row) & 15u) == 0) && ((plane & 3) == 0);
#pragma unroll
    for (int c = 0; c < 6; ++c) {
        if (c >= nch) break;
        const float* v = c < 3 ? va : vb;               // channels 0-2: section k-1, 3-5: section k+1
        float* dst = row + c * plane;
        if (vec) __stcs(reinterpret_cast<float4*>(dst), make_float4(v[0], v[1], v[2], v[3]));
        else
            for (int q = 0; q < 4 && x0 + q < OW; ++q) dst[q] = v[q];
    }
}

__global__ void __launch_bounds__(256)
prediction_to_u8_kernel(const float* __restrict__ pred, uint8_t* __restrict__ out, int H, int W, int pad) {
    const int OW = W + 2 * pad, OH = H + 2 * pad;
    const int x0 = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
    const int y = blockIdx.y;
    const int64_t b = blockIdx.z;
    if (x0 >= W) return;
    const float* src = pred + (b * OH + y + pad) * (int64_t)OW + pad + x0;
    uint8_t r[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const float v = __fmul_rn(__ldcs(src + min(q, W - 1 - x0)), 255.0f);
        r[q] = (uint8_t)(int)v;                         // astype(np.uint8): C truncation (through int32)
    }
    uint8_t* dst = out + (b * H + y) * (int64_t)W + x0;
    if (x0 + 4 <= W && (reinterpret_cast<uintptr_t>(dst) & 3u) == 0) *reinterpret_cast<uchar4*>(dst) = make_uchar4(r[0], r[1], r[2], r[3]);
    else
        for (int q = 0; q < 4 && x0 + q < W; ++q) dst[q] = r[q];
}

// sff_scripts_fusion/inference.py:163-171 -- the correction module's output assembly:
//     warped_sff = (warped * 255).astype(np.uint8)                    [C,H,W] -> transpose -> PIL convert('L')
//     mask = warped_sff >= 2;  stitch = (img_interp * (1 - mask) + warped_sff * mask).astype(np.uint8)
// PIL's 'L' conversion of an RGB triple is (19595 R + 38470 G + 7471 B + 0x8000) >> 16 (exactly R for gray x3).
// One thread = 4 consecutive pixels.  Outputs: the gray warped section and the stitched section, uint8 [B,H,W].
template <int CT>
__global__ void __launch_bounds__(256)
warp_stitch_u8_kernel(const float* __restrict__ warped, const uint8_t* __restrict__ interp,
                      uint8_t* __restrict__ gray_out, uint8_t* __restrict__ stitch_out, int64_t plane, int64_t total4) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total4) return;
    const int64_t pix = 4 * i, b = pix / plane, p = pix - b * plane;   // plane % 4 == 0 (checked by the launcher)
    const float* src = warped + b * CT * plane + p;
    unsigned ch[CT][4];
#pragma unroll
    for (int c = 0; c < CT; ++c) {
        const float4 v = __ldcs(reinterpret_cast<const float4*>(src + c * plane));
        const float f[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int q = 0; q < 4; ++q) ch[c][q] = (unsigned)(uint8_t)(int)__fmul_rn(f[q], 255.0f);   // astype(np.uint8)
    }
    const uchar4 it = *reinterpret_cast<const uchar4*>(interp + pix);
    const uint8_t iv[4] = {it.x, it.y, it.z, it.w};
    uint8_t L[4], st[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const unsigned l = CT == 3 ? (19595u * ch[0][q] + 38470u * ch[1 % CT][q] + 7471u * ch[2 % CT][q] + 0x8000u) >> 16 : ch[0][q];
        L[q] = (uint8_t)l;
        st[q] = l >= 2u ? (uint8_t)l : iv[q];
    }
    if (gray_out) *reinterpret_cast<uchar4*>(gray_out + pix) = make_uchar4(L[0], L[1], L[2], L[3]);
    *reinterpret_cast<uchar4*>(stitch_out + pix) = make_uchar4(st[0], st[1], st[2], st[3]);
}

}  // namespace
}  // namespace sstem

using namespace sstem;

extern "C" int sstem_warp_stitch_u8(const float* warped, const uint8_t* interp, uint8_t* gray_out, uint8_t* stitch_out,
                                    int64_t B, int64_t C, int64_t H, int64_t W, void* stream) {
    if (!warped || !interp || !stitch_out) return SSTEM_E_NULL;
    if (B <= 0 || H <= 0 || W <= 0 || (C != 1 && C != 3)) return SSTEM_E_SHAPE;
    if ((H * W) & 3) return SSTEM_E_SHAPE;             // 4 pixels per thread, never straddling an image
    if (!aligned16(warped) || !aligned4(interp) || !aligned4(stitch_out) || !aligned4(gray_out)) return SSTEM_E_ALIGN;
    DeviceGuard guard(stitch_out);
    if (guard.err) return guard.err;
    const int64_t plane = H * W, total4 = B * plane / 4;
    const unsigned blocks = (unsigned)((total4 + 255) / 256);
    if (C == 3) warp_stitch_u8_kernel<3><<<blocks, 256, 0, (cudaStream_t)stream>>>(warped, interp, gray_out, stitch_out, plane, total4);
    else warp_stitch_u8_kernel<1><<<blocks, 256, 0, (cudaStream_t)stream>>>(warped, interp, gray_out, stitch_out, plane, total4);
    count_launch();
    return finish_launch();
}

extern "C" int sstem_sections_to_input(const uint8_t* section_prev, const uint8_t* section_next, float* inputs,
                                       int64_t B, int64_t H, int64_t W, int32_t pad, void* stream) {
    if (!section_prev || !inputs) return SSTEM_E_NULL;   // section_next == NULL: one section -> [B,3,H+2P,W+2P]
    if (B <= 0 || H <= 0 || W <= 0 || pad < 0 || B > 65535 || H + 2 * (int64_t)pad > 65535 || W + 2 * (int64_t)pad > INT32_MAX / 2)
        return SSTEM_E_SHAPE;
    if (!aligned4(inputs)) return SSTEM_E_ALIGN;
    DeviceGuard guard(inputs);
    if (guard.err) return guard.err;
    const int64_t OW = W + 2 * pad, OH = H + 2 * pad;
    dim3 grid((unsigned)((OW + 1023) / 1024), (unsigned)OH, (unsigned)B);
    sections_to_input_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(section_prev, section_next, inputs, (int)H, (int)W, (int)pad);
    count_launch();
    return finish_launch();
}

// mean over the channel planes + ReplicationPad2d(pad): the one-plane frame the tile-major interpolation tail convolves
// (mean_c sepconv(i_c) = sepconv(mean_c i_c); model_interp.py:46, 90-97).  One thread = one padded pixel.
namespace sstem {
namespace {
__global__ void __launch_bounds__(256)
frame_mean_pad_kernel(const float* __restrict__ frame, int64_t bstride, float* __restrict__ out, int nplanes, float scale,
                      int H, int W, int pad) {
    const int OW = W + 2 * pad, OH = H + 2 * pad;
    const int X = blockIdx.x * blockDim.x + threadIdx.x, Y = blockIdx.y;
    const int64_t b = blockIdx.z;
    if (X >= OW) return;
    const int sy = min(max(Y - pad, 0), H - 1), sx = min(max(X - pad, 0), W - 1);
    const float* p = frame + b * bstride + (int64_t)sy * W + sx;
    float acc = __ldg(p);
    for (int c = 1; c < nplanes; ++c) acc += __ldg(p + (int64_t)c * H * W);
    out[(b * OH + Y) * (int64_t)OW + X] = nplanes > 1 ? acc * scale : acc;
}
}  // namespace
}  // namespace sstem

extern "C" int sstem_frame_mean_pad(const float* frame, int64_t frame_bstride, float* out, int64_t B, int64_t C, int64_t H,
                                    int64_t W, int32_t pad, uint32_t flags, void* stream) {
    if (!frame || !out) return SSTEM_E_NULL;
    if (B <= 0 || C <= 0 || H <= 0 || W <= 0 || pad < 0 || B > 65535 || H + 2 * (int64_t)pad > 65535 || W + 2 * (int64_t)pad > INT32_MAX / 2 ||
        frame_bstride < C * H * W)
        return SSTEM_E_SHAPE;
    if (flags & ~SSTEM_SEPCONV_GRAY_REPLICATED) return SSTEM_E_FLAG;
    if (!aligned4(frame) || !aligned4(out)) return SSTEM_E_ALIGN;
    DeviceGuard guard(out);
    if (guard.err) return guard.err;
    const int nplanes = (flags & SSTEM_SEPCONV_GRAY_REPLICATED) ? 1 : (int)C;
    dim3 grid((unsigned)((W + 2 * pad + 255) / 256), (unsigned)(H + 2 * pad), (unsigned)B);
    frame_mean_pad_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(frame, frame_bstride, out, nplanes, 1.0f / (float)nplanes, (int)H, (int)W, (int)pad);
    count_launch();
    return finish_launch();
}

extern "C" int sstem_prediction_to_u8(const float* pred, uint8_t* section, int64_t B, int64_t H, int64_t W, int32_t pad,
                                      void* stream) {
    if (!pred || !section) return SSTEM_E_NULL;
    if (B <= 0 || H <= 0 || W <= 0 || pad < 0 || B > 65535 || H > 65535 || W + 2 * (int64_t)pad > INT32_MAX / 2) return SSTEM_E_SHAPE;
    if (!aligned4(pred)) return SSTEM_E_ALIGN;
    DeviceGuard guard(section);
    if (guard.err) return guard.err;
    dim3 grid((unsigned)((W + 1023) / 1024), (unsigned)H, (unsigned)B);
    prediction_to_u8_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(pred, section, (int)H, (int)W, (int)pad);
    count_launch();
    return finish_launch();
}
