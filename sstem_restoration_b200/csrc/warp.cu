// Flow-driven bilinear backward warps for sm_100a.
//
//  * warp_torch_kernel  -- zero-padded warp, bit-compatible with
//    SpatialTransformation.forward (sff_scripts_unfolding/utils/image_warp_torch.py:97-113).
//  * image_warp_kernel  -- clamp-border warp with numpy image_warp semantics
//    (simu_sff/image_warp.py:3-111), NHWC, uint8 truncation.
//
// Both are HBM-bound gathers: per output pixel 8 B of flow, 4*C B of image
// (each source pixel is touched by ~4 neighbouring outputs, served by L1/L2) and
// 4*C B of output.  One thread owns PX horizontally adjacent pixels so flow
// loads and output stores are 16-byte vectors when the row is aligned; all
// arithmetic uses the non-contracting intrinsics (__fadd_rn/__fmul_rn/__fsub_rn)
// in the reference's operation order, so results are bit-equal to the CPU run of
// the reference, not merely close.
#include "common.cuh"

namespace sstem {

struct BilinearTap {
    int x0, x1, y0, y1;  // coordinates in the 1-px zero-padded image, clamped to it
    float wa, wb, wc, wd;
};

// image_warp_torch.py:43-57,82-91: x = (fx + j) + 1, floor, +1, clamp, weights from the clamped x1/y1
__device__ __forceinline__ BilinearTap torch_tap(float fx, float fy, int i, int j, int H, int W) {
    BilinearTap t;
    const float x = __fadd_rn(__fadd_rn(fx, (float)j), 1.0f);
    const float y = __fadd_rn(__fadd_rn(fy, (float)i), 1.0f);
    const float xf = floorf(x), yf = floorf(y);
    const float mx = (float)(W + 1), my = (float)(H + 1);
    // clamp(floor(x)) and clamp(floor(x) + 1) -- done in float so huge |flow| cannot overflow
    const float x0c = fminf(fmaxf(xf, 0.f), mx), x1c = fminf(fmaxf(__fadd_rn(xf, 1.0f), 0.f), mx);
    const float y0c = fminf(fmaxf(yf, 0.f), my), y1c = fminf(fmaxf(__fadd_rn(yf, 1.0f), 0.f), my);
    t.x0 = (int)x0c; t.x1 = (int)x1c; t.y0 = (int)y0c; t.y1 = (int)y1c;
    const float dx = __fsub_rn(x1c, x), dy = __fsub_rn(y1c, y);
    const float ex = __fsub_rn(1.0f, dx), ey = __fsub_rn(1.0f, dy);
    t.wa = __fmul_rn(dx, dy);
    t.wb = __fmul_rn(dx, ey);
    t.wc = __fmul_rn(ex, dy);
    t.wd = __fmul_rn(ex, ey);
    return t;
}

// sample of the zero-padded image at padded coordinates (py, px)
__device__ __forceinline__ float padded_at(const float* __restrict__ im, int py, int px, int H, int W) {
    const int yy = py - 1, xx = px - 1;
    return ((unsigned)yy < (unsigned)H && (unsigned)xx < (unsigned)W) ? __ldg(im + (int64_t)yy * W + xx) : 0.f;
}

template <int PX, bool NHWC>
__global__ void __launch_bounds__(256)
warp_torch_kernel(const float* __restrict__ moving, const float* __restrict__ flow,
                  int64_t fs_b, int64_t fs_h, int64_t fs_w, int64_t fs_c,
                  float* __restrict__ out, int C, int H, int W, int groups_per_row, int64_t total_groups) {
    const int64_t gidx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (gidx >= total_groups) return;
    const int gx = (int)(gidx % groups_per_row);
    const int i = (int)((gidx / groups_per_row) % H);
    const int64_t b = gidx / ((int64_t)groups_per_row * H);
    const int j0 = gx * PX;
    const int64_t plane = (int64_t)H * W;

    BilinearTap tap[PX];
    const float* frow = flow + b * fs_b + (int64_t)i * fs_h;
    const bool full = (j0 + PX <= W);
    // flow: vector loads when the two components are separate contiguous planes
    // (the layout at every reference call site) or interleaved pairs
    float fxv[PX], fyv[PX];
    if (PX == 4 && full && fs_w == 1 && (((uintptr_t)(frow + j0) | (uintptr_t)(frow + fs_c + j0)) & 15u) == 0) {
        const float4 a = __ldcs(reinterpret_cast<const float4*>(frow + j0));
        const float4 c = __ldcs(reinterpret_cast<const float4*>(frow + fs_c + j0));
        fxv[0] = a.x; fxv[1 % PX] = a.y; fxv[2 % PX] = a.z; fxv[3 % PX] = a.w;
        fyv[0] = c.x; fyv[1 % PX] = c.y; fyv[2 % PX] = c.z; fyv[3 % PX] = c.w;
    } else if (PX == 4 && full && fs_w == 2 && fs_c == 1 && ((uintptr_t)(frow + 2 * j0) & 15u) == 0) {
        const float4 a = __ldcs(reinterpret_cast<const float4*>(frow + 2 * j0));
        const float4 c = __ldcs(reinterpret_cast<const float4*>(frow + 2 * j0 + 4));
        fxv[0] = a.x; fyv[0] = a.y; fxv[1 % PX] = a.z; fyv[1 % PX] = a.w;
        fxv[2 % PX] = c.x; fyv[2 % PX] = c.y; fxv[3 % PX] = c.z; fyv[3 % PX] = c.w;
    } else {
#pragma unroll
        for (int p = 0; p < PX; ++p) {
            const int j = min(j0 + p, W - 1);
            fxv[p] = __ldg(frow + (int64_t)j * fs_w);
            fyv[p] = __ldg(frow + (int64_t)j * fs_w + fs_c);
        }
    }
#pragma unroll
    for (int p = 0; p < PX; ++p) tap[p] = torch_tap(fxv[p], fyv[p], i, j0 + p, H, W);

    for (int c = 0; c < C; ++c) {
        const float* im = moving + (b * C + c) * plane;
        float res[PX];
#pragma unroll
        for (int p = 0; p < PX; ++p) {
            const BilinearTap& t = tap[p];
            const float Ia = padded_at(im, t.y0, t.x0, H, W);
            const float Ib = padded_at(im, t.y1, t.x0, H, W);
            const float Ic = padded_at(im, t.y0, t.x1, H, W);
            const float Id = padded_at(im, t.y1, t.x1, H, W);
            // image_warp_torch.py:93: sum over the stacked [wa*Ia, wb*Ib, wc*Ic, wd*Id]
            float r = __fadd_rn(__fmul_rn(t.wa, Ia), __fmul_rn(t.wb, Ib));
            r = __fadd_rn(r, __fmul_rn(t.wc, Ic));
            r = __fadd_rn(r, __fmul_rn(t.wd, Id));
            res[p] = r;
        }
        if (NHWC) {
#pragma unroll
            for (int p = 0; p < PX; ++p)
                if (j0 + p < W) out[((b * H + i) * (int64_t)W + j0 + p) * C + c] = res[p];
        } else {
            float* orow = out + (b * C + c) * plane + (int64_t)i * W + j0;
            if (PX == 4 && full && ((uintptr_t)orow & 15u) == 0) {
                __stcs(reinterpret_cast<float4*>(orow), make_float4(res[0], res[1 % PX], res[2 % PX], res[3 % PX]));
            } else {
#pragma unroll
                for (int p = 0; p < PX; ++p)
                    if (j0 + p < W) orow[p] = res[p];
            }
        }
    }
}

// ---- numpy image_warp semantics -------------------------------------------------
template <typename PixT>
__device__ __forceinline__ float pix_load(const PixT* p) { return (float)__ldg(p); }

template <typename PixT, bool NEAREST>
__global__ void __launch_bounds__(256)
image_warp_kernel(const PixT* __restrict__ im, const float* __restrict__ flow,
                  uint8_t* __restrict__ out_u8, float* __restrict__ out_f32,
                  int C, int H, int W, int64_t total) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;  // pixel index over B*H*W
    if (idx >= total) return;
    const int j = (int)(idx % W);
    const int i = (int)((idx / W) % H);
    const int64_t b = idx / ((int64_t)W * H);
    const float2 f = __ldcs(reinterpret_cast<const float2*>(flow) + idx);
    const float ffx = floorf(f.x), ffy = floorf(f.y);
    // image_warp.py:45-57: integer displacement, then clip (bounded first so the add cannot overflow)
    const float lim = 1073741824.0f;
    const int dxi = (int)fminf(fmaxf(ffx, -lim), lim), dyi = (int)fminf(fmaxf(ffy, -lim), lim);
    const int x0 = min(max(j + dxi, 0), W - 1);
    const int y0 = min(max(i + dyi, 0), H - 1);
    const PixT* base = im + b * (int64_t)H * W * C;
    const int64_t o = idx * C;
    if (NEAREST) {
        const PixT* src = base + ((int64_t)y0 * W + x0) * C;
        for (int c = 0; c < C; ++c) {
            const float val = pix_load(src + c);
            if (out_f32) out_f32[o + c] = val;
            if (out_u8) out_u8[o + c] = (uint8_t)(int)val;
        }
        return;
    }
    // :84-88 -- x1 is taken from the CLIPPED x0
    const int x1 = min(x0 + 1, W - 1), y1 = min(y0 + 1, H - 1);
    // :72-82 -- weights from frac(flow), independent of clipping
    const float xw = __fsub_rn(f.x, ffx), yw = __fsub_rn(f.y, ffy);
    const float ax = __fsub_rn(1.0f, xw), ay = __fsub_rn(1.0f, yw);
    const float wa = __fmul_rn(ax, ay), wb = __fmul_rn(ax, yw), wc = __fmul_rn(xw, ay), wd = __fmul_rn(xw, yw);
    const PixT* pa = base + ((int64_t)y0 * W + x0) * C;
    const PixT* pb = base + ((int64_t)y1 * W + x0) * C;
    const PixT* pc = base + ((int64_t)y0 * W + x1) * C;
    const PixT* pd = base + ((int64_t)y1 * W + x1) * C;
    for (int c = 0; c < C; ++c) {
        float r = __fadd_rn(__fmul_rn(wa, pix_load(pa + c)), __fmul_rn(wb, pix_load(pb + c)));
        r = __fadd_rn(r, __fmul_rn(wc, pix_load(pc + c)));
        r = __fadd_rn(r, __fmul_rn(wd, pix_load(pd + c)));
        if (out_f32) out_f32[o + c] = r;
        if (out_u8) out_u8[o + c] = (uint8_t)(int)r;   // :110 astype(uint8): truncation
    }
}

}  // namespace sstem

using namespace sstem;

extern "C" int sstem_warp_forward(const float* moving, const float* flow, const int64_t flow_strides[4],
                                  float* out, int64_t B, int64_t C, int64_t H, int64_t W,
                                  int32_t out_layout, void* stream) {
    if (!moving || !flow || !flow_strides || !out) return SSTEM_E_NULL;
    if (B <= 0 || C <= 0 || H <= 0 || W <= 0 || H > (1 << 24) || W > (1 << 24)) return SSTEM_E_SHAPE;
    if (!aligned4(moving) || !aligned4(flow) || !aligned4(out)) return SSTEM_E_ALIGN;
    if (out_layout != SSTEM_LAYOUT_NCHW && out_layout != SSTEM_LAYOUT_NHWC) return SSTEM_E_FLAG;
    DeviceGuard guard(out);
    if (guard.err) return guard.err;
    cudaStream_t s = (cudaStream_t)stream;
    constexpr int PX = 4;
    const int gpr = (int)((W + PX - 1) / PX);
    const int64_t groups = B * H * gpr;
    const unsigned blocks = (unsigned)((groups + 255) / 256);
    if (out_layout == SSTEM_LAYOUT_NHWC)
        warp_torch_kernel<PX, true><<<blocks, 256, 0, s>>>(moving, flow, flow_strides[0], flow_strides[1], flow_strides[2],
                                                           flow_strides[3], out, (int)C, (int)H, (int)W, gpr, groups);
    else
        warp_torch_kernel<PX, false><<<blocks, 256, 0, s>>>(moving, flow, flow_strides[0], flow_strides[1], flow_strides[2],
                                                            flow_strides[3], out, (int)C, (int)H, (int)W, gpr, groups);
    count_launch();
    return finish_launch();
}

extern "C" int sstem_image_warp(const void* im, int32_t pix_type, const float* flow,
                                uint8_t* out_u8, float* out_f32,
                                int64_t B, int64_t H, int64_t W, int64_t C, int32_t mode, void* stream) {
    if (!im || !flow || (!out_u8 && !out_f32)) return SSTEM_E_NULL;
    if (B <= 0 || C <= 0 || H <= 0 || W <= 0) return SSTEM_E_SHAPE;
    if (pix_type != SSTEM_PIX_U8 && pix_type != SSTEM_PIX_F32) return SSTEM_E_FLAG;
    if (mode != SSTEM_WARP_BILINEAR && mode != SSTEM_WARP_NEAREST) return SSTEM_E_FLAG;
    if ((reinterpret_cast<uintptr_t>(flow) & 7u) != 0) return SSTEM_E_ALIGN;
    if (pix_type == SSTEM_PIX_F32 && !aligned4(im)) return SSTEM_E_ALIGN;
    if (out_f32 && !aligned4(out_f32)) return SSTEM_E_ALIGN;
    DeviceGuard guard(out_u8 ? (const void*)out_u8 : (const void*)out_f32);
    if (guard.err) return guard.err;
    cudaStream_t s = (cudaStream_t)stream;
    const int64_t total = B * H * W;
    const unsigned blocks = (unsigned)((total + 255) / 256);
    const bool nearest = mode == SSTEM_WARP_NEAREST;
    if (pix_type == SSTEM_PIX_U8) {
        if (nearest) image_warp_kernel<uint8_t, true><<<blocks, 256, 0, s>>>((const uint8_t*)im, flow, out_u8, out_f32, (int)C, (int)H, (int)W, total);
        else image_warp_kernel<uint8_t, false><<<blocks, 256, 0, s>>>((const uint8_t*)im, flow, out_u8, out_f32, (int)C, (int)H, (int)W, total);
    } else {
        if (nearest) image_warp_kernel<float, true><<<blocks, 256, 0, s>>>((const float*)im, flow, out_u8, out_f32, (int)C, (int)H, (int)W, total);
        else image_warp_kernel<float, false><<<blocks, 256, 0, s>>>((const float*)im, flow, out_u8, out_f32, (int)C, (int)H, (int)W, total);
    }
    count_launch();
    return finish_launch();
}
