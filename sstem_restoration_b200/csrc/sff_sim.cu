// GPU-resident support-film-fold (SFF) simulation for sm_100a -- SURVEY.md section 8f, N3.
//
//  * sff_degrade_kernel: one pass for simu_sff/simuSFF.py:113-121 --
//        flow, mask = gen_flow(h, w, k, b, line_width, fold_width, dis_k)   (flow_synthesis.py:27-83)
//        deformed   = image_warp(img, flow, mode='bilinear')                (image_warp.py:3-111)
//        deformed   = (deformed * mask).astype(np.uint8)
//        count      = number of zero pixels (the caller's accept test, simuSFF.py:125-130)
//    The fold-line displacement is evaluated per pixel in FP64 with numpy's operation order and
//    without FMA contraction (__dmul_rn / __dadd_rn ...), rounded to float32 exactly where numpy
//    stores into its float32 flow array, then fed to the numpy-semantics bilinear gather; the
//    result is bit-equal to the reference's CPU run.  The flow never has to exist in memory
//    (it is written only when the caller wants it): 1 B read (gather, L1/L2-served) + 1 B written
//    per pixel instead of the ~20 full-size float64 temporaries of the numpy path.
//  * sff_contrast_kernel: the regional-contrast step of simuSFF.py:134-144 (`noise`), in place,
//    touching only the box; the image mean comes from the pixel sum the degrade pass left in `stats`.
//
// Byte / integer work bound by HBM (and by the FP64 pipe for the distance field): no tensor cores.
#include "common.cuh"

namespace sstem {
namespace {

struct FoldLine {                                       // one row of the params array (8 doubles)
    double k, b, norm, line_width, fold_width, dis_k, sin_p, cos_p;
};

// flow_synthesis.py:32-82 for pixel (row i, column j); returns the float32 flow and the 0/1 mask
// (fx2, fy2): the second flow of the data providers' gen_flow variant
// (sff_scripts_unfolding/utils/flow_synthesis.py:44-61 -- displacement kept beyond fold_width, opposite sign)
template <bool FLOW2>
__device__ __forceinline__ void fold_flow_at(const FoldLine& p, int i, int j, float& fx, float& fy, bool& mask,
                                             float& fx2, float& fy2) {
    // dis = (k * pos_x - pos_y + b) / sqrt(k**2 + 1)
    const double dis = __ddiv_rn(__dadd_rn(__dsub_rn(__dmul_rn(p.k, (double)j), (double)i), p.b), p.norm);
    const double sign = dis > 0.0 ? 1.0 : (dis < 0.0 ? -1.0 : 0.0);
    const double dis_abs = fabs(dis);
    mask = dis_abs > p.line_width;                      // :41-42
    const bool outside = !(dis_abs < p.line_width);     // mask_dis, :50-52
    const double dk = -p.dis_k;                         // :56
    const double dis_b = __dsub_rn(__dsub_rn(p.fold_width, p.line_width), __dmul_rn(dk, p.line_width));   // :49,57
    double s = __dadd_rn(__dmul_rn(dk, dis_abs), dis_b);                                                   // :58
    if (s < 0.0) s = 0.0;                               // :59
    if (FLOW2) {
        // unfolding flow_synthesis.py:48-49,58,62: s * mask_dis2 + dis_abs * (1 - mask_dis2), times (-sign)
        const double s2 = !(dis_abs < p.fold_width) ? __dadd_rn(s, 0.0) : dis_abs;
        const double d2 = __dmul_rn(s2, -sign);
        const double dc2 = __dmul_rn(d2, p.cos_p), ds2 = __dmul_rn(d2, p.sin_p);
        if (p.k > 0.0) { fx2 = __double2float_rn(dc2); fy2 = __double2float_rn(-ds2); }
        else { fx2 = __double2float_rn(-dc2); fy2 = __double2float_rn(ds2); }
    }
    // :60  s * mask_dis + dis_abs * (1 - mask_dis): one of the two products is exactly 0
    s = outside ? __dadd_rn(s, 0.0) : dis_abs;
    const double d = __dmul_rn(s, sign);                // :62
    const double dc = __dmul_rn(d, p.cos_p), ds = __dmul_rn(d, p.sin_p);
    if (p.k > 0.0) { fx = __double2float_rn(dc); fy = __double2float_rn(-ds); }      // :74-76
    else { fx = __double2float_rn(-dc); fy = __double2float_rn(ds); }                // :77-79
}

// numpy image_warp, bilinear, one uint8 plane (image_warp.py:35-110); same arithmetic as image_warp_kernel
__device__ __forceinline__ uint8_t numpy_warp_u8(const uint8_t* __restrict__ im, float fx, float fy, int i, int j, int H, int W) {
    const float ffx = floorf(fx), ffy = floorf(fy);
    const float lim = 1073741824.0f;
    const int dxi = (int)fminf(fmaxf(ffx, -lim), lim), dyi = (int)fminf(fmaxf(ffy, -lim), lim);
    const int x0 = min(max(j + dxi, 0), W - 1), y0 = min(max(i + dyi, 0), H - 1);
    const int x1 = min(x0 + 1, W - 1), y1 = min(y0 + 1, H - 1);          // from the CLIPPED x0 / y0 (:84-88)
    const float xw = __fsub_rn(fx, ffx), yw = __fsub_rn(fy, ffy);
    const float ax = __fsub_rn(1.0f, xw), ay = __fsub_rn(1.0f, yw);
    const float wa = __fmul_rn(ax, ay), wb = __fmul_rn(ax, yw), wc = __fmul_rn(xw, ay), wd = __fmul_rn(xw, yw);
    const float Ia = (float)__ldg(im + (int64_t)y0 * W + x0), Ib = (float)__ldg(im + (int64_t)y1 * W + x0);
    const float Ic = (float)__ldg(im + (int64_t)y0 * W + x1), Id = (float)__ldg(im + (int64_t)y1 * W + x1);
    float r = __fadd_rn(__fmul_rn(wa, Ia), __fmul_rn(wb, Ib));
    r = __fadd_rn(r, __fmul_rn(wc, Ic));
    r = __fadd_rn(r, __fmul_rn(wd, Id));
    return (uint8_t)(int)r;                             // :110 astype(uint8)
}

constexpr int DG_TX = 64, DG_TY = 4, DG_PX = 4;         // block = 64 x 4 threads, 4 pixels along x per thread

// `border`: pixels closer than this to the image border are left out of the statistics (the data
// providers count zeros on the centre crop, data_provider.py:231-238); 0 for simuSFF.
template <bool FLOW2>
__global__ void __launch_bounds__(DG_TX * DG_TY)
sff_degrade_kernel(const uint8_t* __restrict__ img, const double* __restrict__ params,
                   uint8_t* __restrict__ out, float* __restrict__ flow_out, float* __restrict__ flow2_out,
                   uint8_t* __restrict__ mask_out, unsigned long long* __restrict__ stats, int H, int W, int border) {
    const int b = blockIdx.z;
    const int i = blockIdx.y * DG_TY + threadIdx.y;
    const int j0 = (blockIdx.x * DG_TX + threadIdx.x) * DG_PX;
    const double* pp = params + 8 * (int64_t)b;
    const FoldLine p = {pp[0], pp[1], pp[2], pp[3], pp[4], pp[5], pp[6], pp[7]};
    const int64_t plane = (int64_t)H * W;
    const uint8_t* im = img + b * plane;
    unsigned zeros = 0, sum = 0;
    if (i < H && j0 < W) {
        uint8_t res[DG_PX], msk[DG_PX];
        float fx[DG_PX], fy[DG_PX], fx2[DG_PX], fy2[DG_PX];
        const bool row_counts = i >= border && i < H - border;
#pragma unroll
        for (int q = 0; q < DG_PX; ++q) {
            const int j = min(j0 + q, W - 1);
            bool m;
            fold_flow_at<FLOW2>(p, i, j, fx[q], fy[q], m, fx2[q], fy2[q]);
            const uint8_t v = numpy_warp_u8(im, fx[q], fy[q], i, j, H, W);
            res[q] = m ? v : (uint8_t)0;                // (deformed * mask).astype(uint8)
            msk[q] = m ? 1 : 0;
            if (j0 + q < W && row_counts && j0 + q >= border && j0 + q < W - border) { zeros += (res[q] == 0); sum += res[q]; }
        }
        const int64_t o = b * plane + (int64_t)i * W + j0;
        const bool full = (j0 + DG_PX <= W);
        if (full && ((o & 3) == 0) && ((reinterpret_cast<uintptr_t>(out) & 3u) == 0)) {
            *reinterpret_cast<uchar4*>(out + o) = make_uchar4(res[0], res[1], res[2], res[3]);
        } else {
            for (int q = 0; q < DG_PX && j0 + q < W; ++q) out[o + q] = res[q];
        }
        if (mask_out) {
            if (full && ((o & 3) == 0) && ((reinterpret_cast<uintptr_t>(mask_out) & 3u) == 0))
                *reinterpret_cast<uchar4*>(mask_out + o) = make_uchar4(msk[0], msk[1], msk[2], msk[3]);
            else
                for (int q = 0; q < DG_PX && j0 + q < W; ++q) mask_out[o + q] = msk[q];
        }
        auto store_flow = [&](float* base, const float (&ax)[DG_PX], const float (&ay)[DG_PX]) {
            float* fo = base + 2 * o;
            if (full && ((reinterpret_cast<uintptr_t>(fo) & 15u) == 0)) {
                __stcs(reinterpret_cast<float4*>(fo), make_float4(ax[0], ay[0], ax[1], ay[1]));
                __stcs(reinterpret_cast<float4*>(fo) + 1, make_float4(ax[2], ay[2], ax[3], ay[3]));
            } else {
                for (int q = 0; q < DG_PX && j0 + q < W; ++q) { fo[2 * q] = ax[q]; fo[2 * q + 1] = ay[q]; }
            }
        };
        if (flow_out) store_flow(flow_out, fx, fy);
        if (FLOW2 && flow2_out) store_flow(flow2_out, fx2, fy2);
    }
    // zero count and pixel sum of image b: warp reduce, then one atomic pair per block
    zeros = __reduce_add_sync(0xffffffffu, zeros);
    sum = __reduce_add_sync(0xffffffffu, sum);
    __shared__ unsigned s_z[DG_TX * DG_TY / 32], s_s[DG_TX * DG_TY / 32];
    const int tid = threadIdx.y * DG_TX + threadIdx.x;
    if ((tid & 31) == 0) { s_z[tid >> 5] = zeros; s_s[tid >> 5] = sum; }
    __syncthreads();
    if (tid == 0) {
        unsigned z = 0, s = 0;
#pragma unroll
        for (int w = 0; w < DG_TX * DG_TY / 32; ++w) { z += s_z[w]; s += s_s[w]; }
        if (z) atomicAdd(stats + 2 * b, (unsigned long long)z);
        if (s) atomicAdd(stats + 2 * b + 1, (unsigned long long)s);
    }
}

// simuSFF.py:134-144: inside the box, p <- uint8(ran * (p - mean) + mean); pixels that were 0 stay 0
__global__ void __launch_bounds__(256)
sff_contrast_kernel(uint8_t* __restrict__ img, const unsigned long long* __restrict__ stats,
                    const double* __restrict__ params, int H, int W) {
    const int b = blockIdx.z;
    const double* pp = params + 8 * (int64_t)b;
    const double ran = pp[0];
    const int r0 = (int)pp[1], c0 = (int)pp[2], bh = (int)pp[3], bw = (int)pp[4];
    const int r = blockIdx.y * 8 + (threadIdx.x >> 5), c = blockIdx.x * 32 + (threadIdx.x & 31);
    if (r >= bh || c >= bw) return;
    const int y = r0 + r, x = c0 + c;
    if (y >= H || x >= W) return;                       // numpy slicing clips the box at the border
    const double mean = __ddiv_rn((double)stats[2 * b + 1], (double)((int64_t)H * W));   // np.mean of a uint8 image
    uint8_t* px = img + ((int64_t)b * H + y) * W + x;
    const uint8_t p = *px;
    if (p == 0) return;                                 // mask[img == 0] = 0 ... np.multiply(img, mask)
    const double v = __dadd_rn(__dmul_rn(ran, __dsub_rn((double)p, mean)), mean);
    *px = (uint8_t)(int)v;                              // float64 -> uint8 store truncates
}

}  // namespace
}  // namespace sstem

using namespace sstem;

extern "C" int sstem_sff_degrade(const uint8_t* img, const double* params, uint8_t* out, float* flow_out,
                                 float* flow2_out, uint8_t* mask_out, int64_t* stats, int64_t B, int64_t H, int64_t W,
                                 int64_t stats_border, void* stream) {
    if (!img || !params || !out || !stats) return SSTEM_E_NULL;
    if (B <= 0 || H <= 0 || W <= 0 || B > 65535 || H * W > INT32_MAX) return SSTEM_E_SHAPE;
    if (stats_border < 0 || 2 * stats_border >= H || 2 * stats_border >= W) return SSTEM_E_SHAPE;
    if ((reinterpret_cast<uintptr_t>(params) & 7u) || (reinterpret_cast<uintptr_t>(stats) & 7u)) return SSTEM_E_ALIGN;
    if ((flow_out && !aligned4(flow_out)) || (flow2_out && !aligned4(flow2_out))) return SSTEM_E_ALIGN;
    DeviceGuard guard(out);
    if (guard.err) return guard.err;
    cudaStream_t s = (cudaStream_t)stream;
    cudaError_t e = cudaMemsetAsync(stats, 0, (size_t)B * 2 * sizeof(int64_t), s);
    if (e != cudaSuccess) return (int)e;
    dim3 block(DG_TX, DG_TY);
    dim3 grid((unsigned)((W + DG_TX * DG_PX - 1) / (DG_TX * DG_PX)), (unsigned)((H + DG_TY - 1) / DG_TY), (unsigned)B);
    if (grid.y > 65535) return SSTEM_E_SHAPE;
    if (flow2_out)
        sff_degrade_kernel<true><<<grid, block, 0, s>>>(img, params, out, flow_out, flow2_out, mask_out,
                                                        reinterpret_cast<unsigned long long*>(stats), (int)H, (int)W, (int)stats_border);
    else
        sff_degrade_kernel<false><<<grid, block, 0, s>>>(img, params, out, flow_out, flow2_out, mask_out,
                                                         reinterpret_cast<unsigned long long*>(stats), (int)H, (int)W, (int)stats_border);
    count_launch();
    return finish_launch();
}

extern "C" int sstem_sff_contrast(uint8_t* img, const int64_t* stats, const double* params,
                                  int64_t B, int64_t H, int64_t W, int64_t max_box_h, int64_t max_box_w, void* stream) {
    if (!img || !params || !stats) return SSTEM_E_NULL;
    if (B <= 0 || H <= 0 || W <= 0 || B > 65535 || max_box_h <= 0 || max_box_w <= 0) return SSTEM_E_SHAPE;
    if ((reinterpret_cast<uintptr_t>(params) & 7u) || (reinterpret_cast<uintptr_t>(stats) & 7u)) return SSTEM_E_ALIGN;
    DeviceGuard guard(img);
    if (guard.err) return guard.err;
    dim3 grid((unsigned)((max_box_w + 31) / 32), (unsigned)((max_box_h + 7) / 8), (unsigned)B);
    if (grid.y > 65535) return SSTEM_E_SHAPE;
    sff_contrast_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(img, reinterpret_cast<const unsigned long long*>(stats), params,
                                                                (int)H, (int)W);
    count_launch();
    return finish_launch();
}
