// Launcher of the second-generation tap-gradient kernel (device code in sepconv_k51_v2.cuh): repack the input
// channel-interleaved into a stream-ordered workspace, encode the two tensor maps, launch, release the workspace.
#include "sepconv_k51_v3.cuh"

#include <stdlib.h>

namespace sstem {
namespace {

template <bool WV, bool WH, bool ACCUM>
int launch_v2_kernel(const CUtensorMap& min, const CUtensorMap& mv, const float* g, const float* h, float* gv, float* gh,
                     int64_t B, int C, int c0, int H, int W, cudaStream_t s) {
    static PerDeviceOnce done;
    auto kern = sepconv_bwd_taps_k51_v2_kernel<WV, WH, ACCUM>;
    if (int e = set_smem_once(kern, V2_BWD_SMEM, done)) return e;
    dim3 grid((unsigned)((W + 31) / 32), (unsigned)((H + V2_R - 1) / V2_R), (unsigned)B);
    kern<<<grid, 128, V2_BWD_SMEM, s>>>(min, mv, g, h, gv, gh, C, c0, H, W);
    count_launch();
    return finish_launch();
}

}  // namespace

// Stream-ordered scratch memory (released with cudaFreeAsync on the same stream).  The device's default pool hands
// freed blocks back to the driver at every synchronisation unless told otherwise -- each call then paid a real
// allocation (measured: +1.5 ms per 80 MB) -- so its release threshold is raised once per device.
int workspace_alloc(void** p, size_t bytes, cudaStream_t s) {
    static PerDeviceOnce tuned;
    int dev = 0;
    cudaGetDevice(&dev);
    if (!tuned.test(dev)) {
        cudaMemPool_t pool;
        if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
            uint64_t keep = UINT64_MAX;
            cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
        }
        cudaGetLastError();
        tuned.set(dev);
    }
    if (cudaMallocAsync(p, bytes, s) != cudaSuccess) {
        cudaGetLastError();
        return -1;
    }
    return 0;
}

// Repacks channels c0..c0+2 of in [B,C,IH,IW] into ws [B,IH,IW,4].
int launch_repack_nhwc4(const float* in, float* ws, int64_t B, int C, int c0, int64_t IH, int64_t IW, cudaStream_t s) {
    const int64_t plane = IH * IW, total = B * plane;
    const bool vec = (plane % 4 == 0) && aligned16(in) && ((int64_t)C * plane % 4 == 0) && ((int64_t)c0 * plane % 4 == 0);
    const int64_t work = vec ? total / 4 : total;
    const unsigned blocks = (unsigned)std::min<int64_t>((work + 255) / 256, (int64_t)sm_count() * 16);
    if (vec) repack_nchw3_to_nhwc4_kernel<true><<<blocks, 256, 0, s>>>(in, reinterpret_cast<float4*>(ws), C, c0, plane, total, g_gate.ptr, g_gate.want);
    else repack_nchw3_to_nhwc4_kernel<false><<<blocks, 256, 0, s>>>(in, reinterpret_cast<float4*>(ws), C, c0, plane, total, g_gate.ptr, g_gate.want);
    count_launch();
    return finish_launch();
}

// Tensor maps of the v2 kernels.  false: the shapes break a tensor-map rule (the caller falls back).
bool make_v2_maps(CUtensorMap* min, CUtensorMap* mv, const float* ws, const float* v, int64_t B, int64_t H, int64_t W, int rows) {
    const int64_t IH = H + K51 - 1, IW = W + K51 - 1;
    {   // repacked input [B][IH][IW][4]: box = 4 channels x 84 columns x (rows + 50) rows
        const int64_t dims[4] = {4, IW, IH, B}, strides[4] = {1, 4, 4 * IW, 4 * IW * IH};
        const int box[4] = {4, V2_WIN_W, rows + K51 - 1, 1};
        if (!make_map_f32(min, ws, 4, dims, strides, box)) return false;
    }
    if (v) { // taps [B][51][H][W]: box = 32 columns x rows x (51 + 2 (rows - 1)) planes, requested from plane -(rows - 1)
        const int64_t dims[4] = {W, H, K51, B}, strides[4] = {1, W, H * W, (int64_t)K51 * H * W};
        const int box[4] = {32, rows, K51 + 2 * (rows - 1), 1};
        if (!make_map_f32(mv, v, 4, dims, strides, box)) return false;
    }
    return true;
}

namespace {

template <bool WV, bool WH, bool ACCUM>
int launch_v3_kernel(const CUtensorMap& min, const CUtensorMap& mv, const CUtensorMap& mh, const CUtensorMap& mg,
                     float* gv, float* gh, int* counter, const V3Shape& sh, cudaStream_t s) {
    static PerDeviceOnce done;
    auto kern = sepconv_bwd_taps_k51_v3_kernel<WV, WH, ACCUM>;
    if (int e = set_smem_once(kern, V3_SMEM, done)) return e;
    const int ctas = std::min<int64_t>(2 * (int64_t)sm_count(), (int64_t)sh.ntiles);
    kern<<<ctas, V3_WARPS * 32, V3_SMEM, s>>>(min, mv, mh, mg, gv, gh, counter, sh, g_gate.ptr, g_gate.want);
    count_launch();
    return finish_launch();
}

// tensor maps of the persistent kernel: window groups, vertical-tap groups, horizontal taps, upstream gradient
bool make_v3_maps(CUtensorMap* min, CUtensorMap* mv, CUtensorMap* mh, CUtensorMap* mg, const float* ws, const float* v,
                  const float* h, const float* g, int64_t B, int64_t C, int64_t c0, int64_t H, int64_t W) {
    const int64_t IH = H + K51 - 1, IW = W + K51 - 1, plane = H * W;
    {
        // a repacked row is 4 * IW contiguous floats: the box's inner extent is a whole 960-byte row segment (an inner
        // extent of one pixel = 16 bytes makes the TMA unit issue 240 tiny requests per box: measured 1.5x slower)
        const int64_t dims[3] = {4 * IW, IH, B}, strides[3] = {1, 4 * IW, 4 * IW * IH};
        const int box[3] = {4 * V3_WIN_COLS, V3_GROUP, 1};
        if (!make_map_f32(min, ws, 3, dims, strides, box)) return false;
    }
    const int64_t tdims[4] = {W, H, K51, B}, tstrides[4] = {1, W, plane, (int64_t)K51 * plane};
    if (v) {
        const int box[4] = {V3_COLS, V2_R, V3_GROUP, 1};
        if (!make_map_f32(mv, v, 4, tdims, tstrides, box)) return false;
    }
    if (h) {
        const int box[4] = {V3_COLS, V2_R, K51, 1};
        if (!make_map_f32(mh, h, 4, tdims, tstrides, box)) return false;
    }
    {
        const int64_t dims[4] = {W, H, 3, B}, strides[4] = {1, W, plane, C * plane};
        const int box[4] = {V3_COLS, V2_R, 3, 1};
        if (!make_map_f32(mg, g + c0 * plane, 4, dims, strides, box)) return false;
    }
    return true;
}

}  // namespace

// returns 0 on launch, > 0 CUDA error, -1000 when the path does not apply (caller runs the first-generation kernel)
int try_launch_bwd_taps_k51_v3(const float* g, const float* in, const float* v, const float* h, float* gv, float* gh,
                               int64_t B, int C, int c0, int H, int W, int accumulate, cudaStream_t s) {
    if ((W & 3) || !aligned16(v) || !aligned16(h) || !aligned16(g) || ((int64_t)H * W * c0 & 3)) return -1000;
    const int64_t tiles_x = (W + V3_COLS * V3_WARPS - 1) / (V3_COLS * V3_WARPS), tiles_y = (H + V2_R - 1) / V2_R;
    if (tiles_x * tiles_y * B > INT32_MAX / 2) return -1000;
    // a persistent grid needs several tiles per warp to balance; small problems stay on the CTA-per-tile kernel
    static const int64_t min_tiles = getenv("SSTEM_V3_MIN_TILES") ? atoll(getenv("SSTEM_V3_MIN_TILES")) : 6;    // measured crossover: ~6 tiles per warp (tap gradients), ~10 (forward)
    if (tiles_x * tiles_y * B < min_tiles * 2 * sm_count()) return -1000;
    const int64_t IH = H + K51 - 1, IW = W + K51 - 1;
    float* ws = nullptr;
    const size_t ws_bytes = (size_t)(B * IH * IW) * 16;
    if (workspace_alloc(reinterpret_cast<void**>(&ws), ws_bytes + 256, s)) return -1000;
    int* counter = reinterpret_cast<int*>(reinterpret_cast<char*>(ws) + ws_bytes);   // dynamic tile scheduler ticket
    cudaMemsetAsync(counter, 0, 256, s);
    int e = launch_repack_nhwc4(in, ws, B, C, c0, IH, IW, s);
    CUtensorMap min, mv, mh, mg;
    if (!e && !make_v3_maps(&min, &mv, &mh, &mg, ws, gh ? v : nullptr, gv ? h : nullptr, g, B, C, c0, H, W)) e = -1000;
    if (!e) {
        if (!gh) mv = min;                                 // unused by the kernel, but must be a valid object to copy
        if (!gv) mh = min;
        V3Shape sh{H, W, (int)tiles_x, (int)tiles_y, (int)(tiles_x * tiles_y * B), C, c0};
#define SSTEM_V3_LAUNCH(WV_, WH_)                                                                   \
        e = accumulate ? launch_v3_kernel<WV_, WH_, true>(min, mv, mh, mg, gv, gh, counter, sh, s)           \
                       : launch_v3_kernel<WV_, WH_, false>(min, mv, mh, mg, gv, gh, counter, sh, s)
        if (gv && gh) SSTEM_V3_LAUNCH(true, true);
        else if (gv) SSTEM_V3_LAUNCH(true, false);
        else SSTEM_V3_LAUNCH(false, true);
#undef SSTEM_V3_LAUNCH
    }
    cudaFreeAsync(ws, s);
    return e;
}

// returns 0 on launch, > 0 CUDA error, -1000 when the v2 path does not apply (caller runs the first-generation kernel)
int try_launch_bwd_taps_k51_v2(const float* g, const float* in, const float* v, const float* h, float* gv, float* gh,
                               int64_t B, int C, int c0, int H, int W, int accumulate, cudaStream_t s) {
    static const int gen = getenv("SSTEM_BWD_GEN") ? atoi(getenv("SSTEM_BWD_GEN")) : 3;   // experiments: force a generation
    if (gen >= 3) {
        const int e3 = try_launch_bwd_taps_k51_v3(g, in, v, h, gv, gh, B, C, c0, H, W, accumulate, s);
        if (e3 != -1000) return e3;
    }
    const bool off = gen != 2 || g_gate.ptr != nullptr;    // generation 2 is kept for A/B runs only (SSTEM_BWD_GEN=2); not gated
    if (off || (W & 3) || !aligned16(v) || B > 65535 || (H + V2_R - 1) / V2_R > 65535) return -1000;
    const int64_t IH = H + K51 - 1, IW = W + K51 - 1;
    float* ws = nullptr;
    if (workspace_alloc(reinterpret_cast<void**>(&ws), (size_t)(B * IH * IW) * 16, s)) return -1000;
    int e = launch_repack_nhwc4(in, ws, B, C, c0, IH, IW, s);
    CUtensorMap min, mv;
    if (!e && !make_v2_maps(&min, &mv, ws, v, B, H, W, V2_R)) e = -1000;
    if (!e) {
#define SSTEM_V2_LAUNCH(WV_, WH_)                                                                          \
        e = accumulate ? launch_v2_kernel<WV_, WH_, true>(min, mv, g, h, gv, gh, B, C, c0, H, W, s)         \
                       : launch_v2_kernel<WV_, WH_, false>(min, mv, g, h, gv, gh, B, C, c0, H, W, s)
        if (gv && gh) SSTEM_V2_LAUNCH(true, true);
        else if (gv) SSTEM_V2_LAUNCH(true, false);
        else SSTEM_V2_LAUNCH(false, true);
#undef SSTEM_V2_LAUNCH
    }
    cudaFreeAsync(ws, s);
    return e;
}

}  // namespace sstem
