// Launcher of the third-generation forward kernel (device code in sepconv_k51_v3.cuh): repack the input
// channel-interleaved into a stream-ordered workspace, encode the tensor maps, launch the persistent grid.
#include "sepconv_k51_v3.cuh"

#include <stdlib.h>

namespace sstem {

// returns 0 on launch, > 0 CUDA error, -1000 when the path does not apply (caller runs the first-generation kernel)
int try_launch_fwd_k51_v3(const float* in, const float* v, const float* h, float* out,
                          int64_t B, int C, int c0, int H, int W, cudaStream_t s) {
    static const int gen = getenv("SSTEM_FWD_GEN") ? atoi(getenv("SSTEM_FWD_GEN")) : 3;   // experiments: force generation 1
    if (gen < 3 || (W & 3) || !aligned16(v) || !aligned16(h)) return -1000;
    const int64_t tiles_x = (W + V3_COLS * V3_WARPS - 1) / (V3_COLS * V3_WARPS), tiles_y = (H + F3_R - 1) / F3_R;
    if (tiles_x * tiles_y * B > INT32_MAX / 2) return -1000;
    // a persistent grid needs several tiles per warp to balance; small problems stay on the CTA-per-tile kernel
    static const int64_t min_tiles = getenv("SSTEM_V3_MIN_TILES") ? atoll(getenv("SSTEM_V3_MIN_TILES")) : 10;   // measured crossover: ~10 tiles per warp (forward), ~6 (tap gradients)
    if (tiles_x * tiles_y * B < min_tiles * 2 * sm_count()) return -1000;
    const int64_t IH = H + K51 - 1, IW = W + K51 - 1, plane = (int64_t)H * W;
    float* ws = nullptr;
    const size_t ws_bytes = (size_t)(B * IH * IW) * 16;
    if (workspace_alloc(reinterpret_cast<void**>(&ws), ws_bytes + 256, s)) return -1000;
    int* counter = reinterpret_cast<int*>(reinterpret_cast<char*>(ws) + ws_bytes);   // dynamic tile scheduler ticket
    cudaMemsetAsync(counter, 0, 256, s);
    int e = launch_repack_nhwc4(in, ws, B, C, c0, IH, IW, s);
    CUtensorMap min, mv, mh;
    if (!e) {
        const int64_t dims[3] = {4 * IW, IH, B}, strides[3] = {1, 4 * IW, 4 * IW * IH};
        const int box[3] = {4 * V3_WIN_COLS, V3_GROUP, 1};
        const int64_t tdims[4] = {W, H, K51, B}, tstrides[4] = {1, W, plane, (int64_t)K51 * plane};
        const int vbox[4] = {V3_COLS, F3_R, V3_GROUP, 1}, hbox[4] = {V3_COLS, F3_R, K51, 1};
        if (!make_map_f32(&min, ws, 3, dims, strides, box) || !make_map_f32(&mv, v, 4, tdims, tstrides, vbox) ||
            !make_map_f32(&mh, h, 4, tdims, tstrides, hbox))
            e = -1000;
    }
    if (!e) {
        static PerDeviceOnce done;
        auto kern = sepconv_fwd_k51_v3_kernel;
        e = set_smem_once(kern, F3_SMEM, done);
        if (!e) {
            V3Shape sh{H, W, (int)tiles_x, (int)tiles_y, (int)(tiles_x * tiles_y * B), C, c0};
            const int ctas = (int)std::min<int64_t>(2 * (int64_t)sm_count(), (int64_t)sh.ntiles);
            kern<<<ctas, V3_WARPS * 32, F3_SMEM, s>>>(min, mv, mh, out, counter, sh);
            count_launch();
            e = finish_launch();
        }
    }
    cudaFreeAsync(ws, s);
    return e;
}

}  // namespace sstem
