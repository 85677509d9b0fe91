// Second-generation 51-tap kernels for the common case (3 channels, W % 4 == 0): TMA-staged, channel-interleaved.
//
// What limits the first-generation kernels (sepconv_k51.cuh) is not the FMA pipe itself but the issue port: per
// step a lane issues 130 FFMA2 (two port cycles each) plus ~100 other instructions -- 39 LDS.32 for the window
// (one per tap and channel), the per-lane cp.async ring of vertical taps with its address arithmetic, and the
// gv reduce -- so the pipe cannot be busier than 260 / 361 of the time.  This version removes most of the "other":
//
//   * the input is first repacked NCHW -> [B][H+50][W+50][4] (channel-interleaved, 4th lane zero) by a small
//     streaming kernel (input is 5 % of the op's bytes); its rows are 16-byte multiples, so the window of a tile
//     arrives with ONE TMA box load, and the lane reads all three channels of a tap with ONE LDS.128
//     (13 per step instead of 39);
//   * the vertical taps of the whole tile -- v[0..50][R rows][32 columns], plus R-1 out-of-range planes on either
//     side that the TMA unit zero-fills -- arrive with ONE more TMA box load; a step reads its diagonal
//     (row p uses fy = s - p) with R LDS at compile-time offsets.  No ring, no per-lane copies, no commit / wait
//     groups in the loop;
//   * one mbarrier wait per tile.
//
// Lane mapping, FFMA2 row-pair packing and the transpose-reduce are those of sepconv_k51.cuh (G = 4 tap groups,
// R = 4 rows); results are bit-identical to the first-generation kernel (same operations in the same order).
#pragma once
#include "sepconv_k51.cuh"
#include "tma.cuh"

namespace sstem {
namespace {

constexpr int V2_G = 4, V2_R = 4;
constexpr int V2_WIN_W = 84;                               // 32 columns + 52 tap slots
constexpr int V2_WIN_H = V2_R + K51 - 1;                   // 54 input rows
constexpr int V2_VPAD = V2_R - 1;                          // zero planes before tap 0 and after tap 50
constexpr int V2_VPLANES = K51 + 2 * V2_VPAD;              // 57
constexpr unsigned V2_WIN_BYTES = V2_WIN_H * V2_WIN_W * 16;
constexpr unsigned V2_VT_BYTES = V2_VPLANES * V2_R * 32 * 4;
constexpr size_t V2_BWD_SMEM = V2_WIN_BYTES + V2_VT_BYTES + 128;

// NCHW planes c0..c0+2 of `in` [B,C,IH,IW] -> out [B,IH,IW] float4 (x, y, z = the three channels, w = 0)
template <bool VEC4>
__global__ void __launch_bounds__(256)
repack_nchw3_to_nhwc4_kernel(const float* __restrict__ in, float4* __restrict__ out, int C, int c0, int64_t plane, int64_t total,
                             const int* gate, int gate_want) {
    SSTEM_GATE_RETURN(gate, gate_want);
    const int64_t step = (int64_t)gridDim.x * blockDim.x;
    if (VEC4) {                                           // plane % 4 == 0 and 16-byte aligned base: 4 pixels per thread
        const int64_t n4 = total >> 2, p4 = plane >> 2;
        for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += step) {
            const int64_t b = i / p4, p = i - b * p4;
            const float4* s = reinterpret_cast<const float4*>(in + (b * C + c0) * plane) + p;
            const float4 a = __ldcs(s), bb = __ldcs(s + p4), c = __ldcs(s + 2 * p4);
            float4* d = out + 4 * i;
            d[0] = make_float4(a.x, bb.x, c.x, 0.f);
            d[1] = make_float4(a.y, bb.y, c.y, 0.f);
            d[2] = make_float4(a.z, bb.z, c.z, 0.f);
            d[3] = make_float4(a.w, bb.w, c.w, 0.f);
        }
    } else {
        for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += step) {
            const int64_t b = i / plane, p = i - b * plane;
            const float* s = in + (b * C + c0) * plane + p;
            out[i] = make_float4(__ldg(s), __ldg(s + plane), __ldg(s + 2 * plane), 0.f);
        }
    }
}

// One plane (channel c0) of in [B,C,IH,IW] -> out [B,IH,IWP] with the row pitch IWP rounded up to 4 floats, so that
// the copy can be the source of a tensor map (16-byte strides); the pad columns are zero.
__global__ void __launch_bounds__(256)
repack_plane_pitch_kernel(const float* __restrict__ in, float* __restrict__ out, int C, int c0, int IH, int IW, int IWP, int64_t total,
                          const int* gate, int gate_want) {
    SSTEM_GATE_RETURN(gate, gate_want);
    const int64_t step = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += step) {
        const int x = (int)(i % IWP);
        const int64_t r = i / IWP;
        const int y = (int)(r % IH);
        const int64_t b = r / IH;
        out[i] = x < IW ? __ldg(in + ((b * C + c0) * IH + y) * (int64_t)IW + x) : 0.f;
    }
}

// flag = 1 iff every channel plane of in [B,C,plane] is bit-identical to plane 0 of its image (gray sections replicated
// x3: sff_scripts_interp/data/data_provider.py:136-137).  *flag must be 1 on entry; any mismatch clears it.
__global__ void __launch_bounds__(256)
planes_equal_kernel(const float* __restrict__ in, int C, int64_t plane, int64_t total /* B * plane */, int* __restrict__ flag) {
    const int64_t step = (int64_t)gridDim.x * blockDim.x;
    bool same = true;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total && same; i += step) {
        const int64_t b = i / plane, p = i - b * plane;
        const unsigned* s = reinterpret_cast<const unsigned*>(in) + b * C * plane + p;
        const unsigned ref = __ldg(s);
        for (int c = 1; c < C; ++c) same = same && (__ldg(s + c * plane) == ref);
    }
    if (!__all_sync(0xffffffffu, same) && (threadIdx.x & 31) == 0) *flag = 0;
}

// NCHW taps [B,51,H,W] -> tile-major [B][ceil(H/8)][ceil(W/8)][51][8][8] (zero outside the image): the layout the
// TILED kernels consume.  One thread per output element; reads are 32-byte row segments.
__global__ void __launch_bounds__(256)
taps_to_tiled_kernel(const float* __restrict__ taps, float* __restrict__ tiled, int H, int W, int ty8, int tx8, int64_t total) {
    const int64_t step = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += step) {
        const int col = (int)(i & 7), row = (int)((i >> 3) & 7);
        int64_t r = i >> 6;
        const int tap = (int)(r % K51); r /= K51;
        const int tx = (int)(r % tx8); r /= tx8;
        const int ty = (int)(r % ty8);
        const int64_t b = r / ty8;
        const int y = ty * 8 + row, x = tx * 8 + col;
        tiled[i] = (y < H && x < W) ? __ldg(taps + ((b * K51 + tap) * H + y) * (int64_t)W + x) : 0.f;
    }
}

#ifndef SSTEM_BWD2_TB
#define SSTEM_BWD2_TB 7                                    // taps per block: NP * TB independent FFMA2 chains
#endif
#ifndef SSTEM_BWD2_SPLIT
#define SSTEM_BWD2_SPLIT 1                                 // two gv partial sums per row pair (shorter dependent chains)
#endif

// One input row.  S >= 0: compile-time step (prologue / epilogue: row pairs that are entirely outside vanish, rows
// whose fy is out of range are zeroed before they can touch an accumulator); S < 0: steady state.
template <int S, bool WV, bool WH>
__device__ __forceinline__ void bwd2_step(const float4* __restrict__ prow, bool novalid,
                                          const float2 (&g2)[3][2], const float2 (&h2)[2][13], const float2 (&v2)[2],
                                          float2 (&gh2)[2][13], float2 (&gvp)[2]) {
    constexpr int NP = 2, NT = 13, TB = SSTEM_BWD2_TB;
    float2 gva[NP], gvb[NP];
#pragma unroll
    for (int pp = 0; pp < NP; ++pp) gva[pp] = gvb[pp] = make_float2(0.f, 0.f);
#pragma unroll
    for (int tb = 0; tb < NT; tb += TB) {
        float2 t2[TB][NP];
        float4 P[TB];
#pragma unroll
        for (int j = 0; j < TB; ++j) {
            const int t = tb + j;
            if (t >= NT) continue;
            P[j] = prow[4 * t];                            // channels x, y, z of column x + g + 4t
            if (t == NT - 1) {                             // tap 51 does not exist (lanes g == 3)
                P[j].x = novalid ? 0.f : P[j].x;
                P[j].y = novalid ? 0.f : P[j].y;
                P[j].z = novalid ? 0.f : P[j].z;
            }
        }
#pragma unroll
        for (int j = 0; j < TB; ++j) {
            if (tb + j >= NT) continue;
#pragma unroll
            for (int pp = 0; pp < NP; ++pp) {
                if (S >= 0 && (S < 2 * pp || S > 2 * pp + K51)) continue;
                t2[j][pp] = __ffma2_rn(make_float2(P[j].x, P[j].x), g2[0][pp], make_float2(0.f, 0.f));
            }
        }
#pragma unroll
        for (int j = 0; j < TB; ++j) {
            if (tb + j >= NT) continue;
#pragma unroll
            for (int pp = 0; pp < NP; ++pp) {
                if (S >= 0 && (S < 2 * pp || S > 2 * pp + K51)) continue;
                t2[j][pp] = __ffma2_rn(make_float2(P[j].y, P[j].y), g2[1][pp], t2[j][pp]);
            }
        }
#pragma unroll
        for (int j = 0; j < TB; ++j) {
            if (tb + j >= NT) continue;
#pragma unroll
            for (int pp = 0; pp < NP; ++pp) {
                if (S >= 0 && (S < 2 * pp || S > 2 * pp + K51)) continue;
                t2[j][pp] = __ffma2_rn(make_float2(P[j].z, P[j].z), g2[2][pp], t2[j][pp]);
            }
        }
#pragma unroll
        for (int j = 0; j < TB; ++j) {
            const int t = tb + j;
            if (t >= NT) continue;
#pragma unroll
            for (int pp = 0; pp < NP; ++pp) {
                if (S >= 0) {
                    if (S < 2 * pp || S > 2 * pp + K51) continue;
                    if (S - 2 * pp > K51 - 1) t2[j][pp].x = 0.f;
                    if (S - 2 * pp - 1 < 0) t2[j][pp].y = 0.f;
                }
                if (WH) gh2[pp][t] = __ffma2_rn(t2[j][pp], v2[pp], gh2[pp][t]);
                if (WV) {
                    if (SSTEM_BWD2_SPLIT && (t & 1)) gvb[pp] = __ffma2_rn(t2[j][pp], h2[pp][t], gvb[pp]);
                    else gva[pp] = __ffma2_rn(t2[j][pp], h2[pp][t], gva[pp]);
                }
            }
        }
    }
#pragma unroll
    for (int pp = 0; pp < NP; ++pp)
        gvp[pp] = SSTEM_BWD2_SPLIT ? make_float2(gva[pp].x + gvb[pp].x, gva[pp].y + gvb[pp].y) : gva[pp];
}

#ifndef SSTEM_BWD2_MINB
#define SSTEM_BWD2_MINB 2
#endif
template <bool WV, bool WH, bool ACCUM>
__global__ void __launch_bounds__(128, SSTEM_BWD2_MINB)
sepconv_bwd_taps_k51_v2_kernel(const __grid_constant__ CUtensorMap map_in, const __grid_constant__ CUtensorMap map_v,
                               const float* __restrict__ gout, const float* __restrict__ h,
                               float* __restrict__ gv, float* __restrict__ gh, int C, int c0, int H, int W) {
    constexpr int G = V2_G, R = V2_R, NP = 2, NT = 13;
    using Gm = Geo<G, R>;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float4* win = reinterpret_cast<float4*>(smem_raw);                      // [54][84] x (c0, c1, c2, 0)
    float* vt = reinterpret_cast<float*>(smem_raw + V2_WIN_BYTES);          // [57][R][32]: plane q = tap q - 3
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw + V2_WIN_BYTES + V2_VT_BYTES);

    const int tid = threadIdx.x;
    const int x0 = blockIdx.x * Gm::TILE_W, y0 = blockIdx.y * R;
    const int b = blockIdx.z;
    if (tid == 0) {
        mbar_init(bar, 1);
        mbar_fence_init();
        mbar_expect_tx(bar, V2_WIN_BYTES + (WH ? V2_VT_BYTES : 0u));
        tma_load_4d(win, &map_in, bar, 0, x0, y0, b);
        if (WH) tma_load_4d(vt, &map_v, bar, x0, y0, -V2_VPAD, b);
    }
    const int64_t plane = (int64_t)H * W;
    const int warp = tid >> 5, lane = tid & 31;
    const int pg = lane / G, g = lane % G;
    const int xl = warp * Gm::COLS + pg;
    const int x = min(x0 + xl, W - 1);
    const bool novalid = g >= Gm::LAST_VALID_G;
    const bool col_ok = (x0 + xl < W);

    float2 h2[NP][NT], gh2[NP][NT], g2[3][NP];
    if (WV) load_h<G, R>(h2, h, (int64_t)b * K51 * plane, plane, y0, x, H, W, g);
#pragma unroll
    for (int pp = 0; pp < NP; ++pp)
#pragma unroll
        for (int t = 0; t < NT; ++t) {
            gh2[pp][t] = make_float2(0.f, 0.f);
            if (!WV) h2[pp][t] = make_float2(0.f, 0.f);
        }
#pragma unroll
    for (int c = 0; c < 3; ++c)
#pragma unroll
        for (int pp = 0; pp < NP; ++pp) {
            const float* gp = gout + ((int64_t)b * C + c0 + c) * plane + x;
            const int ya = y0 + 2 * pp, yb = ya + 1;
            g2[c][pp].x = (col_ok && ya < H) ? __ldg(gp + (int64_t)ya * W) : 0.f;   // outside the image: no contribution
            g2[c][pp].y = (col_ok && yb < H) ? __ldg(gp + (int64_t)yb * W) : 0.f;
        }
    __syncthreads();                                     // the barrier init is visible to every waiter
    mbar_wait(bar, 0);

    const float4* prow = win + xl + g;                   // column of tap slot t is xl + g + 4t
    const float* vrow = vt + V2_VPAD * (R * 32) + xl;    // step s, row p: vrow[s * 128 - 96 * p]
    float* gv_ptr = WV ? gv + (int64_t)b * K51 * plane + (int64_t)min(y0 + g, H - 1) * W + x - (int64_t)g * plane : nullptr;
    const bool gv_row_ok = col_ok && (y0 + g < H);
    float2 gvp[NP], v2[NP];
    auto read_v = [&]() {
#pragma unroll
        for (int pp = 0; pp < NP; ++pp)
            v2[pp] = WH ? make_float2(vrow[-(R * 32 - 32) * (2 * pp)], vrow[-(R * 32 - 32) * (2 * pp + 1)]) : make_float2(0.f, 0.f);
    };
    auto store_gv = [&](int s) {
        if (WV) {
            float val[R];
#pragma unroll
            for (int pp = 0; pp < NP; ++pp) { val[2 * pp] = gvp[pp].x; val[2 * pp + 1] = gvp[pp].y; }
            group_reduce<G, R>(val, g);                  // lane g now holds the total of row g
            const int fy = s - g;
            if (gv_row_ok && fy >= 0 && fy < K51) *gv_ptr = ACCUM ? (*gv_ptr + val[0]) : val[0];
            gv_ptr += plane;
        }
    };
#define SSTEM_BWD2_EDGE_STEP(S)                                                  \
    if ((S) < R - 1 || ((S) >= K51 && (S) < V2_WIN_H)) {                         \
        read_v();                                                                \
        bwd2_step<S, WV, WH>(prow, novalid, g2, h2, v2, gh2, gvp);               \
        store_gv(S);                                                             \
        prow += V2_WIN_W;                                                        \
        vrow += R * 32;                                                          \
    }
    SSTEM_BWD2_EDGE_STEP(0) SSTEM_BWD2_EDGE_STEP(1) SSTEM_BWD2_EDGE_STEP(2)
#ifndef SSTEM_BWD2_UNROLL
#define SSTEM_BWD2_UNROLL 1
#endif
    constexpr int UNR = SSTEM_BWD2_UNROLL;
#pragma unroll UNR
    for (int s = R - 1; s < K51; ++s) {                  // steady state: all rows active
        read_v();
        bwd2_step<-1, WV, WH>(prow, novalid, g2, h2, v2, gh2, gvp);
        store_gv(s);
        prow += V2_WIN_W;
        vrow += R * 32;
    }
    SSTEM_BWD2_EDGE_STEP(51) SSTEM_BWD2_EDGE_STEP(52) SSTEM_BWD2_EDGE_STEP(53)
#undef SSTEM_BWD2_EDGE_STEP
    static_assert(R == 4 && G == 4, "edge-step list and the per-step gv reduce are written for R = G = 4");

    // ---- gh: complete per lane (the sum over fy happened in registers)
    if (WH && col_ok) {
        float* gp = gh + ((int64_t)b * K51 + g) * plane + x0 + xl;
#pragma unroll
        for (int t = 0; t < NT; ++t) {
            if (t == NT - 1 && novalid) break;
#pragma unroll
            for (int pp = 0; pp < NP; ++pp) {
                const int ya = y0 + 2 * pp, yb = ya + 1;
                float* da = gp + (int64_t)(G * t) * plane + (int64_t)ya * W;
                float* db = gp + (int64_t)(G * t) * plane + (int64_t)yb * W;
                if (ya < H) *da = ACCUM ? (*da + gh2[pp][t].x) : gh2[pp][t].x;
                if (yb < H) *db = ACCUM ? (*db + gh2[pp][t].y) : gh2[pp][t].y;
            }
        }
    }
}

}  // namespace
}  // namespace sstem
