// Tuned 51-tap adaptive separable convolution for sm_100a: forward.
//
//   out[b,c,y,x] = sum_fy v[b,fy,y,x] * ( sum_fx in[b,c,y+fy,x+fx] * h[b,fx,y,x] )
//
// (same sum as libs/sepconv/src/SeparableConvolution_kernel.cu:45-49 of the
// reference, factored: 2*C*K*(K+1) flop per pixel instead of 3*C*K*K.)
//
// This is a per-pixel bilinear form v^T P h with no operand shared between
// pixels except the image window P, so it runs on the FP32 FMA pipe, not on the
// tensor cores.  The design problem is operand bandwidth: one LDS per FMA caps
// at 1/4 of the FMA rate.  Mapping:
//
//   * CTA = 4 warps = one 8-row x 32-column output tile of one image, all
//     channels; 2 CTAs per SM (255 registers each) so one CTA's tile load hides
//     behind the other's arithmetic.
//   * The input window of the tile, (8+50) x (32+50) per channel, is staged once
//     in shared memory with cp.async (the reference layout's row pitch,
//     (W+50)*4 B, is not a multiple of 16 B, which rules out a TMA tensor map).
//   * A warp owns 8 columns x 8 rows.  Lane = (column pg = lane>>2, tap group
//     g = lane&3); the lane keeps h[fx][row][col] for its 13 taps fx = 4t+g and
//     all 8 rows of its column in registers (104 values) for the whole tile.
//   * Step s = 0..57 walks the 58 input rows.  One LDS.32 of P[s][col+fx] feeds
//     the 8 rows' FMAs (row p uses it with fy = s-p): 8 FMAs per shared-memory
//     word, issued as 4 packed FFMA2 (fma.rn.f32x2) on row pairs.
//     part[p] = sum over my taps; out_acc[c][p] += v[fy][p] * part[p].
//   * The 4 tap groups of a pixel are summed once per tile with 2 shuffles.
//
// Per step and channel a lane issues 13 LDS + 13 MOV + 52+4 FFMA2 for 112
// FMA-pipe cycles: the FMA pipe is the limiter, as the roofline says it should be.
#include "common.cuh"

namespace sstem {

namespace {

constexpr int K51 = 51;
constexpr int TILE_H = 8;                  // rows per tile = rows held per lane
constexpr int TILE_W = 32;                 // 4 warps x 8 columns
constexpr int IN_ROWS = TILE_H + K51 - 1;  // 58
constexpr int IN_COLS = TILE_W + K51 - 1;  // 82
constexpr int PITCH = 84;                  // smem row pitch in floats (16 B multiple)
constexpr int NT = 13;                     // taps per lane: fx = 4t + g, t = 0..12 (g == 3: t <= 11)
constexpr int NPAIR = TILE_H / 2;
constexpr int VDEPTH = 8;                  // steps of vertical taps in flight per warp (cp.async ring)
constexpr int VSLOT = TILE_H * 8;          // floats per ring slot: 8 rows x 8 columns
#ifndef SSTEM_BWD_ROWS
#define SSTEM_BWD_ROWS 4                   // rows per lane in the tap-gradient kernel (4 or 6)
#endif

__device__ __forceinline__ void cp_async4(float* dst_smem, const float* src, bool valid) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(dst_smem);
    const int n = valid ? 4 : 0;           // src-size 0 => zero fill
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;\n" ::"r"(d), "l"(src), "r"(n));
}
__device__ __forceinline__ void cp_async8(float* dst_smem, const float* src, bool valid) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(dst_smem);
    const int n = valid ? 8 : 0;
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;\n" ::"r"(d), "l"(src), "r"(n));
}
__device__ __forceinline__ void cp_async16(float* dst_smem, const float* src, bool valid) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(dst_smem);
    const int n = valid ? 16 : 0;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(d), "l"(src), "r"(n));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory"); }
__device__ __forceinline__ void cp_async_wait_all() {
    asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;\n" ::: "memory");
}

// One input row for all CC channels.  S >= 0: compile-time step (prologue / epilogue, where
// some of the 8 rows have fy = s - p outside [0, 50] -- inactive pairs vanish at compile time);
// S < 0: steady state, runtime step, every row active.
template <int CC, int S>
__device__ __forceinline__ void fwd_step(const float* __restrict__ prow0, bool g3,
                                         const float2 (&h2)[NPAIR][NT], const float2 (&v2)[NPAIR],
                                         float2 (&acc)[CC][NPAIR]) {
#pragma unroll
    for (int c = 0; c < CC; ++c) {
        const float* prow = prow0 + c * IN_ROWS * PITCH;
        float2 part[NPAIR];
#pragma unroll
        for (int pp = 0; pp < NPAIR; ++pp) part[pp] = make_float2(0.f, 0.f);
#pragma unroll
        for (int t = 0; t < NT; ++t) {
            float P = prow[4 * t];
            if (t == NT - 1) P = g3 ? 0.f : P;         // tap 51 does not exist (lanes g == 3)
#pragma unroll
            for (int pp = 0; pp < NPAIR; ++pp) {
                if (S >= 0 && (S < 2 * pp || S > 2 * pp + K51)) continue;   // pair entirely outside
                part[pp] = __ffma2_rn(make_float2(P, P), h2[pp][t], part[pp]);
            }
        }
#pragma unroll
        for (int pp = 0; pp < NPAIR; ++pp) {
            if (S >= 0) {
                if (S < 2 * pp || S > 2 * pp + K51) continue;
                // a row whose fy is out of range must not even see part (NaN/Inf safety)
                if (S - 2 * pp > K51 - 1) part[pp].x = 0.f;
                if (S - 2 * pp - 1 < 0) part[pp].y = 0.f;
            }
            acc[c][pp] = __ffma2_rn(v2[pp], part[pp], acc[c][pp]);
        }
    }
}

// VEC: W % 4 == 0 and v 16-byte aligned -> the ring is fed with 16-byte cp.async.
// PAIR: (W + 50) even and `in` 8-byte aligned -> the window is staged with 8-byte cp.async.
template <int CC, bool VEC, bool PAIR>
__global__ void __launch_bounds__(128, 2)
sepconv_fwd_k51_kernel(const float* __restrict__ in, const float* __restrict__ v,
                       const float* __restrict__ h, float* __restrict__ out,
                       int C, int c0, int H, int W) {
    extern __shared__ __align__(16) float tile[];      // [CC][IN_ROWS][PITCH] + 4 warps x v ring
    const int IW = W + K51 - 1, IH = H + K51 - 1;
    const int x0 = blockIdx.x * TILE_W, y0 = blockIdx.y * TILE_H;
    const int64_t b = blockIdx.z;
    const int64_t plane = (int64_t)H * W;
    const int tid = threadIdx.x;

    // ---- stage the input window (zero-filled outside the image) ---------------------
    // A thread owns one column (pair) and walks down the rows: no div/mod in the loop.
    {
        constexpr int CPR = PAIR ? PITCH / 2 : PITCH;   // copies per row
        constexpr int RSTEP = 128 / CPR;                // rows covered per pass (3 or 1)
        const int cidx = tid % CPR, r0 = tid / CPR;
        const int col = PAIR ? 2 * cidx : cidx;
        const int gx = x0 + col;
        if (r0 < RSTEP) {
            const bool colok = gx < IW;                 // PAIR: IW and gx even, a pair never straddles
            const float* src = in + (b * C + c0) * (int64_t)IH * IW + (int64_t)(y0 + r0) * IW + (colok ? gx : 0);
            float* dst = tile + r0 * PITCH + col;
#pragma unroll 1
            for (int c = 0; c < CC; ++c) {
                const float* sp = src;
                float* dp = dst;
#pragma unroll 2
                for (int r = r0; r < IN_ROWS; r += RSTEP) {
                    const bool ok = colok && (y0 + r < IH);
                    if (PAIR) cp_async8(dp, ok ? sp : in, ok); else cp_async4(dp, ok ? sp : in, ok);
                    sp += (int64_t)RSTEP * IW;
                    dp += RSTEP * PITCH;
                }
                src += (int64_t)IH * IW;
                dst += IN_ROWS * PITCH;
            }
        }
        cp_async_commit();                              // group 0: the window
    }

    // ---- per-lane geometry ------------------------------------------------------------
    const int warp = tid >> 5, lane = tid & 31;
    const int pg = lane >> 2, g = lane & 3;
    const int xl = warp * 8 + pg;                       // column inside the tile
    const int x = min(x0 + xl, W - 1);                  // clamped for loads; stores are masked
    const bool g3 = (g == 3);

    // ---- vertical taps: at step s row p needs v[fy = s - p][y0 + p][x] --------------------
    // Streamed through a warp-private shared-memory ring with cp.async, VDEPTH steps ahead
    // (each value is used exactly once; registers would have to cover a DRAM latency of
    // several steps).  Slot layout [row p][8 columns]; invalid fy / rows / columns are zero
    // filled, which also makes the prologue / epilogue contributions vanish.
    float* vring = tile + CC * IN_ROWS * PITCH + warp * (VDEPTH * VSLOT);
    // this lane's copy job(s) per step: VEC: lanes 0..15 move 16 B (row lane>>1, half lane&1);
    // scalar: every lane moves 2 x 4 B (rows lane>>3 and 4 + lane>>3, column lane&7)
    const int vp = VEC ? (lane >> 1) : (lane >> 3);
    const int vcol = VEC ? 4 * (lane & 1) : (lane & 7);
    const bool vactive = VEC ? (lane < 16) : true;
    const int xw = x0 + warp * 8 + vcol;
    const bool vok_a = vactive && (y0 + vp < H) && (xw < W);
    const bool vok_b = !VEC && (y0 + vp + 4 < H) && (xw < W);
    // source of (step 0): fy = -vp, advanced by one plane per step
    const float* vsrc_a = v + b * K51 * plane + (int64_t)min(y0 + vp, H - 1) * W + min(xw, W - 1) - (int64_t)vp * plane;
    const float* vsrc_b = v + b * K51 * plane + (int64_t)min(y0 + vp + 4, H - 1) * W + min(xw, W - 1) - (int64_t)(vp + 4) * plane;
    const int vdst = vp * 8 + vcol;
    int vslot_w = 0;                                    // ring slot the next issue writes
    int vstep_w = 0;                                    // step the next issue fetches
    auto issue_v = [&]() {
        float* slot = vring + vslot_w * VSLOT + vdst;
        {
            const int fy = vstep_w - vp;
            const bool ok = vok_a && fy >= 0 && fy < K51;
            if (VEC) { if (vactive) cp_async16(slot, ok ? vsrc_a : v, ok); }
            else cp_async4(slot, ok ? vsrc_a : v, ok);
        }
        if (!VEC) {
            const int fy = vstep_w - vp - 4;
            const bool ok = vok_b && fy >= 0 && fy < K51;
            cp_async4(slot + 32, ok ? vsrc_b : v, ok);
        }
        cp_async_commit();
        vsrc_a += plane;
        vsrc_b += plane;
        ++vstep_w;
        vslot_w = (vslot_w + 1 == VDEPTH) ? 0 : vslot_w + 1;
    };
#pragma unroll
    for (int st = 0; st < VDEPTH - 1; ++st) issue_v();

    // ---- horizontal taps of my 13 fx for the 8 rows, resident for the whole tile --------
    float2 h2[NPAIR][NT];
    {
        const float* hp[TILE_H];
#pragma unroll
        for (int p = 0; p < TILE_H; ++p) hp[p] = h + (b * K51 + g) * plane + (int64_t)min(y0 + p, H - 1) * W + x;
        const int64_t tstep = 4 * plane;
#pragma unroll
        for (int t = 0; t < NT; ++t) {
            const bool last = (t == NT - 1);            // (g == 3, t == 12) = tap 51: masked by P = 0, read tap 47 again
#pragma unroll
            for (int pp = 0; pp < NPAIR; ++pp) {
                const float* pa = (last && g3) ? hp[2 * pp] - tstep : hp[2 * pp];
                const float* pb = (last && g3) ? hp[2 * pp + 1] - tstep : hp[2 * pp + 1];
                h2[pp][t].x = __ldg(pa);
                h2[pp][t].y = __ldg(pb);
                hp[2 * pp] += tstep;
                hp[2 * pp + 1] += tstep;
            }
        }
    }

    float2 acc[CC][NPAIR];
#pragma unroll
    for (int c = 0; c < CC; ++c)
#pragma unroll
        for (int pp = 0; pp < NPAIR; ++pp) acc[c][pp] = make_float2(0.f, 0.f);

    cp_async_wait<VDEPTH - 2>();                        // window + step 0 have landed (this thread's part)
    __syncthreads();                                    // ... and everybody else's part of the window

    int vslot_r = 0;                                    // ring slot of the step about to be read
    auto read_v = [&](float2 (&dst)[NPAIR]) {
        const float* slot = vring + vslot_r * VSLOT + pg;
#pragma unroll
        for (int pp = 0; pp < NPAIR; ++pp) dst[pp] = make_float2(slot[(2 * pp) * 8], slot[(2 * pp + 1) * 8]);
        vslot_r = (vslot_r + 1 == VDEPTH) ? 0 : vslot_r + 1;
    };
    float2 vcur[NPAIR], vnext[NPAIR];
    read_v(vcur);

    const float* prow = tile + xl + g;                  // P column of tap t is xl + g + 4t
    auto advance = [&]() {                              // make the next step readable, refill the ring
        cp_async_wait<VDEPTH - 3>();
        __syncwarp();
        issue_v();
        read_v(vnext);
    };
#define SSTEM_FWD_EDGE_STEP(S)                                            \
    {                                                                     \
        advance();                                                        \
        fwd_step<CC, S>(prow, g3, h2, vcur, acc);                         \
        _Pragma("unroll") for (int pp = 0; pp < NPAIR; ++pp) vcur[pp] = vnext[pp]; \
        prow += PITCH;                                                    \
    }
    SSTEM_FWD_EDGE_STEP(0) SSTEM_FWD_EDGE_STEP(1) SSTEM_FWD_EDGE_STEP(2) SSTEM_FWD_EDGE_STEP(3)
    SSTEM_FWD_EDGE_STEP(4) SSTEM_FWD_EDGE_STEP(5) SSTEM_FWD_EDGE_STEP(6)
#pragma unroll 1
    for (int s = TILE_H - 1; s < K51; ++s) {            // steady state: all 8 rows active
        advance();
        fwd_step<CC, -1>(prow, g3, h2, vcur, acc);
#pragma unroll
        for (int pp = 0; pp < NPAIR; ++pp) vcur[pp] = vnext[pp];
        prow += PITCH;
    }
    SSTEM_FWD_EDGE_STEP(51) SSTEM_FWD_EDGE_STEP(52) SSTEM_FWD_EDGE_STEP(53) SSTEM_FWD_EDGE_STEP(54)
    SSTEM_FWD_EDGE_STEP(55) SSTEM_FWD_EDGE_STEP(56) SSTEM_FWD_EDGE_STEP(57)
#undef SSTEM_FWD_EDGE_STEP

    // ---- sum the 4 tap groups of each pixel; lane g keeps / stores channel g ---------------
    float2 res[NPAIR];
#pragma unroll
    for (int pp = 0; pp < NPAIR; ++pp) res[pp] = make_float2(0.f, 0.f);
#pragma unroll
    for (int c = 0; c < CC; ++c)
#pragma unroll
        for (int pp = 0; pp < NPAIR; ++pp) {
            float a = acc[c][pp].x, d = acc[c][pp].y;
            a += __shfl_xor_sync(0xffffffffu, a, 1);
            d += __shfl_xor_sync(0xffffffffu, d, 1);
            a += __shfl_xor_sync(0xffffffffu, a, 2);
            d += __shfl_xor_sync(0xffffffffu, d, 2);
            if (c == g) res[pp] = make_float2(a, d);
        }
    if (x0 + xl < W && g < CC) {
        float* ob = out + ((b * C + c0 + g) * (int64_t)H + y0) * W + x0 + xl;
#pragma unroll
        for (int pp = 0; pp < NPAIR; ++pp) {
            if (y0 + 2 * pp < H) ob[(int64_t)(2 * pp) * W] = res[pp].x;
            if (y0 + 2 * pp + 1 < H) ob[(int64_t)(2 * pp + 1) * W] = res[pp].y;
        }
    }
}

template <int CC, bool VEC, bool PAIR>
int launch_fwd_variant(const float* in, const float* v, const float* h, float* out,
                       int64_t B, int C, int c0, int H, int W, cudaStream_t s) {
    const size_t smem = ((size_t)CC * IN_ROWS * PITCH + 4 * VDEPTH * VSLOT) * sizeof(float);
    static bool attr_done[16] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (!attr_done[dev & 15]) {
        cudaError_t e = cudaFuncSetAttribute(sepconv_fwd_k51_kernel<CC, VEC, PAIR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
        cudaFuncSetAttribute(sepconv_fwd_k51_kernel<CC, VEC, PAIR>, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
        attr_done[dev & 15] = true;
    }
    dim3 grid((unsigned)((W + TILE_W - 1) / TILE_W), (unsigned)((H + TILE_H - 1) / TILE_H), (unsigned)B);
    sepconv_fwd_k51_kernel<CC, VEC, PAIR><<<grid, 128, smem, s>>>(in, v, h, out, C, c0, H, W);
    count_launch();
    return finish_launch();
}

template <int CC>
int launch_fwd_chunk(const float* in, const float* v, const float* h, float* out,
                     int64_t B, int C, int c0, int H, int W, cudaStream_t s) {
    const bool vec = ((W & 3) == 0) && aligned16(v);
    const bool pair = (((W + K51 - 1) & 1) == 0) && ((reinterpret_cast<uintptr_t>(in) & 7u) == 0);
    if (vec && pair) return launch_fwd_variant<CC, true, true>(in, v, h, out, B, C, c0, H, W, s);
    if (vec) return launch_fwd_variant<CC, true, false>(in, v, h, out, B, C, c0, H, W, s);
    if (pair) return launch_fwd_variant<CC, false, true>(in, v, h, out, B, C, c0, H, W, s);
    return launch_fwd_variant<CC, false, false>(in, v, h, out, B, C, c0, H, W, s);
}


// =====================================================================================
// Backward w.r.t. the taps, fused:
//   t[fy][fx]  = sum_c g[c] * in[c][y+fy][x+fx]
//   gv[fy]     = sum_fx t[fy][fx] * h[fx]          (kernel.cu:97-111 of the reference)
//   gh[fx]     = sum_fy t[fy][fx] * v[fy]          (kernel.cu:134-149)
// 2*(C+2)*K*K flop per pixel.  Same lane mapping as the forward (8 columns x R rows per
// warp, lane = (column, tap group g), taps fx = 4t+g): a lane keeps h and the gh
// accumulators of its 13 taps for R rows in registers; every step it forms t for its taps
// from 13*C shared-memory words, accumulates gh in place, and reduces the gv partial sums
// of the 4 tap groups with 2 shuffles before storing gv[fy = s-p] for each row p.
// =====================================================================================
template <int CC, int R, int S, bool WV, bool WH>
__device__ __forceinline__ void bwd_step(const float* __restrict__ prow0, bool g3,
                                         const float2 (&g2)[CC][R / 2], const float2 (&h2)[R / 2][NT],
                                         const float2 (&v2)[R / 2], float2 (&gh2)[R / 2][NT],
                                         float2 (&gvp)[R / 2]) {
    constexpr int NP = R / 2;
#pragma unroll
    for (int pp = 0; pp < NP; ++pp) gvp[pp] = make_float2(0.f, 0.f);
#pragma unroll
    for (int t = 0; t < NT; ++t) {
        float P[CC];
#pragma unroll
        for (int c = 0; c < CC; ++c) {
            P[c] = prow0[c * (R + K51 - 1) * PITCH + 4 * t];
            if (t == NT - 1) P[c] = g3 ? 0.f : P[c];    // tap 51 does not exist (lanes g == 3)
        }
#pragma unroll
        for (int pp = 0; pp < NP; ++pp) {
            if (S >= 0 && (S < 2 * pp || S > 2 * pp + K51)) continue;
            float2 t2 = make_float2(0.f, 0.f);
#pragma unroll
            for (int c = 0; c < CC; ++c) t2 = __ffma2_rn(make_float2(P[c], P[c]), g2[c][pp], t2);
            if (S >= 0) {                                // rows whose fy is out of range contribute nothing
                if (S - 2 * pp > K51 - 1) t2.x = 0.f;
                if (S - 2 * pp - 1 < 0) t2.y = 0.f;
            }
            if (WV) gvp[pp] = __ffma2_rn(t2, h2[pp][t], gvp[pp]);
            if (WH) gh2[pp][t] = __ffma2_rn(t2, v2[pp], gh2[pp][t]);
        }
    }
}

template <int CC, int R, bool VEC, bool PAIR, bool WV, bool WH>
__global__ void __launch_bounds__(128, 2)
sepconv_bwd_taps_k51_kernel(const float* __restrict__ gout, const float* __restrict__ in,
                            const float* __restrict__ v, const float* __restrict__ h,
                            float* __restrict__ gv, float* __restrict__ gh,
                            int C, int c0, int H, int W, int accumulate) {
    constexpr int NP = R / 2;
    constexpr int ROWS = R + K51 - 1;
    extern __shared__ __align__(16) float tile[];      // [CC][ROWS][PITCH] + 4 warps x v ring
    const int IW = W + K51 - 1, IH = H + K51 - 1;
    const int x0 = blockIdx.x * TILE_W, y0 = blockIdx.y * R;
    const int64_t b = blockIdx.z;
    const int64_t plane = (int64_t)H * W;
    const int tid = threadIdx.x;

    // ---- stage the input window -----------------------------------------------------------
    {
        constexpr int CPR = PAIR ? PITCH / 2 : PITCH;
        constexpr int RSTEP = 128 / CPR;
        const int cidx = tid % CPR, r0 = tid / CPR;
        const int col = PAIR ? 2 * cidx : cidx;
        const int gx = x0 + col;
        if (r0 < RSTEP) {
            const bool colok = gx < IW;
            const float* src = in + (b * C + c0) * (int64_t)IH * IW + (int64_t)(y0 + r0) * IW + (colok ? gx : 0);
            float* dst = tile + r0 * PITCH + col;
#pragma unroll 1
            for (int c = 0; c < CC; ++c) {
                const float* sp = src;
                float* dp = dst;
#pragma unroll 2
                for (int r = r0; r < ROWS; r += RSTEP) {
                    const bool ok = colok && (y0 + r < IH);
                    if (PAIR) cp_async8(dp, ok ? sp : in, ok); else cp_async4(dp, ok ? sp : in, ok);
                    sp += (int64_t)RSTEP * IW;
                    dp += RSTEP * PITCH;
                }
                src += (int64_t)IH * IW;
                dst += ROWS * PITCH;
            }
        }
        cp_async_commit();
    }

    const int warp = tid >> 5, lane = tid & 31;
    const int pg = lane >> 2, g = lane & 3;
    const int xl = warp * 8 + pg;
    const int x = min(x0 + xl, W - 1);
    const bool g3 = (g == 3);
    const bool col_ok = (x0 + xl < W);

    // ---- v ring (only needed for gh) ----------------------------------------------------------
    constexpr int SLOT = R * 8;
    float* vring = tile + CC * ROWS * PITCH + warp * (VDEPTH * SLOT);
    // VEC: lanes 0..2R-1 move 16 B each (row lane>>1, half lane&1); scalar: lane moves up to
    // ceil(8R/32) x 4 B (rows lane>>3 + 4k, column lane&7)
    constexpr int NJOB = VEC ? 1 : (R * 8 + 31) / 32;
    const int vp = VEC ? (lane >> 1) : (lane >> 3);
    const int vcol = VEC ? 4 * (lane & 1) : (lane & 7);
    const int xw = x0 + warp * 8 + vcol;
    const float* vsrc[NJOB];
    bool vok[NJOB];
#pragma unroll
    for (int j = 0; j < NJOB; ++j) {
        const int p = vp + 4 * j;
        vok[j] = (VEC ? (lane < 2 * R) : (p < R)) && (y0 + p < H) && (xw < W);
        vsrc[j] = v + b * K51 * plane + (int64_t)min(y0 + p, H - 1) * W + min(xw, W - 1) - (int64_t)p * plane;
    }
    const int vdst = vp * 8 + vcol;
    int vslot_w = 0, vstep_w = 0;
    auto issue_v = [&]() {
        if (WH) {
            float* slot = vring + vslot_w * SLOT + vdst;
#pragma unroll
            for (int j = 0; j < NJOB; ++j) {
                const int fy = vstep_w - vp - 4 * j;
                const bool ok = vok[j] && fy >= 0 && fy < K51;
                if (VEC) { if (lane < 2 * R) cp_async16(slot, ok ? vsrc[j] : v, ok); }
                else if (vp + 4 * j < R) cp_async4(slot + 32 * j, ok ? vsrc[j] : v, ok);
                vsrc[j] += plane;
            }
        }
        cp_async_commit();
        ++vstep_w;
        vslot_w = (vslot_w + 1 == VDEPTH) ? 0 : vslot_w + 1;
    };
#pragma unroll
    for (int st = 0; st < VDEPTH - 1; ++st) issue_v();

    // ---- per-tile register state: h taps, upstream gradient, gh accumulators ------------------
    float2 h2[NP][NT], gh2[NP][NT], g2[CC][NP];
    {
        const float* hp[R];
#pragma unroll
        for (int p = 0; p < R; ++p) hp[p] = h + (b * K51 + g) * plane + (int64_t)min(y0 + p, H - 1) * W + x;
        const int64_t tstep = 4 * plane;
#pragma unroll
        for (int t = 0; t < NT; ++t) {
            const bool last = (t == NT - 1);
#pragma unroll
            for (int pp = 0; pp < NP; ++pp) {
                if (WV) {
                    const float* pa = (last && g3) ? hp[2 * pp] - tstep : hp[2 * pp];
                    const float* pb = (last && g3) ? hp[2 * pp + 1] - tstep : hp[2 * pp + 1];
                    h2[pp][t] = make_float2(__ldg(pa), __ldg(pb));
                    hp[2 * pp] += tstep;
                    hp[2 * pp + 1] += tstep;
                } else {
                    h2[pp][t] = make_float2(0.f, 0.f);
                }
                gh2[pp][t] = make_float2(0.f, 0.f);
            }
        }
#pragma unroll
        for (int c = 0; c < CC; ++c)
#pragma unroll
            for (int pp = 0; pp < NP; ++pp) {
                const float* gp = gout + (b * C + c0 + c) * plane + x;
                const int ya = y0 + 2 * pp, yb = ya + 1;
                // rows / columns outside the image get g = 0: they then contribute nothing
                g2[c][pp].x = (col_ok && ya < H) ? __ldg(gp + (int64_t)ya * W) : 0.f;
                g2[c][pp].y = (col_ok && yb < H) ? __ldg(gp + (int64_t)yb * W) : 0.f;
            }
    }

    cp_async_wait<VDEPTH - 2>();
    __syncthreads();

    int vslot_r = 0;
    auto read_v = [&](float2 (&dst)[NP]) {
        if (WH) {
            const float* slot = vring + vslot_r * SLOT + pg;
#pragma unroll
            for (int pp = 0; pp < NP; ++pp) dst[pp] = make_float2(slot[(2 * pp) * 8], slot[(2 * pp + 1) * 8]);
        } else {
#pragma unroll
            for (int pp = 0; pp < NP; ++pp) dst[pp] = make_float2(0.f, 0.f);
        }
        vslot_r = (vslot_r + 1 == VDEPTH) ? 0 : vslot_r + 1;
    };
    float2 vcur[NP], vnext[NP];
    read_v(vcur);

    const float* prow = tile + xl + g;
    // gv[fy = s - p][y0 + p][x]: pointer of (s = 0, row p), advanced by one plane per step.
    // After the butterfly every tap-group lane holds the full sums; lane g stores rows p = g, g + 4.
    float* gvp_ptr[2];
#pragma unroll
    for (int j = 0; j < 2; ++j) {
        const int p = g + 4 * j;
        gvp_ptr[j] = gv + b * K51 * plane + (int64_t)min(y0 + p, H - 1) * W + x - (int64_t)p * plane;
    }
    auto advance = [&]() {
        cp_async_wait<VDEPTH - 3>();
        __syncwarp();
        issue_v();
        read_v(vnext);
    };
    auto store_gv = [&](int s, float2 (&gvp)[NP]) {
        if (!WV) return;
        float mine[2] = {0.f, 0.f};
#pragma unroll
        for (int pp = 0; pp < NP; ++pp) {
            float a = gvp[pp].x, d = gvp[pp].y;
            a += __shfl_xor_sync(0xffffffffu, a, 1);
            d += __shfl_xor_sync(0xffffffffu, d, 1);
            a += __shfl_xor_sync(0xffffffffu, a, 2);
            d += __shfl_xor_sync(0xffffffffu, d, 2);
            // row 2pp -> lane (2pp)&3, slot (2pp)>>2 ; row 2pp+1 -> lane (2pp+1)&3
            if (((2 * pp) & 3) == g) mine[(2 * pp) >> 2] = a;
            if (((2 * pp + 1) & 3) == g) mine[(2 * pp + 1) >> 2] = d;
        }
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const int p = g + 4 * j;
            const int fy = s - p;
            if (p < R && col_ok && y0 + p < H && fy >= 0 && fy < K51) {
                float* dst = gvp_ptr[j] + (int64_t)s * plane;
                *dst = accumulate ? (*dst + mine[j]) : mine[j];
            }
        }
    };
    float2 gvp[NP];
#define SSTEM_BWD_EDGE_STEP(S)                                                        \
    if ((S) < R - 1 || ((S) >= K51 && (S) < R + K51 - 1)) {                           \
        advance();                                                                    \
        bwd_step<CC, R, S, WV, WH>(prow, g3, g2, h2, vcur, gh2, gvp);                 \
        store_gv(S, gvp);                                                             \
        _Pragma("unroll") for (int pp = 0; pp < NP; ++pp) vcur[pp] = vnext[pp];       \
        prow += PITCH;                                                                \
    }
    SSTEM_BWD_EDGE_STEP(0) SSTEM_BWD_EDGE_STEP(1) SSTEM_BWD_EDGE_STEP(2) SSTEM_BWD_EDGE_STEP(3)
    SSTEM_BWD_EDGE_STEP(4) SSTEM_BWD_EDGE_STEP(5) SSTEM_BWD_EDGE_STEP(6)
#pragma unroll 1
    for (int s = R - 1; s < K51; ++s) {
        advance();
        bwd_step<CC, R, -1, WV, WH>(prow, g3, g2, h2, vcur, gh2, gvp);
        store_gv(s, gvp);
#pragma unroll
        for (int pp = 0; pp < NP; ++pp) vcur[pp] = vnext[pp];
        prow += PITCH;
    }
    SSTEM_BWD_EDGE_STEP(51) SSTEM_BWD_EDGE_STEP(52) SSTEM_BWD_EDGE_STEP(53) SSTEM_BWD_EDGE_STEP(54)
    SSTEM_BWD_EDGE_STEP(55) SSTEM_BWD_EDGE_STEP(56) SSTEM_BWD_EDGE_STEP(57)
#undef SSTEM_BWD_EDGE_STEP

    // ---- gh: complete per lane (sum over fy happened in registers) ------------------------------
    if (WH && col_ok) {
        float* gp = gh + (b * K51 + g) * plane + x0 + xl;
#pragma unroll
        for (int t = 0; t < NT; ++t) {
            if (t == NT - 1 && g3) break;
#pragma unroll
            for (int pp = 0; pp < NP; ++pp) {
                const int ya = y0 + 2 * pp, yb = ya + 1;
                float* da = gp + (int64_t)(4 * t) * plane + (int64_t)ya * W;
                float* db = gp + (int64_t)(4 * t) * plane + (int64_t)yb * W;
                if (ya < H) *da = accumulate ? (*da + gh2[pp][t].x) : gh2[pp][t].x;
                if (yb < H) *db = accumulate ? (*db + gh2[pp][t].y) : gh2[pp][t].y;
            }
        }
    }
}

template <int CC, int R, bool VEC, bool PAIR, bool WV, bool WH>
int launch_bwd_variant(const float* g, const float* in, const float* v, const float* h, float* gv, float* gh,
                       int64_t B, int C, int c0, int H, int W, int accumulate, cudaStream_t s) {
    const size_t smem = ((size_t)CC * (R + K51 - 1) * PITCH + 4 * VDEPTH * R * 8) * sizeof(float);
    static bool attr_done[16] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    auto kern = sepconv_bwd_taps_k51_kernel<CC, R, VEC, PAIR, WV, WH>;
    if (!attr_done[dev & 15]) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
        cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
        attr_done[dev & 15] = true;
    }
    dim3 grid((unsigned)((W + TILE_W - 1) / TILE_W), (unsigned)((H + R - 1) / R), (unsigned)B);
    kern<<<grid, 128, smem, s>>>(g, in, v, h, gv, gh, C, c0, H, W, accumulate);
    count_launch();
    return finish_launch();
}

template <int CC, bool WV, bool WH>
int launch_bwd_chunk(const float* g, const float* in, const float* v, const float* h, float* gv, float* gh,
                     int64_t B, int C, int c0, int H, int W, int accumulate, cudaStream_t s) {
    constexpr int R = SSTEM_BWD_ROWS;
    const bool vec = ((W & 3) == 0) && aligned16(v);
    const bool pair = (((W + K51 - 1) & 1) == 0) && ((reinterpret_cast<uintptr_t>(in) & 7u) == 0);
    if (vec && pair) return launch_bwd_variant<CC, R, true, true, WV, WH>(g, in, v, h, gv, gh, B, C, c0, H, W, accumulate, s);
    if (vec) return launch_bwd_variant<CC, R, true, false, WV, WH>(g, in, v, h, gv, gh, B, C, c0, H, W, accumulate, s);
    if (pair) return launch_bwd_variant<CC, R, false, true, WV, WH>(g, in, v, h, gv, gh, B, C, c0, H, W, accumulate, s);
    return launch_bwd_variant<CC, R, false, false, WV, WH>(g, in, v, h, gv, gh, B, C, c0, H, W, accumulate, s);
}

template <bool WV, bool WH>
int launch_bwd_all(const float* g, const float* in, const float* v, const float* h, float* gv, float* gh,
                   int64_t B, int C, int H, int W, cudaStream_t s) {
    int c0 = 0;
    while (c0 < C) {                                       // channel chunks of <= 3; later chunks accumulate
        const int cc = (C - c0) < 3 ? (C - c0) : 3;
        const int acc = c0 > 0;
        int e;
        if (cc == 3) e = launch_bwd_chunk<3, WV, WH>(g, in, v, h, gv, gh, B, C, c0, H, W, acc, s);
        else if (cc == 2) e = launch_bwd_chunk<2, WV, WH>(g, in, v, h, gv, gh, B, C, c0, H, W, acc, s);
        else e = launch_bwd_chunk<1, WV, WH>(g, in, v, h, gv, gh, B, C, c0, H, W, acc, s);
        if (e) return e;
        c0 += cc;
    }
    return 0;
}

}  // namespace

int launch_sepconv_fwd_k51(const float* in, const float* v, const float* h, float* out,
                           int64_t B, int64_t C, int64_t H, int64_t W, cudaStream_t s) {
    if (B > 65535 || (H + TILE_H - 1) / TILE_H > 65535)   // grid.y / grid.z limits
        return launch_sepconv_fwd_generic(in, v, h, out, B, C, H, W, 51, false, s);
    int c0 = 0;
    while (c0 < C) {                                       // channels in chunks of <= 3 (taps re-read per chunk)
        const int cc = (C - c0) < 3 ? (int)(C - c0) : 3;
        int e;
        if (cc == 3) e = launch_fwd_chunk<3>(in, v, h, out, B, (int)C, c0, (int)H, (int)W, s);
        else if (cc == 2) e = launch_fwd_chunk<2>(in, v, h, out, B, (int)C, c0, (int)H, (int)W, s);
        else e = launch_fwd_chunk<1>(in, v, h, out, B, (int)C, c0, (int)H, (int)W, s);
        if (e) return e;
        c0 += cc;
    }
    return 0;
}

int launch_sepconv_bwd_taps_k51(const float* g, const float* in, const float* v, const float* h,
                                float* gv, float* gh, int64_t B, int64_t C, int64_t H, int64_t W, cudaStream_t s) {
    if (B > 65535 || (H + SSTEM_BWD_ROWS - 1) / SSTEM_BWD_ROWS > 65535)
        return launch_sepconv_bwd_taps_generic(g, in, v, h, gv, gh, B, C, H, W, 51, s);
    if (gv && gh) return launch_bwd_all<true, true>(g, in, v, h, gv, gh, B, (int)C, (int)H, (int)W, s);
    if (gv) return launch_bwd_all<true, false>(g, in, v, h, gv, gh, B, (int)C, (int)H, (int)W, s);
    return launch_bwd_all<false, true>(g, in, v, h, gv, gh, B, (int)C, (int)H, (int)W, s);
}

}  // namespace sstem
