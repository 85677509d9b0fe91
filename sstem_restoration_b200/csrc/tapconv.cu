// Tap producer (SURVEY 8f, N2 -- the producer side): the last layer pair of the reference's `_kernel_module`,
//     nn.Upsample(scale_factor=2, mode='bilinear', align_corners=True)  ->  nn.Conv2d(51, 51, 3, 1, 1)
// (sff_scripts_interp/model/model_interp.py:18, 130-137; called four times per forward, :86-89), as ONE kernel that
// reads the half-resolution activations, never materialises the upsampled tensor, and writes the taps either NCHW or
// straight into the tile-major layout the consumer streams ([B][H/8][W/8][51][8][8], sstem_sepconv_forward_tiled).
//
// The reference runs this layer through cuDNN with torch's default `allow_tf32 = True`, i.e. in TF32 on the tensor
// cores; so does this kernel, hand-written for sm_100a:
//   * implicit GEMM per output tile of 16 rows x 8 columns: D[128 pixels x 64] += A_tap[128 x 56] * W_tap[64 x 56]^T for
//     the nine taps, 7 K-steps of 8 each = 63 tcgen05.mma (kind::tf32, M = 128, N = 64), accumulator in TMEM
//     (two stages of 64 columns, so the epilogue of tile i overlaps the MMAs of tile i + 1);
//   * the A operand is never gathered per tap: the upsampled 18 x 10 input patch lives in shared memory ONCE, as
//     [channel chunk of 4][patch pixel][4 channels] (16 bytes per pixel and chunk).  In the no-swizzle K-major
//     canonical layout a core matrix is 8 rows x 16 bytes, contiguous: 8 neighbouring pixels of a patch row.  The
//     descriptor of tap (dy, dx) is the same patch with its start address moved by (dy * 10 + dx) * 16 bytes,
//     stride-byte-offset = one patch row (160 B), leading-byte-offset = one chunk (2880 B)
//     (tools/microbench/umma_probe.cu checks exactly this encoding, shifted starts included);
//   * the weights of all nine taps stay resident in shared memory for the life of the CTA (118 KB, bulk copies);
//   * warp roles: 4 epilogue warps (TMEM -> registers -> + bias -> global), 1 MMA warp (uniform descriptor arithmetic,
//     one lane issues), 12 producer warps (thread = patch pixel x chunk parity: half-resolution source window ->
//     shared memory, bilinear blend with ATen's index / weight expressions, round to TF32, write the patch);
//     mbarriers between them, a persistent grid of one CTA per SM.
// Shared memory bandwidth bounds the kernel (each MMA reads 4 KB of A and 2 KB of W for 65 536 FMAs); see DESIGN 4.11
// and profiles/tapconv_r2.md.
#include <string.h>

#include <algorithm>

#include "common.cuh"
#include "tma.cuh"   // mbar_fence_init

#ifndef SSTEM_TC_DIAG
#define SSTEM_TC_DIAG 0                                    // kernel-tuning experiments: 1 = no blend, 2 = no output stores, 4 = no MMAs
#endif

namespace sstem {
namespace {

constexpr int TC_MAXC = 52;                                // input channels: 13 chunks of 4
constexpr int TC_WCHUNKS = TC_MAXC / 4;                    // 13 weight chunks per tap (+ one zero chunk shared by all taps)
constexpr int TC_CHUNKS = TC_WCHUNKS + 1;                  // 14 patch chunks = 7 K-steps of 8; chunk 13 is always zero
constexpr int TC_N = 64;                                   // output channels, padded (UMMA N)
constexpr int TC_TH = 16, TC_TW = 8;                       // output tile: 128 pixels = UMMA M
constexpr int TC_PH = TC_TH + 2, TC_PW = TC_TW + 2, TC_NPIX = TC_PH * TC_PW;   // 18 x 10 = 180
// Half-resolution source window: 11 rows x up to 7 columns are used; it is fetched with 16-byte loads (and, in an
// earlier variant, by TMA, which has the same rule), so the first column is rounded down to a multiple of 4 and 12
// columns are loaded (48-byte rows).
constexpr int TC_WIN_H = 11, TC_WIN_W = 12;
constexpr int TC_WIN_CH = TC_WIN_H * TC_WIN_W;             // 132 floats = 528 bytes per channel: [channel][row][column]
constexpr int TC_WIN_GROUPS = 2, TC_WIN_GROUP_CH = 28;     // two channel groups (chunks 0-6, 7-12; a 128-byte aligned start each)
constexpr unsigned TC_WIN_GROUP_BYTES = TC_WIN_GROUP_CH * TC_WIN_CH * 4;       // 14784
constexpr unsigned TC_WIN_GROUP_PITCH = (TC_WIN_GROUP_BYTES + 127) / 128 * 128;   // 14848
constexpr unsigned TC_W_TAP_BYTES = TC_WCHUNKS * TC_N * 16;                    // 13312
constexpr unsigned TC_W_ZERO_OFF = 9 * TC_W_TAP_BYTES;                         // 119808: the zero chunk
constexpr unsigned TC_W_BYTES = TC_W_ZERO_OFF + TC_N * 16;                     // 120832
constexpr unsigned TC_A_BYTES = TC_CHUNKS * TC_NPIX * 16;                      // 40320
constexpr unsigned TC_WIN_BYTES = TC_WIN_GROUPS * TC_WIN_GROUP_PITCH;          // 29696
constexpr unsigned TC_OFF_A = TC_W_BYTES;
constexpr unsigned TC_OFF_WIN = TC_OFF_A + 2 * TC_A_BYTES;
constexpr unsigned TC_OFF_BIAS = TC_OFF_WIN + TC_WIN_BYTES;
constexpr unsigned TC_OFF_BAR = TC_OFF_BIAS + TC_N * 4;
constexpr unsigned TC_SMEM = TC_OFF_BAR + 128 + 128;       // + barriers + alignment slack = 231 680 (of 232 448)
static_assert(TC_SMEM <= 232448, "shared memory budget");
constexpr int TC_EPI_WARPS = 4, TC_PROD_WARPS = 12;
constexpr int TC_THREADS = (TC_EPI_WARPS + 1 + TC_PROD_WARPS) * 32;            // 544
constexpr int TC_PROD_THREADS = TC_PROD_WARPS * 32;                            // 384 = 2 x 192: (patch pixel, chunk parity)
constexpr int TC_PIX_THREADS = TC_PROD_THREADS / 2;                            // 192 >= 180 patch pixels
constexpr unsigned TC_TMEM_COLS = 128;                     // two accumulator stages of 64 columns

struct TapConvShape {
    int B, cin, cout, h, w, H, W;                          // h, w: source; H, W: output (2h, 2w when upsampling)
    int tiles_x, tiles_y, ntiles;
    int k_steps;                                           // ceil(cin / 8)
    float ry, rx;                                          // align_corners scales (in - 1) / (out - 1)
};

// a wait that cannot hang the device: a protocol error traps (the launch fails loudly) instead of spinning for ever.
// try_wait suspends the thread in hardware for up to the hinted time, so a waiting role issues next to nothing.
__device__ __forceinline__ void tc_wait(unsigned bar, unsigned parity) {
    unsigned ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    if (ok) return;
    const long long t0 = clock64();
    for (unsigned spins = 1;; ++spins) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(bar), "r"(parity), "r"(20000u) : "memory");
        if (ok) return;
        if ((spins & 63u) == 0 && clock64() - t0 > 6000000000ll) __trap();   // ~3 s: a deadlock, not a slow tile
    }
}
__device__ __forceinline__ void tc_arrive(unsigned bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_commit(unsigned bar) {   // arrives on `bar` when every MMA issued so far has completed
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ uint64_t tc_desc(unsigned addr, unsigned lbo, unsigned sbo) {
    // K-major, no swizzle: start >> 4 | leading byte offset (next 16-byte K chunk) | stride byte offset (next 8 rows) | version 1
    return (uint64_t)((addr >> 4) & 0x3fffu) | ((uint64_t)((lbo >> 4) & 0x3fffu) << 16) |
           ((uint64_t)((sbo >> 4) & 0x3fffu) << 32) | ((uint64_t)1 << 46);
}
__device__ __forceinline__ void tc_bulk_load(unsigned dst, const void* src, unsigned bytes, unsigned bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ float to_tf32(float x) {        // round to nearest (the tensor core itself would truncate)
    unsigned u;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x));
    return __uint_as_float(u);
}
__device__ __forceinline__ void sts128(unsigned addr, float a, float b, float c, float d) {
    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
__device__ __forceinline__ float lds_f32(unsigned addr) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
    return v;
}

// UPS: fold the x2 upsample into the producer.  WVEC (UPS only): w % 4 == 0 and x is 16-byte aligned, so the source
// window is fetched with 16-byte loads; otherwise element by element.
// FULL: cin is 49..52, i.e. all 13 chunks exist -- the MMA and blend loops then have no per-chunk checks.
template <bool UPS, bool TILED, bool WVEC, bool FULL>
__global__ void __launch_bounds__(TC_THREADS, 1)
tap_conv3x3_kernel(const float* __restrict__ x, const float* __restrict__ wpacked,
                   const float* __restrict__ bias, float* __restrict__ out, const TapConvShape sh) {
    extern __shared__ __align__(128) unsigned char tc_smem_raw[];
    const unsigned base = ((unsigned)__cvta_generic_to_shared(tc_smem_raw) + 127u) & ~127u;
    unsigned char* gen = tc_smem_raw + (base - (unsigned)__cvta_generic_to_shared(tc_smem_raw));
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    // barriers: 0 weights, 1-2 a_full, 3-4 a_empty, 5-6 acc_full, 7-8 acc_empty; the TMEM base address after them
    const unsigned bar = base + TC_OFF_BAR;
    const unsigned b_w = bar, b_afull = bar + 8, b_aempty = bar + 24, b_accfull = bar + 40, b_accempty = bar + 56;
    volatile unsigned* tmem_slot = reinterpret_cast<volatile unsigned*>(gen + TC_OFF_BAR + 96);

    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(b_w));
        for (int s = 0; s < 2; ++s) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(b_afull + 8 * s), "r"(TC_PROD_THREADS));
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(b_aempty + 8 * s));
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(b_accfull + 8 * s));
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(b_accempty + 8 * s), "r"(TC_EPI_WARPS * 32));
        }
        mbar_fence_init();
    }
    // zero both patches once (the chunks past cin and the pad pixels a tile never writes must be finite) and the bias pad
    for (unsigned i = tid; i < 2 * TC_A_BYTES / 16; i += TC_THREADS) sts128(base + TC_OFF_A + i * 16, 0.f, 0.f, 0.f, 0.f);
        if (tid < TC_N) reinterpret_cast<float*>(gen + TC_OFF_BIAS)[tid] = (bias != nullptr && tid < sh.cout) ? bias[tid] : 0.f;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (warp == 0) {
        __syncwarp();
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(bar + 96), "n"(TC_TMEM_COLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const unsigned tmem = *tmem_slot;

    // A CTA walks tiles blockIdx.x, + gridDim.x, ...; the tile coordinates advance by a fixed (column, row) step, so
    // the loop needs no division (two 32-bit divisions per tile and warp were 15 % of all instructions issued).
    struct TileIter {
        int tile, tx, ty, b;
    };
    const int step_tx = (int)gridDim.x % sh.tiles_x, step_r = (int)gridDim.x / sh.tiles_x;
    auto first_tile = [&]() {
        TileIter t;
        t.tile = blockIdx.x;
        const int r = t.tile / sh.tiles_x;
        t.tx = t.tile - r * sh.tiles_x;
        t.b = r / sh.tiles_y;
        t.ty = r - t.b * sh.tiles_y;
        return t;
    };
    auto advance = [&](TileIter& t) {
        t.tile += gridDim.x;
        t.tx += step_tx;
        int dr = step_r;
        if (t.tx >= sh.tiles_x) { t.tx -= sh.tiles_x; ++dr; }
        t.ty += dr;
        while (t.ty >= sh.tiles_y) { t.ty -= sh.tiles_y; ++t.b; }
    };

    if (warp < TC_EPI_WARPS) {
        // ===================== epilogue: TMEM lanes 32 * warp .. + 31 = tile pixels (row 4 * warp + lane / 8, column lane % 8)
        const float* sbias = reinterpret_cast<const float*>(gen + TC_OFF_BIAS);
        const int r = warp * 4 + (lane >> 3), c = lane & 7;
        const int64_t plane = (int64_t)sh.H * sh.W;
        const int tiles_y8 = (sh.H + 7) / 8, tiles_x8 = (sh.W + 7) / 8;
        int it = 0;
        for (TileIter ti = first_tile(); ti.tile < sh.ntiles; advance(ti), ++it) {
            const int as = it & 1;
            const int b = ti.b, Y0 = ti.ty * TC_TH, X0 = ti.tx * TC_TW;
            tc_wait(b_accfull + 8 * as, (it >> 1) & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            unsigned v[64];
#pragma unroll
            for (int hh = 0; hh < 2; ++hh) {
                const unsigned ta = tmem + ((unsigned)(warp * 32) << 16) + as * TC_N + hh * 32;
                asm volatile(
                    "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
                    "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                    : "=r"(v[hh*32+0]), "=r"(v[hh*32+1]), "=r"(v[hh*32+2]), "=r"(v[hh*32+3]), "=r"(v[hh*32+4]), "=r"(v[hh*32+5]),
                      "=r"(v[hh*32+6]), "=r"(v[hh*32+7]), "=r"(v[hh*32+8]), "=r"(v[hh*32+9]), "=r"(v[hh*32+10]), "=r"(v[hh*32+11]),
                      "=r"(v[hh*32+12]), "=r"(v[hh*32+13]), "=r"(v[hh*32+14]), "=r"(v[hh*32+15]), "=r"(v[hh*32+16]), "=r"(v[hh*32+17]),
                      "=r"(v[hh*32+18]), "=r"(v[hh*32+19]), "=r"(v[hh*32+20]), "=r"(v[hh*32+21]), "=r"(v[hh*32+22]), "=r"(v[hh*32+23]),
                      "=r"(v[hh*32+24]), "=r"(v[hh*32+25]), "=r"(v[hh*32+26]), "=r"(v[hh*32+27]), "=r"(v[hh*32+28]), "=r"(v[hh*32+29]),
                      "=r"(v[hh*32+30]), "=r"(v[hh*32+31])
                    : "r"(ta));
            }
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            tc_arrive(b_accempty + 8 * as);                // the accumulator stage is free: the MMAs of tile it + 2 may start
            const int Y = Y0 + r, X = X0 + c;
            if (SSTEM_TC_DIAG & 2) continue;
            if (TILED) {
                // the 16 x 8 tile is two 8 x 8 blocks of the tile-major layout; a warp writes 32 consecutive floats per tap
                // (cout == 51 here).  Pad pixels of an existing block (ragged H / W) are written as zeros.
                if ((Y >> 3) < tiles_y8) {
                    float* p = out + (((int64_t)b * tiles_y8 + (Y >> 3)) * tiles_x8 + (X0 >> 3)) * (51 * 64) + (Y & 7) * 8 + c;
                    if (Y0 + TC_TH <= sh.H && X0 + TC_TW <= sh.W) {   // the whole tile is inside the image: no masks
#pragma unroll
                        for (int n4 = 0; n4 < 52; n4 += 4) {
                            const float4 b4 = *reinterpret_cast<const float4*>(sbias + n4);
                            p[(n4 + 0) * 64] = __uint_as_float(v[n4 + 0]) + b4.x;
                            p[(n4 + 1) * 64] = __uint_as_float(v[n4 + 1]) + b4.y;
                            p[(n4 + 2) * 64] = __uint_as_float(v[n4 + 2]) + b4.z;
                            if (n4 + 3 < 51) p[(n4 + 3) * 64] = __uint_as_float(v[n4 + 3]) + b4.w;
                        }
                    } else {
                        const bool ok = Y < sh.H && X < sh.W;
#pragma unroll
                        for (int n = 0; n < 51; ++n) p[n * 64] = ok ? __uint_as_float(v[n]) + sbias[n] : 0.f;
                    }
                }
            } else if (Y < sh.H && X < sh.W) {
                float* p = out + (int64_t)b * sh.cout * plane + (int64_t)Y * sh.W + X;
#pragma unroll
                for (int n = 0; n < TC_N; ++n)
                    if (n < sh.cout) p[(int64_t)n * plane] = __uint_as_float(v[n]) + sbias[n];
            }
        }
    } else if (warp == TC_EPI_WARPS) {
        // ===================== MMA issuer.  The whole warp walks the loop, so every descriptor is warp-uniform arithmetic in
        // uniform registers; one lane issues.  (Issued from inside an `if (lane == 0)` region the descriptors were
        // per-thread values moved to uniform registers one by one: ~110 cycles per MMA, 3.4x the tensor core's own time.)
        const unsigned tmem_u = __shfl_sync(0xffffffffu, tmem, 0);
        if (lane == 0) {
            // weights: resident for the life of the CTA
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b_w), "r"(TC_W_BYTES) : "memory");
            for (int t = 0; t < 9; ++t)
                tc_bulk_load(base + t * TC_W_TAP_BYTES, reinterpret_cast<const char*>(wpacked) + (size_t)t * TC_W_TAP_BYTES, TC_W_TAP_BYTES, b_w);
            tc_bulk_load(base + TC_W_ZERO_OFF, reinterpret_cast<const char*>(wpacked) + TC_W_ZERO_OFF, TC_N * 16, b_w);
        }
        __syncwarp();
        tc_wait(b_w, 0);
        // instruction descriptor: D = f32, A = B = tf32, both K-major, N = 64, M = 128
        constexpr unsigned idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((unsigned)(TC_N >> 3) << 17) | ((unsigned)(128 >> 4) << 24);
        // Descriptors differ only in their start-address field (units of 16 bytes, no carry out of the field: shared memory
        // addresses are < 2^18): per tap + (dy * 10 + dx) pixels for the patch, + 13 chunks for the weights; per K-step
        // + 2 chunks each.  The last K-step's second weight chunk is the zero chunk all taps share: its leading-byte
        // offset field is (zero chunk - chunk 12 of the tap) instead of one chunk.
        const uint64_t db0 = tc_desc(base, TC_N * 16, 128);
        const int k_steps = FULL ? TC_CHUNKS / 2 : sh.k_steps;
        int it = 0;
        for (int tile = blockIdx.x; tile < sh.ntiles; tile += gridDim.x, ++it) {   // (needs no coordinates)
            const int s = it & 1;
            const unsigned ph = (it >> 1) & 1;
            tc_wait(b_accempty + 8 * s, ph ^ 1);           // accumulator stage drained by the epilogue (passes at once the first time)
            tc_wait(b_afull + 8 * s, ph);                  // patch written and fenced by the producers
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint64_t da0 = (SSTEM_TC_DIAG & 32) ? tc_desc(base + TC_OFF_A, 3072, 256)       // timing experiment: 128-byte aligned core matrices
                                                      : tc_desc(base + TC_OFF_A + s * TC_A_BYTES, TC_NPIX * 16, TC_PW * 16);
            const unsigned d_tmem = tmem_u + s * TC_N;
#pragma unroll
            for (int tap = 0; tap < 9; ++tap) {
#pragma unroll
                for (int ks = 0; ks < TC_CHUNKS / 2; ++ks) {
                    if ((SSTEM_TC_DIAG & 4) || ks >= k_steps) break;
                    const uint64_t da = (SSTEM_TC_DIAG & 32) ? da0 + (uint64_t)((tap / 3) * 16 + ks * 2 * 192)
                                      : (SSTEM_TC_DIAG & 16) ? da0 + (uint64_t)((tap / 3) * TC_PW + ks * 2 * TC_NPIX)
                                                             : da0 + (uint64_t)((tap / 3) * TC_PW + (tap % 3) + ks * 2 * TC_NPIX);
                    uint64_t db = db0 + (uint64_t)((tap * TC_W_TAP_BYTES + ks * 2 * TC_N * 16) >> 4);
                    if (2 * ks + 1 >= TC_WCHUNKS)
                        db += (uint64_t)(((TC_W_ZERO_OFF - tap * TC_W_TAP_BYTES - ks * 2 * TC_N * 16) >> 4) - TC_N) << 16;
                    if (lane == 0)
                        asm volatile(
                            "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                            "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                            ::"r"(d_tmem), "l"(da), "l"(db), "r"(idesc), "r"((unsigned)((tap | ks) != 0)) : "memory");
                }
            }
            __syncwarp();
            if (lane == 0) {
                tc_commit(b_aempty + 8 * s);               // patch stage may be overwritten
                tc_commit(b_accfull + 8 * s);              // accumulator complete
            }
        }
    } else {
        // ===================== producers: thread = (patch pixel, chunk parity); 180 of each 192 threads have a pixel
        const int ptid = tid - (TC_EPI_WARPS + 1) * 32;
        const int half = ptid >= TC_PIX_THREADS ? 1 : 0, pix = ptid - half * TC_PIX_THREADS;
        const int py = pix / TC_PW, px = pix % TC_PW;
        const bool has_pixel = pix < TC_NPIX;
        const int nchunks = (sh.cin + 3) >> 2;
        const int64_t src_plane = (int64_t)sh.h * sh.w;
        const unsigned win = base + TC_OFF_WIN;
        float* win_gen = reinterpret_cast<float*>(gen + TC_OFF_WIN);
        auto window_origin = [&](int Y0, int X0, int& sy0, int& sx0) {
            // PyTorch upsample_bilinear2d, align_corners = True: source index = scale * dst in float, cut to int
            sy0 = (int)(sh.ry * (float)max(Y0 - 1, 0));
            sx0 = (int)(sh.rx * (float)max(X0 - 1, 0)) & ~3;       // 16-byte aligned rows
        };
        // The window of tile i + 1 is fetched into registers while tile i's patch is computed (a whole tile of lookahead hides
        // the L2 / DRAM latency) and moves to shared memory once every producer is done reading the previous window.  A TMA
        // box per channel group was tried first: its 48-byte rows cost ~2 us per box and shared memory has no room for a
        // second window, so half of that latency stayed exposed (0.85 ms vs the register prefetch).
        constexpr int ROW4 = TC_WIN_W / 4;                 // 16-byte pieces per window row: 3
        constexpr int NFILL = WVEC ? (TC_MAXC * TC_WIN_H * ROW4 + TC_PROD_THREADS - 1) / TC_PROD_THREADS      // 5 x float4
                                   : (TC_MAXC * TC_WIN_CH + TC_PROD_THREADS - 1) / TC_PROD_THREADS;          // 18 x float
        float4 pre4[WVEC ? NFILL : 1];
        float pre[WVEC ? 1 : NFILL];
        // WVEC: which 16-byte piece this thread fetches in round k does not depend on the tile: its channel offset (elements;
        // the host checks cin * h * w < 2^31), window row and column piece are computed once
        int f_choff[WVEC ? NFILL : 1], f_wy[WVEC ? NFILL : 1], f_c4[WVEC ? NFILL : 1], f_dst[WVEC ? NFILL : 1];
        if (WVEC) {
#pragma unroll
            for (int k = 0; k < NFILL; ++k) {
                const int i = ptid + k * TC_PROD_THREADS;
                const int ch = i / (TC_WIN_H * ROW4), rem = i - ch * (TC_WIN_H * ROW4);
                f_wy[k] = rem / ROW4;
                f_c4[k] = 4 * (rem - f_wy[k] * ROW4);
                f_choff[k] = ch < sh.cin ? ch * (int)src_plane : -1;
                const int g = ch >= TC_WIN_GROUP_CH ? 1 : 0;
                f_dst[k] = ch < TC_MAXC ? (int)(g * TC_WIN_GROUP_PITCH + (ch - g * TC_WIN_GROUP_CH) * (TC_WIN_CH * 4) + rem * 16) : -1;
            }
        }
        auto load_window = [&](const TileIter& t) {
            int sy0, sx0;
            window_origin(t.ty * TC_TH, t.tx * TC_TW, sy0, sx0);
            const float* xb = x + (int64_t)t.b * sh.cin * src_plane;
#pragma unroll
            for (int k = 0; k < NFILL; ++k) {
                const int i = ptid + k * TC_PROD_THREADS;
                if (WVEC) {
                    const int gy = min(sy0 + f_wy[k], sh.h - 1), gx = sx0 + f_c4[k];   // columns past the edge are never blended in
                    pre4[k] = (f_choff[k] >= 0 && gx < sh.w) ? __ldg(reinterpret_cast<const float4*>(xb + (f_choff[k] + gy * sh.w + gx)))
                                                             : make_float4(0.f, 0.f, 0.f, 0.f);
                } else {
                    const int ch = i / TC_WIN_CH, p = i - ch * TC_WIN_CH;
                    const int wy = p / TC_WIN_W, wx = p - wy * TC_WIN_W;
                    const int gy = min(sy0 + wy, sh.h - 1), gx = min(sx0 + wx, sh.w - 1);
                    pre[k] = ch < sh.cin ? __ldg(xb + (int64_t)ch * src_plane + (int64_t)gy * sh.w + gx) : 0.f;
                }
            }
        };
        auto store_window = [&]() {
#pragma unroll
            for (int k = 0; k < NFILL; ++k) {
                const int i = ptid + k * TC_PROD_THREADS;
                if (WVEC) {
                    if (f_dst[k] >= 0) sts128(win + f_dst[k], pre4[k].x, pre4[k].y, pre4[k].z, pre4[k].w);
                } else {
                    const int ch = i / TC_WIN_CH, p = i - ch * TC_WIN_CH;
                    const int g = ch >= TC_WIN_GROUP_CH ? 1 : 0;
                    if (ch < TC_MAXC) win_gen[g * (TC_WIN_GROUP_PITCH / 4) + (ch - g * TC_WIN_GROUP_CH) * TC_WIN_CH + p] = pre[k];
                }
            }
        };
        TileIter ti = first_tile(), tn = ti;               // this tile and the next one of this CTA
        advance(tn);
        if (UPS && ti.tile < sh.ntiles) load_window(ti);
        // per-thread constants of the blend: chunk parity folded into the base addresses, so every shared-memory offset
        // below is an immediate
        const unsigned half_win = half * 4 * (TC_WIN_CH * 4), half_dst = half * TC_NPIX * 16;
        int it = 0;
        for (; ti.tile < sh.ntiles; ti = tn, advance(tn), ++it) {
            const int s = it & 1;
            const unsigned ph = (it >> 1) & 1;
            const int b = ti.b, Y0 = ti.ty * TC_TH, X0 = ti.tx * TC_TW;
            const int Y = Y0 - 1 + py, X = X0 - 1 + px;
            const bool inside = has_pixel && Y >= 0 && Y < sh.H && X >= 0 && X < sh.W;   // outside: the convolution's zero padding
            const unsigned a_dst = base + TC_OFF_A + s * TC_A_BYTES + pix * 16 + half_dst;
            const bool has_next = tn.tile < sh.ntiles;
            if (UPS) {
                int sy0, sx0;
                window_origin(Y0, X0, sy0, sx0);
                // neighbour = + 1 unless at the last row / column; weights from the fraction.  The four weights are multiplied
                // out first (ATen nests them); the difference is an fp32 rounding, far below the TF32 rounding that follows.
                const float h1r = sh.ry * (float)Y, w1r = sh.rx * (float)X;
                const int h1 = (int)h1r, w1 = (int)w1r;
                const int h1p = h1 < sh.h - 1 ? 1 : 0, w1p = w1 < sh.w - 1 ? 1 : 0;
                const float h1l = h1r - (float)h1, h0l = 1.f - h1l, w1l = w1r - (float)w1, w0l = 1.f - w1l;
                const float wa = h0l * w0l, wb = h0l * w1l, wc = h1l * w0l, wd = h1l * w1l;
                const int cell = min(max(h1 - sy0, 0), TC_WIN_H - 2) * TC_WIN_W + min(max(w1 - sx0, 0), TC_WIN_W - 2);
                const unsigned a00 = win + cell * 4 + half_win, a01 = a00 + w1p * 4, a10 = a00 + h1p * TC_WIN_W * 4, a11 = a10 + w1p * 4;
                asm volatile("bar.sync 2, %0;" ::"n"(TC_PROD_THREADS) : "memory");   // everyone is done reading the previous window
                store_window();
                asm volatile("bar.sync 2, %0;" ::"n"(TC_PROD_THREADS) : "memory");
                if (has_next) load_window(tn);
                tc_wait(b_aempty + 8 * s, ph ^ 1);         // the MMAs that read this patch stage two tiles ago are complete
#pragma unroll
                for (int g = 0; g < TC_WIN_GROUPS; ++g) {
                    if (has_pixel && !(SSTEM_TC_DIAG & 1)) {
                        // this thread's chunks of the group: 7 g + half + 2 k.  With 13 chunks (cin 49..52, FULL) that is
                        // k < 3, plus k = 3 for the even chunks of group 0; otherwise each chunk is checked against cin.
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            const int ck = g * (TC_CHUNKS / 2) + half + 2 * k;
                            const bool mine = FULL ? (k < 3 || (g == 0 && half == 0)) : (ck < (g + 1) * (TC_CHUNKS / 2) && ck < nchunks);
                            if (!mine) continue;
                            const unsigned dst = a_dst + (g * (TC_CHUNKS / 2) + 2 * k) * TC_NPIX * 16;
                            if (inside) {
                                float q[4];
#pragma unroll
                                for (int j = 0; j < 4; ++j) {
                                    const unsigned o = g * TC_WIN_GROUP_PITCH + (unsigned)(8 * k + j) * (TC_WIN_CH * 4);
                                    q[j] = to_tf32(wa * lds_f32(a00 + o) + wb * lds_f32(a01 + o) + wc * lds_f32(a10 + o) + wd * lds_f32(a11 + o));
                                }
                                sts128(dst, q[0], q[1], q[2], q[3]);
                            } else {
                                sts128(dst, 0.f, 0.f, 0.f, 0.f);
                            }
                        }
                    }
                }
            } else {
                const float* xb = x + (int64_t)b * sh.cin * src_plane;
                tc_wait(b_aempty + 8 * s, ph ^ 1);
                if (has_pixel) {
                    const float* xp = xb + (int64_t)min(max(Y, 0), sh.H - 1) * sh.W + min(max(X, 0), sh.W - 1);
#pragma unroll 1
                    for (int ck0 = half; ck0 < nchunks; ck0 += 8) {    // chunks half, half + 2, ...: 16 independent loads in flight per thread
                        const unsigned dst0 = a_dst - half_dst;
                        float q[16];
#pragma unroll
                        for (int j = 0; j < 16; ++j) {
                            const int ch = (ck0 + 2 * (j >> 2)) * 4 + (j & 3);
                            q[j] = (inside && ch < sh.cin) ? __ldg(xp + (int64_t)ch * src_plane) : 0.f;
                        }
#pragma unroll
                        for (int c4 = 0; c4 < 4; ++c4)
                            if (ck0 + 2 * c4 < nchunks)
                                sts128(dst0 + (ck0 + 2 * c4) * TC_NPIX * 16, to_tf32(q[4 * c4]), to_tf32(q[4 * c4 + 1]), to_tf32(q[4 * c4 + 2]), to_tf32(q[4 * c4 + 3]));
                    }
                }
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy stores -> visible to the tensor core's reads
            tc_arrive(b_afull + 8 * s);
        }
    }

    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(TC_TMEM_COLS));
}

// weights [cout][cin][3][3] -> [tap][chunk < 13][n = 64][4] + one zero chunk, zero padded, rounded to TF32
__global__ void tap_conv3x3_pack_kernel(const float* __restrict__ w, float* __restrict__ packed, int cin, int cout) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (int)(TC_W_BYTES / 4)) return;
    if (i >= (int)(TC_W_ZERO_OFF / 4)) { packed[i] = 0.f; return; }
    const int j = i & 3, n = (i >> 2) % TC_N, ck = (i >> 2) / TC_N % TC_WCHUNKS, tap = (i >> 2) / TC_N / TC_WCHUNKS;
    const int ch = ck * 4 + j;
    packed[i] = (n < cout && ch < cin) ? to_tf32(w[((int64_t)n * cin + ch) * 9 + tap]) : 0.f;
}

template <bool UPS, bool TILED, bool WVEC, bool FULL>
int launch_tap_conv(const float* x, const float* wpacked, const float* bias, float* out, const TapConvShape& sh, cudaStream_t s) {
    static PerDeviceOnce done;
    auto kern = tap_conv3x3_kernel<UPS, TILED, WVEC, FULL>;
    int dev = 0;
    cudaGetDevice(&dev);
    if (!done.test(dev)) {
        if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TC_SMEM) != cudaSuccess) return (int)cudaGetLastError();
        done.set(dev);
    }
    const int ctas = std::min(sh.ntiles, sm_count());
    kern<<<ctas, TC_THREADS, TC_SMEM, s>>>(x, wpacked, bias, out, sh);
    count_launch();
    return finish_launch();
}

}  // namespace

}  // namespace sstem

using namespace sstem;

extern "C" int64_t sstem_tap_conv3x3_packed_elems(void) { return (int64_t)TC_W_BYTES / 4; }

extern "C" int sstem_tap_conv3x3_pack_weights(const float* weight, float* packed, int32_t cin, int32_t cout, void* stream) {
    if (!weight || !packed) return SSTEM_E_NULL;
    if (cin <= 0 || cin > TC_MAXC || cout <= 0 || cout > TC_N) return SSTEM_E_SHAPE;
    if (!aligned4(weight) || !aligned16(packed)) return SSTEM_E_ALIGN;
    DeviceGuard guard(packed);
    if (guard.err) return guard.err;
    const int total = (int)(TC_W_BYTES / 4);
    tap_conv3x3_pack_kernel<<<(total + 255) / 256, 256, 0, (cudaStream_t)stream>>>(weight, packed, cin, cout);
    count_launch();
    return finish_launch();
}

extern "C" int sstem_tap_conv3x3(const float* x, const float* packed_weight, const float* bias, float* out,
                                 int64_t B, int32_t cin, int32_t cout, int64_t h, int64_t w, uint32_t flags, void* stream) {
    if (!x || !packed_weight || !out) return SSTEM_E_NULL;
    if (flags & ~(SSTEM_TAPCONV_UPSAMPLE2X | SSTEM_TAPCONV_TILED)) return SSTEM_E_FLAG;
    const bool ups = flags & SSTEM_TAPCONV_UPSAMPLE2X, tiled = flags & SSTEM_TAPCONV_TILED;
    if (B <= 0 || h <= 0 || w <= 0 || h > (1 << 22) || w > (1 << 22)) return SSTEM_E_SHAPE;
    if (cin <= 0 || cin > TC_MAXC || cout <= 0 || cout > TC_N) return SSTEM_E_SHAPE;
    if (tiled && cout != 51) return SSTEM_E_SHAPE;         // the tile-major layout is the 51-tap consumer's
    if (!aligned4(x) || !aligned16(packed_weight) || !aligned4(out) || (bias && !aligned4(bias))) return SSTEM_E_ALIGN;
    DeviceGuard guard(out);
    if (guard.err) return guard.err;
    TapConvShape sh;
    sh.B = (int)B; sh.cin = cin; sh.cout = cout; sh.h = (int)h; sh.w = (int)w;
    sh.H = ups ? 2 * (int)h : (int)h;
    sh.W = ups ? 2 * (int)w : (int)w;
    sh.tiles_x = (sh.W + TC_TW - 1) / TC_TW;
    sh.tiles_y = (sh.H + TC_TH - 1) / TC_TH;
    const int64_t nt = (int64_t)sh.tiles_x * sh.tiles_y * B;
    if (nt > INT32_MAX / 2 || (int64_t)cin * h * w > INT32_MAX) return SSTEM_E_SHAPE;
    sh.ntiles = (int)nt;
    sh.k_steps = (cin + 7) / 8;
    sh.ry = sh.H > 1 ? (float)(sh.h - 1) / (float)(sh.H - 1) : 0.f;
    sh.rx = sh.W > 1 ? (float)(sh.w - 1) / (float)(sh.W - 1) : 0.f;
    cudaStream_t s = (cudaStream_t)stream;
    const bool full = cin > 48;                            // all 13 channel chunks exist: the common (51-channel) case, no per-chunk checks
#define SSTEM_TC_LAUNCH(UPS_, WVEC_)                                                                                       \
    return full ? (tiled ? launch_tap_conv<UPS_, true, WVEC_, true>(x, packed_weight, bias, out, sh, s)                    \
                         : launch_tap_conv<UPS_, false, WVEC_, true>(x, packed_weight, bias, out, sh, s))                  \
                : (tiled ? launch_tap_conv<UPS_, true, WVEC_, false>(x, packed_weight, bias, out, sh, s)                   \
                         : launch_tap_conv<UPS_, false, WVEC_, false>(x, packed_weight, bias, out, sh, s))
    if (ups) {
        if ((w % 4 == 0) && aligned16(x)) { SSTEM_TC_LAUNCH(true, true); }
        SSTEM_TC_LAUNCH(true, false);
    }
    SSTEM_TC_LAUNCH(false, false);
#undef SSTEM_TC_LAUNCH
}
