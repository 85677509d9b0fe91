// FP32 pipe microbenchmarks, round 2: operand-reuse patterns and true SM clock.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

__device__ __forceinline__ unsigned long long clk_after(float dep) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%clock64;" : "=l"(t) : "f"(dep) : "memory");
    return t;
}

template <int NACC, int REUSE>
__global__ void __launch_bounds__(256) k_ffma(float* out, const float* in, int iters, unsigned long long* cyc) {
    float acc[NACC], a[4], b[NACC];
#pragma unroll
    for (int i = 0; i < NACC; ++i) { acc[i] = in[i]; b[i] = in[64 + i + threadIdx.x % 3]; }
#pragma unroll
    for (int i = 0; i < 4; ++i) a[i] = in[32 + i + threadIdx.x % 2];
    unsigned long long t0 = clk_after(acc[0]);
    a[0] += (t0 == 123ull) ? 1.f : 0.f;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int i = 0; i < NACC; ++i) acc[i] = fmaf(a[(i / REUSE + r) % 4], b[i], acc[i]);
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < NACC; ++i) s += acc[i];
    unsigned long long t1 = clk_after(s);
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int NACC, int REUSE>
__global__ void __launch_bounds__(256) k_ffma2(float* out, const float* in, int iters, unsigned long long* cyc) {
    float2 acc[NACC], a[4], b[NACC];
#pragma unroll
    for (int i = 0; i < NACC; ++i) { acc[i] = make_float2(in[i], in[i + 1]); b[i] = make_float2(in[64 + i + threadIdx.x % 3], in[96 + i]); }
#pragma unroll
    for (int i = 0; i < 4; ++i) { float x = in[32 + i + threadIdx.x % 2]; a[i] = make_float2(x, x); }
    unsigned long long t0 = clk_after(acc[0].x);
    a[0].x += (t0 == 123ull) ? 1.f : 0.f;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int i = 0; i < NACC; ++i) acc[i] = __ffma2_rn(a[(i / REUSE + r) % 4], b[i], acc[i]);
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < NACC; ++i) s += acc[i].x + acc[i].y;
    unsigned long long t1 = clk_after(s);
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

// sepconv-like step: P from smem (LDS.32), dup, 4 pixel-pairs x NT taps FFMA2, h2 resident
template <int NT>
__global__ void __launch_bounds__(128, 2) k_sepstep(float* out, const float* in, int iters, unsigned long long* cyc) {
    __shared__ float sm[3 * 32 * 84];
    for (int i = threadIdx.x; i < 3 * 32 * 84; i += blockDim.x) sm[i] = in[i % 200];
    __syncthreads();
    float2 h2[4][NT], acc[3][4];
#pragma unroll
    for (int p = 0; p < 4; ++p)
#pragma unroll
        for (int t = 0; t < NT; ++t) h2[p][t] = make_float2(in[p * NT + t], in[p * NT + t + 60]);
#pragma unroll
    for (int c = 0; c < 3; ++c)
#pragma unroll
        for (int p = 0; p < 4; ++p) acc[c][p] = make_float2(0.f, 0.f);
    const int lane = threadIdx.x & 31, pg = lane >> 2, g = lane & 3;
    unsigned long long t0 = clk_after(h2[0][0].x);
    float2 v2[4];
#pragma unroll
    for (int p = 0; p < 4; ++p) v2[p] = make_float2(in[300 + p] + ((t0 == 123ull) ? 1.f : 0.f), in[310 + p]);
    for (int it = 0; it < iters; ++it) {
        const int s = it & 31;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const float* prow = sm + (c * 32 + s) * 84 + pg + g;
            float2 part[4];
#pragma unroll
            for (int p = 0; p < 4; ++p) part[p] = make_float2(0.f, 0.f);
#pragma unroll
            for (int t = 0; t < NT; ++t) {
                const float P = prow[4 * t];
                const float2 Pd = make_float2(P, P);
#pragma unroll
                for (int p = 0; p < 4; ++p) part[p] = __ffma2_rn(Pd, h2[p][t], part[p]);
            }
#pragma unroll
            for (int p = 0; p < 4; ++p) acc[c][p] = __ffma2_rn(v2[p], part[p], acc[c][p]);
        }
    }
    float s = 0.f;
#pragma unroll
    for (int c = 0; c < 3; ++c)
#pragma unroll
        for (int p = 0; p < 4; ++p) s += acc[c][p].x + acc[c][p].y;
    unsigned long long t1 = clk_after(s);
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

// same with scalar FFMA (8 pixels)
template <int NT>
__global__ void __launch_bounds__(128, 2) k_sepstep_scalar(float* out, const float* in, int iters, unsigned long long* cyc) {
    __shared__ float sm[3 * 32 * 84];
    for (int i = threadIdx.x; i < 3 * 32 * 84; i += blockDim.x) sm[i] = in[i % 200];
    __syncthreads();
    float h[8][NT], acc[3][8];
#pragma unroll
    for (int p = 0; p < 8; ++p)
#pragma unroll
        for (int t = 0; t < NT; ++t) h[p][t] = in[p * NT + t];
#pragma unroll
    for (int c = 0; c < 3; ++c)
#pragma unroll
        for (int p = 0; p < 8; ++p) acc[c][p] = 0.f;
    const int lane = threadIdx.x & 31, pg = lane >> 2, g = lane & 3;
    unsigned long long t0 = clk_after(h[0][0]);
    float v[8];
#pragma unroll
    for (int p = 0; p < 8; ++p) v[p] = in[300 + p] + ((t0 == 123ull) ? 1.f : 0.f);
    for (int it = 0; it < iters; ++it) {
        const int s = it & 31;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const float* prow = sm + (c * 32 + s) * 84 + pg + g;
            float part[8];
#pragma unroll
            for (int p = 0; p < 8; ++p) part[p] = 0.f;
#pragma unroll
            for (int t = 0; t < NT; ++t) {
                const float P = prow[4 * t];
#pragma unroll
                for (int p = 0; p < 8; ++p) part[p] = fmaf(P, h[p][t], part[p]);
            }
#pragma unroll
            for (int p = 0; p < 8; ++p) acc[c][p] = fmaf(v[p], part[p], acc[c][p]);
        }
    }
    float s = 0.f;
#pragma unroll
    for (int c = 0; c < 3; ++c)
#pragma unroll
        for (int p = 0; p < 8; ++p) s += acc[c][p];
    unsigned long long t1 = clk_after(s);
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <typename F>
static void run(const char* name, F launch, double fma_per_thread_iter, int iters, int blocks, int threads, unsigned long long* dcyc) {
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    launch(iters); launch(iters);
    CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int rep = 0; rep < 5; ++rep) {
        CK(cudaEventRecord(e0)); launch(iters); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        if (ms < best) best = ms;
    }
    unsigned long long hc[4096];
    CK(cudaMemcpy(hc, dcyc, sizeof(unsigned long long) * blocks, cudaMemcpyDeviceToHost));
    double mean = 0; for (int i = 0; i < blocks; ++i) mean += hc[i]; mean /= blocks;
    double fma = fma_per_thread_iter * iters * (double)blocks * threads;
    double tflops = 2.0 * fma / (best * 1e-3) / 1e12;
    // FMA lane-ops per clock per SM from in-kernel cycles (blocks/148 CTAs per SM concurrently)
    double fma_per_clk_sm = fma_per_thread_iter * iters * threads * (blocks / 148.0) / mean;
    printf("{\"bench\": \"%s\", \"blocks\": %d, \"threads\": %d, \"ms\": %.4f, \"tflops\": %.2f, \"cycles\": %.0f, \"mhz_est\": %.0f, \"fma_per_clk_sm\": %.1f}\n",
           name, blocks, threads, best, tflops, mean, mean / (best * 1e-3) / 1e6, fma_per_clk_sm);
    fflush(stdout);
}

int main() {
    CK(cudaSetDevice(0));
    float *in, *out; unsigned long long* cyc;
    CK(cudaMalloc(&in, 1 << 20)); CK(cudaMalloc(&out, 1 << 26)); CK(cudaMalloc(&cyc, 4096 * 8));
    float* h = (float*)malloc(1 << 20);
    for (int i = 0; i < (1 << 18); ++i) h[i] = 1e-3f * (float)((i * 2654435761u) % 1000) - 0.5f;
    CK(cudaMemcpy(in, h, 1 << 20, cudaMemcpyHostToDevice));
    const int it = 4096;
    int blocks = 148 * 4, threads = 256;
#define RUN_FFMA(N, R) run("ffma_n" #N "_reuse" #R, [&](int i) { k_ffma<N, R><<<blocks, threads>>>(out, in, i, cyc); }, 4.0 * N, it, blocks, threads, cyc)
#define RUN_FFMA2(N, R) run("ffma2_n" #N "_reuse" #R, [&](int i) { k_ffma2<N, R><<<blocks, threads>>>(out, in, i, cyc); }, 8.0 * N, it, blocks, threads, cyc)
    RUN_FFMA(8, 1); RUN_FFMA(8, 2); RUN_FFMA(8, 4); RUN_FFMA(8, 8);
    RUN_FFMA(16, 1); RUN_FFMA(16, 4); RUN_FFMA(16, 8); RUN_FFMA(16, 16);
    RUN_FFMA(24, 8); RUN_FFMA(32, 8);
    RUN_FFMA2(4, 4); RUN_FFMA2(8, 1); RUN_FFMA2(8, 4); RUN_FFMA2(8, 8); RUN_FFMA2(16, 4); RUN_FFMA2(16, 8);
    blocks = 148 * 2; threads = 128;
    run("sepstep_ffma2_nt13", [&](int i) { k_sepstep<13><<<blocks, threads>>>(out, in, i, cyc); }, 3.0 * (13 * 8 + 8), 2048, blocks, threads, cyc);
    run("sepstep_ffma2_nt12", [&](int i) { k_sepstep<12><<<blocks, threads>>>(out, in, i, cyc); }, 3.0 * (12 * 8 + 8), 2048, blocks, threads, cyc);
    run("sepstep_scalar_nt13", [&](int i) { k_sepstep_scalar<13><<<blocks, threads>>>(out, in, i, cyc); }, 3.0 * (13 * 8 + 8), 2048, blocks, threads, cyc);
    blocks = 148 * 1;
    run("sepstep_ffma2_nt13_1cta", [&](int i) { k_sepstep<13><<<blocks, threads>>>(out, in, i, cyc); }, 3.0 * (13 * 8 + 8), 2048, blocks, threads, cyc);
    return 0;
}
