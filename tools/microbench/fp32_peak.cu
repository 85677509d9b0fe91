// FP32 pipe microbenchmarks for B200 (sm_100a): establishes the FMA roofline
// denominator used by bench.py and answers design questions for the sepconv
// kernels (3-register FFMA rate, packed FFMA2 rate, FFMA+LDS co-issue).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp32_peak fp32_peak.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

constexpr int NACC = 16;

// 1. plain FFMA, 3 distinct register sources, NACC independent chains
__global__ void __launch_bounds__(256) k_ffma(float* out, const float* in, int iters) {
    float acc[NACC], a[NACC], b[NACC];
#pragma unroll
    for (int i = 0; i < NACC; ++i) { acc[i] = in[i]; a[i] = in[NACC + i + threadIdx.x % 2]; b[i] = in[2 * NACC + i + threadIdx.x % 3]; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 8; ++r)
#pragma unroll
            for (int i = 0; i < NACC; ++i) acc[i] = fmaf(a[i], b[(i + r) % NACC], acc[i]);
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < NACC; ++i) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// 2. packed FFMA2 (fma.rn.f32x2), NACC/2 chains of float2
__global__ void __launch_bounds__(256) k_ffma2(float* out, const float* in, int iters) {
    float2 acc[NACC], a[NACC], b[NACC];
#pragma unroll
    for (int i = 0; i < NACC; ++i) {
        acc[i] = make_float2(in[i], in[i + 1]);
        a[i] = make_float2(in[NACC + i + threadIdx.x % 2], in[NACC + i + 3]);
        b[i] = make_float2(in[2 * NACC + i + threadIdx.x % 3], in[2 * NACC + i + 5]);
    }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 8; ++r)
#pragma unroll
            for (int i = 0; i < NACC; ++i) acc[i] = __ffma2_rn(a[i], b[(i + r) % NACC], acc[i]);
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < NACC; ++i) s += acc[i].x + acc[i].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// 3. FFMA with one operand coming from shared memory every RATIO FMAs (LDS.32, conflict-free)
template <int RATIO>
__global__ void __launch_bounds__(256) k_ffma_lds(float* out, const float* in, int iters) {
    __shared__ float sm[4096];
    for (int i = threadIdx.x; i < 4096; i += blockDim.x) sm[i] = in[i % 64];
    __syncthreads();
    float acc[NACC], b[NACC];
#pragma unroll
    for (int i = 0; i < NACC; ++i) { acc[i] = in[i]; b[i] = in[2 * NACC + i + threadIdx.x % 3]; }
    int base = threadIdx.x;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            float p[NACC / RATIO];
#pragma unroll
            for (int j = 0; j < NACC / RATIO; ++j) p[j] = sm[(base + (r * 16 + j) * 32 + it) & 4095];
#pragma unroll
            for (int i = 0; i < NACC; ++i) acc[i] = fmaf(p[i / RATIO], b[(i + r) % NACC], acc[i]);
        }
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < NACC; ++i) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// 4. FFMA2 with LDS.64 operand every RATIO packed FMAs
template <int RATIO>
__global__ void __launch_bounds__(256) k_ffma2_lds(float* out, const float* in, int iters) {
    __shared__ float2 sm[2048];
    for (int i = threadIdx.x; i < 2048; i += blockDim.x) sm[i] = make_float2(in[i % 64], in[(i + 1) % 64]);
    __syncthreads();
    float2 acc[NACC], b[NACC];
#pragma unroll
    for (int i = 0; i < NACC; ++i) { acc[i] = make_float2(in[i], in[i + 1]); b[i] = make_float2(in[2 * NACC + i + threadIdx.x % 3], in[i + 7]); }
    int base = threadIdx.x;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            float2 p[NACC / RATIO];
#pragma unroll
            for (int j = 0; j < NACC / RATIO; ++j) p[j] = sm[(base + (r * 16 + j) * 32 + it) & 2047];
#pragma unroll
            for (int i = 0; i < NACC; ++i) acc[i] = __ffma2_rn(p[i / RATIO], b[(i + r) % NACC], acc[i]);
        }
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < NACC; ++i) s += acc[i].x + acc[i].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// 5. FFMA where the multiplicand register is shared by 4 consecutive FMAs (operand reuse cache)
__global__ void __launch_bounds__(256) k_ffma_reuse(float* out, const float* in, int iters) {
    float acc[NACC], a[4], b[NACC];
#pragma unroll
    for (int i = 0; i < NACC; ++i) { acc[i] = in[i]; b[i] = in[2 * NACC + i + threadIdx.x % 3]; }
#pragma unroll
    for (int i = 0; i < 4; ++i) a[i] = in[NACC + i + threadIdx.x % 2];
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 8; ++r)
#pragma unroll
            for (int i = 0; i < NACC; ++i) acc[i] = fmaf(a[(i / 4 + r) % 4], b[i], acc[i]);
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < NACC; ++i) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename F>
static void run(const char* name, F launch, double fma_per_thread_iter, int iters, int blocks, int threads) {
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    launch(iters / 8);  // warm-up
    CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int rep = 0; rep < 5; ++rep) {
        CK(cudaEventRecord(e0));
        launch(iters);
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        if (ms < best) best = ms;
    }
    double fma = fma_per_thread_iter * iters * (double)blocks * threads;
    double tflops = 2.0 * fma / (best * 1e-3) / 1e12;
    printf("{\"bench\": \"%s\", \"blocks\": %d, \"threads\": %d, \"ms\": %.4f, \"tflops\": %.2f}\n", name, blocks, threads, best, tflops);
    fflush(stdout);
}

int main(int argc, char** argv) {
    int dev = 0; CK(cudaSetDevice(dev));
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, dev));
    int clk = 0; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, dev);
    printf("{\"device\": \"%s\", \"sms\": %d, \"clock_khz_attr\": %d}\n", p.name, p.multiProcessorCount, clk);
    float *in, *out;
    CK(cudaMalloc(&in, 1 << 20)); CK(cudaMalloc(&out, 1 << 26));
    float* h = (float*)malloc(1 << 20);
    for (int i = 0; i < (1 << 18); ++i) h[i] = 1e-3f * (float)((i * 2654435761u) % 1000) - 0.5f;
    CK(cudaMemcpy(in, h, 1 << 20, cudaMemcpyHostToDevice));
    const int iters = 4096;
    for (int bps = 1; bps <= 8; bps *= 2) {
        int blocks = p.multiProcessorCount * bps, threads = 256;
        run("ffma_3reg", [&](int it) { k_ffma<<<blocks, threads>>>(out, in, it); }, 8.0 * NACC, iters, blocks, threads);
        run("ffma2_packed", [&](int it) { k_ffma2<<<blocks, threads>>>(out, in, it); }, 2 * 8.0 * NACC, iters, blocks, threads);
    }
    int blocks = p.multiProcessorCount * 4, threads = 256;
    run("ffma_reuse4", [&](int it) { k_ffma_reuse<<<blocks, threads>>>(out, in, it); }, 8.0 * NACC, iters, blocks, threads);
    run("ffma_lds_1per16", [&](int it) { k_ffma_lds<16><<<blocks, threads>>>(out, in, it); }, 8.0 * NACC, iters, blocks, threads);
    run("ffma_lds_1per8", [&](int it) { k_ffma_lds<8><<<blocks, threads>>>(out, in, it); }, 8.0 * NACC, iters, blocks, threads);
    run("ffma_lds_1per4", [&](int it) { k_ffma_lds<4><<<blocks, threads>>>(out, in, it); }, 8.0 * NACC, iters, blocks, threads);
    run("ffma_lds_1per2", [&](int it) { k_ffma_lds<2><<<blocks, threads>>>(out, in, it); }, 8.0 * NACC, iters, blocks, threads);
    run("ffma2_lds64_1per16", [&](int it) { k_ffma2_lds<16><<<blocks, threads>>>(out, in, it); }, 2 * 8.0 * NACC, iters, blocks, threads);
    run("ffma2_lds64_1per8", [&](int it) { k_ffma2_lds<8><<<blocks, threads>>>(out, in, it); }, 2 * 8.0 * NACC, iters, blocks, threads);
    run("ffma2_lds64_1per4", [&](int it) { k_ffma2_lds<4><<<blocks, threads>>>(out, in, it); }, 2 * 8.0 * NACC, iters, blocks, threads);
    run("ffma2_lds64_1per2", [&](int it) { k_ffma2_lds<2><<<blocks, threads>>>(out, in, it); }, 2 * 8.0 * NACC, iters, blocks, threads);
    CK(cudaDeviceSynchronize());
    return 0;
}
