// FP32 FMA-pipe probe: the measured denominator of the sepconv roofline
// (MEASURED_PEAKS.json carries only HBM and bf16-GEMM rows).
//
// Register-resident FMA loops, one multiplicand shared by REUSE consecutive FMAs (the
// operand pattern of the sepconv inner loops; with no operand reuse the register file
// feeds the pipe at ~2/3 rate), 4 CTAs of 256 threads per SM, one wave.  Three variants
// run (scalar FFMA with reuse 16 and 8, packed FFMA2 with reuse 8) and the best rate is
// reported: the result is sensitive to ptxas' register-bank assignment, and the chip is
// power-managed -- a trivial loop like this sustains a higher clock than a kernel that
// also moves data, so this is an upper bound, which is what a roofline denominator is.
#include "common.cuh"

namespace sstem {

__device__ __forceinline__ unsigned long long clock_after(float dep) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%clock64;" : "=l"(t) : "f"(dep) : "memory");   // ordered after `dep`
    return t;
}

template <int NACC, int REUSE>
__global__ void __launch_bounds__(256) probe_ffma(float* out, const float* in, int iters, unsigned long long* cyc) {
    float acc[NACC], a[4], b[NACC];
#pragma unroll
    for (int i = 0; i < NACC; ++i) { acc[i] = in[i]; b[i] = in[64 + i + threadIdx.x % 3]; }
#pragma unroll
    for (int i = 0; i < 4; ++i) a[i] = in[32 + i + threadIdx.x % 2];
    const unsigned long long t0 = clock_after(acc[0]);
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
    const unsigned long long t1 = clock_after(s);
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int NACC, int REUSE>
__global__ void __launch_bounds__(256) probe_ffma2(float* out, const float* in, int iters, unsigned long long* cyc) {
    float2 acc[NACC], a[4], b[NACC];
#pragma unroll
    for (int i = 0; i < NACC; ++i) { acc[i] = make_float2(in[i], in[i + 1]); b[i] = make_float2(in[64 + i + threadIdx.x % 3], in[96 + i]); }
#pragma unroll
    for (int i = 0; i < 4; ++i) { const float x = in[32 + i + threadIdx.x % 2]; a[i] = make_float2(x, x); }
    const unsigned long long t0 = clock_after(acc[0].x);
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
    const unsigned long long t1 = clock_after(s);
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

}  // namespace sstem

using namespace sstem;

extern "C" int sstem_fp32_peak_probe(double* tflops_out, double* sm_mhz_out) {
    if (!tflops_out) return SSTEM_E_NULL;
    const int sms = sm_count();
    const int blocks = sms * 4, threads = 256, iters = 4096;
    float *in = nullptr, *out = nullptr;
    unsigned long long* cyc = nullptr;
    cudaError_t e;
    if ((e = cudaMalloc(&in, 256 * sizeof(float))) != cudaSuccess) return (int)e;
    if ((e = cudaMalloc(&out, (size_t)blocks * threads * sizeof(float))) != cudaSuccess) { cudaFree(in); return (int)e; }
    if ((e = cudaMalloc(&cyc, (size_t)blocks * sizeof(unsigned long long))) != cudaSuccess) { cudaFree(in); cudaFree(out); return (int)e; }
    float hin[256];
    for (int i = 0; i < 256; ++i) hin[i] = 1e-3f * (float)((i * 2654435761u) % 1000) - 0.5f;
    cudaMemcpy(in, hin, sizeof(hin), cudaMemcpyHostToDevice);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    double best_tflops = 0.0, best_mhz = 0.0;
    for (int variant = 0; variant < 3; ++variant) {
        auto launch = [&]() {
            if (variant == 0) probe_ffma<16, 16><<<blocks, threads>>>(out, in, iters, cyc);
            else if (variant == 1) probe_ffma<16, 8><<<blocks, threads>>>(out, in, iters, cyc);
            else probe_ffma2<16, 8><<<blocks, threads>>>(out, in, iters, cyc);
        };
        const double fma_per_thread = (variant == 2 ? 8.0 : 4.0) * 16 * iters;
        launch(); launch();                             // warm-up, clocks ramp
        cudaDeviceSynchronize();
        float best_ms = 1e30f;
        for (int rep = 0; rep < 5; ++rep) {
            cudaEventRecord(e0);
            launch();
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
            float ms = 0.f;
            cudaEventElapsedTime(&ms, e0, e1);
            if (ms < best_ms) best_ms = ms;
        }
        count_launch(7);
        unsigned long long hc = 0;
        cudaMemcpy(&hc, cyc, sizeof(hc), cudaMemcpyDeviceToHost);
        const double tflops = 2.0 * fma_per_thread * blocks * threads / (best_ms * 1e-3) / 1e12;
        if (tflops > best_tflops) best_tflops = tflops;
        // variant 0 (44 registers) runs as ONE wave of co-resident CTAs: its loop spans the kernel, so
        // cycles / time = the SM clock under this load
        if (variant == 0) best_mhz = (double)hc / (best_ms * 1e-3) / 1e6;
    }
    e = cudaGetLastError();
    *tflops_out = best_tflops;
    if (sm_mhz_out) *sm_mhz_out = best_mhz;
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    cudaFree(in); cudaFree(out); cudaFree(cyc);
    return (int)e;
}
