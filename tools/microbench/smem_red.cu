// Shared-memory accumulate: red.shared.add.f32 vs LDS + FADD + STS, distinct words per lane (with the 2-way bank overlap of
// the grad_input scatter pattern).  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o smem_red smem_red.cu
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE>
__global__ void __launch_bounds__(128, 2) k(float* out, int iters) {
    __shared__ float sm[4096];
    for (int i = threadIdx.x; i < 4096; i += blockDim.x) sm[i] = 0.f;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, pg = lane >> 2, g = lane & 3;
    float v = 1.0f + lane;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int t = 0; t < 13; ++t) {
            const int k = (t + 2 * g) % 13;
            float* w = sm + warp * 1024 + (it & 7) * 84 + pg + g + 4 * k;
            if (MODE == 0) { asm volatile("red.shared.add.f32 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(w)), "f"(v) : "memory"); }
            else { float o = *(volatile float*)w; *(volatile float*)w = o + v; __syncwarp(); }
            v += 0.5f;
        }
    }
    long long t1 = clock64();
    __syncthreads();
    if (threadIdx.x == 0) out[blockIdx.x] = (float)(t1 - t0) / (iters * 13.0f) + sm[5] * 0.f;
}
int main() {
    float* d; cudaMalloc(&d, 4096);
    for (int mode = 0; mode < 2; ++mode) {
        if (mode == 0) k<0><<<296, 128>>>(d, 2000); else k<1><<<296, 128>>>(d, 2000);
        cudaDeviceSynchronize();
        float h[296]; cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
        double s = 0; for (float x : h) s += x;
        printf("{\"mode\": \"%s\", \"cycles_per_warp_update_instr\": %.2f}\n", mode == 0 ? "red.shared.add.f32" : "lds+fadd+sts+syncwarp", s / 296);
    }
    return 0;
}
