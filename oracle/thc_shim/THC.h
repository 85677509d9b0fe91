/*
 * ORACLE -- TEST INFRASTRUCTURE ONLY.
 * Minimal stand-in for the PyTorch-0.4 THC headers, just enough for the
 * reference's libs/sepconv/src/SeparableConvolution_kernel.cu to compile
 * UNMODIFIED for sm_100a (THC no longer exists in PyTorch >= 1.0).  It provides
 * the tensor struct and the five accessors/macros that file uses
 * (kernel.cu:62-72,160-205).  Written from the call sites, not from THC.
 */
#pragma once
#include <cuda_runtime.h>
#include <stdio.h>

typedef struct THCudaTensor {
    float* data;
    long size[4];
    long stride[4];
} THCudaTensor;

typedef struct THCState {
    cudaStream_t stream;
    int last_error;
} THCState;

static inline long THCudaTensor_nElement(THCState*, THCudaTensor* t) {
    return t->size[0] * t->size[1] * t->size[2] * t->size[3];
}
static inline float* THCudaTensor_data(THCState*, THCudaTensor* t) { return t->data; }
static inline cudaStream_t THCState_getCurrentStream(THCState* s) { return s->stream; }

#define THCudaCheck(expr)                                                        \
    do {                                                                         \
        cudaError_t thc_shim_err = (expr);                                       \
        if (thc_shim_err != cudaSuccess) {                                       \
            fprintf(stderr, "THCudaCheck(shim): %s\n", cudaGetErrorString(thc_shim_err)); \
            state->last_error = (int)thc_shim_err;                               \
        }                                                                        \
    } while (0)
