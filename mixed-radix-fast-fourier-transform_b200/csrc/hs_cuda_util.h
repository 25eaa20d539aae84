// hs_cuda_util.h -- error check shared by the CUDA translation units (hs_device.cu, hs_pipe_inst.cu).
#pragma once
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define HS_CUDA_CHECK(expr)                                                                            \
    do {                                                                                               \
        cudaError_t e_ = (expr);                                                                       \
        if (e_ != cudaSuccess) {                                                                       \
            fprintf(stderr, "Error: CUDA failure %s at %s:%d (%s)\n", cudaGetErrorString(e_), __FILE__, \
                    __LINE__, #expr);                                                                  \
            exit(EXIT_FAILURE);                                                                        \
        }                                                                                              \
    } while (0)

extern "C" void hs_cu_count_launch(void); // bumps the library's kernel-launch counter (hs_device.cu)
