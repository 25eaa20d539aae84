// hs_cuda_util.h -- error check shared by the CUDA translation units (hs_device.cu, hs_pipe_inst.cu).
#pragma once
#include <cstdio>
#include <cstdlib>
#include <cuda.h>
#include <cuda_runtime.h>
#include "hs_abi.h"

#define HS_CUDA_CHECK(expr)                                                                            \
    do {                                                                                               \
        cudaError_t e_ = (expr);                                                                       \
        if (e_ != cudaSuccess) {                                                                       \
            fprintf(stderr, "Error: CUDA failure %s at %s:%d (%s)\n", cudaGetErrorString(e_), __FILE__, \
                    __LINE__, #expr);                                                                  \
            exit(EXIT_FAILURE);                                                                        \
        }                                                                                              \
    } while (0)

extern "C" void hs_cu_timing_mark(const char *tag, void *stream); // per-launch timing hook (no-op unless bench.py turned it on)
extern "C" void hs_cu_count_launch(void); // bumps the library's kernel-launch counter (hs_device.cu)

// ---- tensor maps for the TMA-staged tiles (cuTensorMapEncodeTiled through the runtime's driver entry point)
typedef CUresult (*hs_encode_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                 const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                 CUtensorMapFloatOOBfill);
static inline hs_encode_fn hs_encode()
{
    static hs_encode_fn fn = nullptr;
    if (!fn) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        HS_CUDA_CHECK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q));
        if (!p || q != cudaDriverEntryPointSuccess) { fprintf(stderr, "Error: cuTensorMapEncodeTiled is not available\n"); exit(EXIT_FAILURE); }
        fn = (hs_encode_fn)p;
    }
    return fn;
}

// tensor map of one pass's input: [howmany][rows][2*ncols doubles], row stride in_e, item stride in_xf
static inline CUtensorMap make_tmap(const void *in, const hs_tile_geom &g, int rows, int C, long long howmany)
{
    CUtensorMap m;
    const int boxr = rows < 256 ? rows : 256;
    cuuint64_t dims[3] = {(cuuint64_t)2 * g.ncols, (cuuint64_t)rows, (cuuint64_t)howmany};
    cuuint64_t strides[2] = {(cuuint64_t)g.in_e * 16, (cuuint64_t)g.in_xf * 16};
    cuuint32_t box[3] = {(cuuint32_t)2 * C, (cuuint32_t)boxr, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = hs_encode()(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, const_cast<void *>(in), dims, strides, box, estr,
                             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { fprintf(stderr, "Error: cuTensorMapEncodeTiled failed (%d)\n", (int)r); exit(EXIT_FAILURE); }
    return m;
}

