// hs_common.cuh -- complex type, load/store helpers and the execution-model shim.
//
// The kernels are written against a handful of macros (HS_TID, HS_SYNC, ...) so the very same
// kernel bodies can be compiled by g++ for tests/emu (one OS thread per CUDA thread, a barrier for
// __syncthreads) to check index arithmetic without a GPU.  The emulation is test infrastructure;
// the product is always the nvcc build (HS_EMU undefined).
#pragma once

#ifdef HS_EMU
// ---------------------------------------------------------------- host emulation (tests only)
#include <cmath>
#include <cstdint>
#include <cstring>
struct double2 { double x, y; };
static inline double2 make_double2(double x, double y) { double2 r; r.x = x; r.y = y; return r; }
struct hs_dim3 { unsigned x = 1, y = 1, z = 1; };
extern thread_local hs_dim3 hs_emu_threadIdx;
extern thread_local hs_dim3 hs_emu_blockIdx;
extern hs_dim3 hs_emu_blockDim, hs_emu_gridDim;
extern thread_local unsigned char *hs_emu_smem;
void hs_emu_barrier();
#define HS_GLOBAL
#define HS_DEVICE inline
#define HS_HD inline
#define HS_LAUNCH_BOUNDS(t, b)
#define HS_RESTRICT
#define HS_TID ((int)hs_emu_threadIdx.x)
#define HS_BID ((long long)hs_emu_blockIdx.x)
#define HS_NTHREADS ((int)hs_emu_blockDim.x)
#define HS_NBLOCKS ((long long)hs_emu_gridDim.x)
#define HS_SYNC() hs_emu_barrier()
#define HS_DYN_SMEM(type, name) type *name = reinterpret_cast<type *>(hs_emu_smem)
#define HS_UNROLL
static inline double2 hs_ld_stream(const double2 *p) { return *p; }
static inline void hs_st_stream(double2 *p, double2 v) { *p = v; }
static inline double2 hs_ld_tab(const double2 *p) { return *p; }
static inline double hs_ld_stream_d(const double *p) { return *p; }
static inline void hs_st_stream_d(double *p, double v) { *p = v; }
#else
// ---------------------------------------------------------------------------- CUDA (product)
#include <cuda_runtime.h>
#include <cstdint>
#define HS_GLOBAL __global__
#define HS_DEVICE __device__ __forceinline__
#define HS_HD __host__ __device__ __forceinline__
#define HS_LAUNCH_BOUNDS(t, b) __launch_bounds__(t, b)
#define HS_RESTRICT __restrict__
#define HS_TID ((int)threadIdx.x)
#define HS_BID ((long long)blockIdx.x)
#define HS_NTHREADS ((int)blockDim.x)
#define HS_NBLOCKS ((long long)gridDim.x)
#define HS_SYNC() __syncthreads()
#define HS_DYN_SMEM(type, name)                                  \
    extern __shared__ __align__(16) unsigned char hs_smem_raw[]; \
    type *name = reinterpret_cast<type *>(hs_smem_raw)
#define HS_UNROLL _Pragma("unroll")
// Signal samples are touched exactly once per pass: keep them out of L1 (no_allocate) so that L1
// holds the stage-twiddle / chirp tables, which every CTA on the SM re-reads (evict_last).
HS_DEVICE double2 hs_ld_stream(const double2 *p)
{
    double2 v;
    asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p));
    return v;
}
HS_DEVICE void hs_st_stream(double2 *p, double2 v)
{
    asm volatile("st.global.L1::no_allocate.v2.f64 [%0], {%1, %2};" ::"l"(p), "d"(v.x), "d"(v.y) : "memory");
}
HS_DEVICE double2 hs_ld_tab(const double2 *p)
{
    double2 v;
    asm volatile("ld.global.nc.L1::evict_last.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p));
    return v;
}
HS_DEVICE double hs_ld_stream_d(const double *p) { return __ldcs(p); }
HS_DEVICE void hs_st_stream_d(double *p, double v) { __stcs(p, v); }
#endif

typedef double2 cplx; // x = re, y = im; 16-byte interleaved, same bytes as the API's fft_data

HS_HD cplx hs_c(double re, double im) { return make_double2(re, im); }
HS_HD cplx hs_add(cplx a, cplx b) { return make_double2(a.x + b.x, a.y + b.y); }
HS_HD cplx hs_sub(cplx a, cplx b) { return make_double2(a.x - b.x, a.y - b.y); }
// x * w, written like the reference's combine loops (src/highSpeedFFT.c:741-742)
HS_HD cplx hs_mul(cplx a, cplx w) { return make_double2(a.x * w.x - a.y * w.y, a.y * w.x + a.x * w.y); }
// x * conj(w)
HS_HD cplx hs_mulc(cplx a, cplx w) { return make_double2(a.x * w.x + a.y * w.y, a.y * w.x - a.x * w.y); }
HS_HD cplx hs_conj(cplx a) { return make_double2(a.x, -a.y); }
HS_HD cplx hs_scale(cplx a, double s) { return make_double2(a.x * s, a.y * s); }
// multiply by -i (forward quarter turn) when SGN=+1, by +i when SGN=-1
template <int SGN>
HS_HD cplx hs_rot(cplx a) { return SGN == 1 ? make_double2(a.y, -a.x) : make_double2(-a.y, a.x); }
