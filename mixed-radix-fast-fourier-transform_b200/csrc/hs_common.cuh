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
#include <cstdio>
#include <cstdlib>
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
#define HS_NOINLINE
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
static inline int hs_ld_idx(const int *p) { return *p; }
static inline double2 hs_ld_gather(const double2 *p) { return *p; }
static inline void hs_st_stream_d(double *p, double v) { *p = v; }
// coherent (non read-only path) global load: data this same thread stored earlier in the launch
static inline double2 hs_ld_coherent(const double2 *p) { return *p; }
static inline void hs_st_keep(double2 *p, double2 v) { *p = v; }
template <int MP> static inline double2 hs_ld_mp(const double2 *p) { return *p; }
template <int MP> static inline void hs_st_mp(double2 *p, double2 v) { *p = v; }
// grid-level dataflow (persistent pipeline kernels): tickets and completion counters
static inline unsigned long long hs_ticket_take(unsigned long long *p) { return __atomic_fetch_add(p, 1ULL, __ATOMIC_ACQ_REL); }
static inline void hs_signal_add(unsigned *p) { __atomic_fetch_add(p, 1u, __ATOMIC_RELEASE); }
static inline unsigned hs_peek(const unsigned *p) { return __atomic_load_n(p, __ATOMIC_ACQUIRE); }
static inline void hs_wait_geq(const unsigned *p, unsigned need)
{
    // CTAs run one after another in the emulation: a dependency that is not already met never will be
    if (__atomic_load_n(p, __ATOMIC_ACQUIRE) < need) { fprintf(stderr, "hs_emu: pipeline dependency not met (ticket order bug)\n"); abort(); }
}
// bulk-copy engine (TMA) and mbarrier stand-ins with the hardware's phase semantics: a phase completes when the
// expected arrivals have arrived and the pending transaction bytes are back to zero; copies are done by the
// issuing thread at issue time
#include <mutex>
struct HsMbar { unsigned phase; int count, arrivals; long long tx; };
extern std::mutex hs_emu_mbar_mu;
static inline void hs_emu_mbar_check(HsMbar *m) { if (m->arrivals >= m->count && m->tx == 0) { m->phase++; m->arrivals = 0; } }
static inline void hs_mbar_init_n(HsMbar *m, int count) { m->phase = 0; m->count = count; m->arrivals = 0; m->tx = 0; }
static inline void hs_mbar_init(HsMbar *m) { hs_mbar_init_n(m, 1); }
static inline void hs_mbar_expect(HsMbar *m, unsigned bytes) { std::lock_guard<std::mutex> g(hs_emu_mbar_mu); m->tx += bytes; m->arrivals++; hs_emu_mbar_check(m); }
static inline void hs_mbar_arrive(HsMbar *m) { std::lock_guard<std::mutex> g(hs_emu_mbar_mu); m->arrivals++; hs_emu_mbar_check(m); }
static inline void hs_emu_mbar_complete_tx(HsMbar *m, long long bytes) { std::lock_guard<std::mutex> g(hs_emu_mbar_mu); m->tx -= bytes; hs_emu_mbar_check(m); }
static inline bool hs_mbar_test(HsMbar *m, unsigned parity) { std::lock_guard<std::mutex> g(hs_emu_mbar_mu); return (m->phase & 1u) != parity; }
static inline void hs_mbar_wait(HsMbar *m, unsigned parity) { while (!hs_mbar_test(m, parity)) { } }
static inline void hs_bulk_g2s(void *dst, const void *src, unsigned bytes, HsMbar *m) { memcpy(dst, src, bytes); hs_emu_mbar_complete_tx(m, bytes); }
static inline void hs_bulk_s2g(void *dst, const void *src, unsigned bytes) { memcpy(dst, src, bytes); }
// asynchronous 16-byte copy global -> shared (cp.async): done at once here
static inline void hs_cp_async16(void *dst_smem, const void *src) { memcpy(dst_smem, src, 16); }
static inline void hs_cp_async_wait_all() {}
// tensor-memory stand-in: every CUDA thread is one OS thread here, and a thread only ever reads back what it parked itself
extern thread_local double hs_emu_tmem[128];
static inline unsigned hs_tmem_alloc(unsigned *, int) { return 0u; }
static inline void hs_tmem_free(unsigned, int) {}
static inline unsigned hs_tmem_my_addr(unsigned base, int words_per_thread) { (void)words_per_thread; return base; }
static inline void hs_tmem_st4(unsigned taddr, int chunk, const double2 *v) { memcpy(hs_emu_tmem + 8 * chunk + taddr, v, 64); }
static inline void hs_tmem_ld4(unsigned taddr, int chunk, double2 *v) { memcpy(v, hs_emu_tmem + 8 * chunk + taddr, 64); }
static inline void hs_tmem_wait_st() {}
static inline void hs_tmem_wait_ld() {}
static inline void hs_bulk_commit_wait_read() {}
static inline void hs_fence_async_proxy() {}
static inline void hs_fence_proxy_async_all() {}
static inline void hs_prefetch_l2(const void *, unsigned) {}
// 3-D tensor-map stand-in: [z][row][x doubles], strides in bytes
struct HsTmap { const unsigned char *base; long long row_stride, z_stride; };
static inline void hs_tma_load_3d(void *dst, const HsTmap *t, int x_dbl, int row0, long long z, int rows, int row_bytes, HsMbar *m)
{
    for (int r = 0; r < rows; ++r)
        memcpy((unsigned char *)dst + (size_t)r * row_bytes, t->base + z * t->z_stride + (long long)(row0 + r) * t->row_stride + (long long)x_dbl * 8,
               row_bytes);
    hs_emu_mbar_complete_tx(m, (long long)rows * row_bytes);
}
template <int MP>
static inline void hs_tma_load_3d_mp(void *dst, const HsTmap *t, int x_dbl, int row0, long long z, int rows, int row_bytes, HsMbar *m)
{
    hs_tma_load_3d(dst, t, x_dbl, row0, z, rows, row_bytes, m);
}
// CTA-scope flags in shared memory between warp roles
static inline void hs_flag_add(unsigned *p) { __atomic_fetch_add(p, 1u, __ATOMIC_RELEASE); }
static inline unsigned hs_flag_read(const unsigned *p) { return __atomic_load_n(p, __ATOMIC_ACQUIRE); }
static inline void hs_flag_set(unsigned *p, unsigned v) { __atomic_store_n(p, v, __ATOMIC_RELEASE); }
#include <thread>
static inline void hs_backoff() { std::this_thread::yield(); }
void hs_emu_barrier2();
#define HS_SYNC_COMPUTE(n) hs_emu_barrier2()
#define HS_TRAP() ((void)0) /* spin counts mean nothing with OS threads; a real deadlock shows as a test timeout */
#define HS_GRID_CONSTANT
#else
// ---------------------------------------------------------------------------- CUDA (product)
#ifndef __CUDACC_RTC__ // the plan-time JIT (hs_jit.cu) compiles these headers with NVRTC, which has its own built-ins
#include <cuda_runtime.h>
#include <cstdint>
#endif
#define HS_GLOBAL __global__
#define HS_DEVICE __device__ __forceinline__
#define HS_NOINLINE __device__ __noinline__
#define HS_HD __host__ __device__ __forceinline__
#define HS_LAUNCH_BOUNDS(t, b) __launch_bounds__(t, b)
#define HS_RESTRICT __restrict__
#define HS_TID ((int)threadIdx.x)
#define HS_BID ((long long)blockIdx.x)
#define HS_NTHREADS ((int)blockDim.x)
#define HS_NBLOCKS ((long long)gridDim.x)
#define HS_SYNC() __syncthreads()
#define HS_DYN_SMEM(type, name)                                  \
    extern __shared__ __align__(128) unsigned char hs_smem_raw[]; \
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
// A value this same thread stored earlier in the launch and reads back (the first half-convolution of the split
// Bluestein kernel parks its result in the output row): the load must not take the read-only (.nc) path, the store
// asks the L2 to keep the line until it is read back.
HS_DEVICE double2 hs_ld_coherent(const double2 *p)
{
    double2 v;
    asm volatile("ld.global.cg.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p) : "memory");
    return v;
}
HS_DEVICE void hs_st_keep(double2 *p, double2 v)
{
    asm volatile("st.global.cg.v2.f64 [%0], {%1, %2};" ::"l"(p), "d"(v.x), "d"(v.y) : "memory");
}
HS_DEVICE double hs_ld_stream_d(const double *p) { return __ldcs(p); }
// index tables (re-read by every CTA) and samples gathered through them (the other half of each 32-byte sector is
// some other lane's sample a moment later: let L1 keep the line)
HS_DEVICE int hs_ld_idx(const int *p) { return __ldg(p); }
HS_DEVICE double2 hs_ld_gather(const double2 *p) { return __ldg(p); }
HS_DEVICE void hs_st_stream_d(double *p, double v) { __stcs(p, v); }

// ---- L2 residency classes for the persistent pipeline kernels (hs_pipe.cuh).  Intermediates that one
// phase writes and the next phase reads are tagged evict_last so that they live and die in the 126 MB L2
// (the ring of workspace slots is overwritten in place, its dirty lines never reach HBM); the signal
// streaming in and the spectrum streaming out are tagged evict_first so they do not push the ring out.
//   MP 0: stream, read-only path (as hs_ld_stream / hs_st_stream)
//   MP 1: keep  -- loads are coherent (ld.relaxed.gpu: another CTA of this launch wrote the data)
//   MP 2: stream + L2 evict_first
HS_DEVICE unsigned long long hs_pol_keep()
{
    unsigned long long p;
    asm("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    return p;
}
HS_DEVICE unsigned long long hs_pol_first()
{
    unsigned long long p;
    asm("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
template <int MP>
HS_DEVICE double2 hs_ld_mp(const double2 *p)
{
    double2 v;
    if (MP == 1)
        asm volatile("ld.relaxed.gpu.global.L1::no_allocate.L2::cache_hint.v2.f64 {%0, %1}, [%2], %3;"
                     : "=d"(v.x), "=d"(v.y) : "l"(p), "l"(hs_pol_keep()) : "memory");
    else if (MP == 2)
        asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v2.f64 {%0, %1}, [%2], %3;" : "=d"(v.x), "=d"(v.y) : "l"(p), "l"(hs_pol_first()));
    else
        v = hs_ld_stream(p);
    return v;
}
template <int MP>
HS_DEVICE void hs_st_mp(double2 *p, double2 v)
{
    if (MP == 1)
        asm volatile("st.global.L1::no_allocate.L2::cache_hint.v2.f64 [%0], {%1, %2}, %3;" ::"l"(p), "d"(v.x), "d"(v.y), "l"(hs_pol_keep()) : "memory");
    else if (MP == 2)
        asm volatile("st.global.L1::no_allocate.L2::cache_hint.v2.f64 [%0], {%1, %2}, %3;" ::"l"(p), "d"(v.x), "d"(v.y), "l"(hs_pol_first()) : "memory");
    else
        hs_st_stream(p, v);
}
// ---- grid-level dataflow: work tickets and completion counters in global memory
HS_DEVICE unsigned long long hs_ticket_take(unsigned long long *p) { return atomicAdd(p, 1ULL); }
// publish this CTA's global writes, then count the tile as done: a RELEASE reduction (cumulative over what the
// CTA-scope synchronisation ordered before it).  Not __threadfence() + relaxed: that fence is acq_rel, and its
// acquire half makes ptxas invalidate the SM's whole L1 (MEMBAR.SC.GPU + CCTL.IVALL) -- once per tile, that threw
// the stage-twiddle tables out of L1 and put their L2 latency on every tile's critical path (ncu, round 1).
HS_DEVICE void hs_signal_add(unsigned *p)
{
    asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(p) : "memory");
}
// Completion counters are read with ld.relaxed.gpu, not ld.acquire: an acquire at gpu scope makes ptxas
// invalidate the SM's whole L1 (CCTL.IVALL), which would throw out the twiddle tables on every poll.  The data
// the counter guards is only ever read with ld.relaxed.gpu as well (hs_ld_mp<1>: served by the L2, the point
// of coherence), after a CTA barrier that follows the poll, and was published by hs_signal_add's fence.
HS_DEVICE unsigned hs_peek(const unsigned *p)
{
    unsigned v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
// spin until *p >= need; a dependency that never completes traps instead of hanging the GPU
HS_DEVICE void hs_wait_geq(const unsigned *p, unsigned need)
{
    for (unsigned spin = 0;; ++spin) {
        const unsigned v = hs_peek(p);
        if (v >= need) return;
        if (spin > 64) __nanosleep(spin > 4096 ? 1000 : 100);
        if (spin > (1u << 22)) __trap();
    }
}

// ---- bulk asynchronous copies (the TMA engine, cp.async.bulk -> UBLKCP in SASS) and their mbarrier
struct HsMbar { unsigned long long word; };
HS_DEVICE unsigned hs_smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
HS_DEVICE void hs_mbar_init(HsMbar *m)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(hs_smem_u32(m)) : "memory");
}
HS_DEVICE void hs_mbar_init_n(HsMbar *m, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(hs_smem_u32(m)), "r"(count) : "memory");
}
HS_DEVICE void hs_mbar_arrive(HsMbar *m)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(hs_smem_u32(m)) : "memory");
}
// non-blocking: has the phase of the given parity completed?
HS_DEVICE bool hs_mbar_test(HsMbar *m, unsigned parity)
{
    unsigned ok;
    asm volatile("{\n .reg .pred p;\n mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                 : "=r"(ok)
                 : "r"(hs_smem_u32(m)), "r"(parity)
                 : "memory");
    return ok != 0;
}
// one arrival (of the single expected one) announcing `bytes` of asynchronous writes for this phase
HS_DEVICE void hs_mbar_expect(HsMbar *m, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(hs_smem_u32(m)), "r"(bytes) : "memory");
}
// global -> shared, 16-byte aligned, size a multiple of 16; completion is credited to the mbarrier
HS_DEVICE void hs_bulk_g2s(void *dst_smem, const void *src_gmem, unsigned bytes, HsMbar *m)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(hs_smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(hs_smem_u32(m))
                 : "memory");
}
// shared -> global (the destination may be a peer GPU's memory), 16-byte aligned, size a multiple of 16; part of the
// issuing thread's current bulk group
HS_DEVICE void hs_bulk_s2g(void *dst_gmem, const void *src_smem, unsigned bytes)
{
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst_gmem), "r"(hs_smem_u32(src_smem)), "r"(bytes) : "memory");
}
// close the thread's bulk group and wait until the copy engine has READ its shared-memory sources (the CTA may then
// overwrite them or exit; the writes themselves complete like any other store, by the end of the kernel)
HS_DEVICE void hs_bulk_commit_wait_read()
{
    asm volatile("cp.async.bulk.commit_group;\n cp.async.bulk.wait_group.read 0;" ::: "memory");
}
// ask the copy engine to bring [p, p + bytes) into the L2 (16-byte aligned, a multiple of 16 bytes)
HS_DEVICE void hs_prefetch_l2(const void *p, unsigned bytes)
{
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
}
// wait for the phase of the given parity; a copy that never lands traps instead of hanging the GPU
HS_DEVICE void hs_mbar_wait(HsMbar *m, unsigned parity)
{
    const unsigned a = hs_smem_u32(m);
    for (unsigned spin = 0;; ++spin) {
        unsigned ok;
        asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                     : "=r"(ok)
                     : "r"(a), "r"(parity)
                     : "memory");
        if (ok) return;
        if (spin > (1u << 24)) __trap();
    }
}
// 2-D tile of a 3-D tensor [z][row][x] described by a CUtensorMap (cuTensorMapEncodeTiled on the host): one
// instruction moves box_rows x box_x elements, the TMA engine walks the strides (UTMALDG in SASS)
#ifndef __CUDACC_RTC__
#include <cuda.h>
typedef CUtensorMap HsTmap;
#else
struct alignas(64) HsTmap { unsigned long long opaque[16]; };
#endif
#define HS_GRID_CONSTANT __grid_constant__
HS_DEVICE void hs_tma_load_3d(void *dst_smem, const HsTmap *t, int x_dbl, int row0, long long z, int /*rows*/, int /*row_bytes*/, HsMbar *m)
{
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(
                     hs_smem_u32(dst_smem)),
                 "l"(t), "r"(x_dbl), "r"(row0), "r"((int)z), "r"(hs_smem_u32(m))
                 : "memory");
}
// same with an L2 residency class for the tile's lines: MP 1 keep (evict_last), MP 2 stream (evict_first)
template <int MP>
HS_DEVICE void hs_tma_load_3d_mp(void *dst_smem, const HsTmap *t, int x_dbl, int row0, long long z, int rows, int row_bytes, HsMbar *m)
{
    if (MP == 0) { hs_tma_load_3d(dst_smem, t, x_dbl, row0, z, rows, row_bytes, m); return; }
    const unsigned long long pol = MP == 1 ? hs_pol_keep() : hs_pol_first();
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%2, %3, %4}], [%5], %6;" ::"r"(
            hs_smem_u32(dst_smem)),
        "l"(t), "r"(x_dbl), "r"(row0), "r"((int)z), "r"(hs_smem_u32(m)), "l"(pol)
        : "memory");
}

// asynchronous 16-byte copy global -> shared that bypasses the registers and L1 (LDGSTS): the issuing thread goes on, the data
// lands later; hs_cp_async_wait_all waits for this thread's copies (other threads need a barrier on top)
HS_DEVICE void hs_cp_async16(void *dst_smem, const void *src)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(hs_smem_u32(dst_smem)), "l"(src) : "memory");
}
HS_DEVICE void hs_cp_async_wait_all() { asm volatile("cp.async.commit_group;\n cp.async.wait_group 0;" ::: "memory"); }

// ---- tensor memory (TMEM, 256 KB per SM: 128 lanes x 512 columns x 32 bit) used as a per-thread parking lot.  The split
// chirp-z kernel has to keep 160 KB of intermediate results per transform while its shared memory and registers are busy
// with the next half-convolution; tcgen05.st / tcgen05.ld move registers to and from TMEM without touching L1 / L2.
// A warp reaches the 32 lanes of its quadrant (warp id mod 4), one lane per thread, and any column.
HS_DEVICE unsigned hs_tmem_alloc(unsigned *smem_slot, int ncols) // whole CTA calls; warp 0 allocates ncols columns (a power of two, 32..512)
{
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(hs_smem_u32(smem_slot)), "r"(ncols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    return *(volatile unsigned *)smem_slot;
}
HS_DEVICE void hs_tmem_free(unsigned base, int ncols) // whole CTA calls after its last tcgen05.ld has been waited for
{
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(base), "r"(ncols) : "memory");
}
// this thread's slot of `words_per_thread` columns: lane quadrant of its warp, columns by the warp's rank inside the quadrant
HS_DEVICE unsigned hs_tmem_my_addr(unsigned base, int words_per_thread)
{
    const unsigned warp = threadIdx.x >> 5;
    return base + (((warp & 3u) * 32u) << 16) + (warp >> 2) * (unsigned)words_per_thread;
}
// four complex doubles (16 words) of chunk `chunk` of the thread's slot
HS_DEVICE void hs_tmem_st4(unsigned taddr, int chunk, const double2 *v)
{
    unsigned r[16];
    HS_UNROLL
    for (int i = 0; i < 4; ++i) {
        r[4 * i + 0] = (unsigned)__double2loint(v[i].x); r[4 * i + 1] = (unsigned)__double2hiint(v[i].x);
        r[4 * i + 2] = (unsigned)__double2loint(v[i].y); r[4 * i + 3] = (unsigned)__double2hiint(v[i].y);
    }
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(
                     taddr + 16u * (unsigned)chunk),
                 "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]),
                 "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
                 : "memory");
}
HS_DEVICE void hs_tmem_ld4(unsigned taddr, int chunk, double2 *v)
{
    unsigned r[16];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
                   "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(taddr + 16u * (unsigned)chunk)
                 : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); // the registers are valid from here on
    HS_UNROLL
    for (int i = 0; i < 4; ++i) {
        v[i].x = __hiloint2double((int)r[4 * i + 1], (int)r[4 * i + 0]);
        v[i].y = __hiloint2double((int)r[4 * i + 3], (int)r[4 * i + 2]);
    }
}
HS_DEVICE void hs_tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
HS_DEVICE void hs_tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// order this thread's generic-proxy accesses to shared memory before later asynchronous-proxy ones
HS_DEVICE void hs_fence_async_proxy() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
// same for every state space: a counter read through the generic proxy gates a TMA load of global data
HS_DEVICE void hs_fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }
// CTA-scope counters in shared memory between warp roles (producer warp <-> compute warps)
HS_DEVICE void hs_flag_add(unsigned *p)
{
    asm volatile("red.release.cta.shared::cta.add.u32 [%0], 1;" ::"r"(hs_smem_u32(p)) : "memory");
}
HS_DEVICE unsigned hs_flag_read(const unsigned *p)
{
    unsigned v;
    asm volatile("ld.acquire.cta.shared::cta.u32 %0, [%1];" : "=r"(v) : "r"(hs_smem_u32(p)) : "memory");
    return v;
}
HS_DEVICE void hs_flag_set(unsigned *p, unsigned v)
{
    asm volatile("st.release.cta.shared::cta.u32 [%0], %1;" ::"r"(hs_smem_u32(p)), "r"(v) : "memory");
}
HS_DEVICE void hs_backoff() { __nanosleep(64); }
// barrier among the first n threads of the CTA only (the compute warps of a warp-specialised kernel)
#define HS_SYNC_COMPUTE(n) asm volatile("bar.sync 1, %0;" ::"r"(n) : "memory")
#define HS_TRAP() __trap()
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

// r2c split (reference src/real.c:109-125): bin k of the real transform from Z[k], Z[H-k] and twiddle2[k]
HS_DEVICE cplx hs_r2c_bin(cplx a, cplx c, cplx w) // the reference's expression for bin k from (Z[k], Z[H-k], twiddle2[k])
{
    const double t1 = a.y + c.y, t2 = c.x - a.x;
    return hs_c((a.x + c.x + (t1 * w.x) + (t2 * w.y)) / 2.0, (a.y - c.y + (t2 * w.x) - (t1 * w.y)) / 2.0);
}
