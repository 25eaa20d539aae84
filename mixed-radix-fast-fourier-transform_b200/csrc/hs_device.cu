// hs_device.cu -- the CUDA side of the thin C ABI (hs_abi.h): device/stream/memory helpers and
// the launchers of the sm_100a kernels.  Built only by nvcc for sm_100a; there is no other backend.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <atomic>
#include <string>

#include "hs_abi.h"
#include "hs_cuda_util.h"
#include "hs_conv_fft.cuh"
#include "hs_fused.cuh"
#include "hs_smem_fft.cuh"
#include "hs_tiled.cuh"

// The launchers are compiled as three translation units (build.py: -DHS_TU_PART=1,2,3) so that nvcc works on them in
// parallel: 1 = this file's plain functions plus the one-kernel families, 2 = the tiled passes, 3 = the tiled passes
// with fused ends and the long-Bluestein middle kernel.  HS_TU_PART undefined = everything in one unit.
#ifndef HS_TU_PART
#define HS_TU_PART 0
#endif
#define HS_PART(n) (HS_TU_PART == 0 || HS_TU_PART == (n))
#if HS_PART(1)
#include "hs_generic.cuh" // plain (non-template) kernels: defined in one unit only
#endif
#if HS_PART(1)
std::atomic<long long> g_launches{0};
#else
extern std::atomic<long long> g_launches;
#endif

typedef cudaStream_t hs_stream_t;

#if HS_PART(1)
// ----------------------------------------------------------------------------- device management
static int g_dev_count = -2;

extern "C" int hs_cu_device_count(void)
{
    if (g_dev_count == -2) {
        int n = 0;
        if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); n = 0; }
        g_dev_count = n;
    }
    return g_dev_count;
}
extern "C" void hs_cu_require_device(void)
{
    if (hs_cu_device_count() <= 0) {
        fprintf(stderr, "Error: hsfft-b200 needs a CUDA device (sm_100a); there is no CPU fallback\n");
        exit(EXIT_FAILURE);
    }
}
extern "C" int hs_cu_get_device(void) { int d = 0; HS_CUDA_CHECK(cudaGetDevice(&d)); return d; }
extern "C" void hs_cu_set_device(int dev) { HS_CUDA_CHECK(cudaSetDevice(dev)); }
extern "C" int hs_cu_sm_count(void)
{
    int dev = 0, n = 0;
    HS_CUDA_CHECK(cudaGetDevice(&dev));
    HS_CUDA_CHECK(cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev));
    return n;
}
extern "C" void *hs_cu_malloc(size_t bytes)
{
    void *p = nullptr;
    HS_CUDA_CHECK(cudaMalloc(&p, bytes ? bytes : 16));
    return p;
}
extern "C" void hs_cu_free(void *d) { if (d) HS_CUDA_CHECK(cudaFree(d)); }
extern "C" void *hs_cu_malloc_async(size_t bytes, void *stream)
{
    void *p = nullptr;
    HS_CUDA_CHECK(cudaMallocAsync(&p, bytes ? bytes : 16, (cudaStream_t)stream));
    return p;
}
extern "C" void hs_cu_free_async(void *d, void *stream) { if (d) HS_CUDA_CHECK(cudaFreeAsync(d, (cudaStream_t)stream)); }
extern "C" void hs_cu_device_sync(void) { HS_CUDA_CHECK(cudaDeviceSynchronize()); }
extern "C" int hs_cu_can_access_peer(int dev, int peer)
{
    int ok = 0;
    HS_CUDA_CHECK(cudaDeviceCanAccessPeer(&ok, dev, peer));
    return ok;
}
extern "C" void hs_cu_enable_peer_access(int peer)
{
    const cudaError_t e = cudaDeviceEnablePeerAccess(peer, 0);
    if (e == cudaErrorPeerAccessAlreadyEnabled) { cudaGetLastError(); return; }
    HS_CUDA_CHECK(e);
}
extern "C" void hs_cu_h2d_2d(void *d, size_t dpitch, const void *h, size_t hpitch, size_t width_bytes, size_t rows, void *stream)
{
    if (width_bytes && rows) HS_CUDA_CHECK(cudaMemcpy2DAsync(d, dpitch, h, hpitch, width_bytes, rows, cudaMemcpyHostToDevice, (cudaStream_t)stream));
}
extern "C" void hs_cu_d2h_2d(void *h, size_t hpitch, const void *d, size_t dpitch, size_t width_bytes, size_t rows, void *stream)
{
    if (width_bytes && rows) HS_CUDA_CHECK(cudaMemcpy2DAsync(h, hpitch, d, dpitch, width_bytes, rows, cudaMemcpyDeviceToHost, (cudaStream_t)stream));
}
extern "C" void *hs_cu_malloc_host(size_t bytes)
{
    void *p = nullptr;
    HS_CUDA_CHECK(cudaMallocHost(&p, bytes ? bytes : 16));
    return p;
}
extern "C" void hs_cu_free_host(void *h) { if (h) HS_CUDA_CHECK(cudaFreeHost(h)); }
extern "C" void hs_cu_h2d(void *d, const void *h, size_t bytes, void *stream)
{
    if (!bytes) return;
    if (stream) HS_CUDA_CHECK(cudaMemcpyAsync(d, h, bytes, cudaMemcpyHostToDevice, (cudaStream_t)stream));
    else HS_CUDA_CHECK(cudaMemcpy(d, h, bytes, cudaMemcpyHostToDevice));
}
extern "C" void hs_cu_d2h(void *h, const void *d, size_t bytes, void *stream)
{
    if (!bytes) return;
    if (stream) HS_CUDA_CHECK(cudaMemcpyAsync(h, d, bytes, cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    else HS_CUDA_CHECK(cudaMemcpy(h, d, bytes, cudaMemcpyDeviceToHost));
}
extern "C" void hs_cu_d2d(void *dst, const void *src, size_t bytes, void *stream)
{
    if (bytes) HS_CUDA_CHECK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
}
extern "C" void hs_cu_memset0(void *d, size_t bytes, void *stream)
{
    if (bytes) HS_CUDA_CHECK(cudaMemsetAsync(d, 0, bytes, (cudaStream_t)stream));
}
extern "C" void *hs_cu_stream_create(void)
{
    cudaStream_t s;
    HS_CUDA_CHECK(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
    return (void *)s;
}
extern "C" void hs_cu_stream_destroy(void *stream) { if (stream) HS_CUDA_CHECK(cudaStreamDestroy((cudaStream_t)stream)); }
extern "C" void hs_cu_stream_sync(void *stream) { HS_CUDA_CHECK(cudaStreamSynchronize((cudaStream_t)stream)); }
extern "C" void *hs_cu_event_create(void)
{
    cudaEvent_t e;
    HS_CUDA_CHECK(cudaEventCreate(&e));
    return (void *)e;
}
extern "C" void hs_cu_event_destroy(void *ev) { if (ev) HS_CUDA_CHECK(cudaEventDestroy((cudaEvent_t)ev)); }
extern "C" void hs_cu_event_record(void *ev, void *stream) { HS_CUDA_CHECK(cudaEventRecord((cudaEvent_t)ev, (cudaStream_t)stream)); }
extern "C" void hs_cu_stream_wait_event(void *stream, void *ev)
{
    HS_CUDA_CHECK(cudaStreamWaitEvent((cudaStream_t)stream, (cudaEvent_t)ev, 0));
}
extern "C" float hs_cu_event_elapsed_ms(void *start, void *stop)
{
    float ms = 0.f;
    HS_CUDA_CHECK(cudaEventElapsedTime(&ms, (cudaEvent_t)start, (cudaEvent_t)stop));
    return ms;
}
extern "C" void hs_cu_event_sync(void *ev) { HS_CUDA_CHECK(cudaEventSynchronize((cudaEvent_t)ev)); }
extern "C" void hs_cu_ipc_get(void *d, void *handle64)
{
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    HS_CUDA_CHECK(cudaIpcGetMemHandle((cudaIpcMemHandle_t *)handle64, d));
}
extern "C" void *hs_cu_ipc_open(const void *handle64)
{
    void *p = nullptr;
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, sizeof h);
    HS_CUDA_CHECK(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    return p;
}
extern "C" void hs_cu_ipc_close(void *mapped) { if (mapped) HS_CUDA_CHECK(cudaIpcCloseMemHandle(mapped)); }
extern "C" long long hs_cu_launch_count(void) { return g_launches.load(); }
extern "C" void hs_cu_count_launch(void) { ++g_launches; }

// Per-launch timing for bench.py's roofline record (hsfft_timing_begin / hsfft_timing_report in include/hsfft_b200.h): while it
// is on, every launcher records a CUDA event on its stream right after its kernel; the time between consecutive events
// of a stream is that kernel's duration (launches of one call run back to back).  Off (the default) it costs a branch.
#include <mutex>
#include <string>
#include <vector>
static std::atomic<int> g_timing{0};
static std::mutex g_timing_mu;
struct HsTimedLaunch { const char *tag; cudaStream_t st; cudaEvent_t ev; };
static std::vector<HsTimedLaunch> g_timed;
extern "C" void hs_cu_timing_mark(const char *tag, void *stream)
{
    if (!g_timing.load(std::memory_order_relaxed)) return;
    cudaEvent_t e;
    HS_CUDA_CHECK(cudaEventCreate(&e));
    HS_CUDA_CHECK(cudaEventRecord(e, (cudaStream_t)stream));
    std::lock_guard<std::mutex> lk(g_timing_mu);
    g_timed.push_back({tag, (cudaStream_t)stream, e});
}
extern "C" void hs_cu_timing_begin(void *stream)
{
    {
        std::lock_guard<std::mutex> lk(g_timing_mu);
        for (auto &t : g_timed) cudaEventDestroy(t.ev);
        g_timed.clear();
    }
    g_timing.store(1);
    hs_cu_timing_mark("", stream); // origin of the first launch on this stream
}
// stops the recording, waits for the events and writes "tag:count:total_ms;..." (launch order of first appearance)
extern "C" int hs_cu_timing_report(char *out, int cap)
{
    g_timing.store(0);
    std::lock_guard<std::mutex> lk(g_timing_mu);
    std::vector<std::string> tags;
    std::vector<double> ms;
    std::vector<long long> cnt;
    for (size_t i = 0; i < g_timed.size(); ++i) {
        HS_CUDA_CHECK(cudaEventSynchronize(g_timed[i].ev));
        if (!g_timed[i].tag[0]) continue;
        long long j = (long long)i - 1;
        while (j >= 0 && g_timed[j].st != g_timed[i].st) --j; // previous event of the same stream
        if (j < 0) continue;
        float t = 0.f;
        HS_CUDA_CHECK(cudaEventElapsedTime(&t, g_timed[j].ev, g_timed[i].ev));
        size_t k = 0;
        while (k < tags.size() && tags[k] != g_timed[i].tag) ++k;
        if (k == tags.size()) { tags.push_back(g_timed[i].tag); ms.push_back(0.0); cnt.push_back(0); }
        ms[k] += t; cnt[k] += 1;
    }
    std::string r;
    for (size_t k = 0; k < tags.size(); ++k) {
        char b[160];
        snprintf(b, sizeof b, "%s%s:%lld:%.6f", k ? ";" : "", tags[k].c_str(), cnt[k], ms[k]);
        r += b;
    }
    for (auto &t : g_timed) cudaEventDestroy(t.ev);
    g_timed.clear();
    if ((int)r.size() + 1 > cap) return -1;
    memcpy(out, r.c_str(), r.size() + 1);
    return (int)r.size();
}
#else
extern "C" int hs_cu_sm_count(void);
#endif

extern "C" void hs_cu_timing_mark(const char *tag, void *stream);
static inline void hs_launched(const char *tag, cudaStream_t st)
{
    ++g_launches;
    hs_cu_timing_mark(tag, (void *)st);
}
// --------------------------------------------------------------------------------------- launchers
// Function attributes (dynamic shared memory limit, carve-out) belong to a (function, device) pair, so "done" is one bit
// per device ordinal: a process that drives several GPUs sets them once on each.
static int cur_device()
{
    int dev = 0;
    HS_CUDA_CHECK(cudaGetDevice(&dev));
    return dev;
}
static bool attr_todo(const std::atomic<unsigned long long> &mask) { return !((mask.load(std::memory_order_acquire) >> (cur_device() & 63)) & 1ULL); }
static void attr_mark(std::atomic<unsigned long long> &mask) { mask.fetch_or(1ULL << (cur_device() & 63), std::memory_order_release); }
static void set_carveout(const void *kern, int smem_per_cta, int ctas_per_sm);
// the usual pair for a kernel with `smem` bytes of dynamic shared memory meant to run `ctas_per_sm` CTAs per SM
#define hs_func_attr_once(kern, smem, ctas_per_sm)                                                                          \
    do {                                                                                                                    \
        static std::atomic<unsigned long long> attr_done_{0};                                                               \
        if (attr_todo(attr_done_)) {                                                                                        \
            HS_CUDA_CHECK(cudaFuncSetAttribute((kern), cudaFuncAttributeMaxDynamicSharedMemorySize, (smem)));               \
            set_carveout((kern), (smem), (ctas_per_sm));                                                                    \
            attr_mark(attr_done_);                                                                                          \
        }                                                                                                                   \
    } while (0)

static void check_grid(long long blocks)
{
    if (blocks <= 0 || blocks > 2147483647LL) {
        fprintf(stderr, "Error: batch too large for one launch (%lld CTAs)\n", blocks);
        exit(EXIT_FAILURE);
    }
}

// Ask for just enough shared memory for the intended CTAs per SM, so that the rest of the 228 KB
// stays L1 and keeps the twiddle / chirp tables close (HSFFT_CARVEOUT=0 leaves the driver default).
static void set_carveout(const void *kern, int smem_per_cta, int ctas_per_sm)
{
    static int mode = -1;
    if (mode < 0) { const char *e = getenv("HSFFT_CARVEOUT"); mode = e ? atoi(e) : 1; }
    if (!mode) return;
    int pct = (int)(((long long)(smem_per_cta + 1024) * ctas_per_sm * 100 + 228 * 1024 - 1) / (228 * 1024));
    if (pct > 100) pct = 100;
    HS_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, pct));
}

// L2 prefetch distance of the one-shot kernels, in CTAs: the number of CTAs resident at once (HSFFT_PREFETCH=0 turns it off,
// HSFFT_PREFETCH=<n> sets it in units of resident waves).  Measured on B200 (fraction of the HBM peak, prefetch off -> on):
// N=8192 56 -> 65 %, N=256 90 -> 104 %, N=512 104 -> 107 %, Bluestein N=1009 25.0 -> 26.2 %; N=4096 / 1024 / 2048 / 3000 lose
// 1-3 points, so only the first group asks for it.
static int prefetch_dist(int ctas_per_sm)
{
    static int waves = -1;
    if (waves < 0) { const char *e = getenv("HSFFT_PREFETCH"); waves = e ? atoi(e) : 1; }
    return waves * ctas_per_sm * hs_cu_sm_count();
}

template <class P, int SGN, int MINB, int PFW = 0> // PFW: L2 prefetch distance in resident waves (0 = none)
static void launch_c2c(const void *in, void *out, const void *tw, long long howmany, cudaStream_t st)
{
    auto kern = hs_c2c_smem_kernel<P, SGN, MINB>;
    static std::atomic<unsigned long long> attr_done{0}; // one bit per device: function attributes are per device
    if (attr_todo(attr_done)) {
        HS_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, P::smem_bytes));
        set_carveout((const void *)kern, P::smem_bytes, MINB);
        attr_mark(attr_done);
    }
    const long long blocks = (howmany + P::X - 1) / P::X;
    check_grid(blocks);
    kern<<<(unsigned)blocks, P::threads, P::smem_bytes, st>>>((const cplx *)in, (cplx *)out, (const cplx *)tw, howmany, PFW ? prefetch_dist(MINB) : 0);
    HS_CUDA_CHECK(cudaGetLastError());
    hs_launched("c2c", st);
}

template <class P, int SGN, int MINB>
static void launch_r2c(const void *in, void *out, const void *tw, const void *tw2, long long out_stride, int full, long long howmany, cudaStream_t st)
{
    auto kern = hs_r2c_smem_kernel<P, SGN, MINB>;
    hs_func_attr_once((const void *)kern, P::smem_bytes, MINB);
    const long long blocks = (howmany + P::X - 1) / P::X;
    check_grid(blocks);
    kern<<<(unsigned)blocks, P::threads, P::smem_bytes, st>>>((const cplx *)in, (cplx *)out, (const cplx *)tw, (const cplx *)tw2, out_stride, full, howmany);
    HS_CUDA_CHECK(cudaGetLastError());
    hs_launched("r2c", st);
}
#ifndef HS_C2R_STAGE_DEFAULT
#define HS_C2R_STAGE_DEFAULT 1
#endif
template <class P, int SGN, int MINB>
static void launch_c2r(const void *in, void *out, const void *tw, const void *tw2, long long in_stride, long long howmany, cudaStream_t st)
{
    // The merge reads a copy of the row staged in the team's buffer (every bin fetched once, in order) unless HSFFT_C2R_STAGE=0
    // (every bin loaded twice straight from global memory); =2 stages for every plan.  Measured on B200, fraction of the HBM
    // peak staged / direct by real length N (plan length H = N/2): N = 512 91 / 60 %, 2048 65 / 46 %, 4096 72 / 70 %, 8192
    // 72 / 70 %; the two plans that do not gain keep the direct form: N = 1024 (H = 512) 79.5 / 80.8 %, N = 16384 (H = 8192,
    // one CTA per SM) 45.0 / 46.7 %.  profiles/r2_c2r_staged_merge_one_kernel_ab.txt
    static int want = -1;
    if (want < 0) { const char *e = getenv("HSFFT_C2R_STAGE"); want = e ? atoi(e) : HS_C2R_STAGE_DEFAULT; }
    const int stage = want == 2 || (want == 1 && P::N != 512 && P::N != 8192);
    const long long blocks = (howmany + P::X - 1) / P::X;
    check_grid(blocks);
    if (stage) {
        auto kern = hs_c2r_smem_kernel<P, SGN, MINB, true>;
        hs_func_attr_once((const void *)kern, P::smem_bytes, MINB);
        kern<<<(unsigned)blocks, P::threads, P::smem_bytes, st>>>((const cplx *)in, (cplx *)out, (const cplx *)tw, (const cplx *)tw2, in_stride, howmany);
    } else {
        auto kern = hs_c2r_smem_kernel<P, SGN, MINB>;
        hs_func_attr_once((const void *)kern, P::smem_bytes, MINB);
        kern<<<(unsigned)blocks, P::threads, P::smem_bytes, st>>>((const cplx *)in, (cplx *)out, (const cplx *)tw, (const cplx *)tw2, in_stride, howmany);
    }
    HS_CUDA_CHECK(cudaGetLastError());
    hs_launched(stage ? "c2r (staged merge)" : "c2r", st);
}

template <class PF, class PI, int SGN, int MINB>
static void launch_blu(int N, const void *in, void *out, const void *twf, const void *twi, const void *chirp, const void *hk,
                       long long howmany, cudaStream_t st)
{
    auto kern = hs_bluestein_kernel<PF, PI, SGN, MINB>;
    static std::atomic<unsigned long long> attr_done{0}; // one bit per device: function attributes are per device
    if (attr_todo(attr_done)) {
        HS_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, PF::smem_bytes));
        set_carveout((const void *)kern, PF::smem_bytes, MINB);
        attr_mark(attr_done);
    }
    const long long blocks = (howmany + PF::X - 1) / PF::X;
    check_grid(blocks);
    kern<<<(unsigned)blocks, PF::threads, PF::smem_bytes, st>>>((const cplx *)in, (cplx *)out, (const cplx *)twf, (const cplx *)twi,
                                                                 (const cplx *)chirp, (const cplx *)hk, N, howmany, prefetch_dist(MINB));
    HS_CUDA_CHECK(cudaGetLastError());
    hs_launched("blu", st);
}

template <class PF, class PI, int SGN, int MINB>
static void launch_blu_split(int N, const void *in, void *out, const void *twf, const void *twi, const void *chirp3, const void *hk2,
                             long long howmany, cudaStream_t st)
{
    auto kern = hs_bluestein_split_kernel<PF, PI, SGN, MINB>;
    hs_func_attr_once((const void *)kern, PF::smem_bytes, MINB);
    check_grid(howmany);
    kern<<<(unsigned)howmany, PF::threads, PF::smem_bytes, st>>>((const cplx *)in, (cplx *)out, (const cplx *)twf, (const cplx *)twi,
                                                                  (const cplx *)chirp3, (const cplx *)hk2, N, howmany, prefetch_dist(MINB));
    HS_CUDA_CHECK(cudaGetLastError());
    hs_launched("blu_split", st);
}

// HSFFT_SPLIT_TMEM=0 / 1: the split chirp-z kernel with its intermediates in the L2 / in tensor memory (see hs_conv_fft.cuh)
template <class P, int SGN, int MINB>
static void launch_conv_split_tmem(int N, const void *in, void *out, const void *tw, const void *h, const void *wp, const void *hkt2,
                                   long long howmany, cudaStream_t st)
{
    auto kern = hs_conv_bluestein_split_tmem_kernel<P, SGN, MINB>;
    constexpr int smem = P::smem_bytes + 16;
    hs_func_attr_once((const void *)kern, smem, MINB);
    // persistent: one CTA per SM (it owns the SM's tensor memory and 175 KB of its shared memory) walks the batch
    long long blocks = (long long)hs_cu_sm_count() * MINB;
    if (blocks > howmany) blocks = howmany;
    check_grid(blocks);
    kern<<<(unsigned)blocks, P::threads, smem, st>>>((const cplx *)in, (cplx *)out, (const cplx *)tw, (const cplx *)h, (const cplx *)wp,
                                                     (const cplx *)hkt2, N, howmany, 0);
    HS_CUDA_CHECK(cudaGetLastError());
    hs_launched("conv_split_tmem", st);
}

template <class P, int SGN, int MINB>
static void launch_conv_split(int N, const void *in, void *out, const void *tw, const void *chirp3, const void *hkt2, long long howmany,
                              cudaStream_t st)
{
    auto kern = hs_conv_bluestein_split_kernel<P, SGN, MINB>;
    hs_func_attr_once((const void *)kern, P::smem_bytes, MINB);
    check_grid(howmany);
    kern<<<(unsigned)howmany, P::threads, P::smem_bytes, st>>>((const cplx *)in, (cplx *)out, (const cplx *)tw, (const cplx *)chirp3,
                                                                (const cplx *)hkt2, N, howmany, prefetch_dist(MINB));
    HS_CUDA_CHECK(cudaGetLastError());
    hs_launched("conv_split", st);
}

#ifndef HS_RADER_STAGE_DEFAULT
#define HS_RADER_STAGE_DEFAULT 2
#endif
template <class PF, class PI, int MINB, int STG>
static void launch_rader_v(const void *in, void *out, const void *twf, const void *twi, const void *perm, const void *bk, long long howmany,
                           cudaStream_t st)
{
    constexpr int smem = HsRaderSmem<PF, PI>::bytes;
    auto kern = hs_rader_kernel<PF, PI, MINB, STG>;
    static std::atomic<unsigned long long> attr_done{0}; // one bit per device: function attributes are per device
    if (attr_todo(attr_done)) {
        HS_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        set_carveout((const void *)kern, smem, MINB);
        attr_mark(attr_done);
    }
    const long long blocks = (howmany + PF::X - 1) / PF::X;
    check_grid(blocks);
    kern<<<(unsigned)blocks, PF::threads, smem, st>>>((const cplx *)in, (cplx *)out, (const cplx *)twf, (const cplx *)twi, (const int *)perm,
                                                      (const cplx *)bk, howmany, prefetch_dist(MINB));
    HS_CUDA_CHECK(cudaGetLastError());
    hs_launched("rader_v", st);
}

// HSFFT_RADER_STAGE=0 / 2: scatter the outputs straight to HBM / through shared memory (default; measured on B200 at
// N = 1009: 0.238 -> 0.228 ms, N = 257: 0.173 -> 0.161 ms).  Staging the gather as well (bit 0 of the kernel's STG)
// costs a barrier before the first pass and was slower everywhere (0.250 ms), so it is not instantiated here.
template <class PF, class PI, int MINB>
static void launch_rader(const void *in, void *out, const void *twf, const void *twi, const void *perm, const void *bk, long long howmany,
                         cudaStream_t st)
{
    static int stg = -1;
    if (stg < 0) { const char *e = getenv("HSFFT_RADER_STAGE"); stg = e ? atoi(e) & 3 : HS_RADER_STAGE_DEFAULT; }
    if (stg & 2) launch_rader_v<PF, PI, MINB, 2>(in, out, twf, twi, perm, bk, howmany, st);
    else launch_rader_v<PF, PI, MINB, 0>(in, out, twf, twi, perm, bk, howmany, st);
}

// persistent, TMA-staged variant (hs_tiled_tma_kernel).  Measured on B200 (two-pass c2c, same box, TMA vs plain):
// 2^17 43.8 vs 38.8 %, 2^18 40.5 vs 29.6 %, 2^19 36.4 vs 32.3 %, 2^20 35.9 vs 35.6 % of the HBM peak, but
// 2^14..2^16 (128/256-point sub-plans, already at the two-pass cap) 45-49 vs 52-54 %.  So it is the default for
// sub-plans of 512 points and more; HSFFT_TMA=0 / 1 forces it off / on everywhere.
// (later: the plain kernel at three CTAs per SM -- 80 registers, no spill -- beats the persistent one for the 512-point
// sub-plan: 2^18 44.9 vs 41.1 %, 2^17 47.2 vs 46.5 %; the 1024-point one would spill there and keeps TMA for the first pass,
// but its second pass (inter-pass twiddle) is faster plain: 2^19 40.3 vs 38.7 %.  profiles/r1b_tiled_occupancy_ab.txt)
static bool tma_enabled(int sub_len, int tw)
{
    static int v = -1;
    if (v < 0) { const char *e = getenv("HSFFT_TMA"); v = e ? atoi(e) : 2; }
    return v == 1 || (v == 2 && sub_len >= 1024 && tw == 0) || (v == 3 && sub_len >= 512 && tw == 0);
}

template <class P, int SGN, int C, int TW, bool STCOL, int IO = 0>
static void launch_tiled_tma(const void *in, void *out, const void *tw, const HsTileGeom &g, long long howmany, cudaStream_t st)
{
    constexpr int smem = HsTmaTileSmem<P, C>::bytes;
    static_assert(smem <= 227 * 1024, "staging buffers do not fit");
    auto kern = hs_tiled_tma_kernel<P, SGN, C, TW, STCOL, IO>;
    static std::atomic<unsigned long long> attr_done{0}; // one bit per device: function attributes are per device
    static int per_sm = 1;
    if (attr_todo(attr_done)) {
        HS_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        HS_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, P::T * C, smem));
        if (per_sm < 1) per_sm = 1;
        if (getenv("HSFFT_VERBOSE")) fprintf(stderr, "hsfft: TMA tiled pass n=%d C=%d: %d CTAs/SM, %d bytes smem\n", P::N, C, per_sm, smem);
        attr_mark(attr_done);
    }
    const long long total = howmany * (g.ncols / C);
    long long blocks = (long long)hs_cu_sm_count() * per_sm;
    if (blocks > total) blocks = total;
    check_grid(blocks);
    const CUtensorMap tm = make_tmap(in, g, P::N, C, howmany);
    kern<<<(unsigned)blocks, P::T * C, smem, st>>>(tm, (cplx *)out, (const cplx *)tw, g, howmany);
    HS_CUDA_CHECK(cudaGetLastError());
    static const std::string tag = "tiled_tma n=" + std::to_string(P::N) + (TW ? " pass2" : " pass1") + (IO ? " io" + std::to_string(IO) : std::string());
    hs_launched(tag.c_str(), st);
}

#ifndef HS_R4_TMA_DEFAULT
#define HS_R4_TMA_DEFAULT 2
#endif
#ifndef HS_TILED_MINB_SMALL
#define HS_TILED_MINB_SMALL (P::N == 160 ? 5 : 2) /* the 160-point plan fits 72 registers: five 160-thread CTAs per SM (N = 10007: 5.70 -> 5.45 ms) */
#endif
#ifndef HS_BLUMID_MINB
#define HS_BLUMID_MINB ((PF::N == 128 || (PF::N == 512 && PF::T == 64)) ? 4 : (PF::T * C <= 256) ? 2 : 1) /* 128 / 512 points: no spill at 4 CTAs per SM, +2-3 %; 256 loses 3 %, 1024 spills */
#endif
#ifndef HS_TILED_MINB_MID
#define HS_TILED_MINB_MID (P::N == 512 && P::T == 64 ? 3 : 2) /* 512-point tiles need 80 registers: three 256-thread CTAs per SM without a spill */
#endif
template <class P, int SGN, int C, int TW, bool STCOL, int IO = 0, int CS = 1>
static void launch_tiled(const void *in, void *out, const void *tw, const HsTileGeom &g, long long howmany, cudaStream_t st)
{
    if constexpr (IO == 8 && !STCOL && TW == 0) {
        // the real four-step's first pass: same staging, the split as epilogue.  HSFFT_R4_TMA: 0 off, 1 for 1024 points, 2 (default)
        // for 512 points as well.  Measured at 2^28 reals: 1024 points 0.84 ms staged vs 1.05 plain, 512 points 0.86 vs 1.29
        constexpr int CT = HsTileColsTma<P, C>::value;
        static int r4tma = -1;
        if (r4tma < 0) { const char *e = getenv("HSFFT_R4_TMA"); r4tma = e ? atoi(e) : HS_R4_TMA_DEFAULT; }
        if (r4tma && P::N >= (r4tma == 2 ? 512 : 1024) && g.in_c == 1 && g.blk_log == 0 && g.ncols % CT == 0) { launch_tiled_tma<P, SGN, CT, 0, false, 8>(in, out, tw, g, howmany, st); return; }
    }
    if constexpr (IO == 0) {
        // (half as many columns per tile and two CTAs per SM -- independent phase streams -- measured slower: 2^20 34.3 % vs
        // 39.3 %, 2^18 35.6 % vs 40.3 % of the HBM peak; 32-byte runs cost more than the overlap gains)
        constexpr int CT = HsTileColsTma<P, C>::value;
        if (tma_enabled(P::N, TW) && g.in_c == 1 && g.blk_log == 0 && g.ncols % CT == 0) { launch_tiled_tma<P, SGN, CT, TW, STCOL>(in, out, tw, g, howmany, st); return; }
    }
    constexpr int MINB = (P::T * C <= 160) ? HS_TILED_MINB_SMALL : (P::T * C <= 256 || (P::T * C <= 320 && P::N / P::T <= 10)) ? HS_TILED_MINB_MID : 1; // 320 threads of 8-10 points fit 100 registers
    constexpr int smem = C * hs_col_stride(P::NPAD, C) * (int)sizeof(cplx);
    auto kern = hs_tiled_kernel<P, SGN, C, TW, STCOL, MINB, IO, CS>;
    static std::atomic<unsigned long long> attr_done{0}; // one bit per device: function attributes are per device
    if (attr_todo(attr_done)) {
        HS_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        attr_mark(attr_done);
    }
    if (!HS_TILE_PARTIAL(IO) && g.ncols % C != 0) { fprintf(stderr, "hsfft-b200: tiled pass with %d columns in tiles of %d\n", g.ncols, C); exit(EXIT_FAILURE); }
    const long long blocks = howmany * ((g.ncols + C - 1) / C);
    check_grid(blocks);
    kern<<<(unsigned)blocks, P::T * C, smem, st>>>((const cplx *)in, (cplx *)out, (const cplx *)tw, g, howmany);
    HS_CUDA_CHECK(cudaGetLastError());
    static const std::string tag = "tiled n=" + std::to_string(P::N) + (TW ? " pass2" : " pass1") + (IO ? " io" + std::to_string(IO) : std::string());
    hs_launched(tag.c_str(), st);
}

template <class PF, class PI, int CS>
static void launch_blu_mid(const void *in, void *out, const void *twf, const void *twi, const HsTileGeom &g, const void *hk,
                           long long howmany, cudaStream_t st)
{
    constexpr int C = HsTileCols<PF>::C;
    constexpr int MINB = HS_BLUMID_MINB;
    constexpr int smem = C * hs_col_stride(PF::NPAD, C) * (int)sizeof(cplx);
    auto kern = hs_blu_mid_kernel<PF, PI, CS, C, MINB>;
    static std::atomic<unsigned long long> attr_done{0}; // one bit per device: function attributes are per device
    if (attr_todo(attr_done)) {
        HS_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        attr_mark(attr_done);
    }
    if (g.ncols % C != 0) { fprintf(stderr, "Error: tiled pass needs a multiple of %d columns\n", C); exit(EXIT_FAILURE); }
    const long long blocks = howmany * (g.ncols / C);
    check_grid(blocks);
    kern<<<(unsigned)blocks, PF::T * C, smem, st>>>((const cplx *)in, (cplx *)out, (const cplx *)twf, (const cplx *)twi, g, (const cplx *)hk,
                                                    howmany);
    HS_CUDA_CHECK(cudaGetLastError());
    static const std::string tag = "blu_mid n=" + std::to_string(PF::N);
    hs_launched(tag.c_str(), st);
}

#include "hs_ct_table.cuh"

#if HS_PART(1)
extern "C" void hs_cu_launch_generic(int sgn, const void *d_in, void *d_out, const hs_gen_params *p, void *stream)
{
    cudaStream_t st = (cudaStream_t)stream;
    const long long ntile = (p->n_c + p->C - 1) / p->C;
    const long long blocks = ntile * p->n_o * p->n_batch;
    check_grid(blocks);
    const size_t smem = (size_t)2 * p->C * p->npad * sizeof(cplx);
    static std::atomic<unsigned long long> attr_done{0}; // one bit per device: function attributes are per device
    if (attr_todo(attr_done)) {
        HS_CUDA_CHECK(cudaFuncSetAttribute(hs_generic_pass_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        HS_CUDA_CHECK(cudaFuncSetAttribute(hs_generic_pass_kernel<-1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        attr_mark(attr_done);
    }
    if (smem > 227 * 1024) { fprintf(stderr, "Error: pass needs %zu bytes of shared memory\n", smem); exit(EXIT_FAILURE); }
    if (sgn == 1) hs_generic_pass_kernel<1><<<(unsigned)blocks, 256, smem, st>>>((const cplx *)d_in, (cplx *)d_out, *p);
    else hs_generic_pass_kernel<-1><<<(unsigned)blocks, 256, smem, st>>>((const cplx *)d_in, (cplx *)d_out, *p);
    HS_CUDA_CHECK(cudaGetLastError());
    hs_launched("generic", st);
}

static unsigned ew_grid(long long total)
{
    long long b = (total + 255) / 256;
    const long long cap = 148LL * 32;
    if (b > cap) b = cap;
    if (b < 1) b = 1;
    return (unsigned)b;
}

extern "C" void hs_cu_launch_chirp_pad(int sgn, const void *x, void *a, const void *h, int N, int M, long long howmany, void *stream)
{
    cudaStream_t st = (cudaStream_t)stream;
    const unsigned g = ew_grid(howmany * M);
    if (sgn == 1) hs_chirp_pad_kernel<1><<<g, 256, 0, st>>>((const cplx *)x, (cplx *)a, (const cplx *)h, N, M, howmany);
    else hs_chirp_pad_kernel<-1><<<g, 256, 0, st>>>((const cplx *)x, (cplx *)a, (const cplx *)h, N, M, howmany);
    HS_CUDA_CHECK(cudaGetLastError());
    hs_launched("chirp_pad", st);
}
extern "C" void hs_cu_launch_spectrum_mul(int sgn, void *A, const void *Hk, int M, long long howmany, void *stream)
{
    cudaStream_t st = (cudaStream_t)stream;
    const unsigned g = ew_grid(howmany * M);
    if (sgn == 1) hs_spectrum_mul_kernel<1><<<g, 256, 0, st>>>((cplx *)A, (const cplx *)Hk, M, howmany);
    else hs_spectrum_mul_kernel<-1><<<g, 256, 0, st>>>((cplx *)A, (const cplx *)Hk, M, howmany);
    HS_CUDA_CHECK(cudaGetLastError());
    hs_launched("spectrum_mul", st);
}
extern "C" void hs_cu_launch_chirp_crop(int sgn, const void *y, void *out, const void *h, int N, int M, long long howmany, void *stream)
{
    cudaStream_t st = (cudaStream_t)stream;
    const unsigned g = ew_grid(howmany * N);
    if (sgn == 1) hs_chirp_crop_kernel<1><<<g, 256, 0, st>>>((const cplx *)y, (cplx *)out, (const cplx *)h, N, M, howmany);
    else hs_chirp_crop_kernel<-1><<<g, 256, 0, st>>>((const cplx *)y, (cplx *)out, (const cplx *)h, N, M, howmany);
    HS_CUDA_CHECK(cudaGetLastError());
    hs_launched("chirp_crop", st);
}
extern "C" void hs_cu_launch_r2c_post(const void *Z, void *out, const void *tw2, int H, long long out_stride, int full, long long howmany,
                                      void *stream)
{
    hs_r2c_post_kernel<<<ew_grid(howmany * (H / 2 + 1)), 256, 0, (cudaStream_t)stream>>>((const cplx *)Z, (cplx *)out, (const cplx *)tw2, H,
                                                                                      out_stride, full, howmany);
    HS_CUDA_CHECK(cudaGetLastError());
    hs_launched("r2c_post", (cudaStream_t)stream);
}
extern "C" void hs_cu_launch_c2r_pre(const void *in, void *z, const void *tw2, int H, long long in_stride, long long howmany, void *stream)
{
    hs_c2r_pre_kernel<<<ew_grid(howmany * H), 256, 0, (cudaStream_t)stream>>>((const cplx *)in, (cplx *)z, (const cplx *)tw2, H, in_stride,
                                                                               howmany);
    HS_CUDA_CHECK(cudaGetLastError());
    hs_launched("c2r_pre", (cudaStream_t)stream);
}
extern "C" void hs_cu_launch_conv_mul(const void *A, const void *B, void *C, long long total, double scale, void *stream)
{
    hs_conv_mul_kernel<<<ew_grid(total), 256, 0, (cudaStream_t)stream>>>((const cplx *)A, (const cplx *)B, (cplx *)C, total, scale);
    HS_CUDA_CHECK(cudaGetLastError());
    hs_launched("conv_mul", (cudaStream_t)stream);
}
extern "C" void hs_cu_launch_pad_rows(const void *src, int len, void *dst, int P, long long howmany, void *stream)
{
    hs_pad_rows_kernel<<<ew_grid(howmany * P), 256, 0, (cudaStream_t)stream>>>((const double *)src, len, (double *)dst, P, howmany);
    HS_CUDA_CHECK(cudaGetLastError());
    hs_launched("pad_rows", (cudaStream_t)stream);
}
extern "C" void hs_cu_launch_slice_rows(const void *src, int P, int start, int len, void *dst, long long dst_stride, long long howmany,
                                        void *stream)
{
    hs_slice_rows_kernel<<<ew_grid(howmany * len), 256, 0, (cudaStream_t)stream>>>((const double *)src, P, start, len, (double *)dst,
                                                                                    dst_stride, howmany);
    HS_CUDA_CHECK(cudaGetLastError());
    hs_launched("slice_rows", (cudaStream_t)stream);
}
extern "C" void hs_cu_launch_transpose(const void *in, void *out, int rows, int cols, long long howmany, void *stream)
{
    const long long tiles = (long long)((rows + 31) / 32) * ((cols + 31) / 32) * howmany;
    long long blocks = tiles < 148LL * 16 ? tiles : 148LL * 16;
    if (blocks < 1) blocks = 1;
    hs_transpose_kernel<<<(unsigned)blocks, 256, 32 * 33 * sizeof(cplx), (cudaStream_t)stream>>>((const cplx *)in, (cplx *)out, rows, cols, howmany);
    HS_CUDA_CHECK(cudaGetLastError());
    hs_launched("transpose", (cudaStream_t)stream);
}
extern "C" void hs_cu_launch_dist_twiddle(void *b, const void *hi, const void *lo, long long rows, long long cols, long long c0, int logN,
                                          int logS, void *stream)
{
    hs_dist_twiddle_kernel<0><<<ew_grid(rows * cols), 256, 0, (cudaStream_t)stream>>>((cplx *)b, (const cplx *)hi, (const cplx *)lo, rows, cols,
                                                                                       c0, logN, logS);
    HS_CUDA_CHECK(cudaGetLastError());
    hs_launched("dist_twiddle", (cudaStream_t)stream);
}
#endif // HS_PART(1)
