// hs_device_emu.cpp -- CPU emulation of the CUDA side of hs_abi.h.  TEST INFRASTRUCTURE ONLY.
//
// Linked with the product's own host C files (hsfft_plan.c, hsfft_real.c, hsfft_convolve.c) and
// the product's own kernel headers (csrc/*.cuh compiled with HS_EMU) into
// tests/emu/_build/libhsfft_emu.so, it runs the REAL kernel bodies with one OS thread per CUDA
// thread.  It exists so the index arithmetic, barrier placement, planner and numerics can be
// checked against the oracle on a box without a GPU.  It is never shipped, never imported by the
// package, and is orders of magnitude slower than anything one would call a fallback.
#include "hs_emu.h"

#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "hs_abi.h"
#include "hs_conv_fft.cuh"
#include "hs_fused.cuh"
#include "hs_generic.cuh"
#include "hs_smem_fft.cuh"
#include "hs_tiled.cuh"

typedef void *hs_stream_t;
static std::atomic<long long> g_launches{0};

template <class P, int SGN, int MINB, int PFW = 0>
static void launch_c2c(const void *in, void *out, const void *tw, long long howmany, hs_stream_t)
{
    const long long blocks = (howmany + P::X - 1) / P::X;
    hs_emu_launch(hs_c2c_smem_kernel<P, SGN, MINB>, blocks, P::threads, P::smem_bytes, (const cplx *)in, (cplx *)out, (const cplx *)tw,
                  howmany, 2);
    ++g_launches;
}

template <class P, int SGN, int MINB>
static void launch_r2c(const void *in, void *out, const void *tw, const void *tw2, long long out_stride, int full, long long howmany, hs_stream_t)
{
    hs_emu_launch(hs_r2c_smem_kernel<P, SGN, MINB>, (howmany + P::X - 1) / P::X, P::threads, P::smem_bytes, (const cplx *)in, (cplx *)out,
                  (const cplx *)tw, (const cplx *)tw2, out_stride, full, howmany);
    ++g_launches;
}
template <class P, int SGN, int MINB>
static void launch_c2r(const void *in, void *out, const void *tw, const void *tw2, long long in_stride, long long howmany, hs_stream_t)
{
    const char *e = getenv("HSFFT_C2R_STAGE");
    if (e ? atoi(e) != 0 : (P::N != 512 && P::N != 8192)) { // the product's rule (hs_device.cu launch_c2r)
        hs_emu_launch(hs_c2r_smem_kernel<P, SGN, MINB, true>, (howmany + P::X - 1) / P::X, P::threads, P::smem_bytes, (const cplx *)in, (cplx *)out,
                      (const cplx *)tw, (const cplx *)tw2, in_stride, howmany);
        ++g_launches;
        return;
    }
    hs_emu_launch(hs_c2r_smem_kernel<P, SGN, MINB>, (howmany + P::X - 1) / P::X, P::threads, P::smem_bytes, (const cplx *)in, (cplx *)out,
                  (const cplx *)tw, (const cplx *)tw2, in_stride, howmany);
    ++g_launches;
}

template <class PF, class PI, int SGN, int MINB>
static void launch_blu(int N, const void *in, void *out, const void *twf, const void *twi, const void *chirp, const void *hk,
                       long long howmany, hs_stream_t)
{
    const long long blocks = (howmany + PF::X - 1) / PF::X;
    hs_emu_launch(hs_bluestein_kernel<PF, PI, SGN, MINB>, blocks, PF::threads, PF::smem_bytes, (const cplx *)in, (cplx *)out,
                  (const cplx *)twf, (const cplx *)twi, (const cplx *)chirp, (const cplx *)hk, N, howmany, 2);
    ++g_launches;
}

template <class PF, class PI, int SGN, int MINB>
static void launch_blu_split(int N, const void *in, void *out, const void *twf, const void *twi, const void *chirp3, const void *hk2,
                             long long howmany, hs_stream_t)
{
    hs_emu_launch(hs_bluestein_split_kernel<PF, PI, SGN, MINB>, howmany, PF::threads, PF::smem_bytes, (const cplx *)in, (cplx *)out,
                  (const cplx *)twf, (const cplx *)twi, (const cplx *)chirp3, (const cplx *)hk2, N, howmany, 2);
    ++g_launches;
}

template <class P, int SGN, int MINB>
static void launch_conv_split(int N, const void *in, void *out, const void *tw, const void *chirp3, const void *hkt2, long long howmany,
                              hs_stream_t)
{
    hs_emu_launch(hs_conv_bluestein_split_kernel<P, SGN, MINB>, howmany, P::threads, P::smem_bytes, (const cplx *)in, (cplx *)out,
                  (const cplx *)tw, (const cplx *)chirp3, (const cplx *)hkt2, N, howmany, 2);
    ++g_launches;
}

template <class P, int SGN, int MINB>
static void launch_conv_split_tmem(int N, const void *in, void *out, const void *tw, const void *h, const void *wp, const void *hkt2,
                                   long long howmany, hs_stream_t)
{
    hs_emu_launch(hs_conv_bluestein_split_tmem_kernel<P, SGN, MINB>, howmany < 2 ? howmany : 2, P::threads, P::smem_bytes + 16, (const cplx *)in, (cplx *)out,
                  (const cplx *)tw, (const cplx *)h, (const cplx *)wp, (const cplx *)hkt2, N, howmany, 2);
    ++g_launches;
}

template <class PF, class PI, int MINB>
static void launch_rader(const void *in, void *out, const void *twf, const void *twi, const void *perm, const void *bk, long long howmany,
                         hs_stream_t)
{
    const long long blocks = (howmany + PF::X - 1) / PF::X;
    const char *e = getenv("HSFFT_RADER_STAGE");
    const int stg = e ? atoi(e) & 3 : 2;
    auto k0 = hs_rader_kernel<PF, PI, MINB, 0>;
    auto k1 = hs_rader_kernel<PF, PI, MINB, 1>;
    auto k2 = hs_rader_kernel<PF, PI, MINB, 2>;
    auto k3 = hs_rader_kernel<PF, PI, MINB, 3>;
    hs_emu_launch(stg == 0 ? k0 : stg == 1 ? k1 : stg == 2 ? k2 : k3, blocks, PF::threads, HsRaderSmem<PF, PI>::bytes, (const cplx *)in,
                  (cplx *)out, (const cplx *)twf, (const cplx *)twi, (const int *)perm, (const cplx *)bk, howmany, 2);
    ++g_launches;
}

template <class PF, class PI, int CS>
static void launch_blu_mid(const void *in, void *out, const void *twf, const void *twi, const HsTileGeom &g, const void *hk,
                           long long howmany, hs_stream_t)
{
    constexpr int C = HsTileCols<PF>::C;
    constexpr int smem = C * hs_col_stride(PF::NPAD, C) * (int)sizeof(cplx);
    hs_emu_launch(hs_blu_mid_kernel<PF, PI, CS, C, 1>, howmany * (g.ncols / C), PF::T * C, smem, (const cplx *)in, (cplx *)out,
                  (const cplx *)twf, (const cplx *)twi, g, (const cplx *)hk, howmany);
    ++g_launches;
}

template <class P, int SGN, int C, int TW, bool STCOL, int IO = 0, int CS = 1>
static void launch_tiled(const void *in, void *out, const void *tw, const HsTileGeom &g, long long howmany, hs_stream_t)
{
    if constexpr (IO == 8 && !STCOL && TW == 0) {
        const char *e = getenv("HSFFT_R4_TMA");
        if (e && atoi(e) != 0 && g.in_c == 1 && g.blk_log == 0 && g.ncols % C == 0) { // persistent TMA-staged kernel with the split epilogue, 3 "SMs"
            const long long total = howmany * (g.ncols / C);
            const HsTmap tm{(const unsigned char *)in, (long long)g.in_e * 16, (long long)g.in_xf * 16};
            hs_emu_launch(hs_tiled_tma_kernel<P, SGN, C, 0, false, 8>, total < 3 ? total : 3, P::T * C, HsTmaTileSmem<P, C>::bytes, tm,
                          (cplx *)out, (const cplx *)tw, g, howmany);
            ++g_launches;
            return;
        }
    }
    if constexpr (IO == 0) {
        const char *e = getenv("HSFFT_TMA");
        if ((!e || atoi(e) != 0) && g.in_c == 1 && g.blk_log == 0) { // persistent TMA-staged kernel, 3 "SMs"
            const long long total = howmany * (g.ncols / C);
            const HsTmap tm{(const unsigned char *)in, (long long)g.in_e * 16, (long long)g.in_xf * 16};
            hs_emu_launch(hs_tiled_tma_kernel<P, SGN, C, TW, STCOL>, total < 3 ? total : 3, P::T * C, HsTmaTileSmem<P, C>::bytes, tm,
                          (cplx *)out, (const cplx *)tw, g, howmany);
            ++g_launches;
            return;
        }
    }
    constexpr int smem = C * hs_col_stride(P::NPAD, C) * (int)sizeof(cplx);
    hs_emu_launch(hs_tiled_kernel<P, SGN, C, TW, STCOL, 1, IO, CS>, howmany * ((g.ncols + C - 1) / C), P::T * C, smem, (const cplx *)in, (cplx *)out,
                  (const cplx *)tw, g, howmany);
    ++g_launches;
}

#include "hs_ct_table.cuh"

extern "C" {
// HSFFT_EMU_DEVICES=n pretends to have n devices (one address space; the "current device" is per host thread, as in CUDA):
// exercises the (plan, device) registry and the batch sharding of hsfft_set_devices on a box without a GPU
static thread_local int t_dev = 0;
int hs_cu_device_count(void) { const char *e = getenv("HSFFT_EMU_DEVICES"); const int n = e ? atoi(e) : 1; return n > 0 ? n : 1; }
void hs_cu_require_device(void) {}
int hs_cu_get_device(void) { return t_dev; }
void hs_cu_set_device(int d) { t_dev = d; }
int hs_cu_sm_count(void) { return 148; }
void *hs_cu_malloc(size_t bytes) { return malloc(bytes ? bytes : 16); }
void hs_cu_free(void *d) { free(d); }
void *hs_cu_malloc_async(size_t bytes, void *) { return malloc(bytes ? bytes : 16); }
void hs_cu_free_async(void *d, void *) { free(d); }
void hs_cu_device_sync(void) {}
int hs_cu_can_access_peer(int, int) { return 1; }
void hs_cu_enable_peer_access(int) {}
void hs_cu_h2d_2d(void *d, size_t dpitch, const void *h, size_t hpitch, size_t width_bytes, size_t rows, void *)
{
    for (size_t r = 0; r < rows; ++r) memcpy((char *)d + r * dpitch, (const char *)h + r * hpitch, width_bytes);
}
void hs_cu_d2h_2d(void *h, size_t hpitch, const void *d, size_t dpitch, size_t width_bytes, size_t rows, void *)
{
    for (size_t r = 0; r < rows; ++r) memcpy((char *)h + r * hpitch, (const char *)d + r * dpitch, width_bytes);
}
void *hs_cu_malloc_host(size_t bytes) { return malloc(bytes ? bytes : 16); }
void hs_cu_free_host(void *h) { free(h); }
void hs_cu_h2d(void *d, const void *h, size_t bytes, void *) { if (bytes) memcpy(d, h, bytes); }
void hs_cu_d2h(void *h, const void *d, size_t bytes, void *) { if (bytes) memcpy(h, d, bytes); }
void hs_cu_d2d(void *dst, const void *src, size_t bytes, void *) { if (bytes) memmove(dst, src, bytes); }
void hs_cu_memset0(void *d, size_t bytes, void *) { if (bytes) memset(d, 0, bytes); }
void *hs_cu_stream_create(void) { return (void *)1; }
void hs_cu_stream_destroy(void *) {}
void hs_cu_stream_sync(void *) {}
void *hs_cu_event_create(void) { return (void *)1; }
void hs_cu_event_destroy(void *) {}
void hs_cu_event_record(void *, void *) {}
void hs_cu_stream_wait_event(void *, void *) {}
float hs_cu_event_elapsed_ms(void *, void *) { return 0.f; }
void hs_cu_event_sync(void *) {}
long long hs_cu_launch_count(void) { return g_launches.load(); }
void hs_cu_timing_begin(void *) {}
int hs_cu_timing_report(char *out, int cap) { if (cap > 0) out[0] = 0; return 0; }
void hs_cu_ipc_get(void *d, void *handle64) { memset(handle64, 0, 64); memcpy(handle64, &d, sizeof d); } // one address space here
void *hs_cu_ipc_open(const void *handle64) { void *p; memcpy(&p, handle64, sizeof p); return p; }
void hs_cu_ipc_close(void *) {}
int hs_cu_jit_c2c_create(int, int, const int *, int, hs_ct_desc *) { return -1; } // no NVRTC here: the run-time engine serves these lengths
void hs_cu_jit_c2c_launch(int, const void *, void *, const void *, long long, void *) {}
int hs_cu_jit_source(int, int, const int *, int, char *, int) { return 0; }
int hs_cu_jit_has_real(int) { return 0; }
int hs_cu_jit_is_staged(int) { return 0; }
int hs_cu_jit_rader_create(int, const int *, int, hs_ct_desc *) { return -1; }
void hs_cu_jit_rader_launch(int, const void *, void *, const void *, const void *, const void *, const void *, long long, void *) {}
int hs_cu_jit_twopass_create(long long, int, const int *, int, int, hs_ct_desc *) { return -1; }
void hs_cu_jit_twopass_launch(int, int, const void *, void *, const void *, const void *, const void *, long long, void *) {}
void hs_cu_jit_real_launch(int, int, const void *, void *, const void *, const void *, long long, int, long long, void *) {}

void hs_cu_launch_generic(int sgn, const void *d_in, void *d_out, const hs_gen_params *p, void *)
{
    const long long ntile = (p->n_c + p->C - 1) / p->C;
    const long long blocks = ntile * p->n_o * p->n_batch;
    const size_t smem = (size_t)2 * p->C * p->npad * sizeof(cplx);
    const int threads = 64; // the kernel strides by blockDim; fewer OS threads keep the emulation quick
    if (sgn == 1) hs_emu_launch(hs_generic_pass_kernel<1>, blocks, threads, smem, (const cplx *)d_in, (cplx *)d_out, *p);
    else hs_emu_launch(hs_generic_pass_kernel<-1>, blocks, threads, smem, (const cplx *)d_in, (cplx *)d_out, *p);
    ++g_launches;
}

static long long ew_blocks(long long total) { long long b = (total + 63) / 64; return b > 4 ? 4 : (b < 1 ? 1 : b); }

void hs_cu_launch_chirp_pad(int sgn, const void *x, void *a, const void *h, int N, int M, long long howmany, void *)
{
    if (sgn == 1) hs_emu_launch(hs_chirp_pad_kernel<1>, ew_blocks(howmany * M), 64, 0, (const cplx *)x, (cplx *)a, (const cplx *)h, N, M, howmany);
    else hs_emu_launch(hs_chirp_pad_kernel<-1>, ew_blocks(howmany * M), 64, 0, (const cplx *)x, (cplx *)a, (const cplx *)h, N, M, howmany);
    ++g_launches;
}
void hs_cu_launch_spectrum_mul(int sgn, void *A, const void *Hk, int M, long long howmany, void *)
{
    if (sgn == 1) hs_emu_launch(hs_spectrum_mul_kernel<1>, ew_blocks(howmany * M), 64, 0, (cplx *)A, (const cplx *)Hk, M, howmany);
    else hs_emu_launch(hs_spectrum_mul_kernel<-1>, ew_blocks(howmany * M), 64, 0, (cplx *)A, (const cplx *)Hk, M, howmany);
    ++g_launches;
}
void hs_cu_launch_chirp_crop(int sgn, const void *y, void *out, const void *h, int N, int M, long long howmany, void *)
{
    if (sgn == 1) hs_emu_launch(hs_chirp_crop_kernel<1>, ew_blocks(howmany * N), 64, 0, (const cplx *)y, (cplx *)out, (const cplx *)h, N, M, howmany);
    else hs_emu_launch(hs_chirp_crop_kernel<-1>, ew_blocks(howmany * N), 64, 0, (const cplx *)y, (cplx *)out, (const cplx *)h, N, M, howmany);
    ++g_launches;
}
void hs_cu_launch_r2c_post(const void *Z, void *out, const void *tw2, int H, long long out_stride, int full, long long howmany, void *)
{
    hs_emu_launch(hs_r2c_post_kernel, ew_blocks(howmany * (H / 2 + 1)), 64, 0, (const cplx *)Z, (cplx *)out, (const cplx *)tw2, H, out_stride, full,
                  howmany);
    ++g_launches;
}
void hs_cu_launch_c2r_pre(const void *in, void *z, const void *tw2, int H, long long in_stride, long long howmany, void *)
{
    hs_emu_launch(hs_c2r_pre_kernel, ew_blocks(howmany * H), 64, 0, (const cplx *)in, (cplx *)z, (const cplx *)tw2, H, in_stride, howmany);
    ++g_launches;
}
void hs_cu_launch_conv_mul(const void *A, const void *B, void *C, long long total, double scale, void *)
{
    hs_emu_launch(hs_conv_mul_kernel, ew_blocks(total), 64, 0, (const cplx *)A, (const cplx *)B, (cplx *)C, total, scale);
    ++g_launches;
}
void hs_cu_launch_pad_rows(const void *src, int len, void *dst, int P, long long howmany, void *)
{
    hs_emu_launch(hs_pad_rows_kernel, ew_blocks(howmany * P), 64, 0, (const double *)src, len, (double *)dst, P, howmany);
    ++g_launches;
}
void hs_cu_launch_dist_twiddle(void *b, const void *hi, const void *lo, long long rows, long long cols, long long c0, int logN, int logS,
                               void *)
{
    hs_emu_launch(hs_dist_twiddle_kernel<0>, ew_blocks(rows * cols), 64, 0, (cplx *)b, (const cplx *)hi, (const cplx *)lo, rows, cols, c0, logN,
                  logS);
    ++g_launches;
}
void hs_cu_launch_transpose(const void *in, void *out, int rows, int cols, long long howmany, void *)
{
    hs_emu_launch(hs_transpose_kernel, 3, 64, 32 * 33 * sizeof(cplx), (const cplx *)in, (cplx *)out, rows, cols, howmany);
    ++g_launches;
}
void hs_cu_launch_slice_rows(const void *src, int P, int start, int len, void *dst, long long dst_stride, long long howmany, void *)
{
    hs_emu_launch(hs_slice_rows_kernel, ew_blocks(howmany * len), 64, 0, (const double *)src, P, start, len, (double *)dst, dst_stride, howmany);
    ++g_launches;
}
}

// unit hook for tests: one odd-radix butterfly with compile-time tables (hs_butterflies.cuh), v[r] in place
extern "C" int hs_emu_bfly_odd(int r, int sgn, double *v)
{
    cplx *c = reinterpret_cast<cplx *>(v);
    if (r == 11) { if (sgn == 1) hs_bfly<11, 1>(c); else hs_bfly<11, -1>(c); }
    else if (r == 13) { if (sgn == 1) hs_bfly<13, 1>(c); else hs_bfly<13, -1>(c); }
    else if (r == 17) { if (sgn == 1) hs_bfly<17, 1>(c); else hs_bfly<17, -1>(c); }
    else return -1;
    return 0;
}

