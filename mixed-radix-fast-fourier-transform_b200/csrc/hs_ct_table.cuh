// hs_ct_table.cuh -- the table of compile-time specialised plans and their dispatch.
// Included by hs_device.cu (CUDA launchers) and by tests/emu/hs_device_emu.cpp (CPU emulation of the
// same kernels); the including file provides launch_c2c<P,SGN,MINB>() / launch_blu<PF,PI,SGN,MINB>()
// and the hs_stream_t type before including this header.
#pragma once
#include <cstdio>
#include <cstdlib>
#include "hs_abi.h"
#include "hs_fused.cuh"
#include "hs_smem_fft.cuh"
#include "hs_tiled.cuh"

#include "hs_plans.cuh"

#ifndef HS_PART
#define HS_PART(n) 1
#endif

#if HS_PART(1)
template <class P>
static void fill_desc(hs_ct_desc &d, int id, int family)
{
    d.id = id;
    d.family = family;
    d.n = P::N;
    d.nst = P::NST;
    for (int s = 0; s < 8; ++s) d.radix[s] = s < P::NST ? P::radix(s) : 0;
    d.split_min = P::SPLIT;
    d.threads = P::threads;
    d.smem_bytes = P::smem_bytes;
    d.n1 = d.n2 = d.nst2 = d.split_min2 = 0;
    d.exact = P::EXACT ? 1 : 0;
    for (int s = 0; s < 8; ++s) d.radix2[s] = 0;
}

template <class P1, class P2>
static void fill_desc2(hs_ct_desc &d, int id)
{
    fill_desc<P1>(d, id, HS_CT_TWOPASS);
    d.n = P1::N * P2::N;
    d.n1 = P1::N;
    d.n2 = P2::N;
    d.nst2 = P2::NST;
    for (int s = 0; s < 8; ++s) d.radix2[s] = s < P2::NST ? P2::radix(s) : 0;
    d.split_min2 = P2::SPLIT;
    d.exact = (P1::EXACT || P2::EXACT) ? 1 : 0;
}

static hs_ct_desc g_ct[CT_COUNT];
static int g_ct_order[CT_COUNT];
static bool g_ct_ready = false;

static void ct_init()
{
    if (g_ct_ready) return;
    fill_desc<P4096a>(g_ct[CT_4096A], CT_4096A, HS_CT_C2C);
    fill_desc<P4096b>(g_ct[CT_4096B], CT_4096B, HS_CT_C2C);
    fill_desc<P3000>(g_ct[CT_3000], CT_3000, HS_CT_C2C);
    fill_desc<P1024>(g_ct[CT_1024], CT_1024, HS_CT_C2C);
    fill_desc<P2048>(g_ct[CT_2048], CT_2048, HS_CT_C2C);
    fill_desc<P512>(g_ct[CT_512], CT_512, HS_CT_C2C);
    fill_desc<P256>(g_ct[CT_256], CT_256, HS_CT_C2C);
    fill_desc<P8192>(g_ct[CT_8192], CT_8192, HS_CT_C2C);
    fill_desc<PB2048f>(g_ct[CT_B2048], CT_B2048, HS_CT_BLUESTEIN);
    fill_desc<PB256>(g_ct[CT_B256], CT_B256, HS_CT_BLUESTEIN);
    fill_desc<PB512>(g_ct[CT_B512], CT_B512, HS_CT_BLUESTEIN);
    fill_desc<PB1024>(g_ct[CT_B1024], CT_B1024, HS_CT_BLUESTEIN);
    fill_desc<PB4096>(g_ct[CT_B4096], CT_B4096, HS_CT_BLUESTEIN);
    fill_desc<PB8192f>(g_ct[CT_B8192], CT_B8192, HS_CT_BLUESTEIN);
    fill_desc<PR1008f>(g_ct[CT_R1008], CT_R1008, HS_CT_RADER);
    fill_desc<PR1008af>(g_ct[CT_R1008A], CT_R1008A, HS_CT_RADER);
    fill_desc<PR1008gf>(g_ct[CT_R1008G], CT_R1008G, HS_CT_RADER);
    fill_desc<PB256>(g_ct[CT_R256], CT_R256, HS_CT_RADER);
    fill_desc<PB2048x2f>(g_ct[CT_B2048X2], CT_B2048X2, HS_CT_BLUESTEIN);
    fill_desc<PB256x16>(g_ct[CT_B256X16], CT_B256X16, HS_CT_BLUESTEIN);
    fill_desc<PB4096>(g_ct[CT_H4096], CT_H4096, HS_CT_BLUESTEIN_SPLIT);
    fill_desc<PB8192f>(g_ct[CT_H8192], CT_H8192, HS_CT_BLUESTEIN_SPLIT);
    fill_desc<PC10240>(g_ct[CT_C10240], CT_C10240, HS_CT_CONV_SPLIT);
    fill_desc<PC8192>(g_ct[CT_C8192], CT_C8192, HS_CT_CONV_SPLIT);
    fill_desc<P4096x>(g_ct[CT_4096X], CT_4096X, HS_CT_C2C);
    fill_desc<P3000x>(g_ct[CT_3000X], CT_3000X, HS_CT_C2C);
    fill_desc<P4096y>(g_ct[CT_4096Y], CT_4096Y, HS_CT_C2C);
    fill_desc<P3000y>(g_ct[CT_3000Y], CT_3000Y, HS_CT_C2C);
    fill_desc<P3000z>(g_ct[CT_3000Z], CT_3000Z, HS_CT_C2C);
    fill_desc<P3000t>(g_ct[CT_3000T], CT_3000T, HS_CT_C2C);
    fill_desc<P3000v>(g_ct[CT_3000V], CT_3000V, HS_CT_C2C);
    fill_desc<P3000u>(g_ct[CT_3000U], CT_3000U, HS_CT_C2C);
    fill_desc<P8192q>(g_ct[CT_8192Q], CT_8192Q, HS_CT_C2C);
    fill_desc<P8192b>(g_ct[CT_8192B], CT_8192B, HS_CT_C2C);
    fill_desc2<S128, S128>(g_ct[CT_T14], CT_T14);
    fill_desc2<S128, S256>(g_ct[CT_T15], CT_T15);
    fill_desc2<S256, S256>(g_ct[CT_T16], CT_T16);
    fill_desc2<HS_T17_P1, HS_T17_P2>(g_ct[CT_T17], CT_T17);
    fill_desc2<HS_T18_P1, HS_T18_P2>(g_ct[CT_T18], CT_T18);
    fill_desc2<S1024, S512b>(g_ct[CT_T19], CT_T19); // (512 x 1024 before: 3.73 ms per 2^19 x 512 against ... with the 1024-point pass staged first)
    fill_desc2<S1024, S1024>(g_ct[CT_T20], CT_T20);
    fill_desc<S128>(g_ct[CT_S128], CT_S128, HS_CT_SUB);
    fill_desc<S256>(g_ct[CT_S256], CT_S256, HS_CT_SUB);
    fill_desc<S512>(g_ct[CT_S512], CT_S512, HS_CT_SUB);
    fill_desc<S1024>(g_ct[CT_S1024], CT_S1024, HS_CT_SUB);
    fill_desc2<S160x, S128>(g_ct[CT_TX20480], CT_TX20480);
    fill_desc2<S160x, S256>(g_ct[CT_TX40960], CT_TX40960);
    fill_desc2<S160x, S512b>(g_ct[CT_TX81920], CT_TX81920);
    fill_desc2<S160x, S1024>(g_ct[CT_TX163840], CT_TX163840);
    fill_desc<S160x>(g_ct[CT_S160X], CT_S160X, HS_CT_SUB);
    fill_desc<S512b>(g_ct[CT_S512B], CT_S512B, HS_CT_SUB2);
    for (int i = 0; i < CT_COUNT; ++i) g_ct_order[i] = i;
    // HSFFT_PREFER_CT=<id>: list that instance first (A/B runs of alternative plans for one N)
    if (const char *e = getenv("HSFFT_PREFER_CT")) {
        const int want = atoi(e);
        if (want > 0 && want < CT_COUNT) { g_ct_order[0] = want; for (int i = 1; i <= want; ++i) g_ct_order[i] = i - 1; }
    }
    // HSFFT_NO_CT=1: hide the compile-time kernels (exercise the run-time planned engine)
    g_ct_ready = true;
}

extern "C" int hs_cu_ct_count(void)
{
    ct_init();
    if (const char *e = getenv("HSFFT_NO_CT")) if (atoi(e) != 0) return 0;
    return CT_COUNT;
}
extern "C" const hs_ct_desc *hs_cu_ct_get(int idx)
{
    ct_init();
    return (idx >= 0 && idx < CT_COUNT) ? &g_ct[g_ct_order[idx]] : nullptr;
}

template <class P, int MINB, int PFW = 0>
static void launch_c2c_sgn(int sgn, const void *in, void *out, const void *tw, long long howmany, hs_stream_t st)
{
    if (sgn == 1) launch_c2c<P, 1, MINB, PFW>(in, out, tw, howmany, st);
    else launch_c2c<P, -1, MINB, PFW>(in, out, tw, howmany, st);
}

extern "C" void hs_cu_launch_ct_c2c(int id, int sgn, const void *d_in, void *d_out, const void *d_tw, long long howmany, void *stream)
{
    hs_stream_t st = (hs_stream_t)stream;
    switch (id) {
    case CT_4096A: launch_c2c_sgn<P4096a, 2>(sgn, d_in, d_out, d_tw, howmany, st); break;
    case CT_4096B: launch_c2c_sgn<P4096b, 2>(sgn, d_in, d_out, d_tw, howmany, st); break;
    case CT_3000: launch_c2c_sgn<P3000, 3>(sgn, d_in, d_out, d_tw, howmany, st); break;
    case CT_1024: launch_c2c_sgn<P1024, 2>(sgn, d_in, d_out, d_tw, howmany, st); break;
    case CT_2048: launch_c2c_sgn<P2048, 2>(sgn, d_in, d_out, d_tw, howmany, st); break;
    case CT_512: launch_c2c_sgn<P512, 3, 1>(sgn, d_in, d_out, d_tw, howmany, st); break;
    case CT_256: launch_c2c_sgn<P256, 3, 1>(sgn, d_in, d_out, d_tw, howmany, st); break;
    case CT_8192: launch_c2c_sgn<P8192, 1, 1>(sgn, d_in, d_out, d_tw, howmany, st); break;
    case CT_4096X: launch_c2c_sgn<P4096x, 2>(sgn, d_in, d_out, d_tw, howmany, st); break;
    case CT_3000X: launch_c2c_sgn<P3000x, 2>(sgn, d_in, d_out, d_tw, howmany, st); break;
    case CT_4096Y: launch_c2c_sgn<P4096y, 2>(sgn, d_in, d_out, d_tw, howmany, st); break;
    case CT_3000Y: launch_c2c_sgn<P3000y, 3>(sgn, d_in, d_out, d_tw, howmany, st); break;
    case CT_3000Z: launch_c2c_sgn<P3000z, 3>(sgn, d_in, d_out, d_tw, howmany, st); break;
    case CT_3000T: launch_c2c_sgn<P3000t, 2>(sgn, d_in, d_out, d_tw, howmany, st); break;
    case CT_3000V: launch_c2c_sgn<P3000v, 3>(sgn, d_in, d_out, d_tw, howmany, st); break;
    case CT_3000U: launch_c2c_sgn<P3000u, 3>(sgn, d_in, d_out, d_tw, howmany, st); break;
    case CT_8192Q: launch_c2c_sgn<P8192q, 1, 1>(sgn, d_in, d_out, d_tw, howmany, st); break;
    case CT_8192B: launch_c2c_sgn<P8192b, 1, 1>(sgn, d_in, d_out, d_tw, howmany, st); break;
    default: fprintf(stderr, "Error: unknown compile-time kernel id %d\n", id); exit(EXIT_FAILURE);
    }
}

template <class P, int MINB>
static void launch_real_sgn(int mode, int sgn, const void *in, void *out, const void *tw, const void *tw2, long long stride, int full,
                            long long howmany, hs_stream_t st)
{
    if (mode == 1) { if (sgn == 1) launch_r2c<P, 1, MINB>(in, out, tw, tw2, stride, full, howmany, st); else launch_r2c<P, -1, MINB>(in, out, tw, tw2, stride, full, howmany, st); }
    else           { if (sgn == 1) launch_c2r<P, 1, MINB>(in, out, tw, tw2, stride, howmany, st); else launch_c2r<P, -1, MINB>(in, out, tw, tw2, stride, howmany, st); }
}

extern "C" void hs_cu_launch_ct_real(int id, int mode, int sgn, const void *d_in, void *d_out, const void *d_tw, const void *d_tw2, long long stride,
                                     int full, long long howmany, void *stream)
{
    hs_stream_t st = (hs_stream_t)stream;
    switch (id) {
    case CT_4096A: launch_real_sgn<P4096a, 2>(mode, sgn, d_in, d_out, d_tw, d_tw2, stride, full, howmany, st); break;
    case CT_3000: // c2r keeps the 200-thread plan (same stages, same tables): N = 6000 59.1 % of the HBM peak against 53.9 % at 125 threads; r2c 74.8 -> 81.0 %
        if (mode == 1) launch_real_sgn<P3000, 3>(mode, sgn, d_in, d_out, d_tw, d_tw2, stride, full, howmany, st);
        else launch_real_sgn<P3000t, 2>(mode, sgn, d_in, d_out, d_tw, d_tw2, stride, full, howmany, st);
        break;
    case CT_1024: launch_real_sgn<P1024, 2>(mode, sgn, d_in, d_out, d_tw, d_tw2, stride, full, howmany, st); break;
    case CT_2048: launch_real_sgn<P2048, 2>(mode, sgn, d_in, d_out, d_tw, d_tw2, stride, full, howmany, st); break;
    case CT_512: launch_real_sgn<P512, 3>(mode, sgn, d_in, d_out, d_tw, d_tw2, stride, full, howmany, st); break;
    case CT_256: launch_real_sgn<P256, 3>(mode, sgn, d_in, d_out, d_tw, d_tw2, stride, full, howmany, st); break;
    case CT_8192: launch_real_sgn<P8192, 1>(mode, sgn, d_in, d_out, d_tw, d_tw2, stride, full, howmany, st); break;
    default: fprintf(stderr, "Error: compile-time kernel %d has no real-transform form\n", id); exit(EXIT_FAILURE);
    }
}
/* which one-kernel plans hs_cu_launch_ct_real serves */
extern "C" int hs_cu_ct_has_real(int id)
{
    return id == CT_4096A || id == CT_3000 || id == CT_1024 || id == CT_2048 || id == CT_512 || id == CT_256 || id == CT_8192;
}

extern "C" void hs_cu_launch_ct_rader(int id, const void *d_in, void *d_out, const void *d_tw_f, const void *d_tw_i, const void *d_perm,
                                      const void *d_bk, long long howmany, void *stream)
{
    hs_stream_t st = (hs_stream_t)stream;
    switch (id) {
    case CT_R1008: launch_rader<PR1008f, PR1008i, 5>(d_in, d_out, d_tw_f, d_tw_i, d_perm, d_bk, howmany, st); break;
    case CT_R1008A: launch_rader<PR1008af, PR1008ai, 8>(d_in, d_out, d_tw_f, d_tw_i, d_perm, d_bk, howmany, st); break;
    case CT_R1008G: launch_rader<PR1008gf, PR1008gi, 6>(d_in, d_out, d_tw_f, d_tw_i, d_perm, d_bk, howmany, st); break;
    case CT_R256: launch_rader<PB256, PB256, 8>(d_in, d_out, d_tw_f, d_tw_i, d_perm, d_bk, howmany, st); break;
    default: fprintf(stderr, "Error: unknown Rader kernel id %d\n", id); exit(EXIT_FAILURE);
    }
}


#define HS_BLU_CASE(ID, PF, PI, MINB)                                                                                     \
    case ID:                                                                                                              \
        if (sgn == 1) launch_blu<PF, PI, 1, MINB>(N, d_in, d_out, d_tw_f, d_tw_i, d_chirp, d_hk, howmany, st);            \
        else launch_blu<PF, PI, -1, MINB>(N, d_in, d_out, d_tw_f, d_tw_i, d_chirp, d_hk, howmany, st);                    \
        break;

extern "C" void hs_cu_launch_ct_bluestein(int id, int sgn, int N, const void *d_in, void *d_out, const void *d_tw_f, const void *d_tw_i,
                                          const void *d_chirp, const void *d_hk, long long howmany, void *stream)
{
    hs_stream_t st = (hs_stream_t)stream;
    switch (id) {
    HS_BLU_CASE(CT_B2048X2, PB2048x2f, PB2048x2i, 2)
    HS_BLU_CASE(CT_B256X16, PB256x16, PB256x16, 2)
    case CT_B2048:
        if (sgn == 1) launch_blu<PB2048f, PB2048i, 1, 4>(N, d_in, d_out, d_tw_f, d_tw_i, d_chirp, d_hk, howmany, st);
        else launch_blu<PB2048f, PB2048i, -1, 4>(N, d_in, d_out, d_tw_f, d_tw_i, d_chirp, d_hk, howmany, st);
        break;
    case CT_B256:
        if (sgn == 1) launch_blu<PB256, PB256, 1, 8>(N, d_in, d_out, d_tw_f, d_tw_i, d_chirp, d_hk, howmany, st);
        else launch_blu<PB256, PB256, -1, 8>(N, d_in, d_out, d_tw_f, d_tw_i, d_chirp, d_hk, howmany, st);
        break;
    case CT_B512:
        if (sgn == 1) launch_blu<PB512, PB512, 1, 12>(N, d_in, d_out, d_tw_f, d_tw_i, d_chirp, d_hk, howmany, st);
        else launch_blu<PB512, PB512, -1, 12>(N, d_in, d_out, d_tw_f, d_tw_i, d_chirp, d_hk, howmany, st);
        break;
    case CT_B1024:
        if (sgn == 1) launch_blu<PB1024, PB1024, 1, 8>(N, d_in, d_out, d_tw_f, d_tw_i, d_chirp, d_hk, howmany, st);
        else launch_blu<PB1024, PB1024, -1, 8>(N, d_in, d_out, d_tw_f, d_tw_i, d_chirp, d_hk, howmany, st);
        break;
    case CT_B4096:
        if (sgn == 1) launch_blu<PB4096, PB4096, 1, 2>(N, d_in, d_out, d_tw_f, d_tw_i, d_chirp, d_hk, howmany, st);
        else launch_blu<PB4096, PB4096, -1, 2>(N, d_in, d_out, d_tw_f, d_tw_i, d_chirp, d_hk, howmany, st);
        break;
    case CT_B8192:
        if (sgn == 1) launch_blu<PB8192f, PB8192i, 1, 1>(N, d_in, d_out, d_tw_f, d_tw_i, d_chirp, d_hk, howmany, st);
        else launch_blu<PB8192f, PB8192i, -1, 1>(N, d_in, d_out, d_tw_f, d_tw_i, d_chirp, d_hk, howmany, st);
        break;
    default: fprintf(stderr, "Error: unknown fused Bluestein kernel id %d\n", id); exit(EXIT_FAILURE);
    }
}


#define HS_BLUSPLIT_CASE(ID, PF, PI, MINB)                                                                                 \
    case ID:                                                                                                              \
        if (sgn == 1) launch_blu_split<PF, PI, 1, MINB>(N, d_in, d_out, d_tw_f, d_tw_i, d_chirp3, d_hk2, howmany, st);    \
        else launch_blu_split<PF, PI, -1, MINB>(N, d_in, d_out, d_tw_f, d_tw_i, d_chirp3, d_hk2, howmany, st);            \
        break;

extern "C" void hs_cu_launch_ct_bluestein_split(int id, int sgn, int N, const void *d_in, void *d_out, const void *d_tw_f, const void *d_tw_i,
                                                const void *d_chirp3, const void *d_hk2, long long howmany, void *stream)
{
    hs_stream_t st = (hs_stream_t)stream;
    switch (id) {
    HS_BLUSPLIT_CASE(CT_H4096, PB4096, PB4096, 2)
    HS_BLUSPLIT_CASE(CT_H8192, PB8192f, PB8192i, 1)
    default: fprintf(stderr, "Error: unknown split Bluestein kernel id %d\n", id); exit(EXIT_FAILURE);
    }
}

// intermediates of the split chirp-z kernel in tensor memory (default) or in the L2: N = 10007 x 16384 4.33 vs 5.12 ms
#ifndef HS_SPLIT_TMEM_DEFAULT
#define HS_SPLIT_TMEM_DEFAULT 1
#endif
template <class P, int SGN, int MINB>
static void launch_conv_split_any(int tmem, int N, const void *d_in, void *d_out, const void *d_tw, const void *d_chirp3, const void *d_wp,
                                  const void *d_hkt2, long long howmany, hs_stream_t st)
{
    if constexpr (P::N == 16 * P::T && P::radix(0) == 16) { // the tensor-memory variant wants one radix-16 butterfly per thread on the outer stage
        if (tmem) { launch_conv_split_tmem<P, SGN, MINB>(N, d_in, d_out, d_tw, d_chirp3, d_wp, d_hkt2, howmany, st); return; }
    }
    launch_conv_split<P, SGN, MINB>(N, d_in, d_out, d_tw, d_chirp3, d_hkt2, howmany, st);
}
#define HS_CONVSPLIT_CASE(ID, P, MINB)                                                                                        \
    case ID:                                                                                                                 \
        if (sgn == 1) launch_conv_split_any<P, 1, MINB>(tmem, N, d_in, d_out, d_tw, d_chirp3, d_wp, d_hkt2, howmany, st);   \
        else launch_conv_split_any<P, -1, MINB>(tmem, N, d_in, d_out, d_tw, d_chirp3, d_wp, d_hkt2, howmany, st);           \
        break;

extern "C" void hs_cu_launch_ct_conv_split(int id, int sgn, int N, const void *d_in, void *d_out, const void *d_tw, const void *d_chirp3,
                                           const void *d_wp, const void *d_hkt2, long long howmany, void *stream)
{
    hs_stream_t st = (hs_stream_t)stream;
    static int tmem = -1; // intermediates parked in tensor memory (d_wp = the w[p] table) or in the L2
    if (tmem < 0) { const char *e = getenv("HSFFT_SPLIT_TMEM"); tmem = e ? atoi(e) : HS_SPLIT_TMEM_DEFAULT; }
    switch (id) {
    HS_CONVSPLIT_CASE(CT_C10240, PC10240, 1)
    HS_CONVSPLIT_CASE(CT_C8192, PC8192, 1)
    default: fprintf(stderr, "Error: unknown split Bluestein kernel id %d\n", id); exit(EXIT_FAILURE);
    }
}

#endif // HS_PART(1)

#if HS_PART(2)
// ---- two-pass transforms: C columns per CTA, chosen so that a CTA has 128..512 threads

template <class P1, class P2>
static void launch_twopass(int pass, int sgn, const void *in, void *out, const void *tw, const void *twb, const void *tws,
                           long long howmany, hs_stream_t st)
{
    constexpr int N = P1::N * P2::N;
    HsTileGeom g;
    g.in_xf = g.out_xf = N;
    g.blk_log = 0; g.blk_stride = 0; g.tw_ld = 0;
    if (pass == 1) {
        g.ncols = P2::N; g.in_c = 1; g.in_e = P2::N; g.out_c = P1::N; g.out_e = 1; g.twb = nullptr; g.tws = nullptr;
        if (sgn == 1) launch_tiled<P1, 1, HsTileCols<P1>::C, 0, false>(in, out, tw, g, howmany, st);
        else launch_tiled<P1, -1, HsTileCols<P1>::C, 0, false>(in, out, tw, g, howmany, st);
    } else {
        g.ncols = P1::N; g.in_c = 1; g.in_e = P1::N; g.out_c = 1; g.out_e = P1::N; g.twb = twb; g.tws = tws; g.tw_ld = P1::N;
        if (sgn == 1) launch_tiled<P2, 1, HsTileCols<P2>::C, 1, true>(in, out, tw, g, howmany, st);
        else launch_tiled<P2, -1, HsTileCols<P2>::C, 1, true>(in, out, tw, g, howmany, st);
    }
}

extern "C" void hs_cu_launch_ct_twopass(int id, int pass, int sgn, const void *d_in, void *d_out, const void *d_tw, const void *d_twb,
                                        const void *d_tws, long long howmany, void *stream)
{
    hs_stream_t st = (hs_stream_t)stream;
    switch (id) {
    case CT_T14: launch_twopass<S128, S128>(pass, sgn, d_in, d_out, d_tw, d_twb, d_tws, howmany, st); break;
    case CT_T15: launch_twopass<S128, S256>(pass, sgn, d_in, d_out, d_tw, d_twb, d_tws, howmany, st); break;
    case CT_T16: launch_twopass<S256, S256>(pass, sgn, d_in, d_out, d_tw, d_twb, d_tws, howmany, st); break;
    case CT_T17: launch_twopass<HS_T17_P1, HS_T17_P2>(pass, sgn, d_in, d_out, d_tw, d_twb, d_tws, howmany, st); break;
    case CT_T18: launch_twopass<HS_T18_P1, HS_T18_P2>(pass, sgn, d_in, d_out, d_tw, d_twb, d_tws, howmany, st); break;
    case CT_T19: launch_twopass<S1024, S512b>(pass, sgn, d_in, d_out, d_tw, d_twb, d_tws, howmany, st); break;
    case CT_T20: launch_twopass<S1024, S1024>(pass, sgn, d_in, d_out, d_tw, d_twb, d_tws, howmany, st); break;
    case CT_TX20480: launch_twopass<S160x, S128>(pass, sgn, d_in, d_out, d_tw, d_twb, d_tws, howmany, st); break;
    case CT_TX40960: launch_twopass<S160x, S256>(pass, sgn, d_in, d_out, d_tw, d_twb, d_tws, howmany, st); break;
    case CT_TX81920: launch_twopass<S160x, S512b>(pass, sgn, d_in, d_out, d_tw, d_twb, d_tws, howmany, st); break;
    case CT_TX163840: launch_twopass<S160x, S1024>(pass, sgn, d_in, d_out, d_tw, d_twb, d_tws, howmany, st); break;
    default: fprintf(stderr, "Error: unknown two-pass kernel id %d\n", id); exit(EXIT_FAILURE);
    }
}

// ---- sub-plans with explicit geometry (strided batches, distributed four-step)
template <class P>
static void launch_sub(int sgn, int tw_mode, int stcol, const void *in, void *out, const void *tw, const hs_tile_geom &g, long long howmany,
                       hs_stream_t st)
{
    constexpr int C = HsTileCols<P>::C;
    if (g.ncols % C != 0) { fprintf(stderr, "Error: tiled pass needs a multiple of %d columns (got %d)\n", C, g.ncols); exit(EXIT_FAILURE); }
    if (tw_mode == 0 && !stcol) { if (sgn == 1) launch_tiled<P, 1, C, 0, false>(in, out, tw, g, howmany, st); else launch_tiled<P, -1, C, 0, false>(in, out, tw, g, howmany, st); }
    else if (tw_mode == 0)      { if (sgn == 1) launch_tiled<P, 1, C, 0, true>(in, out, tw, g, howmany, st); else launch_tiled<P, -1, C, 0, true>(in, out, tw, g, howmany, st); }
    else if (tw_mode == 1)      { if (sgn == 1) launch_tiled<P, 1, C, 1, true>(in, out, tw, g, howmany, st); else launch_tiled<P, -1, C, 1, true>(in, out, tw, g, howmany, st); }
    else                        { if (sgn == 1) launch_tiled<P, 1, C, 2, true>(in, out, tw, g, howmany, st); else launch_tiled<P, -1, C, 2, true>(in, out, tw, g, howmany, st); }
}

// second passes only (family HS_CT_SUB2)
template <class P>
static void launch_sub2(int sgn, int tw_mode, int stcol, const void *in, void *out, const void *tw, const hs_tile_geom &g, long long howmany,
                        hs_stream_t st)
{
    constexpr int C = HsTileCols<P>::C;
    if (g.ncols % C != 0 || !stcol) { fprintf(stderr, "Error: this sub-plan runs passes with strided stores over a multiple of %d columns only\n", C); exit(EXIT_FAILURE); }
    if (tw_mode == 0) { if (sgn == 1) launch_tiled<P, 1, C, 0, true>(in, out, tw, g, howmany, st); else launch_tiled<P, -1, C, 0, true>(in, out, tw, g, howmany, st); } // column pass of a 2-D transform
    else if (tw_mode == 1) { if (sgn == 1) launch_tiled<P, 1, C, 1, true>(in, out, tw, g, howmany, st); else launch_tiled<P, -1, C, 1, true>(in, out, tw, g, howmany, st); }
    else              { if (sgn == 1) launch_tiled<P, 1, C, 2, true>(in, out, tw, g, howmany, st); else launch_tiled<P, -1, C, 2, true>(in, out, tw, g, howmany, st); }
}

extern "C" void hs_cu_launch_tiled(int sub_id, int sgn, int tw_mode, int stcol, const void *d_in, void *d_out, const void *d_tw,
                                   const hs_tile_geom *g, long long howmany, void *stream)
{
    hs_stream_t st = (hs_stream_t)stream;
    switch (sub_id) {
    case CT_S128: launch_sub<S128>(sgn, tw_mode, stcol, d_in, d_out, d_tw, *g, howmany, st); break;
    case CT_S256: launch_sub<S256>(sgn, tw_mode, stcol, d_in, d_out, d_tw, *g, howmany, st); break;
    case CT_S512: launch_sub<S512>(sgn, tw_mode, stcol, d_in, d_out, d_tw, *g, howmany, st); break;
    case CT_S1024: launch_sub<S1024>(sgn, tw_mode, stcol, d_in, d_out, d_tw, *g, howmany, st); break;
    case CT_S160X: launch_sub<S160x>(sgn, tw_mode, stcol, d_in, d_out, d_tw, *g, howmany, st); break;
    case CT_S512B: launch_sub2<S512b>(sgn, tw_mode, stcol, d_in, d_out, d_tw, *g, howmany, st); break;
    default: fprintf(stderr, "Error: unknown sub-plan id %d\n", sub_id); exit(EXIT_FAILURE);
    }
}

#endif // HS_PART(2)

#if HS_PART(3)
// ---- long Bluestein (padded length 2^14..2^20): fused ends and the middle kernel
template <class P>
static void launch_sub_io(int sgn, int tw_mode, int stcol, int io_mode, int cs, const void *in, void *out, const void *tw,
                          const hs_tile_geom &g, long long howmany, hs_stream_t st)
{
    constexpr int C = HsTileCols<P>::C;
    if (g.ncols % C != 0 && io_mode != 9) { fprintf(stderr, "Error: tiled pass needs a multiple of %d columns (got %d)\n", C, g.ncols); exit(EXIT_FAILURE); }
    if (io_mode == 8 && tw_mode == 0 && !stcol) {
        if (sgn == 1) launch_tiled<P, 1, C, 0, false, 8, 1>(in, out, tw, g, howmany, st); else launch_tiled<P, -1, C, 0, false, 8, 1>(in, out, tw, g, howmany, st);
    } else if (io_mode == 10 && tw_mode == 0 && !stcol && sgn == 1) {
        launch_tiled<P, 1, C, 0, false, 10, 1>(in, out, tw, g, howmany, st);
    } else if (io_mode == 9 && tw_mode == 1 && stcol) {
        constexpr int C2 = HsTileColsMirror<P>::C;
        if (sgn == 1) launch_tiled<P, 1, C2, 1, true, 9, 1>(in, out, tw, g, howmany, st); else launch_tiled<P, -1, C2, 1, true, 9, 1>(in, out, tw, g, howmany, st);
    } else if (io_mode == 1 && sgn == 1 && tw_mode == 0 && !stcol) {
        if (cs == 1) launch_tiled<P, 1, C, 0, false, 1, 1>(in, out, tw, g, howmany, st);
        else launch_tiled<P, 1, C, 0, false, 1, -1>(in, out, tw, g, howmany, st);
    } else if (io_mode == 2 && sgn == -1 && tw_mode == 1 && stcol) {
        if (cs == 1) launch_tiled<P, -1, C, 1, true, 2, 1>(in, out, tw, g, howmany, st);
        else launch_tiled<P, -1, C, 1, true, 2, -1>(in, out, tw, g, howmany, st);
    } else if (io_mode == 3 && tw_mode == 0 && !stcol) {
        if (sgn == 1) launch_tiled<P, 1, C, 0, false, 3, 1>(in, out, tw, g, howmany, st);
        else launch_tiled<P, -1, C, 0, false, 3, 1>(in, out, tw, g, howmany, st);
    } else if ((io_mode == 4 || io_mode == 5) && tw_mode == 0 && !stcol) {
        if (io_mode == 4) { if (sgn == 1) launch_tiled<P, 1, C, 0, false, 4, 1>(in, out, tw, g, howmany, st); else launch_tiled<P, -1, C, 0, false, 4, 1>(in, out, tw, g, howmany, st); }
        else { if (sgn == 1) launch_tiled<P, 1, C, 0, false, 5, 1>(in, out, tw, g, howmany, st); else launch_tiled<P, -1, C, 0, false, 5, 1>(in, out, tw, g, howmany, st); }
    } else if (io_mode == 6 && tw_mode == 1 && stcol) {
        if (sgn == 1) launch_tiled<P, 1, C, 1, true, 6, 1>(in, out, tw, g, howmany, st); else launch_tiled<P, -1, C, 1, true, 6, 1>(in, out, tw, g, howmany, st);
    } else if (io_mode == 7 && stcol && (tw_mode == 0 || tw_mode == 2)) {
        // the exchange pass: wider tiles mean longer contiguous runs per peer store (C * 16 bytes per row); HSFFT_DIST_C picks
        // 8 / 16 / 32 columns for the short sub-plans (default: 16 where the CTA stays within 256 threads)
        constexpr int CW = (P::T <= 16) ? 16 : C;
        constexpr int CX = (P::T <= 16) ? 32 : C;
        static int want = -1;
        if (want < 0) { const char *e = getenv("HSFFT_DIST_C"); want = e ? atoi(e) : 0; }
        const int use = (want == 8 || want == C) ? C : (want == 32 ? CX : (want == 16 ? CW : CW));
        if (tw_mode == 2) {
            if (use == C || g.ncols % use != 0) { if (sgn == 1) launch_tiled<P, 1, C, 2, true, 7, 1>(in, out, tw, g, howmany, st); else launch_tiled<P, -1, C, 2, true, 7, 1>(in, out, tw, g, howmany, st); }
            else if (use == CW) { if (sgn == 1) launch_tiled<P, 1, CW, 2, true, 7, 1>(in, out, tw, g, howmany, st); else launch_tiled<P, -1, CW, 2, true, 7, 1>(in, out, tw, g, howmany, st); }
            else { if (sgn == 1) launch_tiled<P, 1, CX, 2, true, 7, 1>(in, out, tw, g, howmany, st); else launch_tiled<P, -1, CX, 2, true, 7, 1>(in, out, tw, g, howmany, st); }
        }
        else { if (sgn == 1) launch_tiled<P, 1, C, 0, true, 7, 1>(in, out, tw, g, howmany, st); else launch_tiled<P, -1, C, 0, true, 7, 1>(in, out, tw, g, howmany, st); }
    } else if (io_mode == 11 && stcol && (tw_mode == 0 || tw_mode == 2)) {
        // the same with bulk copies (see hs_tiled.cuh IO 11)
        constexpr int CW = (P::T <= 16) ? 16 : C;
        constexpr int CX = (P::T <= 16) ? 32 : C;
        static int want = -1;
        if (want < 0) { const char *e = getenv("HSFFT_DIST_C"); want = e ? atoi(e) : 0; }
        const int use = (want == 8 || want == C) ? C : (want == 32 ? CX : (want == 16 ? CW : CW));
        if (tw_mode == 2) {
            if (use == C || g.ncols % use != 0) { if (sgn == 1) launch_tiled<P, 1, C, 2, true, 11, 1>(in, out, tw, g, howmany, st); else launch_tiled<P, -1, C, 2, true, 11, 1>(in, out, tw, g, howmany, st); }
            else if (use == CW) { if (sgn == 1) launch_tiled<P, 1, CW, 2, true, 11, 1>(in, out, tw, g, howmany, st); else launch_tiled<P, -1, CW, 2, true, 11, 1>(in, out, tw, g, howmany, st); }
            else { if (sgn == 1) launch_tiled<P, 1, CX, 2, true, 11, 1>(in, out, tw, g, howmany, st); else launch_tiled<P, -1, CX, 2, true, 11, 1>(in, out, tw, g, howmany, st); }
        }
        else { if (sgn == 1) launch_tiled<P, 1, C, 0, true, 11, 1>(in, out, tw, g, howmany, st); else launch_tiled<P, -1, C, 0, true, 11, 1>(in, out, tw, g, howmany, st); }
    } else { fprintf(stderr, "Error: unsupported fused tiled pass\n"); exit(EXIT_FAILURE); }
}

// second passes only (family HS_CT_SUB2): the real four-step's mirrored pass and the convolution's slice
template <class P>
static void launch_sub_io2(int sgn, int tw_mode, int stcol, int io_mode, const void *in, void *out, const void *tw, const hs_tile_geom &g,
                           long long howmany, hs_stream_t st)
{
    constexpr int C = HsTileCols<P>::C;
    if (g.ncols % C != 0 && io_mode != 9) { fprintf(stderr, "Error: tiled pass needs a multiple of %d columns (got %d)\n", C, g.ncols); exit(EXIT_FAILURE); }
    if (io_mode == 9 && tw_mode == 1 && stcol) {
        constexpr int C2 = HsTileColsMirror<P>::C;
        if (sgn == 1) launch_tiled<P, 1, C2, 1, true, 9, 1>(in, out, tw, g, howmany, st); else launch_tiled<P, -1, C2, 1, true, 9, 1>(in, out, tw, g, howmany, st);
    } else if (io_mode == 6 && tw_mode == 1 && stcol) {
        if (sgn == 1) launch_tiled<P, 1, C, 1, true, 6, 1>(in, out, tw, g, howmany, st); else launch_tiled<P, -1, C, 1, true, 6, 1>(in, out, tw, g, howmany, st);
    } else { fprintf(stderr, "Error: unsupported fused tiled pass\n"); exit(EXIT_FAILURE); }
}

extern "C" void hs_cu_launch_tiled_io(int sub_id, int sgn, int tw_mode, int stcol, int io_mode, int chirp_sgn, const void *d_in, void *d_out,
                                      const void *d_tw, const hs_tile_geom *g, long long howmany, void *stream)
{
    hs_stream_t st = (hs_stream_t)stream;
    switch (sub_id) {
    case CT_S128: launch_sub_io<S128>(sgn, tw_mode, stcol, io_mode, chirp_sgn, d_in, d_out, d_tw, *g, howmany, st); break;
    case CT_S256: launch_sub_io<S256>(sgn, tw_mode, stcol, io_mode, chirp_sgn, d_in, d_out, d_tw, *g, howmany, st); break;
    case CT_S512: launch_sub_io<S512>(sgn, tw_mode, stcol, io_mode, chirp_sgn, d_in, d_out, d_tw, *g, howmany, st); break;
    case CT_S1024: launch_sub_io<S1024>(sgn, tw_mode, stcol, io_mode, chirp_sgn, d_in, d_out, d_tw, *g, howmany, st); break;
    case CT_S160X: launch_sub_io<S160x>(sgn, tw_mode, stcol, io_mode, chirp_sgn, d_in, d_out, d_tw, *g, howmany, st); break;
    case CT_S512B: launch_sub_io2<S512b>(sgn, tw_mode, stcol, io_mode, d_in, d_out, d_tw, *g, howmany, st); break;
    default: fprintf(stderr, "Error: unknown sub-plan id %d\n", sub_id); exit(EXIT_FAILURE);
    }
}

extern "C" void hs_cu_launch_blu_mid(int sub_id, int chirp_sgn, const void *d_in, void *d_out, const void *d_twf, const void *d_twi,
                                     const hs_tile_geom *g, const void *d_hk, long long howmany, void *stream)
{
    hs_stream_t st = (hs_stream_t)stream;
    switch (sub_id) {
    case CT_S128: if (chirp_sgn == 1) launch_blu_mid<S128, S128i, 1>(d_in, d_out, d_twf, d_twi, *g, d_hk, howmany, st); else launch_blu_mid<S128, S128i, -1>(d_in, d_out, d_twf, d_twi, *g, d_hk, howmany, st); break;
    case CT_S256: if (chirp_sgn == 1) launch_blu_mid<S256, S256i, 1>(d_in, d_out, d_twf, d_twi, *g, d_hk, howmany, st); else launch_blu_mid<S256, S256i, -1>(d_in, d_out, d_twf, d_twi, *g, d_hk, howmany, st); break;
    case CT_S512: if (chirp_sgn == 1) launch_blu_mid<S512, S512i, 1>(d_in, d_out, d_twf, d_twi, *g, d_hk, howmany, st); else launch_blu_mid<S512, S512i, -1>(d_in, d_out, d_twf, d_twi, *g, d_hk, howmany, st); break;
    case CT_S160X: if (chirp_sgn == 1) launch_blu_mid<S160x, S160xi, 1>(d_in, d_out, d_twf, d_twi, *g, d_hk, howmany, st); else launch_blu_mid<S160x, S160xi, -1>(d_in, d_out, d_twf, d_twi, *g, d_hk, howmany, st); break;
    case CT_S1024: if (chirp_sgn == 1) launch_blu_mid<S1024, S1024i, 1>(d_in, d_out, d_twf, d_twi, *g, d_hk, howmany, st); else launch_blu_mid<S1024, S1024i, -1>(d_in, d_out, d_twf, d_twi, *g, d_hk, howmany, st); break;
    default: fprintf(stderr, "Error: unknown sub-plan id %d\n", sub_id); exit(EXIT_FAILURE);
    }
}
#endif // HS_PART(3)
