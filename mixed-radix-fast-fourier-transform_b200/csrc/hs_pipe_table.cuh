// hs_pipe_table.cuh -- which persistent pipelines (hs_pipe.cuh) exist and how they are dispatched.
// Included by hs_pipe_inst.cu (one translation unit per two-pass size, -DHS_PIPE_ID=<CT_Txx>) and by
// tests/emu/hs_device_emu.cpp; the including file provides
//     template <class PH0, class PH1, class PH2> void launch_pipe(hs_pipe_args &a, hs_stream_t st);
#pragma once
#include <cstdio>
#include <cstdlib>
#include "hs_abi.h"
#include "hs_pipe.cuh"
#include "hs_plans.cuh"

// per two-pass size: the sub-plans, their inverse-order partners (Bluestein middle) and the CTA size.
// V selects the CTA shape: 0 = 256 compute threads (512 for 2^18), 1 = half of that (smaller tiles, more CTAs per SM);
// chosen at run time by HSFFT_PIPE_SHAPE.
template <int ID> struct HsPipeSubs;
template <> struct HsPipeSubs<CT_T14> { using P1 = S128;  using P2 = S128;  using P2i = S128i;  static constexpr int NT0 = 256; };
template <> struct HsPipeSubs<CT_T15> { using P1 = S128;  using P2 = S256;  using P2i = S256i;  static constexpr int NT0 = 256; };
template <> struct HsPipeSubs<CT_T16> { using P1 = S256;  using P2 = S256;  using P2i = S256i;  static constexpr int NT0 = 256; };
template <> struct HsPipeSubs<CT_T17> { using P1 = S256;  using P2 = S512;  using P2i = S512i;  static constexpr int NT0 = 256; };
template <> struct HsPipeSubs<CT_T18> { using P1 = S512;  using P2 = S512;  using P2i = S512i;  static constexpr int NT0 = 512; };
template <> struct HsPipeSubs<CT_T19> { using P1 = S512;  using P2 = S1024; using P2i = S1024i; static constexpr int NT0 = 256; };
template <> struct HsPipeSubs<CT_T20> { using P1 = S1024; using P2 = S1024; using P2i = S1024i; static constexpr int NT0 = 256; };
template <int ID, int V>
struct HsPipeCfg : HsPipeSubs<ID> {
    static constexpr int NT = HsPipeSubs<ID>::NT0 >> V; // compute threads (the producer warp comes on top)
};

// residency classes: 2 = streamed from / to HBM once (L2 evict_first), 1 = ring slot kept in L2 (evict_last, coherent loads)
template <int ID, int V, int SGN>
struct HsPipeC2C {
    using K = HsPipeCfg<ID, V>;
    using A = HsPhTiled<typename K::P1, SGN, K::NT, 0, false, 0, 1, 2, 1>;
    using B = HsPhTiled<typename K::P2, SGN, K::NT, 1, true, 0, 1, 1, 2>;
};
template <int ID, int V, int SGN>
struct HsPipeC2R { // the Hermitian merge (reference src/real.c:171-179) rides on pass 1's loads
    using K = HsPipeCfg<ID, V>;
    using A = HsPhTiled<typename K::P1, SGN, K::NT, 0, false, 3, 1, 2, 1>;
    using B = HsPhTiled<typename K::P2, SGN, K::NT, 1, true, 0, 1, 1, 2>;
};
template <int ID, int V, int SGN>
struct HsPipeR2C { // half-length transform, then the split (reference src/real.c:109-132) as a third phase
    using K = HsPipeCfg<ID, V>;
    using A = HsPhTiled<typename K::P1, SGN, K::NT, 0, false, 0, 1, 2, 1>;
    using B = HsPhTiled<typename K::P2, SGN, K::NT, 1, true, 0, 1, 1, 1>;
    using C = HsPhR2CSplit<K::NT, 16, 1, 2>;
};
template <int ID, int V, int CS>
struct HsPipeBlu { // chirp+pad+forward pass 1 | forward pass 2, *Hk, inverse pass 1 | inverse pass 2, chirp, crop
    using K = HsPipeCfg<ID, V>;
    using A = HsPhTiled<typename K::P1, 1, K::NT, 0, false, 1, CS, 2, 1>;
    using B = HsPhBluMid<typename K::P2, typename K::P2i, CS, K::NT, 1, 1>;
    using C = HsPhTiled<typename K::P1, -1, K::NT, 1, true, 2, CS, 1, 2>;
};

enum { HS_PIPE_C2C = 0, HS_PIPE_C2R = 1, HS_PIPE_R2C = 2, HS_PIPE_BLU = 3 };

// tiles per transform of every phase, from the geometry the host filled in
template <class PH0, class PH1, class PH2>
static void pipe_fill_tiles(hs_pipe_args &a, int r2c_kt)
{
    a.ph[0].tiles = a.ph[0].g.ncols / PH0::C;
    a.ph[1].tiles = a.ph[1].g.ncols / PH1::C;
    if (a.nphase == 3) a.ph[2].tiles = r2c_kt ? (a.ph[2].g.nlog / 2 + 1 + r2c_kt - 1) / r2c_kt : a.ph[2].g.ncols / PH2::C;
    if (a.G < 1) a.G = 1;
    a.ngroups = (a.howmany + a.G - 1) / a.G;
    a.round_items = 0;
    for (int j = 0; j < HS_PIPE_MAX_PHASE; ++j) {
        a.items[j] = j < a.nphase ? (unsigned)a.G * (unsigned)a.ph[j].tiles : 0u;
        a.round_items += a.items[j];
    }
    a.total = (unsigned long long)(a.ngroups + (long long)(a.nphase - 1) * a.lag) * a.round_items;
}

template <int ID, int V>
static int pipe_dispatch_v(int variant, int sgn, hs_pipe_args &a, hs_stream_t st, bool query)
{
    using K = HsPipeCfg<ID, V>;
    switch (variant) {
    case HS_PIPE_C2C:
        if (sgn == 1) { using W = HsPipeC2C<ID, V, 1>; pipe_fill_tiles<typename W::A, typename W::B, HsPhNone>(a, 0); if (!query) launch_pipe<typename W::A, typename W::B, HsPhNone>(a, st); }
        else { using W = HsPipeC2C<ID, V, -1>; pipe_fill_tiles<typename W::A, typename W::B, HsPhNone>(a, 0); if (!query) launch_pipe<typename W::A, typename W::B, HsPhNone>(a, st); }
        return 1;
    case HS_PIPE_C2R:
        if (sgn != -1) return 0;
        { using W = HsPipeC2R<ID, V, -1>; pipe_fill_tiles<typename W::A, typename W::B, HsPhNone>(a, 0); if (!query) launch_pipe<typename W::A, typename W::B, HsPhNone>(a, st); }
        return 1;
    case HS_PIPE_R2C:
        if (sgn != 1) return 0;
        { using W = HsPipeR2C<ID, V, 1>; pipe_fill_tiles<typename W::A, typename W::B, typename W::C>(a, W::C::KT); if (!query) launch_pipe<typename W::A, typename W::B, typename W::C>(a, st); }
        return 1;
    case HS_PIPE_BLU: // sgn = the chirp direction (the plan's sgn)
        if (sgn == 1) { using W = HsPipeBlu<ID, V, 1>; pipe_fill_tiles<typename W::A, typename W::B, typename W::C>(a, 0); if (!query) launch_pipe<typename W::A, typename W::B, typename W::C>(a, st); }
        else { using W = HsPipeBlu<ID, V, -1>; pipe_fill_tiles<typename W::A, typename W::B, typename W::C>(a, 0); if (!query) launch_pipe<typename W::A, typename W::B, typename W::C>(a, st); }
        return 1;
    }
    return 0;
}

template <int ID>
static int pipe_dispatch(int variant, int sgn, hs_pipe_args &a, hs_stream_t st, bool query)
{
    static int shape = -1;
    if (shape < 0) { const char *e = getenv("HSFFT_PIPE_SHAPE"); shape = e ? atoi(e) : 0; }
#ifdef HS_EMU
    return pipe_dispatch_v<ID, 0>(variant, sgn, a, st, query);
#else
    return shape == 1 ? pipe_dispatch_v<ID, 1>(variant, sgn, a, st, query) : pipe_dispatch_v<ID, 0>(variant, sgn, a, st, query);
#endif
}
