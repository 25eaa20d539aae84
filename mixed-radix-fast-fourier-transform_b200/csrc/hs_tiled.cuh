// hs_tiled.cuh -- one pass of a transform too long for one SM: C adjacent strided sub-transforms per CTA.
//
// A transform of length N = R1*R2 whose states do not fit shared memory is run as two of the
// reference's "super stages" (the recursion of reference
// src/highSpeedFFT.c:318 unrolled two levels at a time):
//   pass 1:  for every q < R2:   A[q][j] = FFT_R1{ x[q + R2 i] }            in stride R2, out contiguous
//   pass 2:  for every k < R1:   X[k + R1 j] = FFT_R2{ W_N^{i k} A[i][k] }  in stride R1, out stride R1
// Each sub-transform is a compile-time plan of hs_smem_fft.cuh run by a team of threads.  A CTA
// owns C adjacent columns (q or k), so every global access is a run of C*16 contiguous bytes:
//   * first pass of the plan: teams across lanes (lane = column) -> the strided loads coalesce;
//   * later passes: pass 1 switches to threads-of-a-team across lanes so that its contiguous output
//     rows coalesce; pass 2 keeps lane = column because its stores are strided like its loads.
// The inter-pass twiddle W_N^{i k} is applied in registers right after the load.  Element
// i = b + NB*ii of column k needs W_N^{k b} * (W_N^{k NB})^ii: one table load per butterfly plus one
// per thread, the powers are formed by multiplication (the L1 data pipe, not FP64, limits these
// kernels -- see DESIGN.md section 3).
#pragma once
#include "hs_smem_fft.cuh"

struct HsTileGeom {
    long long in_xf, out_xf; // elements between consecutive transforms
    long long in_c, in_e;    // input: column stride, element stride
    long long out_c, out_e;  // output: column stride, element stride
    int ncols;               // columns per transform (a multiple of C)
    const cplx *twb;         // [b][k] = W_N^{k b}, b < NB of the plan's first stage, k < ncols; null = no twiddle
    const cplx *tws;         // [k]    = W_N^{k NB}
};

struct HsLoadCol {
    static constexpr bool kPost = false;
    const cplx *HS_RESTRICT base;
    long long es;
    HS_DEVICE cplx operator()(int, int, int n) const { return hs_ld_stream(base + n * es); }
};
struct HsLoadColTw {
    static constexpr bool kPost = true;
    const cplx *HS_RESTRICT base;
    long long es;
    const cplx *HS_RESTRICT twb; // already offset by the column
    int ncols;
    cplx step;
    HS_DEVICE cplx operator()(int, int, int n) const { return hs_ld_stream(base + n * es); }
    template <int R>
    HS_DEVICE void post(int b, cplx *v) const
    {
        const cplx w0 = hs_ld_tab(twb + (long long)b * ncols);
        cplx s[R];
        s[1] = step;
        HS_UNROLL
        for (int i = 2; i < R; i *= 2) s[i] = hs_mul(s[i / 2], s[i / 2]);
        HS_UNROLL
        for (int i = 3; i < R; ++i)
            if (i & (i - 1)) {
                int hb = 1;
                while (hb * 2 <= i) hb *= 2;
                s[i] = hs_mul(s[hb], s[i - hb]);
            }
        v[0] = hs_mul(v[0], w0);
        HS_UNROLL
        for (int i = 1; i < R; ++i) v[i] = hs_mul(v[i], hs_mul(w0, s[i]));
    }
};
struct HsStoreCol {
    cplx *HS_RESTRICT base;
    long long es;
    HS_DEVICE void operator()(int, int, int k, cplx v) const { hs_st_stream(base + k * es, v); }
};

// P: plan of the sub-transform (P::X is ignored); C: columns per CTA; TW: apply the inter-pass twiddle
// on load; STCOL: stores keep lane = column (pass 2) instead of switching to lane = thread (pass 1).
template <class P, int SGN, int C, bool TW, bool STCOL, int MINB>
HS_GLOBAL void HS_LAUNCH_BOUNDS(P::T *C, MINB)
hs_tiled_kernel(const cplx *HS_RESTRICT in, cplx *HS_RESTRICT out, const cplx *HS_RESTRICT tw, const HsTileGeom g, long long howmany)
{
    constexpr int NPADC = P::NPAD | 1; // odd: columns side by side in lanes hit different banks
    HS_DYN_SMEM(cplx, smem);
    const int tiles = g.ncols / C;
    const long long xf = HS_BID / tiles;
    const int c0 = (int)(HS_BID % tiles) * C;
    const int colA = HS_TID % C, tA = HS_TID / C; // lane = column
    const int colB = HS_TID / P::T, tB = HS_TID % P::T; // lane = thread of a team
    const HsLane first{smem + colA * NPADC, tA, true};
    const HsLane rest = STCOL ? first : HsLane{smem + colB * NPADC, tB, true};
    const int colS = STCOL ? colA : colB;
    HsStoreCol st{out + xf * g.out_xf + (long long)(c0 + colS) * g.out_c, g.out_e};
    if constexpr (TW) {
        HsLoadColTw ld{in + xf * g.in_xf + (long long)(c0 + colA) * g.in_c, g.in_e, g.twb + (c0 + colA), g.ncols,
                       hs_ld_tab(g.tws + (c0 + colA))};
        hs_fft_team2<P, SGN>(ld, st, first, rest, tw);
    } else {
        HsLoadCol ld{in + xf * g.in_xf + (long long)(c0 + colA) * g.in_c, g.in_e};
        hs_fft_team2<P, SGN>(ld, st, first, rest, tw);
    }
}
