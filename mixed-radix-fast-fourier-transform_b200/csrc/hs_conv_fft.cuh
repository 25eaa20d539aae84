// hs_conv_fft.cuh -- in-place shared-memory transforms for the cyclic convolutions inside Bluestein's algorithm.
//
// The reference computes y = IFFT_M(FFT_M(a) * Hk) (bluestein_fft, src/highSpeedFFT.c:1797-1866) with its general
// out-of-place recursion.  A convolution does not care in which order the spectrum is stored, so here the forward
// transform is a decimation-in-frequency pass structure that leaves the spectrum in digit-reversed order, the table Hk is
// stored in that same order when the plan is built, and the inverse transform is the mirror decimation-in-time
// structure that takes digit-reversed input back to natural order.  Both work IN PLACE: a butterfly reads and writes
// the same r positions, so a thread can run its butterflies one after the other with only r points in registers (the
// auto-sort schedule of hs_smem_fft.cuh must hold all of a thread's points of a pass at once -- 16 to 20 points at the
// 640 threads a 10240-point transform wants, which does not fit 96 registers).  One CTA barrier per stage.
//
// Stage s of a length-n plan with radices r_0 .. r_{S-1} works on blocks of length L_s = n / (r_0 .. r_{s-1}), m = L/r:
//     forward  (DIF):  b_j = W_L^(p j) * sum_i a_i w_r^(i j),      a_i = X[base + p + m i]   ->   X[base + p + m j] = b_j
//     inverse  (DIT):  a_i = sum_j (conj W_L^(p j)) b_j conj(w_r)^(i j)                       (stages in reverse order)
// After the forward stages position  sum_s j_s m_s  holds bin  j_0 + r_0 (j_1 + r_1 (j_2 + ...)).  The last forward
// stage and the first inverse stage touch the same r positions from the same thread (m = 1, no twiddles), so the
// spectrum is multiplied by Hk and turned around in registers.
#pragma once
#include "hs_smem_fft.cuh"

// N_: length; T_: threads of the CTA (one transform per CTA); LOGPAD_: one pad slot per 2^LOGPAD_ points;
// TWM_: how a stage gets its r-1 twiddles from the [j-1][p] table -- 0 load every row, -1 rows 1,2,4,8 and products,
// -2 row 1 and a product tree, -3 row 1 and successive powers (fewest live registers); RS...: DIF stage radices in order.
template <int N_, int T_, int LOGPAD_, int TWM_, int... RS>
struct HsConvPlan {
    using radices = HsRadixList<RS...>;
    static constexpr int N = N_, T = T_, LOGPAD = LOGPAD_, TWM = TWM_, X = 1;
    static constexpr int NST = radices::count;
    static constexpr int radix(int s) { return radices::at(s); }
    static constexpr int L(int s) { return s == 0 ? N_ : L(s - 1) / radix(s - 1); } // block length stage s works on
    static constexpr int m(int s) { return L(s) / radix(s); }
    // twiddle table: stages 0 .. NST-2 (the last one has m = 1: no twiddles), [j-1][p] per stage
    static constexpr int tw_off(int s) { return s == 0 ? 0 : tw_off(s - 1) + (radix(s - 1) - 1) * m(s - 1); }
    static constexpr int tw_count = tw_off(NST - 1);
    static constexpr int NPAD = LOGPAD_ ? N_ + (N_ >> LOGPAD_) + 1 : N_;
    static constexpr int pad(int p) { return LOGPAD_ ? p + (p >> LOGPAD_) : p; }
    static constexpr int smem_bytes = NPAD * 16;
    static constexpr int threads = T_;
    static constexpr bool EXACT = true; // Bluestein-internal: full-precision radix-3/5/7 constants
    static constexpr int SPLIT = TWM_;  // descriptor field (fill_desc)
    static_assert(L(NST - 1) == radix(NST - 1), "radices must multiply to N");
};

// w[1..R) <- W^(p j), j = 1..R-1, from the stage table rows row1 + (j-1)*M
template <int R, int TWM>
HS_DEVICE void hs_conv_twiddles(const cplx *HS_RESTRICT row1, const int M, cplx *wv)
{
    if (TWM == 0) {
        HS_UNROLL
        for (int j = 1; j < R; ++j) wv[j] = hs_ld_tab(row1 + (j - 1) * M);
    } else {
        if (TWM == -1) {
            HS_UNROLL
            for (int j = 1; j < R; j *= 2) wv[j] = hs_ld_tab(row1 + (j - 1) * M);
        } else {
            wv[1] = hs_ld_tab(row1);
            HS_UNROLL
            for (int j = 2; j < R; j *= 2) wv[j] = hs_mul(wv[j / 2], wv[j / 2]);
        }
        HS_UNROLL
        for (int j = 3; j < R; ++j)
            if (j & (j - 1)) {
                int hb = 1;
                while (hb * 2 <= j) hb *= 2;
                wv[j] = hs_mul(wv[hb], wv[j - hb]);
            }
    }
}
// v[j] *= W^(p j) (CONJ: conj W^(p j)), j = 1..R-1
template <int R, int TWM, bool CONJ>
HS_DEVICE void hs_conv_twiddle_apply(const cplx *HS_RESTRICT row1, const int M, cplx *v)
{
    if (TWM == -3) { // successive powers, applied at once: two live twiddles
        const cplx w1 = hs_ld_tab(row1);
        cplx w = w1;
        if (R == 16) {
            HS_UNROLL
            for (int j = 1; j < 8; ++j) {
                v[j] = CONJ ? hs_mulc(v[j], w) : hs_mul(v[j], w);
                v[8 + j] = CONJ ? hs_mulc(v[8 + j], w) : hs_mul(v[8 + j], w);
                w = hs_mul(w, w1);
            }
            HS_UNROLL
            for (int j = 8; j < 16; ++j) v[j] = CONJ ? hs_mulc(v[j], w) : hs_mul(v[j], w);
        } else {
            HS_UNROLL
            for (int j = 1; j < R; ++j) {
                v[j] = CONJ ? hs_mulc(v[j], w) : hs_mul(v[j], w);
                if (j + 1 < R) w = hs_mul(w, w1);
            }
        }
    } else {
        cplx wv[R];
        hs_conv_twiddles<R, TWM>(row1, M, wv);
        HS_UNROLL
        for (int j = 1; j < R; ++j) v[j] = CONJ ? hs_mulc(v[j], wv[j]) : hs_mul(v[j], wv[j]);
    }
}

// Forward stage S (not the last).  FIRST: the inputs come from ld(n) (global memory), else from shared memory.
template <class P, int S, bool FIRST, class LD>
HS_DEVICE void hs_dif_stage(LD &ld, cplx *sm, const cplx *HS_RESTRICT tw, const int t)
{
    constexpr int R = P::radix(S), LB = P::L(S), M = P::m(S), NB = P::N / R;
    constexpr int B = (NB + P::T - 1) / P::T;
#pragma unroll 1
    for (int u = 0; u < B; ++u) {
        const int beta = t + P::T * u;
        if (NB % P::T != 0 && beta >= NB) break;
        const int p = beta % M, base = (beta / M) * LB + p;
        cplx v[R];
        if constexpr (FIRST) { // all signal loads first (their latency overlaps), then the table loads one by one
            HS_UNROLL
            for (int i = 0; i < R; ++i) v[i] = ld.raw(base + M * i);
            HS_UNROLL
            for (int i = 0; i < R; ++i) v[i] = ld.fin(base + M * i, v[i]);
        } else {
            HS_UNROLL
            for (int i = 0; i < R; ++i) v[i] = sm[P::pad(base + M * i)];
        }
        hs_bfly<R, 1, P::EXACT>(v);
        hs_conv_twiddle_apply<R, P::TWM, false>(tw + P::tw_off(S) + p, M, v);
        HS_UNROLL
        for (int j = 0; j < R; ++j) sm[P::pad(base + M * j)] = v[j];
    }
}

// Inverse stage S (not the last one of the forward order).  LAST (S == 0): the outputs go to st(n, value).
template <class P, int S, bool LAST, class ST>
HS_DEVICE void hs_dit_stage(ST &st, cplx *sm, const cplx *HS_RESTRICT tw, const int t)
{
    constexpr int R = P::radix(S), LB = P::L(S), M = P::m(S), NB = P::N / R;
    constexpr int B = (NB + P::T - 1) / P::T;
#pragma unroll 1
    for (int u = 0; u < B; ++u) {
        const int beta = t + P::T * u;
        if (NB % P::T != 0 && beta >= NB) break;
        const int p = beta % M, base = (beta / M) * LB + p;
        cplx v[R];
        HS_UNROLL
        for (int j = 0; j < R; ++j) v[j] = sm[P::pad(base + M * j)];
        hs_conv_twiddle_apply<R, P::TWM, true>(tw + P::tw_off(S) + p, M, v);
        hs_bfly<R, -1, P::EXACT>(v);
        HS_UNROLL
        for (int i = 0; i < R; ++i) {
            if (LAST) st(base + M * i, v[i]);
            else sm[P::pad(base + M * i)] = v[i];
        }
    }
}

// The turn-around: last forward stage, * Hk (CONJ: * conj Hk), first inverse stage -- same positions, same thread.
// hkt: Hk in the layout [i][beta] (entry of position R beta + i at i*NB + beta: coalesced across the threads).
template <class P, bool CONJ>
HS_DEVICE void hs_conv_middle(cplx *sm, const cplx *HS_RESTRICT hkt, const int t)
{
    constexpr int S = P::NST - 1, R = P::radix(S), NB = P::N / R;
    constexpr int B = (NB + P::T - 1) / P::T;
    static_assert(P::m(S) == 1, "the last stage works on adjacent points");
#pragma unroll 1
    for (int u = 0; u < B; ++u) {
        const int beta = t + P::T * u;
        if (NB % P::T != 0 && beta >= NB) break;
        cplx v[R];
        HS_UNROLL
        for (int i = 0; i < R; ++i) v[i] = sm[P::pad(R * beta + i)];
        hs_bfly<R, 1, P::EXACT>(v);
        HS_UNROLL
        for (int i = 0; i < R; ++i) {
            const cplx hv = hs_ld_tab(hkt + i * NB + beta);
            v[i] = CONJ ? hs_mulc(v[i], hv) : hs_mul(v[i], hv);
        }
        hs_bfly<R, -1, P::EXACT>(v);
        HS_UNROLL
        for (int i = 0; i < R; ++i) sm[P::pad(R * beta + i)] = v[i];
    }
}

template <class P, int S, class LD>
HS_DEVICE void hs_conv_forward_from(LD &ld, cplx *sm, const cplx *tw, const int t)
{
    if constexpr (S < P::NST - 1) {
        hs_dif_stage<P, S, S == 0>(ld, sm, tw, t);
        HS_SYNC();
        hs_conv_forward_from<P, S + 1>(ld, sm, tw, t);
    }
}
template <class P, int S, class ST>
HS_DEVICE void hs_conv_inverse_from(ST &st, cplx *sm, const cplx *tw, const int t)
{
    if constexpr (S >= 0) {
        HS_SYNC();
        hs_dit_stage<P, S, S == 0>(st, sm, tw, t);
        hs_conv_inverse_from<P, S - 1>(st, sm, tw, t);
    }
}

// One cyclic convolution of length P::N:  st(n, y[n]) with y = IFFT(FFT(ld(.)) * Hk), unnormalised (Hk carries 1/N).
// Inputs and outputs are addressed in natural order; a thread loads and stores the same positions p + (N/r_0) i, so
// back-to-back convolutions need no barrier in between.
template <class P, bool CONJ, class LD, class ST>
HS_DEVICE void hs_conv_run(LD &ld, ST &st, cplx *sm, const cplx *tw, const cplx *hkt, const int t)
{
    static_assert(P::NST >= 2, "at least two stages");
    hs_conv_forward_from<P, 0>(ld, sm, tw, t);
    hs_conv_middle<P, CONJ>(sm, hkt, t);
    hs_conv_inverse_from<P, P::NST - 2>(st, sm, tw, t);
}

// ---------------------------------------------------------------------------------------------------------------
// Split chirp-z on top of it (see hs_fused.cuh for the algebra): two convolutions of length M' = P::N >= N per
// transform, one CTA per transform, HBM sees x once and the result once.
template <int SGN>
struct HsConvLoadChirp {
    const cplx *HS_RESTRICT x;
    const cplx *HS_RESTRICT h;
    int N;
    HS_DEVICE cplx raw(int n) const { return n < N ? hs_ld_stream(x + n) : hs_c(0.0, 0.0); }
    HS_DEVICE cplx fin(int n, cplx xv) const
    {
        if (n >= N) return xv;
        const cplx hv = hs_ld_tab(h + n);
        return (SGN == 1) ? hs_mulc(xv, hv) : hs_mul(xv, hv);
    }
};
struct HsConvStorePark {
    cplx *HS_RESTRICT out;
    int N;
    HS_DEVICE void operator()(int n, cplx v) const
    {
        if (n < N) hs_st_keep(out + n, v);
    }
};
template <int SGN>
struct HsConvStoreCombine {
    cplx *HS_RESTRICT out;
    const cplx *HS_RESTRICT c0, *HS_RESTRICT c1;
    int N;
    HS_DEVICE void operator()(int n, cplx v) const
    {
        if (n < N) {
            const cplx y0 = hs_ld_coherent(out + n), h0 = hs_ld_tab(c0 + n), h1 = hs_ld_tab(c1 + n);
            const cplx a = (SGN == 1) ? hs_mulc(y0, h0) : hs_mul(y0, h0), b = (SGN == 1) ? hs_mulc(v, h1) : hs_mul(v, h1);
            hs_st_stream(out + n, hs_add(a, b));
        }
    }
};
template <int SGN>
struct HsConvStoreChirpCrop {
    cplx *HS_RESTRICT out;
    const cplx *HS_RESTRICT h;
    int N;
    HS_DEVICE void operator()(int n, cplx v) const
    {
        if (n < N) {
            const cplx hv = hs_ld_tab(h + n);
            hs_st_stream(out + n, (SGN == 1) ? hs_mulc(v, hv) : hs_mul(v, hv));
        }
    }
};

// chirp3 = h | h conj(w) | h w (N entries each), hkt2 = Hk[even] | Hk[odd], each P::N entries in hs_conv_middle's layout
template <class P, int SGN, int MINB>
HS_GLOBAL void HS_LAUNCH_BOUNDS(P::threads, MINB)
hs_conv_bluestein_split_kernel(const cplx *HS_RESTRICT in, cplx *HS_RESTRICT out, const cplx *HS_RESTRICT tw, const cplx *HS_RESTRICT chirp3,
                               const cplx *HS_RESTRICT hkt2, int N, long long howmany, int pf_dist)
{
    HS_DYN_SMEM(cplx, sm);
    const int t = HS_TID;
    const long long xf = HS_BID;
    if (pf_dist > 0 && t == 0 && xf + pf_dist < howmany) hs_prefetch_l2(in + (xf + pf_dist) * N, (unsigned)(N * 16));
    const cplx *x = in + xf * N;
    cplx *o = out + xf * N;
    const cplx *h = chirp3, *gA = chirp3 + N, *gB = chirp3 + 2 * (long long)N;
    {
        HsConvLoadChirp<SGN> ld{x, h, N};
        HsConvStorePark st{o, N};
        hs_conv_run<P, SGN != 1>(ld, st, sm, tw, hkt2, t);
    }
    {
        HsConvLoadChirp<SGN> ld{x, SGN == 1 ? gA : gB, N};
        HsConvStoreCombine<SGN> st{o, h, SGN == 1 ? gB : gA, N};
        hs_conv_run<P, SGN != 1>(ld, st, sm, tw, hkt2 + P::N, t);
    }
}

// Plain chirp-z (one convolution of length M = P::N >= 2N-1) with the same in-place transforms.
template <class P, int SGN, int MINB>
HS_GLOBAL void HS_LAUNCH_BOUNDS(P::threads, MINB)
hs_conv_bluestein_kernel(const cplx *HS_RESTRICT in, cplx *HS_RESTRICT out, const cplx *HS_RESTRICT tw, const cplx *HS_RESTRICT chirp,
                         const cplx *HS_RESTRICT hkt, int N, long long howmany, int pf_dist)
{
    HS_DYN_SMEM(cplx, sm);
    const int t = HS_TID;
    const long long xf = HS_BID;
    if (pf_dist > 0 && t == 0 && xf + pf_dist < howmany) hs_prefetch_l2(in + (xf + pf_dist) * N, (unsigned)(N * 16));
    HsConvLoadChirp<SGN> ld{in + xf * N, chirp, N};
    HsConvStoreChirpCrop<SGN> st{out + xf * N, chirp, N};
    hs_conv_run<P, SGN != 1>(ld, st, sm, tw, hkt, t);
}
