// hs_smem_fft.cuh -- one transform per thread team, held in registers and shared memory.
//
// Replaces the reference's recursive hot loop mixed_radix_dit_rec (reference
// src/highSpeedFFT.c:318-1629) for lengths that fit one SM's shared memory.
//
// Formulation.  The reference recursion with factor list f[0..lf) computes, bottom up, the states
//     A_m[q][k] = sum_{p<m} x[q + (N/m) p] * exp(-sgn 2 pi i p k / m),      q < N/m, k < m,
// starting at m = 1 (A_1[q][0] = x[q]) and ending at m = N (A_N[0][k] = X[k]).  One stage of
// radix r takes m -> m' = r m:
//     A_m'[q'][k + m j] = sum_{i<r} W_r^{i j} * ( W_m'^{i k} * A_m[q' + (N/m') i][k] ).
// The stages are run in the reference's own order (last factor first), with the reference's
// butterflies and the reference's stage twiddles W_m'^{ik} (same cos/sin expression, built on the
// host), so the result agrees with the reference to rounding even where its 11-digit radix-3/5/7
// constants dominate the error.  What differs is the schedule: the recursion becomes a Stockham
// auto-sort pass structure --
//   * state A_m lives at position  q + (N/m) k  ("q fastest"): the r outputs of butterfly
//     b = q' + (N/m') k go to  b + (N/r) j, i.e. every store is contiguous across threads;
//   * the first stage reads x straight from HBM (q' + (N/r) i: coalesced 16-byte loads), the last
//     stage writes X straight to HBM (k + (N/r) j: coalesced 16-byte stores); only the states in
//     between touch shared memory, one 16-byte LDS/STS per point per exchange;
//   * a thread keeps all its butterflies of a stage in registers, so one buffer suffices.
// Bank conflicts of the strided reads of the late stages are removed by padding one 16-byte slot
// every 2^LOGPAD points (the plans below were checked with tests/emu/bank_check.py).
#pragma once
#include "hs_butterflies.cuh"

// ------------------------------------------------------------------ compile-time plan descriptor
template <int... RS>
struct HsRadixList;
template <int R0, int... REST>
struct HsRadixList<R0, REST...> {
    static constexpr int count = 1 + sizeof...(REST);
    static constexpr int at(int s) { return s == 0 ? R0 : HsRadixList<REST...>::at(s - 1); }
};
template <>
struct HsRadixList<> {
    static constexpr int count = 0;
    static constexpr int at(int) { return 1; }
};

// N_: transform length; T_: threads cooperating on one transform; X_: transforms per CTA;
// LOGPAD_: one pad slot per 2^LOGPAD_ points (0 = none); RS...: stage radices in EXECUTION order
// (= the reference's factor list reversed, runs of power-of-two factors possibly regrouped).
template <int N_, int T_, int X_, int LOGPAD_, int... RS>
struct HsPlan {
    using radices = HsRadixList<RS...>;
    static constexpr int N = N_, T = T_, X = X_, LOGPAD = LOGPAD_;
    static constexpr int NST = radices::count;
    static constexpr int radix(int s) { return radices::at(s); }
    static constexpr int m_before(int s) { return s == 0 ? 1 : m_before(s - 1) * radix(s - 1); }
    // offset of stage s in the device twiddle table ([i-1][k] layout per stage), stage 0 has none
    static constexpr int tw_off(int s) { return s <= 1 ? 0 : tw_off(s - 1) + (radix(s - 1) - 1) * m_before(s - 1); }
    static constexpr int tw_count = tw_off(NST);
    static constexpr int NPAD = LOGPAD_ ? N_ + (N_ >> LOGPAD_) + 1 : N_;
    static constexpr int pad(int p) { return LOGPAD_ ? p + (p >> LOGPAD_) : p; }
    static constexpr int smem_bytes = NST > 1 ? NPAD * X_ * 16 : 0;
    static constexpr int threads = T_ * X_;
};

// ------------------------------------------------------------------------------- one stage
// LD(u, i, n)       -> cplx   : element n of the input signal   (first stage only)
// ST(u, j, k, v)              : element k of the output spectrum (last stage only)
// u (butterfly slot of this thread) and i/j are compile-time constants after unrolling, so
// register-to-register hand-over between two transforms (fused Bluestein) stays in registers.
template <class P, int SGN, int S, class LD, class ST>
HS_DEVICE void hs_stage(LD &ld, ST &st, cplx *sm, const cplx *HS_RESTRICT tw, const int t, const bool active)
{
    constexpr int R = P::radix(S);
    constexpr int M = P::m_before(S);   // sub-transform length already done
    constexpr int QP = P::N / (M * R);  // q' range after this stage
    constexpr int NB = P::N / R;        // butterflies in this stage
    constexpr int B = (NB + P::T - 1) / P::T;
    constexpr bool FIRST = (S == 0), LAST = (S == P::NST - 1);
    cplx v[B][R];

    HS_UNROLL
    for (int u = 0; u < B; ++u) {
        const int b = t + P::T * u;
        if (active && (NB % P::T == 0 || b < NB)) {
            const int q = b % QP, k = b / QP;
            HS_UNROLL
            for (int i = 0; i < R; ++i) {
                if (FIRST) v[u][i] = ld(u, i, b + NB * i);
                else       v[u][i] = sm[P::pad(q + QP * i + (QP * R) * k)];
            }
            if (!FIRST) {
                const cplx *w = tw + P::tw_off(S) + k;
                HS_UNROLL
                for (int i = 1; i < R; ++i) v[u][i] = hs_mul(v[u][i], hs_ld_tab(w + (i - 1) * M));
            }
            hs_bfly<R, SGN>(v[u]);
        }
    }
    if (!FIRST && !LAST) HS_SYNC(); // every thread has read its inputs; the buffer may be overwritten
    HS_UNROLL
    for (int u = 0; u < B; ++u) {
        const int b = t + P::T * u;
        if (active && (NB % P::T == 0 || b < NB)) {
            HS_UNROLL
            for (int j = 0; j < R; ++j) {
                if (LAST) st(u, j, b + NB * j, v[u][j]);
                else      sm[P::pad(b + NB * j)] = v[u][j];
            }
        }
    }
    if (!LAST) HS_SYNC();
}

template <class P, int SGN, int S, class LD, class ST>
HS_DEVICE void hs_stages_from(LD &ld, ST &st, cplx *sm, const cplx *tw, int t, bool active)
{
    hs_stage<P, SGN, S>(ld, st, sm, tw, t, active);
    if constexpr (S + 1 < P::NST) hs_stages_from<P, SGN, S + 1>(ld, st, sm, tw, t, active);
}

// Whole transform for one thread team.  sm: this team's NPAD-point buffer; tw: the plan's device
// twiddle table; t: thread index inside the team; active: false for the padding teams of the last
// CTA (they still take part in the barriers).
template <class P, int SGN, class LD, class ST>
HS_DEVICE void hs_fft_team(LD &ld, ST &st, cplx *sm, const cplx *tw, int t, bool active)
{
    hs_stages_from<P, SGN, 0>(ld, st, sm, tw, t, active);
}

// ------------------------------------------------------------------------- plain c2c kernel
struct HsLoadGlobal {
    const cplx *HS_RESTRICT base;
    HS_DEVICE cplx operator()(int, int, int n) const { return hs_ld_stream(base + n); }
};
struct HsStoreGlobal {
    cplx *HS_RESTRICT base;
    HS_DEVICE void operator()(int, int, int k, cplx v) const { hs_st_stream(base + k, v); }
};

// grid = ceil(howmany / X) CTAs of T*X threads; transform b lives at in + b*N / out + b*N.
template <class P, int SGN, int MINB>
HS_GLOBAL void HS_LAUNCH_BOUNDS(P::threads, MINB)
hs_c2c_smem_kernel(const cplx *HS_RESTRICT in, cplx *HS_RESTRICT out, const cplx *HS_RESTRICT tw, long long howmany)
{
    HS_DYN_SMEM(cplx, smem);
    const int team = HS_TID / P::T, t = HS_TID % P::T;
    const long long xf = HS_BID * P::X + team;
    const bool active = xf < howmany;
    HsLoadGlobal ld{in + (active ? xf : 0) * P::N};
    HsStoreGlobal st{out + (active ? xf : 0) * P::N};
    hs_fft_team<P, SGN>(ld, st, smem + team * P::NPAD, tw, t, active);
}
