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
#include "hs_abi.h"
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
// LOGPAD_: one pad slot per 2^LOGPAD_ points (0 = none); SPLIT_: how a stage gets its r-1 twiddles --
//   0: load all r-1 rows of the [i-1][k] table;  -1: load rows 1,2,4,8 and form the others as products
//   (-2: load row 1 only)
//   (fewer 16-byte loads through the L1 data pipe, which is what limits these kernels, for a few FP64
//   multiplies);  >0: split-table threshold in entries (hs_split_kl in hs_abi.h);
// GROUPS_: how consecutive stages are fused into register-level passes, as decimal digits from the
//   first pass to the last (221 = {s0,s1} {s2,s3} {s4}); 0 = every stage is its own pass.  A fused
//   pass of radices (ra, rb) keeps ra*rb points per butterfly in registers and exchanges through
//   shared memory once instead of twice; the twiddles, butterflies and their order are unchanged.
// RS...: stage radices in EXECUTION order
// (= the reference's factor list reversed, runs of power-of-two factors possibly regrouped).
template <int N_, int T_, int X_, int LOGPAD_, int SPLIT_, int GROUPS_, int... RS>
struct HsPlan {
    using radices = HsRadixList<RS...>;
    static constexpr int N = N_, T = T_, X = X_, LOGPAD = LOGPAD_, SPLIT = SPLIT_;
    static constexpr int NST = radices::count;
    static constexpr int radix(int s) { return radices::at(s); }
    // pass structure
    static constexpr int ndigits(int g) { return g == 0 ? 0 : 1 + ndigits(g / 10); }
    static constexpr int pow10(int e) { return e == 0 ? 1 : 10 * pow10(e - 1); }
    static constexpr int NPASS = GROUPS_ ? ndigits(GROUPS_) : NST;
    static constexpr int pass_size(int p) { return GROUPS_ ? (GROUPS_ / pow10(NPASS - 1 - p)) % 10 : 1; }
    static constexpr int pass_start(int p) { return p == 0 ? 0 : pass_start(p - 1) + pass_size(p - 1); }
    static constexpr int m_before(int s) { return s == 0 ? 1 : m_before(s - 1) * radix(s - 1); }
    // offset of stage s in the device twiddle table ([i-1][k] layout per stage), stage 0 has none
    static constexpr int split_kl(int s) { return s == 0 ? 0 : hs_split_kl(radix(s), m_before(s), SPLIT_); }
    static constexpr int tw_off(int s) { return s <= 1 ? 0 : tw_off(s - 1) + hs_stage_tw_entries(radix(s - 1), m_before(s - 1), SPLIT_); }
    static constexpr int tw_count = tw_off(NST);
    static constexpr int NPAD = LOGPAD_ ? N_ + (N_ >> LOGPAD_) + 1 : N_;
    static constexpr int pad(int p) { return LOGPAD_ ? p + (p >> LOGPAD_) : p; }
    static constexpr int smem_bytes = NPASS > 1 ? NPAD * X_ * 16 : 0;
    static constexpr int threads = T_ * X_;
    static constexpr int ABL = 0; // ablation bits for profiling builds (1: skip twiddle loads -- wrong results)
    static constexpr bool EXACT = false; // radix-3/5/7 butterflies with the reference's 11-digit literals
};
template <class Base, int A>
struct HsAblate : Base { static constexpr int ABL = A; };
// the same plan with full-precision radix-3/5/7 constants: for the padded convolution inside Bluestein, which
// the reference runs with power-of-two stages only (so its result carries no truncation error to reproduce)
template <class Base>
struct HsExact : Base { static constexpr bool EXACT = true; };

// Where a thread sits in its team for one pass: team buffer, index inside the team, and whether the
// team has work (padding teams of the last CTA still take part in the barriers).  Tiled kernels use
// one mapping for the first pass (teams across lanes: coalesced strided loads) and another for the
// rest (threads of a team across lanes: coalesced contiguous stores); the shared-memory exchange
// between passes makes the switch free.
struct HsLane {
    cplx *sm;
    int t;
    bool active;
};

// Hooks a caller may hang into a transform: after the first pass has issued its global loads (their latency
// hides whatever the hook does), after the first pass has finished (its trailing barrier passed: every thread
// has consumed the input) and before the last pass (one shared-memory exchange and the global stores still to
// come).  The persistent pipeline kernels (hs_pipe.cuh) run their dataflow protocol there.
struct HsNoHooks {
    HS_DEVICE void sync() const { HS_SYNC(); } // the barrier between passes: the whole CTA, unless the caller's warps have roles
    HS_DEVICE void after_first_loads() const {}
    HS_DEVICE void after_first_pass() const {}
    HS_DEVICE void before_last_pass() const {}
};

// A loader / storer may itself work on the team's shared-memory buffer (a staged gather or scatter, hs_fused.cuh): it
// says so with `static constexpr bool kFromSmem / kToSmem = true`, and the pass then keeps the read-before-overwrite
// barrier it would otherwise drop for a first / last pass.
template <class T, class = void>
struct HsLdFromSmem { static constexpr bool value = false; };
template <class T>
struct HsLdFromSmem<T, decltype((void)T::kFromSmem)> { static constexpr bool value = T::kFromSmem; };
template <class T, class = void>
struct HsStToSmem { static constexpr bool value = false; };
template <class T>
struct HsStToSmem<T, decltype((void)T::kToSmem)> { static constexpr bool value = T::kToSmem; };

// ---------------------------------------------------------------------------- stage twiddles
// wv[1..R) <- W_{M R}^{i k}, i = 1..R-1, for flat stage S (radix R on sub-transforms of length M).
template <class P, int S>
HS_DEVICE void hs_twiddles(const cplx *HS_RESTRICT tw, const int k, cplx *wv)
{
    constexpr int R = P::radix(S);
    constexpr int M = P::m_before(S);
    constexpr int KL = P::split_kl(S);
    if (KL == 0 && P::SPLIT < 0) {
        // load only the rows i = 1, 2, 4, 8 (or just row 1) and build the others by products (<= 3
        // deep): fewer 16-byte loads through the L1 data pipe, a few more FP64 multiplies
        const cplx *w = tw + P::tw_off(S) + k;
        if (P::SPLIT == -1) {
            HS_UNROLL
            for (int i = 1; i < R; i *= 2) wv[i] = hs_ld_tab(w + (i - 1) * M);
        } else {
            wv[1] = hs_ld_tab(w);
            HS_UNROLL
            for (int i = 2; i < R; i *= 2) wv[i] = hs_mul(wv[i / 2], wv[i / 2]);
        }
        HS_UNROLL
        for (int i = 3; i < R; ++i)
            if (i & (i - 1)) { // not a power of two: highest set bit times the rest
                int hb = 1;
                while (hb * 2 <= i) hb *= 2;
                wv[i] = hs_mul(wv[hb], wv[i - hb]);
            }
    } else if (KL == 0) { // one table, [i-1][k]: load every row
        const cplx *w = tw + P::tw_off(S) + k;
        HS_UNROLL
        for (int i = 1; i < R; ++i) wv[i] = (P::ABL & 1) ? hs_c(1.0, 1e-300 * i) : hs_ld_tab(w + (i - 1) * M);
    } else { // split tables: W^(i k) = lo[i-1][k % KL] * hi[i-1][k / KL]
        constexpr int KH = M / (KL ? KL : 1);
        const cplx *wl = tw + P::tw_off(S) + k % (KL ? KL : 1);
        const cplx *wh = tw + P::tw_off(S) + (R - 1) * KL + k / (KL ? KL : 1);
        HS_UNROLL
        for (int i = 1; i < R; ++i) wv[i] = hs_mul(hs_ld_tab(wl + (i - 1) * KL), hs_ld_tab(wh + (i - 1) * KH));
    }
}

// SPLIT = -3: one table load (row 1) per butterfly, the powers w^2, w^3, ... formed one after the other and applied at
// once, so that only two twiddles are ever live (a radix-16 stage otherwise holds up to 7 of them next to its 16 points:
// with 640 threads per CTA -- 96 registers -- that spills).  R = 16 applies w^i to points i and 8+i and w^8 to 8..15:
// 29 complex products like the tree of SPLIT = -2, but a chain 8 deep; rounding grows with the chain, so this mode is
// for the convolutions inside Bluestein / Rader only (never for a transform whose result the reference defines bit-wise).
template <class P, int S>
HS_DEVICE void hs_twiddle_apply_seq(const cplx *HS_RESTRICT tw, const int k, cplx *v)
{
    constexpr int R = P::radix(S);
    const cplx w1 = hs_ld_tab(tw + P::tw_off(S) + k);
    cplx w = w1;
    if (R == 16) {
        HS_UNROLL
        for (int i = 1; i < 8; ++i) {
            v[i] = hs_mul(v[i], w);
            v[8 + i] = hs_mul(v[8 + i], w);
            w = hs_mul(w, w1);
        }
        HS_UNROLL
        for (int i = 8; i < 16; ++i) v[i] = hs_mul(v[i], w);
    } else {
        HS_UNROLL
        for (int i = 1; i < R; ++i) {
            v[i] = hs_mul(v[i], w);
            if (i + 1 < R) w = hs_mul(w, w1);
        }
    }
}

// ------------------------------------------------------------------------------- one stage
// LD(u, i, n)       -> cplx   : element n of the input signal   (first pass only)
// ST(u, j, k, v)              : element k of the output spectrum (last pass only)
// u (butterfly slot of this thread) and i/j are compile-time constants after unrolling, so
// register-to-register hand-over between two transforms (fused Bluestein) stays in registers.
template <class P, int SGN, int S, class LD, class ST, class HK>
HS_DEVICE void hs_stage(LD &ld, ST &st, const HsLane &ln, const cplx *HS_RESTRICT tw, HK &hk)
{
    cplx *const sm = ln.sm;
    const int t = ln.t;
    const bool active = ln.active;
    constexpr int R = P::radix(S);
    constexpr int M = P::m_before(S);   // sub-transform length already done
    constexpr int QP = P::N / (M * R);  // q' range after this stage
    constexpr int NB = P::N / R;        // butterflies in this stage
    constexpr int B = (NB + P::T - 1) / P::T;
    constexpr bool FIRST = (S == 0), LAST = (S == P::NST - 1);
    cplx v[B][R];

    // phase 1: issue every load of this thread's butterflies back to back (memory-level parallelism)
    HS_UNROLL
    for (int u = 0; u < B; ++u) {
        const int b = t + P::T * u;
        if (active && (NB % P::T == 0 || u < NB / P::T || b < NB)) {
            const int q = b % QP, k = b / QP;
            HS_UNROLL
            for (int i = 0; i < R; ++i) {
                if (FIRST) v[u][i] = ld(u, i, b + NB * i);
                else       v[u][i] = sm[P::pad(q + QP * i + (QP * R) * k)];
            }
        }
    }
    if constexpr (FIRST) hk.after_first_loads();
    // phase 2: twiddle (reference combine loops) and butterfly, in registers
    HS_UNROLL
    for (int u = 0; u < B; ++u) {
        const int b = t + P::T * u;
        if (active && (NB % P::T == 0 || u < NB / P::T || b < NB)) {
            if (!FIRST && P::SPLIT == -3) {
                hs_twiddle_apply_seq<P, S>(tw, b / QP, v[u]);
            } else if (!FIRST) {
                cplx wv[R];
                hs_twiddles<P, S>(tw, b / QP, wv);
                HS_UNROLL
                for (int i = 1; i < R; ++i) v[u][i] = hs_mul(v[u][i], wv[i]);
            } else if constexpr (LD::kPost) {
                ld.template post<R>(b, v[u]); // inter-pass twiddle of a multi-pass transform
            }
            hs_bfly<R, SGN, P::EXACT>(v[u]);
        }
    }
    if ((!FIRST || HsLdFromSmem<LD>::value) && (!LAST || HsStToSmem<ST>::value)) hk.sync(); // every thread has read its inputs; the buffer may be overwritten
    HS_UNROLL
    for (int u = 0; u < B; ++u) {
        const int b = t + P::T * u;
        if (active && (NB % P::T == 0 || u < NB / P::T || b < NB)) {
            HS_UNROLL
            for (int j = 0; j < R; ++j) {
                if (LAST) st(u, j, b + NB * j, v[u][j]);
                else      sm[P::pad(b + NB * j)] = v[u][j];
            }
        }
    }
    if (!LAST) hk.sync();
}

// ------------------------------------------------------------- two consecutive stages, fused
// Stages F (radix RA) and F+1 (radix RB) run back to back on RA*RB points held in registers:
//   in  : A_m0[q' + QB (ib + RB ia)][k0]                      (element e = ib + RB ia)
//   a   : y[ib][ja] = radix-RA over ia of ( W_{m0 RA}^{ia k0} * in )
//   b   : out[jb][ja] = radix-RB over ib of ( W_{m0 RA RB}^{ib (k0 + m0 ja)} * y[ib][ja] )
//   out : A_{m0 RA RB}[q'][k0 + m0 ja + m0 RA jb]              (element o = ja + RA jb)
// i.e. exactly the reference's two combine loops, with the intermediate state never leaving the
// thread.  To shared memory it looks like one stage of radix RA*RB (bank_check.py models it so).
template <class P, int SGN, int F, class LD, class ST, class HK>
HS_DEVICE void hs_stage2(LD &ld, ST &st, const HsLane &ln, const cplx *HS_RESTRICT tw, HK &hk)
{
    static_assert(!(F == 0 && LD::kPost), "a loader with an inter-pass twiddle hook needs an unfused first pass");
    cplx *const sm = ln.sm;
    const int t = ln.t;
    const bool active = ln.active;
    constexpr int RA = P::radix(F), RB = P::radix(F + 1), RC = RA * RB;
    constexpr int M0 = P::m_before(F);
    constexpr int QB = P::N / (M0 * RC);
    constexpr int NB = P::N / RC;
    constexpr int B = (NB + P::T - 1) / P::T;
    constexpr bool FIRST = (F == 0), LAST = (F + 2 == P::NST);
    cplx v[B][RB][RA];

    HS_UNROLL
    for (int u = 0; u < B; ++u) {
        const int b = t + P::T * u;
        if (active && (NB % P::T == 0 || u < NB / P::T || b < NB)) {
            const int q = b % QB, k0 = b / QB;
            HS_UNROLL
            for (int ia = 0; ia < RA; ++ia) {
                HS_UNROLL
                for (int ib = 0; ib < RB; ++ib) {
                    const int e = ib + RB * ia;
                    if (FIRST) v[u][ib][ia] = ld(u, e, b + NB * e);
                    else       v[u][ib][ia] = sm[P::pad(q + QB * e + (QB * RC) * k0)];
                }
            }
        }
    }
    if constexpr (FIRST) hk.after_first_loads();
    HS_UNROLL
    for (int u = 0; u < B; ++u) {
        const int b = t + P::T * u;
        if (active && (NB % P::T == 0 || u < NB / P::T || b < NB)) {
            const int k0 = b / QB;
            if (!FIRST) {
                cplx wa[RA];
                hs_twiddles<P, F>(tw, k0, wa);
                HS_UNROLL
                for (int ib = 0; ib < RB; ++ib) {
                    HS_UNROLL
                    for (int ia = 1; ia < RA; ++ia) v[u][ib][ia] = hs_mul(v[u][ib][ia], wa[ia]);
                }
            }
            HS_UNROLL
            for (int ib = 0; ib < RB; ++ib) hs_bfly<RA, SGN, P::EXACT>(v[u][ib]);
            HS_UNROLL
            for (int ja = 0; ja < RA; ++ja) {
                cplx tt[RB];
                HS_UNROLL
                for (int ib = 0; ib < RB; ++ib) tt[ib] = v[u][ib][ja];
                if (!(FIRST && ja == 0)) { // k1 = 0 in the first pass: W^0 = 1, as in the reference's k = 0 column
                    cplx wb[RB];
                    hs_twiddles<P, F + 1>(tw, k0 + M0 * ja, wb);
                    HS_UNROLL
                    for (int ib = 1; ib < RB; ++ib) tt[ib] = hs_mul(tt[ib], wb[ib]);
                }
                hs_bfly<RB, SGN, P::EXACT>(tt);
                HS_UNROLL
                for (int jb = 0; jb < RB; ++jb) v[u][jb][ja] = tt[jb];
            }
        }
    }
    if ((!FIRST || HsLdFromSmem<LD>::value) && (!LAST || HsStToSmem<ST>::value)) hk.sync();
    HS_UNROLL
    for (int u = 0; u < B; ++u) {
        const int b = t + P::T * u;
        if (active && (NB % P::T == 0 || u < NB / P::T || b < NB)) {
            HS_UNROLL
            for (int jb = 0; jb < RB; ++jb) {
                HS_UNROLL
                for (int ja = 0; ja < RA; ++ja) {
                    const int o = ja + RA * jb;
                    if (LAST) st(u, o, b + NB * o, v[u][jb][ja]);
                    else      sm[P::pad(b + NB * o)] = v[u][jb][ja];
                }
            }
        }
    }
    if (!LAST) hk.sync();
}

template <class P, int SGN, int PASS, class LD, class ST, class HK>
HS_DEVICE void hs_passes_from(LD &ld, ST &st, const HsLane &first, const HsLane &rest, const cplx *tw, HK &hk)
{
    constexpr int F = P::pass_start(PASS);
    static_assert(P::pass_size(PASS) == 1 || P::pass_size(PASS) == 2, "a pass fuses one or two stages");
    const HsLane &ln = (PASS == 0) ? first : rest;
    if constexpr (PASS > 0 && PASS + 1 == P::NPASS) hk.before_last_pass();
    if constexpr (P::pass_size(PASS) == 2) hs_stage2<P, SGN, F>(ld, st, ln, tw, hk);
    else hs_stage<P, SGN, F>(ld, st, ln, tw, hk);
    if constexpr (PASS == 0 && P::NPASS > 1) hk.after_first_pass();
    if constexpr (PASS + 1 < P::NPASS) hs_passes_from<P, SGN, PASS + 1>(ld, st, first, rest, tw, hk);
}

// Whole transform for one thread team.  sm: this team's NPAD-point buffer; tw: the plan's device
// twiddle table; t: thread index inside the team; active: false for the padding teams of the last
// CTA (they still take part in the barriers).
template <class P, int SGN, class LD, class ST>
HS_DEVICE void hs_fft_team(LD &ld, ST &st, cplx *sm, const cplx *tw, int t, bool active)
{
    static_assert(P::pass_start(P::NPASS - 1) + P::pass_size(P::NPASS - 1) == P::NST, "GROUPS must cover every stage");
    const HsLane ln{sm, t, active};
    HsNoHooks hk;
    hs_passes_from<P, SGN, 0>(ld, st, ln, ln, tw, hk);
}
template <class P, int SGN, class LD, class ST, class HK>
HS_DEVICE void hs_fft_team_h(LD &ld, ST &st, cplx *sm, const cplx *tw, int t, bool active, HK &hk)
{
    static_assert(P::pass_start(P::NPASS - 1) + P::pass_size(P::NPASS - 1) == P::NST, "GROUPS must cover every stage");
    const HsLane ln{sm, t, active};
    hs_passes_from<P, SGN, 0>(ld, st, ln, ln, tw, hk);
}
// Same, with a different thread-to-team mapping for the first pass than for the others.
template <class P, int SGN, class LD, class ST, class HK>
HS_DEVICE void hs_fft_team2(LD &ld, ST &st, const HsLane &first, const HsLane &rest, const cplx *tw, HK &hk)
{
    static_assert(P::pass_start(P::NPASS - 1) + P::pass_size(P::NPASS - 1) == P::NST, "GROUPS must cover every stage");
    static_assert(P::NPASS >= 2, "switching the mapping needs a shared-memory exchange");
    hs_passes_from<P, SGN, 0>(ld, st, first, rest, tw, hk);
}
template <class P, int SGN, class LD, class ST>
HS_DEVICE void hs_fft_team2(LD &ld, ST &st, const HsLane &first, const HsLane &rest, const cplx *tw)
{
    HsNoHooks hk;
    hs_fft_team2<P, SGN>(ld, st, first, rest, tw, hk);
}

// ------------------------------------------------------------------------- plain c2c kernel
struct HsLoadGlobal {
    static constexpr bool kPost = false;
    const cplx *HS_RESTRICT base;
    HS_DEVICE cplx operator()(int, int, int n) const { return hs_ld_stream(base + n); }
};
struct HsStoreGlobal {
    cplx *HS_RESTRICT base;
    HS_DEVICE void operator()(int, int, int k, cplx v) const { hs_st_stream(base + k, v); }
};

// grid = ceil(howmany / X) CTAs of T*X threads; transform b lives at in + b*N / out + b*N.
// (hs_jit.cu generates the same kernel for lengths without a static plan.  Keep the body in the kernel itself: wrapped in a
// device function it ran at 43 % instead of 78 % of the HBM peak for N = 3000.)
template <class P, int SGN, int MINB>
HS_GLOBAL void HS_LAUNCH_BOUNDS(P::threads, MINB)
hs_c2c_smem_kernel(const cplx *HS_RESTRICT in, cplx *HS_RESTRICT out, const cplx *HS_RESTRICT tw, long long howmany, int pf_dist)
{
    HS_DYN_SMEM(cplx, smem);
    const int team = HS_TID / P::T, t = HS_TID % P::T;
    const long long xf = HS_BID * P::X + team;
    const bool active = xf < howmany;
    if (pf_dist > 0 && HS_TID == 0) { // warm the L2 for the CTA that takes this one's place: its loads then see L2, not HBM, latency
        const long long pf = (HS_BID + pf_dist) * P::X, left = howmany - pf;
        if (left > 0) hs_prefetch_l2(in + pf * P::N, (unsigned)((left < P::X ? left : P::X) * P::N * 16));
    }
    HsLoadGlobal ld{in + (active ? xf : 0) * P::N};
    HsStoreGlobal st{out + (active ? xf : 0) * P::N};
    hs_fft_team<P, SGN>(ld, st, smem + team * P::NPAD, tw, t, active);
}

// Persistent form of the one-kernel transform for plans that run ONE CTA per SM (plan-time kernels for N = 5000...6600:
// the exchange buffer or the register file leaves room for one).  With a single CTA nothing covers the HBM latency of the
// first pass's loads, so the NEXT transform's input is staged by the copy engine: one cp.async.bulk of N*16 bytes into a
// second shared-memory buffer, completion on an mbarrier, issued as soon as the first pass has consumed the buffer, so the
// copy runs under the remaining passes and the stores.  (At two CTAs per SM the loop does not fit the register cap: N =
// 3000 spills 500 bytes and runs 1.98 vs 1.23 ms, profiles/r2_n3000_staged_persistent_ab.txt.)
struct HsStageRefill { // hook: after the first pass's barrier thread 0 starts the copy of the CTA's next transform
    HsMbar *bar;
    cplx *stage;
    const cplx *HS_RESTRICT in;
    long long howmany;
    int j, n;
    HS_DEVICE void sync() { HS_SYNC(); }
    HS_DEVICE void after_first_loads() {}
    HS_DEVICE void after_first_pass()
    {
        hs_fence_async_proxy(); // this transform's reads of the staging buffer precede the copy that refills it
        if (HS_TID == 0) {
            const long long next = HS_BID + (long long)(j + 1) * HS_NBLOCKS;
            if (next < howmany) {
                hs_mbar_expect(bar, (unsigned)(n * 16));
                hs_bulk_g2s(stage, in + next * n, (unsigned)(n * 16), bar);
            }
        }
    }
    HS_DEVICE void before_last_pass() {}
};
struct HsLoadStage {
    static constexpr bool kPost = false;
    const cplx *stage;
    HS_DEVICE cplx operator()(int, int, int n) const { return stage[n]; }
};
template <class P>
struct HsStagedSmem {
    static constexpr int stage_off = (P::NPAD + 7) / 8 * 8; // 128-byte aligned
    static constexpr int bytes = (stage_off + P::N) * 16 + 16;
};

// ------------------------------------------------------------------ real transforms, one kernel
// reference src/real.c packs N reals as H = N/2 complex samples, runs the half-length complex plan and splits (r2c,
// :109-132) or merges (c2r, :171-179) with twiddle2[k] = exp(+2 pi i k / N).  Here both ride on the complex kernel:
//   c2r: the first pass builds element k of the packed signal from bins k and H-k of the Hermitian input as it loads it;
//   r2c: the last pass leaves the half-length spectrum Z in the team's shared-memory buffer (natural order), and after
//        one barrier every thread turns pairs (Z[k], Z[H-k]) into bins k and H-k -- and their mirrors N-k, H+k when the
//        caller wants the reference's full N-bin spectrum -- with coalesced stores.
// The merged signal / the half-length spectrum never exist in HBM.
struct HsLoadC2RMerge {
    static constexpr bool kPost = false;
    const cplx *HS_RESTRICT x;   // this transform's bins 0..H
    const cplx *HS_RESTRICT tw2; // exp(+2 pi i k / N)
    int H;
    HS_DEVICE cplx operator()(int, int, int k) const
    {
        const cplx a = hs_ld_stream(x + k), c = hs_ld_stream(x + (H - k)), w = hs_ld_tab(tw2 + k);
        const double t1 = -a.y - c.y, t2 = -c.x + a.x; // reference :171-179
        return hs_c(a.x + c.x + (t1 * w.x) - (t2 * w.y), a.y - c.y + (t2 * w.x) + (t1 * w.y));
    }
};
template <class P>
struct HsStoreSmemNatural {
    static constexpr bool kToSmem = true;
    cplx *sm;
    HS_DEVICE void operator()(int, int, int k, cplx v) const { sm[P::pad(k)] = v; }
};
// bins k and H-k (and their mirrors) from the half-length spectrum Z held in sm; pairs k = t, t + T, ... <= H/2.
// In two steps, so that the twiddle2 loads are in flight while the CTA waits at the barrier that ends the
// transform (ncu had 29 % of the r2c kernel's stall samples on these two loads: they were issued after the barrier, at the
// very end of the CTA, where nothing else can cover them).  w[2 i], w[2 i + 1] = twiddle2[k], twiddle2[H - k], k = t + T i.
template <class P>
struct HsR2CSplitTw {
    static constexpr int ITER = (P::N / 2) / P::T + 1;
    cplx w[2 * ITER];
    HS_DEVICE void load(const cplx *HS_RESTRICT tw2, const int t)
    {
        HS_UNROLL
        for (int i = 0; i < ITER; ++i) {
            const int k = t + P::T * i;
            if (k <= P::N / 2) { w[2 * i] = hs_ld_tab(tw2 + k); w[2 * i + 1] = hs_ld_tab(tw2 + (P::N - k)); }
        }
    }
};
template <class P>
HS_DEVICE void hs_r2c_split_rows_w(const cplx *sm, cplx *HS_RESTRICT o, const HsR2CSplitTw<P> &tw, const int full, const int t)
{
    constexpr int H = P::N;
    HS_UNROLL
    for (int i = 0; i < HsR2CSplitTw<P>::ITER; ++i) {
        const int k = t + P::T * i;
        if (k > H / 2) continue;
        if (k == 0) {
            const cplx z0 = sm[0];
            hs_st_stream(o, hs_c(z0.x + z0.y, 0.0));
            hs_st_stream(o + H, hs_c(z0.x - z0.y, 0.0));
        } else {
            const cplx a = sm[P::pad(k)], c = sm[P::pad(H - k)];
            const cplx r = hs_r2c_bin(a, c, tw.w[2 * i]);
            hs_st_stream(o + k, r);
            if (full) hs_st_stream(o + (2 * H - k), hs_c(r.x, -r.y));
            if (2 * k != H) {
                const cplx r2 = hs_r2c_bin(c, a, tw.w[2 * i + 1]);
                hs_st_stream(o + (H - k), r2);
                if (full) hs_st_stream(o + (H + k), hs_c(r2.x, -r2.y));
            }
        }
    }
}

// in: rows of 2H reals read as H complex samples (contiguous); out: rows of out_stride bins
template <class P, int SGN, int MINB>
HS_GLOBAL void HS_LAUNCH_BOUNDS(P::threads, MINB)
hs_r2c_smem_kernel(const cplx *HS_RESTRICT in, cplx *HS_RESTRICT out, const cplx *HS_RESTRICT tw, const cplx *HS_RESTRICT tw2, long long out_stride,
                   int full, long long howmany)
{
    static_assert(P::NPASS > 1, "the split works on the shared-memory buffer");
    HS_DYN_SMEM(cplx, smem);
    const int team = HS_TID / P::T, t = HS_TID % P::T;
    const long long xf = HS_BID * P::X + team;
    const bool active = xf < howmany;
    cplx *sm = smem + team * P::NPAD;
    HsLoadGlobal ld{in + (active ? xf : 0) * P::N};
    HsStoreSmemNatural<P> st{sm};
    hs_fft_team<P, SGN>(ld, st, sm, tw, t, active);
    HsR2CSplitTw<P> w2;
    w2.load(tw2, t); // in flight across the barrier
    HS_SYNC();
    if (active) hs_r2c_split_rows_w<P>(sm, out + xf * out_stride, w2, full, t);
}

// the merge from a copy of bins 0..H-1 in the team's own buffer (bin H, needed for k = 0 only, comes from global memory):
// every bin is then read from HBM / L2 once, in order, by asynchronous 16-byte copies, instead of twice (as k and as H-k)
struct HsLoadC2RMergeSmem {
    static constexpr bool kPost = false, kFromSmem = true;
    const cplx *sm;                // bins 0..H-1, natural order
    const cplx *HS_RESTRICT xh;    // bin H
    const cplx *HS_RESTRICT tw2;
    int H;
    HS_DEVICE cplx operator()(int, int, int k) const
    {
        const cplx a = sm[k], c = k ? sm[H - k] : hs_ld_tab(xh), w = hs_ld_tab(tw2 + k);
        const double t1 = -a.y - c.y, t2 = -c.x + a.x; // reference :171-179
        return hs_c(a.x + c.x + (t1 * w.x) - (t2 * w.y), a.y - c.y + (t2 * w.x) + (t1 * w.y));
    }
};

// in: rows of in_stride bins (0..H are read); out: rows of 2H reals written as H complex samples (contiguous)
// STG: the merge reads a staged copy of the row (HsLoadC2RMergeSmem) instead of loading every bin twice
template <class P, int SGN, int MINB, bool STG = false>
HS_GLOBAL void HS_LAUNCH_BOUNDS(P::threads, MINB)
hs_c2r_smem_kernel(const cplx *HS_RESTRICT in, cplx *HS_RESTRICT out, const cplx *HS_RESTRICT tw, const cplx *HS_RESTRICT tw2, long long in_stride,
                   long long howmany)
{
    HS_DYN_SMEM(cplx, smem);
    const int team = HS_TID / P::T, t = HS_TID % P::T;
    const long long xf = HS_BID * P::X + team;
    const bool active = xf < howmany;
    HsStoreGlobal st{out + (active ? xf : 0) * P::N};
    if constexpr (STG) {
        cplx *sm = smem + team * P::NPAD;
        const cplx *x = in + (active ? xf : 0) * in_stride;
        if (active)
            for (int k = t; k < P::N; k += P::T) hs_cp_async16(sm + k, x + k);
        hs_cp_async_wait_all();
        HS_SYNC();
        HsLoadC2RMergeSmem ld{sm, x + P::N, tw2, P::N};
        hs_fft_team<P, SGN>(ld, st, sm, tw, t, active);
    } else {
        HsLoadC2RMerge ld{in + (active ? xf : 0) * in_stride, tw2, P::N};
        hs_fft_team<P, SGN>(ld, st, smem + team * P::NPAD, tw, t, active);
    }
}
