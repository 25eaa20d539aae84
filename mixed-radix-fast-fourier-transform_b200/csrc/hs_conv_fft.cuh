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

// butterflies of a thread that the compiler may interleave (1: strictly one after the other, fewest registers).  Measured on
// B200 with the tensor-memory kernel, N = 10007 x 16384 (profiles/r2_split_tmem_variants.txt): 1 -> 4.53 ms, 2 -> 4.33 ms,
// 4 -> 6.24 ms (1.4 KB of spill stores per thread); 320 threads with 2 / 4 interleaved butterflies 6.2 / 6.4 ms.
#ifndef HS_CONV_UNROLL
#define HS_CONV_UNROLL 2
#endif
#define HS_PRAGMA_(x) _Pragma(#x)
#define HS_CONV_UNROLL_PRAGMA(n) HS_PRAGMA_(unroll n)

// Forward stage S (not the last).  FIRST: the inputs come from ld(n) (global memory), else from shared memory.
template <class P, int S, bool FIRST, class LD>
HS_DEVICE void hs_dif_stage(LD &ld, cplx *sm, const cplx *HS_RESTRICT tw, const int t)
{
    constexpr int R = P::radix(S), LB = P::L(S), M = P::m(S), NB = P::N / R;
    constexpr int B = (NB + P::T - 1) / P::T;
HS_CONV_UNROLL_PRAGMA(HS_CONV_UNROLL)
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
HS_CONV_UNROLL_PRAGMA(HS_CONV_UNROLL)
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
HS_CONV_UNROLL_PRAGMA(HS_CONV_UNROLL)
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

// inverse stages S, S-1, ... LO (each preceded by its barrier); the caller does stages LO-1 .. 0 itself
template <class P, int S, int LO, class ST>
HS_DEVICE void hs_conv_inverse_down_to(ST &st, cplx *sm, const cplx *tw, const int t)
{
    if constexpr (S >= LO) {
        HS_SYNC();
        hs_dit_stage<P, S, false>(st, sm, tw, t);
        hs_conv_inverse_down_to<P, S - 1, LO>(st, sm, tw, t);
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

// ---------------------------------------------------------------------------------------------------------------
// The same split chirp-z with its intermediates parked in TENSOR MEMORY instead of travelling through the L2.
// ncu on the kernel above (profiles/r2_ncu_blu10007_split_summary.txt, stall timeline): 44 % of its stall samples sit in
// the phases that move data between the SM and the L2 -- the second read of x with its second chirp table, the parked
// first half-convolution going out and coming back, three loads per point in the final combine: 1.77 MB per transform
// through an SM port that sustains about 32 bytes per clock there.  Two observations remove 45 % of it:
//   * the second convolution's input is a1 = a0 * w with a0 = x * conj(h) already in registers in the first stage of the
//     first convolution, and w[p + 640 i] = w[p] * exp(-i pi i / 16): one table load per thread and compile-time constants.
//     a1 is computed there and parked in the thread's TMEM slot (tcgen05.st, 16 points = 64 columns of its lane);
//   * at the turn between the two convolutions the thread that holds y0[p + 640 i] (last inverse stage of the first) is
//     the thread that needs a1[p + 640 i] (first forward stage of the second): it swaps them, four points at a time, so
//     the same 160 KB of TMEM hold a1 during the first convolution and y0 during the second.
// The final stage reads y0 back, four points at a time, next to y1 in registers:  out = (y0 + conj(w) y1) * conj(h).
// L2 traffic per transform: x, h (twice), Hk and the result -- no second read of x, no gA / gB tables, no parked row.
// Needs one radix-16 butterfly per thread on the outer stage (P::N = 16 T) and the whole TMEM of the SM (one CTA per SM).
template <class P, int SGN, int MINB>
HS_GLOBAL void HS_LAUNCH_BOUNDS(P::threads, MINB)
hs_conv_bluestein_split_tmem_kernel(const cplx *HS_RESTRICT in, cplx *HS_RESTRICT out, const cplx *HS_RESTRICT tw, const cplx *HS_RESTRICT h,
                                    const cplx *HS_RESTRICT wp, const cplx *HS_RESTRICT hkt2, int N, long long howmany, int pf_dist)
{
    static_assert(P::radix(0) == 16 && P::N == 16 * P::T && P::NST >= 3 && P::T % 32 == 0, "one radix-16 butterfly per thread on the outer stage");
    constexpr int M0 = P::m(0);
    // exp(-i pi i / 16), i < 16
    constexpr double CR[16] = {1.0, 0.98078528040323044913, 0.92387953251128675613, 0.83146961230254523708, 0.70710678118654752440,
                               0.55557023301960222474, 0.38268343236508977173, 0.19509032201612826785, 0.0, -0.19509032201612826785,
                               -0.38268343236508977173, -0.55557023301960222474, -0.70710678118654752440, -0.83146961230254523708,
                               -0.92387953251128675613, -0.98078528040323044913};
    constexpr double CI[16] = {-0.0, -0.19509032201612826785, -0.38268343236508977173, -0.55557023301960222474, -0.70710678118654752440,
                               -0.83146961230254523708, -0.92387953251128675613, -0.98078528040323044913, -1.0, -0.98078528040323044913,
                               -0.92387953251128675613, -0.83146961230254523708, -0.70710678118654752440, -0.55557023301960222474,
                               -0.38268343236508977173, -0.19509032201612826785};
    HS_DYN_SMEM(cplx, sm);
    unsigned *tm_slot = reinterpret_cast<unsigned *>(sm + P::NPAD); // 16 spare bytes behind the transform buffer
    const int t = HS_TID;
    // 64 columns per thread, the warps of a lane quadrant side by side: a power of two of columns per CTA
    constexpr int QW = (P::T / 32 + 3) / 4, TCOLS = QW * 64 <= 64 ? 64 : (QW * 64 <= 128 ? 128 : (QW * 64 <= 256 ? 256 : 512));
    const unsigned tbase = hs_tmem_alloc(tm_slot, TCOLS), taddr = hs_tmem_my_addr(tbase, 64);
    const cplx *tw0 = tw + P::tw_off(0) + t; // outer-stage twiddles of this thread's butterfly (p = t)
    HsConvStorePark dummy{out, 0};
    cplx v[16];
    (void)pf_dist;

    // Persistent: the CTA walks the transforms xf = bid, bid + grid, ...  The raw samples of a transform are brought into this
    // thread's own sixteen positions of the buffer with asynchronous copies (cp.async) that the PREVIOUS transform's last stage
    // issues as soon as it has taken its values out of those positions: the HBM latency of x hides behind the final combine.
    auto fetch = [&](long long xf) {
        const cplx *x = in + xf * N;
        HS_UNROLL
        for (int i = 0; i < 16; ++i) {
            const int n = t + M0 * i;
            if (n < N) hs_cp_async16(&sm[P::pad(n)], x + n);
            else sm[P::pad(n)] = hs_c(0.0, 0.0);
        }
    };
    if (HS_BID < howmany) fetch(HS_BID);
    for (long long xf = HS_BID; xf < howmany; xf += HS_NBLOCKS) {
        cplx *o = out + xf * N;
        // ---- first convolution, outer forward stage: a0 = x * conj(h); a1 = a0 * w parked in TMEM
        hs_cp_async_wait_all();
        HS_UNROLL
        for (int i = 0; i < 16; ++i) v[i] = sm[P::pad(t + M0 * i)];
        HS_UNROLL
        for (int i = 0; i < 16; ++i)
            if (t + M0 * i < N) {
                const cplx hv = hs_ld_tab(h + t + M0 * i);
                v[i] = (SGN == 1) ? hs_mulc(v[i], hv) : hs_mul(v[i], hv);
            }
        {
            const cplx w_p = hs_ld_tab(wp + t); // loaded where it is used (an opaque load): the sixteen products w_p c_i are not kept alive
            HS_UNROLL
            for (int c = 0; c < 4; ++c) {
                cplx a1[4];
                HS_UNROLL
                for (int k = 0; k < 4; ++k) a1[k] = hs_mul(hs_mul(v[4 * c + k], hs_c(CR[4 * c + k], CI[4 * c + k])), w_p);
                hs_tmem_st4(taddr, c, a1);
            }
        }
        hs_bfly<16, 1, P::EXACT>(v);
        hs_conv_twiddle_apply<16, P::TWM, false>(tw0, M0, v);
        HS_UNROLL
        for (int j = 0; j < 16; ++j) sm[P::pad(t + M0 * j)] = v[j];
        HS_SYNC();
        hs_conv_forward_from<P, 1>(dummy, sm, tw, t);
        hs_conv_middle<P, SGN != 1>(sm, hkt2, t);
        hs_conv_inverse_down_to<P, P::NST - 2, 1>(dummy, sm, tw, t);

        // ---- the turn: last inverse stage of the first convolution, y0 <-> a1 through TMEM, first forward stage of the second
        HS_SYNC();
        HS_UNROLL
        for (int j = 0; j < 16; ++j) v[j] = sm[P::pad(t + M0 * j)];
        hs_conv_twiddle_apply<16, P::TWM, true>(tw0, M0, v);
        hs_bfly<16, -1, P::EXACT>(v);
        hs_tmem_wait_st();
        HS_UNROLL
        for (int c = 0; c < 4; ++c) {
            cplx a1[4];
            hs_tmem_ld4(taddr, c, a1);
            hs_tmem_st4(taddr, c, v + 4 * c);
            HS_UNROLL
            for (int k = 0; k < 4; ++k) v[4 * c + k] = a1[k];
        }
        hs_bfly<16, 1, P::EXACT>(v);
        hs_conv_twiddle_apply<16, P::TWM, false>(tw0, M0, v);
        HS_UNROLL
        for (int j = 0; j < 16; ++j) sm[P::pad(t + M0 * j)] = v[j];
        HS_SYNC();
        hs_conv_forward_from<P, 1>(dummy, sm, tw, t);
        hs_conv_middle<P, SGN != 1>(sm, hkt2 + P::N, t);
        hs_conv_inverse_down_to<P, P::NST - 2, 1>(dummy, sm, tw, t);

        // ---- last inverse stage of the second convolution and the combine: out = (y0 + conj(w) y1) * conj(h)
        HS_SYNC();
        HS_UNROLL
        for (int j = 0; j < 16; ++j) v[j] = sm[P::pad(t + M0 * j)];
        hs_conv_twiddle_apply<16, P::TWM, true>(tw0, M0, v);
        hs_bfly<16, -1, P::EXACT>(v);
        // this thread's positions of the buffer are free now (every v[j] above went through the butterfly): start the next
        // transform's samples on their way
        if (xf + HS_NBLOCKS < howmany) fetch(xf + HS_NBLOCKS);
        hs_tmem_wait_st();
        const cplx w_p = hs_ld_tab(wp + t);
        HS_UNROLL
        for (int c = 0; c < 4; ++c) {
            cplx y0[4];
            hs_tmem_ld4(taddr, c, y0);
            HS_UNROLL
            for (int k = 0; k < 4; ++k) {
                const int i = 4 * c + k, n = t + M0 * i;
                if (n < N) {
                    const cplx s = hs_add(y0[k], hs_mulc(hs_mulc(v[i], hs_c(CR[i], CI[i])), w_p));
                    const cplx hv = hs_ld_tab(h + n);
                    hs_st_stream(o + n, (SGN == 1) ? hs_mulc(s, hv) : hs_mul(s, hv));
                }
            }
        }
    }
    hs_tmem_free(tbase, TCOLS);
}

