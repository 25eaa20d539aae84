// hs_fused.cuh -- fused kernels built on the thread-team transform of hs_smem_fft.cuh.
//
// Bluestein chirp-z (reference bluestein_fft, src/highSpeedFFT.c:1735-1907, and bluestein_exp,
// :1645-1704).  The reference does, per call: build the chirp h (N sincos), FFT_M(h/M), x*conj(h),
// zero-pad, FFT_M, pointwise product, conjugate all twiddles, inverse FFT_M, multiply by conj(h),
// with five malloc'd M-point temporaries.  Here h and Hk = FFT_M(h_padded/M) are plan-time tables
// resident in HBM (built once, by this same engine), and one kernel does
//     load N -> *conj(h) -> zero-pad (registers) -> FFT_M -> *Hk (registers) -> IFFT_M -> *conj(h) -> store N
// with ONE HBM read and ONE HBM write per transform.  The forward transform's last stage and the
// inverse transform's first stage use the same radix and the same butterfly-to-thread map, so the
// M-point spectrum never leaves the registers between them (no shared-memory exchange there).
//
// For the inverse plan direction (sgn = -1) the chirp roles flip exactly as in the reference
// (:1811-1815, :1846-1854, :1879-1885): x*h, *conj(Hk), *h.
#pragma once
#include "hs_smem_fft.cuh"

template <int SGN>
struct HsLoadChirp {
    static constexpr bool kPost = false;
    const cplx *HS_RESTRICT x;
    const cplx *HS_RESTRICT h;
    int N;
    HS_DEVICE cplx operator()(int, int, int n) const
    {
        if (n >= N) return hs_c(0.0, 0.0);
        const cplx xv = hs_ld_stream(x + n), hv = hs_ld_tab(h + n);
        return (SGN == 1) ? hs_mulc(xv, hv) : hs_mul(xv, hv);
    }
};
template <int SGN>
struct HsStoreChirpCrop {
    cplx *HS_RESTRICT out;
    const cplx *HS_RESTRICT h;
    int N;
    HS_DEVICE void operator()(int, int, int k, cplx v) const
    {
        if (k < N) {
            const cplx hv = hs_ld_tab(h + k);
            hs_st_stream(out + k, (SGN == 1) ? hs_mulc(v, hv) : hs_mul(v, hv));
        }
    }
};
template <int SGN, int RL>
struct HsStoreRegsMul {
    cplx (*y)[RL];
    const cplx *HS_RESTRICT hk;
    HS_DEVICE void operator()(int u, int j, int k, cplx v) const
    {
        const cplx hv = hs_ld_tab(hk + k);
        y[u][j] = (SGN == 1) ? hs_mul(v, hv) : hs_mulc(v, hv);
    }
};
template <int RL>
struct HsLoadRegs {
    static constexpr bool kPost = false;
    cplx (*y)[RL];
    HS_DEVICE cplx operator()(int u, int i, int) const { return y[u][i]; }
};

// PF: forward plan of length M; PI: inverse plan of length M with PI::radix(0) == PF::radix(last)
// and the same T / X / padding.  twf: PF's stage twiddles for exp(-); twi: PI's for exp(+).
template <class PF, class PI, int SGN, int MINB>
HS_GLOBAL void HS_LAUNCH_BOUNDS(PF::threads, MINB)
hs_bluestein_kernel(const cplx *HS_RESTRICT in, cplx *HS_RESTRICT out, const cplx *HS_RESTRICT twf, const cplx *HS_RESTRICT twi,
                    const cplx *HS_RESTRICT chirp, const cplx *HS_RESTRICT hk, int N, long long howmany, int pf_dist)
{
    static_assert(PF::N == PI::N && PF::T == PI::T && PF::X == PI::X && PF::NPAD == PI::NPAD, "plans must pair");
    constexpr int RL = PF::radix(PF::NST - 1);
    static_assert(PI::radix(0) == RL, "register hand-over needs matching radices");
    constexpr int NB = PF::N / RL, B = (NB + PF::T - 1) / PF::T;
    HS_DYN_SMEM(cplx, smem);
    const int team = HS_TID / PF::T, t = HS_TID % PF::T;
    const long long xf = HS_BID * PF::X + team;
    const bool active = xf < howmany;
    cplx *sm = smem + team * PF::NPAD;
    cplx y[B][RL];
    if (pf_dist > 0 && HS_TID == 0) { // warm the L2 for the CTA that takes this one's place
        const long long pf = (HS_BID + pf_dist) * PF::X, left = howmany - pf;
        if (left > 0) hs_prefetch_l2(in + pf * N, (unsigned)((left < PF::X ? left : PF::X) * N * 16));
    }

    HsLoadChirp<SGN> ld{in + (active ? xf : 0) * N, chirp, N};
    HsStoreRegsMul<SGN, RL> st{y, hk};
    hs_fft_team<PF, 1>(ld, st, sm, twf, t, active);
    HS_SYNC(); // the forward transform's last reads of sm precede the inverse transform's first writes
    HsLoadRegs<RL> ld2{y};
    HsStoreChirpCrop<SGN> st2{out + (active ? xf : 0) * N, chirp, N};
    hs_fft_team<PI, -1>(ld2, st2, sm, twi, t, active);
}

// ---------------------------------------------------------------------------------------------------------------
// Split chirp-z: lengths whose padded convolution (M >= 2N-1 points) does not fit one SM, but whose N points do.
// With M = 2 M' and M' >= N the zero-padded input occupies only the first half of the M-point buffer, and only the
// first N <= M' outputs are wanted, so one radix-2 step of the M-point convolution is trivial on both ends
// (reference bluestein_fft, src/highSpeedFFT.c:1735-1907, computes the same y = IFFT_M(FFT_M(a) * Hk)):
//     A[2k']   = FFT_M'(a)[k']                 A[2k'+1] = FFT_M'(a * w)[k'],      w[n] = exp(-2 pi i n / M)
//     y[n]     = IFFT_M'(A_even * Hk_even)[n] + conj(w[n]) * IFFT_M'(A_odd * Hk_odd)[n],     n < M'
// i.e. two cyclic convolutions of length M', each of which one CTA holds in shared memory (M' = 10240: 174 KB).
// One kernel, one CTA per transform, does both halves back to back:
//     half 0: load x * conj(h) -> FFT_M' -> * Hk_even (registers) -> IFFT_M' -> park y0 in the output row
//     half 1: load x * conj(h w~) -> FFT_M' -> * Hk_odd -> IFFT_M' -> out = y0 * conj(h) + y1 * conj(h w)
// HBM sees x once and the result once (the second read of x and the parked y0 are served by the L2: the same
// thread that parked y0[n] reads it back).  The w factors are folded into two more chirp tables,
// gA = h * conj(w) and gB = h * w, so each half costs exactly what the one-kernel chirp-z costs per point.
// For the inverse plan direction the chirp roles flip as in the reference (:1811-1815, :1846-1854, :1879-1885).
template <int SGN>
struct HsStoreParkCrop { // half 0: keep y0[n], n < N, where half 1 finds it again
    cplx *HS_RESTRICT out;
    int N;
    HS_DEVICE void operator()(int, int, int k, cplx v) const
    {
        if (k < N) hs_st_keep(out + k, v);
    }
};
template <int SGN>
struct HsStoreCombineChirpCrop { // half 1: out[n] = y0[n] * c0[n] + y1[n] * c1[n]  (conjugated tables for SGN = +1)
    cplx *HS_RESTRICT out;
    const cplx *HS_RESTRICT c0, *HS_RESTRICT c1;
    int N;
    HS_DEVICE void operator()(int, int, int k, cplx v) const
    {
        if (k < N) {
            const cplx y0 = hs_ld_coherent(out + k), h0 = hs_ld_tab(c0 + k), h1 = hs_ld_tab(c1 + k);
            const cplx a = (SGN == 1) ? hs_mulc(y0, h0) : hs_mul(y0, h0), b = (SGN == 1) ? hs_mulc(v, h1) : hs_mul(v, h1);
            hs_st_stream(out + k, hs_add(a, b));
        }
    }
};

// chirp3: the three tables h | gA | gB, N entries each; hk2: Hk_even | Hk_odd, PF::N entries each
template <class PF, class PI, int SGN, int MINB>
HS_GLOBAL void HS_LAUNCH_BOUNDS(PF::threads, MINB)
hs_bluestein_split_kernel(const cplx *HS_RESTRICT in, cplx *HS_RESTRICT out, const cplx *HS_RESTRICT twf, const cplx *HS_RESTRICT twi,
                          const cplx *HS_RESTRICT chirp3, const cplx *HS_RESTRICT hk2, int N, long long howmany, int pf_dist)
{
    static_assert(PF::N == PI::N && PF::T == PI::T && PF::X == 1 && PI::X == 1 && PF::NPAD == PI::NPAD, "plans must pair");
    constexpr int RL = PF::radix(PF::NST - 1);
    static_assert(PI::radix(0) == RL, "register hand-over needs matching radices");
    constexpr int NB = PF::N / RL, B = (NB + PF::T - 1) / PF::T;
    HS_DYN_SMEM(cplx, sm);
    const int t = HS_TID;
    const long long xf = HS_BID;
    cplx y[B][RL];
    if (pf_dist > 0 && t == 0 && xf + pf_dist < howmany) hs_prefetch_l2(in + (xf + pf_dist) * N, (unsigned)(N * 16));
    const cplx *x = in + xf * N;
    cplx *o = out + xf * N;
    const cplx *h = chirp3, *gA = chirp3 + N, *gB = chirp3 + 2 * (long long)N;
    {
        HsLoadChirp<SGN> ld{x, h, N};
        HsStoreRegsMul<SGN, RL> st{y, hk2};
        hs_fft_team<PF, 1>(ld, st, sm, twf, t, true);
        HS_SYNC();
        HsLoadRegs<RL> ld2{y};
        HsStoreParkCrop<SGN> st2{o, N};
        hs_fft_team<PI, -1>(ld2, st2, sm, twi, t, true);
    }
    HS_SYNC(); // the inverse transform's last reads of sm precede the next forward transform's first writes
    {
        HsLoadChirp<SGN> ld{x, SGN == 1 ? gA : gB, N};
        HsStoreRegsMul<SGN, RL> st{y, hk2 + PF::N};
        hs_fft_team<PF, 1>(ld, st, sm, twf, t, true);
        HS_SYNC();
        HsLoadRegs<RL> ld2{y};
        HsStoreCombineChirpCrop<SGN> st2{o, h, SGN == 1 ? gB : gA, N};
        hs_fft_team<PI, -1>(ld2, st2, sm, twi, t, true);
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Rader's algorithm for a prime length N whose N-1 has a shared-memory plan (N = 1009: N-1 = 1008 = 16*9*7).  The
// reference sends every prime to Bluestein with a power-of-two pad M >= 2N-1 (two 2048-point transforms for N = 1009);
// for such a prime the DFT is also a CYCLIC convolution of length L = N-1 exactly, with no padding:
//     X[0]      = sum_n x[n]
//     X[g^-p]   = x[0] + sum_q x[g^q] * W^(g^(q-p))   = x[0] + (a (*) b)[p],    a[q] = x[g^q],  b[m] = W^(g^-m)
// (g a primitive root mod N, W = exp(-sgn 2 pi i / N)).  So one kernel does
//     gather a through perm[q] = g^q mod N -> FFT_L -> * Bk (+ x[0] on the DC bin) -> IFFT_L -> scatter to perm[L - p]
// with Bk = FFT_L(b) / L a plan-time table.  Two facts keep it to the Bluestein kernel's shape: A[0] = sum_{n>0} x[n],
// so X[0] = x[0] + A[0] is written by the thread that holds bin 0; and adding x[0] to bin 0 of the product adds x[0] to
// every output of the unnormalised inverse transform.  Half the arithmetic of the padded chirp-z, same single HBM
// read and write; the result is the same DFT to rounding (full-precision radix-3/7 constants: HsExact plans).
struct HsLoadRader {
    static constexpr bool kPost = false;
    const cplx *HS_RESTRICT x;
    const int *HS_RESTRICT perm;
    HS_DEVICE cplx operator()(int, int, int n) const { return hs_ld_gather(x + hs_ld_idx(perm + n)); }
};
// staged forms: the outputs are scattered into the team's buffer and leave it with coalesced stores (measured faster on
// B200: the default); the inputs can likewise be copied in with coalesced loads and gathered from shared memory, but the
// barrier that needs before the first pass costs more than the uncoalesced gather (measured slower: kept for the
// emulation tests and A/B builds, not instantiated by hs_device.cu)
struct HsLoadRaderSm {
    static constexpr bool kPost = false, kFromSmem = true;
    const cplx *sm;
    const int *HS_RESTRICT perm;
    HS_DEVICE cplx operator()(int, int, int n) const { return sm[hs_ld_idx(perm + n)]; }
};
struct HsStoreRaderSm {
    static constexpr bool kToSmem = true;
    cplx *sm;
    const int *HS_RESTRICT perm;
    int L;
    HS_DEVICE void operator()(int, int, int p, cplx v) const { sm[hs_ld_idx(perm + (p ? L - p : 0))] = v; }
};
template <int RL>
struct HsStoreRegsMulRader {
    cplx (*y)[RL];
    const cplx *HS_RESTRICT bk;
    const cplx *HS_RESTRICT x;
    cplx *HS_RESTRICT out;
    HS_DEVICE void operator()(int u, int j, int k, cplx v) const
    {
        cplx r = hs_mul(v, hs_ld_tab(bk + k));
        if (k == 0) {
            const cplx x0 = hs_ld_gather(x);
            hs_st_stream(out, hs_c(x0.x + v.x, x0.y + v.y));
            r.x += x0.x;
            r.y += x0.y;
        }
        y[u][j] = r;
    }
};
struct HsStoreRader {
    cplx *HS_RESTRICT out;
    const int *HS_RESTRICT perm;
    int L;
    HS_DEVICE void operator()(int, int, int p, cplx v) const { hs_st_stream(out + hs_ld_idx(perm + (p ? L - p : 0)), v); }
};
template <class P>
struct HsPassPoints {
    static constexpr int first = P::pass_size(0) == 2 ? P::radix(0) * P::radix(1) : P::radix(0);
    static constexpr int last = P::pass_size(P::NPASS - 1) == 2 ? P::radix(P::NST - 1) * P::radix(P::NST - 2) : P::radix(P::NST - 1);
};

template <class PF, class PI>
struct HsRaderSmem {
    static constexpr int npad0 = PF::NPAD > PI::NPAD ? PF::NPAD : PI::NPAD;
    static constexpr int npad = npad0 > PF::N + 1 ? npad0 : PF::N + 2; // room for the N = L+1 staged outputs
    static constexpr int bytes = npad * PF::X * 16;
};

// PF / PI: forward / inverse plans of length L = N-1, PI's first pass as wide as PF's last (register hand-over).
// bk already carries the plan direction, so one instantiation serves both.  STG bit 0: gather through the shared-memory
// buffer, bit 1: scatter through it.
template <class PF, class PI, int MINB, int STG>
HS_GLOBAL void HS_LAUNCH_BOUNDS(PF::threads, MINB)
hs_rader_kernel(const cplx *HS_RESTRICT in, cplx *HS_RESTRICT out, const cplx *HS_RESTRICT twf, const cplx *HS_RESTRICT twi,
                const int *HS_RESTRICT perm, const cplx *HS_RESTRICT bk, long long howmany, int pf_dist)
{
    static_assert(PF::N == PI::N && PF::T == PI::T && PF::X == PI::X, "plans must pair");
    constexpr int RL = HsPassPoints<PF>::last;
    static_assert(HsPassPoints<PI>::first == RL, "register hand-over needs passes of the same width");
    constexpr int L = PF::N, N = L + 1;
    constexpr int NB = L / RL, B = (NB + PF::T - 1) / PF::T;
    HS_DYN_SMEM(cplx, smem);
    const int team = HS_TID / PF::T, t = HS_TID % PF::T;
    const long long xf = HS_BID * PF::X + team;
    const bool active = xf < howmany;
    cplx *sm = smem + team * HsRaderSmem<PF, PI>::npad;
    cplx y[B][RL];
    if (pf_dist > 0 && HS_TID == 0) {
        const long long pf = (HS_BID + pf_dist) * PF::X, left = howmany - pf;
        if (left > 0) hs_prefetch_l2(in + pf * N, (unsigned)((left < PF::X ? left : PF::X) * N * 16));
    }
    const cplx *x = in + (active ? xf : 0) * N;
    cplx *o = out + (active ? xf : 0) * N;
    static_assert(HsRaderSmem<PF, PI>::npad >= N, "the staged row must fit the team's buffer");
    HsStoreRegsMulRader<RL> st{y, bk, x, o};
    if constexpr (STG & 1) {
        if (active)
            for (int i = t; i < N; i += PF::T) sm[i] = hs_ld_stream(x + i);
        HS_SYNC();
        HsLoadRaderSm ld{sm, perm};
        hs_fft_team<PF, 1>(ld, st, sm, twf, t, active);
    } else {
        HsLoadRader ld{x, perm};
        hs_fft_team<PF, 1>(ld, st, sm, twf, t, active);
    }
    HS_SYNC();
    HsLoadRegs<RL> ld2{y};
    if constexpr (STG & 2) {
        HsStoreRaderSm st2{sm, perm, L};
        hs_fft_team<PI, -1>(ld2, st2, sm, twi, t, active);
        HS_SYNC();
        if (active)
            for (int i = 1 + t; i < N; i += PF::T) hs_st_stream(o + i, sm[i]);
    } else {
        HsStoreRader st2{o, perm, L};
        hs_fft_team<PI, -1>(ld2, st2, sm, twi, t, active);
    }
}
