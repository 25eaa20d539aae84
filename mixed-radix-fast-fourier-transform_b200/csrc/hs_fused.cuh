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
