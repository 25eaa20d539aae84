// hs_butterflies.cuh -- register-resident radix-r butterflies, r in {2,3,4,5,7,8,16} + generic odd r.
//
// v[0..r) holds the r already-twiddled sub-transform values on entry and the r outputs on exit
// (output j = sum_i v[i] * exp(-SGN * 2 pi i * i j / r)).
//
// Parity contract: the reference's radix-3/5/7
// butterflies use 11-digit constants (reference src/highSpeedFFT.c:378,453,530,763,916,1086) and
// its own error against an exact DFT (1.5e-11 at N=3000) is larger than the parity tolerance, so
// these butterflies MUST use the same literals in the same algebraic form.  Radix 2/4/8 use exact
// constants in the reference (:623,:1325); any regrouping of power-of-two factors (radix 16 below)
// therefore stays within rounding of it.  The *_exact variants are for transforms the reference
// itself computes with power-of-two stages only (Bluestein's padded convolution), where using a
// smooth pad length instead must not introduce the truncation error.
#pragma once
#include "hs_common.cuh"

template <int SGN>
HS_HD void hs_bfly2(cplx *v)
{ // reference src/highSpeedFFT.c:714-749
    cplx a = v[0], b = v[1];
    v[0] = hs_add(a, b);
    v[1] = hs_sub(a, b);
}

template <int SGN>
HS_HD void hs_bfly3_k(cplx *v, const double K)
{ // reference src/highSpeedFFT.c:750-803:  t1 = sgn*K*(b-c), t2 = a - 0.5*(b+c)
    cplx a = v[0], b = v[1], c = v[2];
    cplx t0 = hs_add(b, c);
    const double sk = (SGN == 1) ? K : -K;
    cplx t1 = hs_c(sk * (b.x - c.x), sk * (b.y - c.y));
    cplx t2 = hs_c(a.x - t0.x * 0.5, a.y - t0.y * 0.5);
    v[0] = hs_add(a, t0);
    v[1] = hs_c(t2.x + t1.y, t2.y - t1.x);
    v[2] = hs_c(t2.x - t1.y, t2.y + t1.x);
}
template <int SGN>
HS_HD void hs_bfly3(cplx *v) { hs_bfly3_k<SGN>(v, 0.86602540378); }
template <int SGN>
HS_HD void hs_bfly3_exact(cplx *v) { hs_bfly3_k<SGN>(v, 0.86602540378443864676372317075294); }

template <int SGN>
HS_HD void hs_bfly4(cplx *v)
{ // reference src/highSpeedFFT.c:804-900
    cplx a = v[0], b = v[1], c = v[2], d = v[3];
    cplx t0 = hs_add(a, c), t1 = hs_sub(a, c), t2 = hs_add(b, d);
    cplx t3 = (SGN == 1) ? hs_sub(b, d) : hs_sub(d, b);
    v[0] = hs_add(t0, t2);
    v[1] = hs_c(t1.x + t3.y, t1.y - t3.x);
    v[2] = hs_sub(t0, t2);
    v[3] = hs_c(t1.x - t3.y, t1.y + t3.x);
}

template <int SGN>
HS_HD void hs_bfly5_k(cplx *v, const double c1, const double c2, const double s1, const double s2)
{ // reference src/highSpeedFFT.c:901-1070
    cplx a = v[0], b = v[1], c = v[2], d = v[3], e = v[4];
    cplx t0 = hs_add(b, e), t1 = hs_add(c, d), t2 = hs_sub(b, e), t3 = hs_sub(c, d);
    cplx t4, t5, t6;
    v[0] = hs_c(a.x + t0.x + t1.x, a.y + t0.y + t1.y);
    t4 = hs_c(c1 * t0.x + c2 * t1.x, c1 * t0.y + c2 * t1.y);
    t5 = hs_c(s1 * t2.x + s2 * t3.x, s1 * t2.y + s2 * t3.y); // forward value; the inverse is its negation
    t6 = hs_add(a, t4);
    if (SGN == 1) { v[1] = hs_c(t6.x + t5.y, t6.y - t5.x); v[4] = hs_c(t6.x - t5.y, t6.y + t5.x); }
    else          { v[1] = hs_c(t6.x - t5.y, t6.y + t5.x); v[4] = hs_c(t6.x + t5.y, t6.y - t5.x); }
    t4 = hs_c(c2 * t0.x + c1 * t1.x, c2 * t0.y + c1 * t1.y);
    t5 = hs_c(s2 * t2.x - s1 * t3.x, s2 * t2.y - s1 * t3.y);
    t6 = hs_add(a, t4);
    if (SGN == 1) { v[2] = hs_c(t6.x + t5.y, t6.y - t5.x); v[3] = hs_c(t6.x - t5.y, t6.y + t5.x); }
    else          { v[2] = hs_c(t6.x - t5.y, t6.y + t5.x); v[3] = hs_c(t6.x + t5.y, t6.y - t5.x); }
}
template <int SGN>
HS_HD void hs_bfly5(cplx *v) { hs_bfly5_k<SGN>(v, 0.30901699437, -0.80901699437, 0.95105651629, 0.58778525229); }
template <int SGN>
HS_HD void hs_bfly5_exact(cplx *v)
{
    hs_bfly5_k<SGN>(v, 0.30901699437494742410229341718282, -0.80901699437494742410229341718282,
                    0.95105651629515357211643933337938, 0.58778525229247312916870595463907);
}

template <int SGN>
HS_HD void hs_bfly7_k(cplx *v, const double c1, const double c2, const double c3, const double s1, const double s2,
                      const double s3)
{ // reference src/highSpeedFFT.c:1071-1309
    cplx a = v[0];
    cplx t0 = hs_add(v[1], v[6]), t3 = hs_sub(v[1], v[6]);
    cplx t1 = hs_add(v[2], v[5]), t4 = hs_sub(v[2], v[5]);
    cplx t2 = hs_add(v[3], v[4]), t5 = hs_sub(v[3], v[4]);
    cplx u, w; // w is the INVERSE-direction value (s-sums positive); forward uses its negation
    v[0] = hs_c(a.x + t0.x + t1.x + t2.x, a.y + t0.y + t1.y + t2.y);

    u = hs_c(a.x + c1 * t0.x + c2 * t1.x + c3 * t2.x, a.y + c1 * t0.y + c2 * t1.y + c3 * t2.y);
    w = hs_c(s1 * t3.x + s2 * t4.x + s3 * t5.x, s1 * t3.y + s2 * t4.y + s3 * t5.y);
    if (SGN == 1) { v[1] = hs_c(u.x + w.y, u.y - w.x); v[6] = hs_c(u.x - w.y, u.y + w.x); }
    else          { v[1] = hs_c(u.x - w.y, u.y + w.x); v[6] = hs_c(u.x + w.y, u.y - w.x); }

    u = hs_c(a.x + c2 * t0.x + c3 * t1.x + c1 * t2.x, a.y + c2 * t0.y + c3 * t1.y + c1 * t2.y);
    w = hs_c(s2 * t3.x - s3 * t4.x - s1 * t5.x, s2 * t3.y - s3 * t4.y - s1 * t5.y);
    if (SGN == 1) { v[2] = hs_c(u.x + w.y, u.y - w.x); v[5] = hs_c(u.x - w.y, u.y + w.x); }
    else          { v[2] = hs_c(u.x - w.y, u.y + w.x); v[5] = hs_c(u.x + w.y, u.y - w.x); }

    u = hs_c(a.x + c3 * t0.x + c1 * t1.x + c2 * t2.x, a.y + c3 * t0.y + c1 * t1.y + c2 * t2.y);
    w = hs_c(s3 * t3.x - s1 * t4.x + s2 * t5.x, s3 * t3.y - s1 * t4.y + s2 * t5.y);
    if (SGN == 1) { v[3] = hs_c(u.x + w.y, u.y - w.x); v[4] = hs_c(u.x - w.y, u.y + w.x); }
    else          { v[3] = hs_c(u.x - w.y, u.y + w.x); v[4] = hs_c(u.x + w.y, u.y - w.x); }
}
template <int SGN>
HS_HD void hs_bfly7(cplx *v)
{
    hs_bfly7_k<SGN>(v, 0.62348980185, -0.22252093395, -0.9009688679, 0.78183148246, 0.97492791218, 0.43388373911);
}
template <int SGN>
HS_HD void hs_bfly7_exact(cplx *v)
{
    hs_bfly7_k<SGN>(v, 0.62348980185873353052500488400424, -0.22252093395631440428890256449679,
                    -0.90096886790241912623610231950745, 0.78183148246802980870844452667406,
                    0.97492791218182360701813168299393, 0.43388373911755812047576833284836);
}

template <int SGN>
HS_HD void hs_bfly8(cplx *v)
{ // reference src/highSpeedFFT.c:1310-1474 (pairs 0/4, 1/7, 3/5, 2/6); exact constant
    const double c = 0.70710678118654752440084436210485;
    cplx t0 = hs_add(v[0], v[4]), t4 = hs_sub(v[0], v[4]);
    cplx t1 = hs_add(v[1], v[7]), t5 = hs_sub(v[1], v[7]);
    cplx t2 = hs_add(v[3], v[5]), t6 = hs_sub(v[3], v[5]);
    cplx t3 = hs_add(v[2], v[6]), t7 = hs_sub(v[2], v[6]);
    cplx p1 = hs_sub(t1, t2), p2 = hs_add(t5, t6);
    cplx u, w; // w = inverse-direction value
    v[0] = hs_c(t0.x + t1.x + t2.x + t3.x, t0.y + t1.y + t2.y + t3.y);
    v[4] = hs_c(t0.x - t1.x - t2.x + t3.x, t0.y - t1.y - t2.y + t3.y);

    u = hs_c(t4.x + c * p1.x, t4.y + c * p1.y);
    w = hs_c(c * p2.x + t7.x, c * p2.y + t7.y);
    if (SGN == 1) { v[1] = hs_c(u.x + w.y, u.y - w.x); v[7] = hs_c(u.x - w.y, u.y + w.x); }
    else          { v[1] = hs_c(u.x - w.y, u.y + w.x); v[7] = hs_c(u.x + w.y, u.y - w.x); }

    u = hs_sub(t0, t3);
    w = hs_sub(t5, t6);
    if (SGN == 1) { v[2] = hs_c(u.x + w.y, u.y - w.x); v[6] = hs_c(u.x - w.y, u.y + w.x); }
    else          { v[2] = hs_c(u.x - w.y, u.y + w.x); v[6] = hs_c(u.x + w.y, u.y - w.x); }

    u = hs_c(t4.x - c * p1.x, t4.y - c * p1.y);
    w = hs_c(c * p2.x - t7.x, c * p2.y - t7.y);
    if (SGN == 1) { v[3] = hs_c(u.x + w.y, u.y - w.x); v[5] = hs_c(u.x - w.y, u.y + w.x); }
    else          { v[3] = hs_c(u.x - w.y, u.y + w.x); v[5] = hs_c(u.x + w.y, u.y - w.x); }
}

// Radix 16 = 4 x 4 (decimation in time inside the registers); only exact constants, so it may
// replace any run of power-of-two factors of the reference's list (8*2, 4*4, ...) -- section 8c.
template <int SGN>
HS_HD void hs_bfly16(cplx *v)
{
    const double C8 = 0.70710678118654752440084436210485;  // cos(pi/4)
    const double CA = 0.92387953251128675612818318939679;  // cos(pi/8)
    const double SA = 0.38268343236508977172845998403040;  // sin(pi/8)
    cplx y[4][4]; // y[i0][j0] = sum_i1 v[i0 + 4 i1] W4^(i1 j0)
    HS_UNROLL
    for (int i0 = 0; i0 < 4; ++i0) {
        cplx t[4] = {v[i0], v[i0 + 4], v[i0 + 8], v[i0 + 12]};
        hs_bfly4<SGN>(t);
        y[i0][0] = t[0]; y[i0][1] = t[1]; y[i0][2] = t[2]; y[i0][3] = t[3];
    }
    // internal twiddles W16^(i0 j0), forward value (re, -im*...) ; inverse = conjugate
    const cplx w1 = hs_c(CA, -SA * SGN), w2 = hs_c(C8, -C8 * SGN), w3 = hs_c(SA, -CA * SGN);
    const cplx w6 = hs_c(-C8, -C8 * SGN), w9 = hs_c(-CA, SA * SGN);
    y[1][1] = hs_mul(y[1][1], w1); y[1][2] = hs_mul(y[1][2], w2); y[1][3] = hs_mul(y[1][3], w3);
    y[2][1] = hs_mul(y[2][1], w2); y[2][2] = hs_rot<SGN>(y[2][2]);  y[2][3] = hs_mul(y[2][3], w6);
    y[3][1] = hs_mul(y[3][1], w3); y[3][2] = hs_mul(y[3][2], w6); y[3][3] = hs_mul(y[3][3], w9);
    HS_UNROLL
    for (int j0 = 0; j0 < 4; ++j0) {
        cplx t[4] = {y[0][j0], y[1][j0], y[2][j0], y[3][j0]};
        hs_bfly4<SGN>(t);
        v[j0] = t[0]; v[j0 + 4] = t[1]; v[j0 + 8] = t[2]; v[j0 + 12] = t[3];
    }
}

// Generic odd radix (11, 13, 17, ... 53 and whatever factors() leaves above 31), O(r^2);
// reference src/highSpeedFFT.c:1475-1628.  cs/sn are the half tables cos/sin(i*PI2/r), i=1..(r-1)/2,
// extended by the mirror rule of :1539-1543; they are built on the host with the reference's
// expression and live in HBM.
#ifndef HS_MAX_GENERIC_RADIX
#define HS_MAX_GENERIC_RADIX 128
#endif
template <int SGN>
HS_DEVICE void hs_bfly_generic(cplx *v, int r, const double *cs, const double *sn)
{
    const int h = (r - 1) / 2;
    cplx tp[HS_MAX_GENERIC_RADIX / 2], tm[HS_MAX_GENERIC_RADIX / 2];
    cplx y0 = v[0];
    cplx acc = y0;
    for (int i = 0; i < h; ++i) {
        tp[i] = hs_add(v[i + 1], v[r - 1 - i]);
        tm[i] = hs_sub(v[i + 1], v[r - 1 - i]);
        acc = hs_add(acc, tp[i]);
    }
    v[0] = acc;
    for (int u = 0; u < h; ++u) {
        cplx p = y0, q = hs_c(0.0, 0.0);
        for (int w = 0; w < h; ++w) {
            int t = ((u + 1) * (w + 1)) % r;
            double c = cs[t - 1], s = sn[t - 1];
            p.x += c * tp[w].x; p.y += c * tp[w].y;
            q.x -= s * tm[w].x; q.y -= s * tm[w].y;
        }
        if (SGN != 1) { q.x = -q.x; q.y = -q.y; }
        v[u + 1] = hs_c(p.x - q.y, p.y + q.x);
        v[r - u - 1] = hs_c(p.x + q.y, p.y - q.x);
    }
}


// Odd prime radices 11, 13, 17 with compile-time tables: the reference's generic butterfly (src/highSpeedFFT.c:1475-1628: sums and
// differences of mirrored inputs, then one cosine and one sine sum per output pair) fully unrolled, so that lengths with
// these factors get plan-time kernels like every other length (hs_jit.cu).  The reference evaluates cos / sin(i*PI2/r) in double
// precision; the literals below are those values (no truncation here: only the radix-3/5/7 butterflies carry 11-digit constants).
template <int R> struct HsOddTab;
template <> struct HsOddTab<11> {
    static HS_HD double cs(int i) { constexpr double t[10] = {8.41253532831181168863e-01, 4.15415013001886425489e-01, -1.42314838273285140460e-01, -6.54860733945285064127e-01, -9.59492973614497389901e-01, -9.59492973614497389901e-01, -6.54860733945285064072e-01, -1.42314838273285140243e-01, 4.15415013001886425679e-01, 8.41253532831181168917e-01}; return t[i]; }
    static HS_HD double sn(int i) { constexpr double t[10] = {5.40640817455597582101e-01, 9.09631995354518371418e-01, 9.89821441880932732359e-01, 7.55749574354258283704e-01, 2.81732556841429697652e-01, -2.81732556841429697734e-01, -7.55749574354258283758e-01, -9.89821441880932732414e-01, -9.09631995354518371364e-01, -5.40640817455597581993e-01}; return t[i]; }
};
template <> struct HsOddTab<13> {
    static HS_HD double cs(int i) { constexpr double t[12] = {8.85456025653209895872e-01, 5.68064746731155802487e-01, 1.20536680255323053250e-01, -3.54604887042535626003e-01, -7.48510748171101098684e-01, -9.70941817426052027193e-01, -9.70941817426052027193e-01, -7.48510748171101098630e-01, -3.54604887042535626003e-01, 1.20536680255323053460e-01, 5.68064746731155802378e-01, 8.85456025653209896089e-01}; return t[i]; }
    static HS_HD double sn(int i) { constexpr double t[12] = {4.64723172043768545663e-01, 8.22983865893656394580e-01, 9.92708874098053992835e-01, 9.35016242685414823450e-01, 6.63122658240795202322e-01, 2.39315664287557766952e-01, -2.39315664287557767061e-01, -6.63122658240795202430e-01, -9.35016242685414823450e-01, -9.92708874098053992781e-01, -8.22983865893656394688e-01, -4.64723172043768545310e-01}; return t[i]; }
};
template <> struct HsOddTab<17> {
    static HS_HD double cs(int i) { constexpr double t[16] = {9.32472229404355804593e-01, 7.39008917220659115909e-01, 4.45738355776538267358e-01, 9.22683594633019952527e-02, -2.73662990072082863561e-01, -6.02634636379256389269e-01, -8.50217135729614152131e-01, -9.82973099683901778260e-01, -9.82973099683901778260e-01, -8.50217135729614152076e-01, -6.02634636379256389161e-01, -2.73662990072082863371e-01, 9.22683594633019951375e-02, 4.45738355776538267439e-01, 7.39008917220659116072e-01, 9.32472229404355804538e-01}; return t[i]; }
    static HS_HD double sn(int i) { constexpr double t[16] = {3.61241666187152948744e-01, 6.73695643646557211731e-01, 8.95163291355062322062e-01, 9.95734176295034521895e-01, 9.61825643172819070399e-01, 7.98017227280239503260e-01, 5.26432162877355800247e-01, 1.83749517816570331603e-01, -1.83749517816570331698e-01, -5.26432162877355800301e-01, -7.98017227280239503315e-01, -9.61825643172819070453e-01, -9.95734176295034521895e-01, -8.95163291355062322062e-01, -6.73695643646557211569e-01, -3.61241666187152948798e-01}; return t[i]; }
};
template <int R, int SGN>
HS_HD void hs_bfly_oddct(cplx *v)
{
    constexpr int H = (R - 1) / 2;
    cplx tp[H], tm[H];
    const cplx y0 = v[0];
    cplx acc = y0;
    HS_UNROLL
    for (int i = 0; i < H; ++i) {
        tp[i] = hs_add(v[i + 1], v[R - 1 - i]);
        tm[i] = hs_sub(v[i + 1], v[R - 1 - i]);
        acc = hs_add(acc, tp[i]);
    }
    v[0] = acc;
    HS_UNROLL
    for (int u = 0; u < H; ++u) {
        cplx p = y0, q = hs_c(0.0, 0.0);
        HS_UNROLL
        for (int w = 0; w < H; ++w) {
            const int t = ((u + 1) * (w + 1)) % R;
            const double c = HsOddTab<R>::cs(t - 1), s = HsOddTab<R>::sn(t - 1);
            p.x += c * tp[w].x; p.y += c * tp[w].y;
            q.x -= s * tm[w].x; q.y -= s * tm[w].y;
        }
        if (SGN != 1) { q.x = -q.x; q.y = -q.y; }
        v[u + 1] = hs_c(p.x - q.y, p.y + q.x);
        v[R - u - 1] = hs_c(p.x + q.y, p.y - q.x);
    }
}

// compile-time radix dispatch; EXACT selects the full-precision radix-3/5/7 constants (Bluestein-internal
// transforms only -- a transform the reference itself runs must keep the reference's literals)
template <int R, int SGN, bool EXACT = false>
HS_HD void hs_bfly(cplx *v)
{
    if (R == 2) hs_bfly2<SGN>(v);
    else if (R == 3) { if (EXACT) hs_bfly3_exact<SGN>(v); else hs_bfly3<SGN>(v); }
    else if (R == 4) hs_bfly4<SGN>(v);
    else if (R == 5) { if (EXACT) hs_bfly5_exact<SGN>(v); else hs_bfly5<SGN>(v); }
    else if (R == 7) { if (EXACT) hs_bfly7_exact<SGN>(v); else hs_bfly7<SGN>(v); }
    else if (R == 8) hs_bfly8<SGN>(v);
    else if (R == 16) hs_bfly16<SGN>(v);
    if constexpr (R == 11 || R == 13 || R == 17) hs_bfly_oddct<R, SGN>(v);
}
