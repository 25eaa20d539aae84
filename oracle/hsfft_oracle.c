/*
 * hsfft_oracle.c -- CPU restatement of the reference FFT hot path (plain C).
 *
 * TEST INFRASTRUCTURE ONLY (see hsfft_oracle.h).  Never linked into the product.
 *
 * What is restated, and from where (paths relative to /root/reference):
 *   plan routing / factorisation      src/highSpeedFFT.c:13,18-55,1979-2025,2038-2163
 *   stage twiddle table               src/highSpeedFFT.c:2238-2313 (the dynamic branch
 *                                     :2273-2289 -- the table branch :2256-2272 is the
 *                                     reference's bug 3)
 *   recursive DIT + butterflies       src/highSpeedFFT.c:318-1629
 *   Bluestein chirp-z                 src/highSpeedFFT.c:1645-1704, :1735-1907
 *   r2c / c2r                         src/real.c:26-64, :78-136, :150-193
 *   FFT convolution                   src/convolve.c:20-25, :39-55, :74-214
 *
 * Restatement policy ("repaired" semantics):
 *   - the two-point leaf reads its INPUT (the reference reads the output buffer
 *     before writing it at :354 -- bug 2);
 *   - stage twiddles always come from the cos/sin expression (bug 3);
 *   - Bluestein allocates for the pad length it actually uses (bug 4).
 *   Wherever the unmodified reference is well defined this code follows the same
 *   operation order, so results agree to the last few ulps (tests pin 1e-15).
 *
 * The radix-3/5/7 butterflies use the reference's 11-digit literals
 * (:378,:453,:530,:763,:916,:1086) on purpose: they are part of the contract.
 */
#include "hsfft_oracle.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ORC_PI2 6.28318530717958647692528676655900577 /* src/highspeedFFT.h:13 */

/* ------------------------------------------------------------------ routing */

/* The seventeen radices the router accepts (src/highSpeedFFT.c:13) -- 19 is absent. */
static const int k_router_radices[17] = {2, 3, 4, 5, 7, 8, 11, 13, 17, 23, 29, 31, 37, 41, 43, 47, 53};

int orc_dividebyN(int n)
{
    /* src/highSpeedFFT.c:1979-2025 + lookup constructor :18-55.  Both the table
     * (n < 1024) and the trial-division tail answer the same question: does n
     * reduce to 1 when divided by the accepted radices?  n == 1 -> yes. */
    if (n < 1) return 0;
    int r = n;
    for (int i = 16; i >= 0; --i)
        while (r % k_router_radices[i] == 0) r /= k_router_radices[i];
    return r == 1;
}

int orc_factors(int n, int *out)
{
    /* src/highSpeedFFT.c:2038-2163: descending 53..11 (19 included here), then
     * 8,7,5,4,3,2, then a 6k+-1 sweep for whatever is left above 31. */
    static const int order[18] = {53, 47, 43, 41, 37, 31, 29, 23, 19, 17, 13, 11, 8, 7, 5, 4, 3, 2};
    int cnt = 0, t = n;
    for (int i = 0; i < 18; ++i)
        while (t % order[i] == 0) { out[cnt++] = order[i]; t /= order[i]; }
    if (t > 31) {
        for (int k = 2; t > 1; ++k) {
            int f1 = 6 * k - 1, f2 = 6 * k + 1;
            while (t % f1 == 0) { out[cnt++] = f1; t /= f1; }
            while (t % f2 == 0) { out[cnt++] = f2; t /= f2; }
        }
    }
    return cnt;
}

void orc_stage_twiddles(orc_cplx *tw, int n, const int *fac, int lf)
{
    /* src/highSpeedFFT.c:2238-2313, dynamic branch.  Factors are walked last to
     * first; the stage of combine length L = r*Ls contributes (r-1)*Ls entries
     * exp(-2 pi i (k+1) j / L), j-major.  Offsets telescope to n-1 entries. */
    int L = 1, c = 0;
    for (int s = 0; s < lf; ++s) {
        int r = fac[lf - 1 - s];
        int Ls = L;
        L *= r;
        double theta = -ORC_PI2 / L;
        for (int j = 0; j < Ls; ++j)
            for (int k = 0; k < r - 1; ++k) {
                if (c < n - 1) {
                    double ang = (k + 1) * j * theta;
                    tw[c].re = cos(ang);
                    tw[c].im = sin(ang);
                    ++c;
                }
            }
    }
}

int orc_bluestein_pad(int N)
{
    /* src/highSpeedFFT.c:1750-1751 */
    return (int)pow(2.0, ceil(log2((double)(2 * N - 1))));
}

orc_plan *orc_fft_init(int N, int sgn)
{
    /* src/highSpeedFFT.c:206-286 */
    if (N < 2) { fprintf(stderr, "orc_fft_init: N=%d unsupported\n", N); exit(EXIT_FAILURE); }
    orc_plan *p = (orc_plan *)calloc(1, sizeof(orc_plan));
    if (!p) return NULL;
    p->N = N;
    p->sgn = sgn;
    if (orc_dividebyN(N)) { p->lt = 0; p->len = N; }
    else                  { p->lt = 1; p->len = orc_bluestein_pad(N); } /* bug-4 repair */
    p->lf = orc_factors(p->len, p->factors);
    p->tw = (orc_cplx *)calloc((size_t)p->len + 1, sizeof(orc_cplx));
    if (!p->tw) { free(p); return NULL; }
    orc_stage_twiddles(p->tw, p->len, p->factors, p->lf);
    if (sgn == -1) /* :277-283 */
        for (int i = 0; i < p->len; ++i) p->tw[i].im = -p->tw[i].im;
    return p;
}

void orc_fft_free(orc_plan *p)
{
    if (!p) return;
    free(p->tw);
    free(p);
}

/* -------------------------------------------------------------- butterflies */
/* Each butterfly works on a local array v[0..r) (v[i] = i-th sub-transform value,
 * twiddle already applied) and leaves the r outputs in v.  `leaf` only selects the
 * association the reference uses for output 0 of radix 5/7 at the recursion
 * bottom (`x += t0 + t1` at :517, :611 versus `a + t0 + t1` at :950, :1130). */

static void bfly2(orc_cplx *v)
{ /* :344-363 (leaf, repaired), :714-749 */
    orc_cplx a = v[0], b = v[1];
    v[0].re = a.re + b.re; v[0].im = a.im + b.im;
    v[1].re = a.re - b.re; v[1].im = a.im - b.im;
}

static void bfly3(orc_cplx *v, int sgn)
{ /* :365-398, :750-803 */
    const double K = 0.86602540378;
    orc_cplx a = v[0], b = v[1], c = v[2];
    double t0r = b.re + c.re, t0i = b.im + c.im;
    double t1r = sgn * K * (b.re - c.re), t1i = sgn * K * (b.im - c.im);
    double t2r = a.re - t0r * 0.5, t2i = a.im - t0i * 0.5;
    v[0].re = a.re + t0r; v[0].im = a.im + t0i;
    v[1].re = t2r + t1i;  v[1].im = t2i - t1r;
    v[2].re = t2r - t1i;  v[2].im = t2i + t1r;
}

static void bfly4(orc_cplx *v, int sgn)
{ /* :399-439, :804-900 */
    orc_cplx a = v[0], b = v[1], c = v[2], d = v[3];
    double t0r = a.re + c.re, t0i = a.im + c.im;
    double t1r = a.re - c.re, t1i = a.im - c.im;
    double t2r = b.re + d.re, t2i = b.im + d.im;
    double t3r = sgn * (b.re - d.re), t3i = sgn * (b.im - d.im);
    v[0].re = t0r + t2r; v[0].im = t0i + t2i;
    v[1].re = t1r + t3i; v[1].im = t1i - t3r;
    v[2].re = t0r - t2r; v[2].im = t0i - t2i;
    v[3].re = t1r - t3i; v[3].im = t1i + t3r;
}

static void bfly5(orc_cplx *v, int sgn, int leaf)
{ /* :440-518, :901-1070 */
    const double c1 = 0.30901699437, c2 = -0.80901699437, s1 = 0.95105651629, s2 = 0.58778525229;
    orc_cplx a = v[0], b = v[1], c = v[2], d = v[3], e = v[4];
    double t0r = b.re + e.re, t0i = b.im + e.im;
    double t1r = c.re + d.re, t1i = c.im + d.im;
    double t2r = b.re - e.re, t2i = b.im - e.im;
    double t3r = c.re - d.re, t3i = c.im - d.im;
    double t4r, t4i, t5r, t5i, t6r, t6i;

    t4r = c1 * t0r + c2 * t1r; t4i = c1 * t0i + c2 * t1i;
    if (sgn == 1) { t5r = s1 * t2r + s2 * t3r;  t5i = s1 * t2i + s2 * t3i; }
    else          { t5r = -s1 * t2r - s2 * t3r; t5i = -s1 * t2i - s2 * t3i; }
    t6r = a.re + t4r; t6i = a.im + t4i;
    v[1].re = t6r + t5i; v[1].im = t6i - t5r;
    v[4].re = t6r - t5i; v[4].im = t6i + t5r;

    t4r = c2 * t0r + c1 * t1r; t4i = c2 * t0i + c1 * t1i;
    if (sgn == 1) { t5r = s2 * t2r - s1 * t3r;  t5i = s2 * t2i - s1 * t3i; }
    else          { t5r = -s2 * t2r + s1 * t3r; t5i = -s2 * t2i + s1 * t3i; }
    t6r = a.re + t4r; t6i = a.im + t4i;
    v[2].re = t6r + t5i; v[2].im = t6i - t5r;
    v[3].re = t6r - t5i; v[3].im = t6i + t5r;

    if (leaf) { v[0].re = a.re + (t0r + t1r); v[0].im = a.im + (t0i + t1i); }
    else      { v[0].re = a.re + t0r + t1r;   v[0].im = a.im + t0i + t1i; }
}

static void bfly7(orc_cplx *v, int sgn, int leaf)
{ /* :519-612, :1071-1309 */
    const double c1 = 0.62348980185, c2 = -0.22252093395, c3 = -0.9009688679;
    const double s1 = 0.78183148246, s2 = 0.97492791218, s3 = 0.43388373911;
    orc_cplx a = v[0];
    double t0r = v[1].re + v[6].re, t3r = v[1].re - v[6].re, t0i = v[1].im + v[6].im, t3i = v[1].im - v[6].im;
    double t1r = v[2].re + v[5].re, t4r = v[2].re - v[5].re, t1i = v[2].im + v[5].im, t4i = v[2].im - v[5].im;
    double t2r = v[3].re + v[4].re, t5r = v[3].re - v[4].re, t2i = v[3].im + v[4].im, t5i = v[3].im - v[4].im;
    double ur, ui, wr, wi;

    ur = a.re + c1 * t0r + c2 * t1r + c3 * t2r; ui = a.im + c1 * t0i + c2 * t1i + c3 * t2i;
    if (sgn == 1) { wr = -s1 * t3r - s2 * t4r - s3 * t5r; wi = -s1 * t3i - s2 * t4i - s3 * t5i; }
    else          { wr = s1 * t3r + s2 * t4r + s3 * t5r;  wi = s1 * t3i + s2 * t4i + s3 * t5i; }
    v[1].re = ur - wi; v[6].re = ur + wi; v[1].im = ui + wr; v[6].im = ui - wr;

    ur = a.re + c2 * t0r + c3 * t1r + c1 * t2r; ui = a.im + c2 * t0i + c3 * t1i + c1 * t2i;
    if (sgn == 1) { wr = -s2 * t3r + s3 * t4r + s1 * t5r; wi = -s2 * t3i + s3 * t4i + s1 * t5i; }
    else          { wr = s2 * t3r - s3 * t4r - s1 * t5r;  wi = s2 * t3i - s3 * t4i - s1 * t5i; }
    v[2].re = ur - wi; v[5].re = ur + wi; v[2].im = ui + wr; v[5].im = ui - wr;

    ur = a.re + c3 * t0r + c1 * t1r + c2 * t2r; ui = a.im + c3 * t0i + c1 * t1i + c2 * t2i;
    if (sgn == 1) { wr = -s3 * t3r + s1 * t4r - s2 * t5r; wi = -s3 * t3i + s1 * t4i - s2 * t5i; }
    else          { wr = s3 * t3r - s1 * t4r + s2 * t5r;  wi = s3 * t3i - s1 * t4i + s2 * t5i; }
    v[3].re = ur - wi; v[4].re = ur + wi; v[3].im = ui + wr; v[4].im = ui - wr;

    if (leaf) { v[0].re = a.re + (t0r + t1r + t2r); v[0].im = a.im + (t0i + t1i + t2i); }
    else      { v[0].re = a.re + t0r + t1r + t2r;   v[0].im = a.im + t0i + t1i + t2i; }
}

static void bfly8(orc_cplx *v, int sgn)
{ /* :613-712, :1310-1474 */
    const double c = 0.70710678118654752440084436210485;
    double t0r = v[0].re + v[4].re, t4r = v[0].re - v[4].re, t0i = v[0].im + v[4].im, t4i = v[0].im - v[4].im;
    double t1r = v[1].re + v[7].re, t5r = v[1].re - v[7].re, t1i = v[1].im + v[7].im, t5i = v[1].im - v[7].im;
    double t2r = v[3].re + v[5].re, t6r = v[3].re - v[5].re, t2i = v[3].im + v[5].im, t6i = v[3].im - v[5].im;
    double t3r = v[2].re + v[6].re, t7r = v[2].re - v[6].re, t3i = v[2].im + v[6].im, t7i = v[2].im - v[6].im;
    double p1r = t1r - t2r, p1i = t1i - t2i, p2r = t5r + t6r, p2i = t5i + t6i;
    double ur, ui, wr, wi;

    v[0].re = t0r + t1r + t2r + t3r; v[0].im = t0i + t1i + t2i + t3i;
    v[4].re = t0r - t1r - t2r + t3r; v[4].im = t0i - t1i - t2i + t3i;

    ur = t4r + c * p1r; ui = t4i + c * p1i;
    if (sgn == 1) { wr = -c * p2r - t7r; wi = -c * p2i - t7i; }
    else          { wr = c * p2r + t7r;  wi = c * p2i + t7i; }
    v[1].re = ur - wi; v[1].im = ui + wr; v[7].re = ur + wi; v[7].im = ui - wr;

    ur = t0r - t3r; ui = t0i - t3i;
    if (sgn == 1) { wr = -t5r + t6r; wi = -t5i + t6i; }
    else          { wr = t5r - t6r;  wi = t5i - t6i; }
    v[2].re = ur - wi; v[2].im = ui + wr; v[6].re = ur + wi; v[6].im = ui - wr;

    ur = t4r - c * p1r; ui = t4i - c * p1i;
    if (sgn == 1) { wr = -c * p2r + t7r; wi = -c * p2i + t7i; }
    else          { wr = c * p2r - t7r;  wi = c * p2i - t7i; }
    v[3].re = ur - wi; v[3].im = ui + wr; v[5].re = ur + wi; v[5].im = ui - wr;
}

static void bfly_generic(orc_cplx *v, int r, int sgn)
{ /* :1475-1628: odd radix r > 8 through the half-length cos/sin tables */
    int h = (r - 1) / 2;
    double cs[64], sn[64], tr[64], ti[64];
    orc_cplx y0 = v[0];
    for (int i = 1; i <= h; ++i) { cs[i - 1] = cos(i * ORC_PI2 / r); sn[i - 1] = sin(i * ORC_PI2 / r); }
    for (int i = 0; i < h; ++i) { sn[i + h] = -sn[h - 1 - i]; cs[i + h] = cs[h - 1 - i]; }
    for (int i = 0; i < h; ++i) {
        tr[i] = v[i + 1].re + v[r - 1 - i].re;     ti[i] = v[i + 1].im + v[r - 1 - i].im;
        tr[i + h] = v[i + 1].re - v[r - 1 - i].re; ti[i + h] = v[i + 1].im - v[r - 1 - i].im;
    }
    double ar = y0.re, ai = y0.im;
    for (int i = 0; i < h; ++i) { ar += tr[i]; ai += ti[i]; }
    orc_cplx o0 = {ar, ai};
    orc_cplx outv[64];
    for (int u = 0; u < h; ++u) {
        double pr = y0.re, pi = y0.im, qr = 0.0, qi = 0.0;
        for (int w = 0; w < h; ++w) {
            int t = (u + 1) * (w + 1);
            while (t >= r) t -= r;
            pr += cs[t - 1] * tr[w];     pi += cs[t - 1] * ti[w];
            qr -= sn[t - 1] * tr[w + h]; qi -= sn[t - 1] * ti[w + h];
        }
        qr = sgn * qr; qi = sgn * qi;
        outv[u + 1].re = pr - qi;     outv[u + 1].im = pi + qr;
        outv[r - u - 1].re = pr + qi; outv[r - u - 1].im = pi - qr;
    }
    v[0] = o0;
    for (int i = 1; i < r; ++i) v[i] = outv[i];
}

static void bfly(orc_cplx *v, int r, int sgn, int leaf)
{
    switch (r) {
    case 2: bfly2(v); break;
    case 3: bfly3(v, sgn); break;
    case 4: bfly4(v, sgn); break;
    case 5: bfly5(v, sgn, leaf); break;
    case 7: bfly7(v, sgn, leaf); break;
    case 8: bfly8(v, sgn); break;
    default:
        if (r > 63 || (r & 1) == 0) { fprintf(stderr, "orc: unsupported radix %d\n", r); exit(EXIT_FAILURE); }
        bfly_generic(v, r, sgn);
    }
}

/* ------------------------------------------------------------ recursive DIT */

static void dit_rec(orc_cplx *out, const orc_cplx *in, const orc_plan *p, int len, int stride, int fi)
{
    /* src/highSpeedFFT.c:318-1629.  The r sub-transforms of the decimated input
     * land in out[i*m .. (i+1)*m); the combine loop then twiddles rows 1..r-1 by
     * tw[(m-1) + (r-1)k + (i-1)] and applies the radix-r butterfly per column k. */
    if (len == 1) { out[0] = in[0]; return; }
    int r = p->factors[fi], m = len / r;
    orc_cplx v[64];
    if (m == 1) { /* leaf: gather the strided inputs, butterfly, done */
        for (int i = 0; i < r; ++i) v[i] = in[(long)i * stride];
        bfly(v, r, p->sgn, 1);
        for (int i = 0; i < r; ++i) out[i] = v[i];
        return;
    }
    for (int i = 0; i < r; ++i)
        dit_rec(out + (long)i * m, in + (long)i * stride, p, m, stride * r, fi + 1);
    const orc_cplx *tw = p->tw + (m - 1);
    for (int k = 0; k < m; ++k) {
        v[0] = out[k];
        for (int i = 1; i < r; ++i) {
            orc_cplx x = out[k + (long)i * m], w = tw[(r - 1) * k + (i - 1)];
            v[i].re = x.re * w.re - x.im * w.im;
            v[i].im = x.im * w.re + x.re * w.im;
        }
        bfly(v, r, p->sgn, 0);
        for (int i = 0; i < r; ++i) out[k + (long)i * m] = v[i];
    }
}

/* ---------------------------------------------------------------- Bluestein */

static void bluestein(orc_plan *p, const orc_cplx *in, orc_cplx *out)
{
    /* src/highSpeedFFT.c:1735-1907 with bluestein_exp :1645-1704 inlined.
     * Same three length-M transforms and the same pointwise steps; the plan is
     * not mutated (a local view with conjugated twiddles is used for FFT #3). */
    int N = p->N, M = p->len, sgn = p->sgn;
    orc_cplx *yn = malloc(sizeof(orc_cplx) * M), *hk = malloc(sizeof(orc_cplx) * M);
    orc_cplx *tmp = malloc(sizeof(orc_cplx) * M), *yno = malloc(sizeof(orc_cplx) * M);
    orc_cplx *h = malloc(sizeof(orc_cplx) * N);
    orc_cplx *twc = malloc(sizeof(orc_cplx) * (M + 1));
    if (!yn || !hk || !tmp || !yno || !h || !twc) { fprintf(stderr, "orc: out of memory\n"); exit(EXIT_FAILURE); }

    /* chirp h[n] = exp(i pi n^2 / N), n^2 reduced mod 2N by a running sum (:1676-1689) */
    const double PI = 3.1415926535897932384626433832795;
    double theta = PI / N;
    int l2 = 0, len2 = 2 * N;
    for (int n = 0; n < N; ++n) {
        double ang = theta * l2;
        h[n].re = cos(ang); h[n].im = sin(ang);
        tmp[n] = h[n];
        l2 += 2 * n + 1;
        while (l2 > len2) l2 -= len2;
    }
    for (int n = N; n < M - N + 1; ++n) { tmp[n].re = 0.0; tmp[n].im = 0.0; }
    for (int n = M - N + 1; n < M; ++n) tmp[n] = h[M - n];

    double scale = 1.0 / M;
    for (int n = 0; n < M; ++n) { tmp[n].im *= scale; tmp[n].re *= scale; }

    orc_plan view = *p; /* mixed-radix view of length M */
    view.N = M; view.lt = 0;
    dit_rec(hk, tmp, &view, M, 1, 0); /* FFT #1 :1797 */

    if (sgn == 1)
        for (int n = 0; n < N; ++n) {
            tmp[n].re = in[n].re * h[n].re + in[n].im * h[n].im;
            tmp[n].im = -in[n].re * h[n].im + in[n].im * h[n].re;
        }
    else
        for (int n = 0; n < N; ++n) {
            tmp[n].re = in[n].re * h[n].re - in[n].im * h[n].im;
            tmp[n].im = in[n].re * h[n].im + in[n].im * h[n].re;
        }
    for (int n = N; n < M; ++n) { tmp[n].re = 0.0; tmp[n].im = 0.0; }

    dit_rec(yn, tmp, &view, M, 1, 0); /* FFT #2 :1832 */

    if (sgn == 1)
        for (int n = 0; n < M; ++n) {
            double t = yn[n].re * hk[n].re - yn[n].im * hk[n].im;
            yn[n].im = yn[n].re * hk[n].im + yn[n].im * hk[n].re;
            yn[n].re = t;
        }
    else
        for (int n = 0; n < M; ++n) {
            double t = yn[n].re * hk[n].re + yn[n].im * hk[n].im;
            yn[n].im = -yn[n].re * hk[n].im + yn[n].im * hk[n].re;
            yn[n].re = t;
        }

    for (int n = 0; n < M; ++n) { twc[n].re = p->tw[n].re; twc[n].im = -p->tw[n].im; }
    view.tw = twc; view.sgn = -sgn;
    dit_rec(yno, yn, &view, M, 1, 0); /* FFT #3 :1866 */

    if (sgn == 1)
        for (int n = 0; n < N; ++n) {
            out[n].re = yno[n].re * h[n].re + yno[n].im * h[n].im;
            out[n].im = -yno[n].re * h[n].im + yno[n].im * h[n].re;
        }
    else
        for (int n = 0; n < N; ++n) {
            out[n].re = yno[n].re * h[n].re - yno[n].im * h[n].im;
            out[n].im = yno[n].re * h[n].im + yno[n].im * h[n].re;
        }
    free(yn); free(hk); free(tmp); free(yno); free(h); free(twc);
}

void orc_fft_exec(orc_plan *p, const orc_cplx *in, orc_cplx *out)
{
    /* src/highSpeedFFT.c:1920-1942 */
    if (!p || !in || !out) { fprintf(stderr, "Error: Invalid FFT object or data pointers\n"); exit(EXIT_FAILURE); }
    if (p->lt == 0) dit_rec(out, in, p, p->N, 1, 0);
    else            bluestein(p, in, out);
}

/* --------------------------------------------------------------------- real */

orc_real_plan *orc_real_init(int N, int sgn)
{
    /* src/real.c:26-64 */
    if (N <= 0 || (N & 1)) { fprintf(stderr, "Error: Signal length (%d) must be positive and even\n", N); exit(EXIT_FAILURE); }
    orc_real_plan *p = calloc(1, sizeof(*p));
    int H = N / 2;
    p->N = N;
    p->cplan = orc_fft_init(H, sgn);
    p->tw2 = malloc(sizeof(orc_cplx) * H);
    for (int k = 0; k < H; ++k) {
        double ang = ORC_PI2 * k / N;
        p->tw2[k].re = cos(ang); p->tw2[k].im = sin(ang);
    }
    return p;
}

void orc_real_free(orc_real_plan *p)
{
    if (!p) return;
    orc_fft_free(p->cplan);
    free(p->tw2);
    free(p);
}

void orc_r2c_exec(orc_real_plan *p, const double *in, orc_cplx *out)
{
    /* src/real.c:78-136: even/odd samples packed as one half-length complex
     * signal, transformed, then split; the full N-bin mirrored spectrum is written. */
    int H = p->N / 2, N = p->N;
    orc_cplx *z = malloc(sizeof(orc_cplx) * H), *Z = malloc(sizeof(orc_cplx) * H);
    for (int i = 0; i < H; ++i) { z[i].re = in[2 * i]; z[i].im = in[2 * i + 1]; }
    orc_fft_exec(p->cplan, z, Z);
    out[0].re = Z[0].re + Z[0].im; out[0].im = 0.0;
    for (int k = 1; k < H; ++k) {
        double t1 = Z[k].im + Z[H - k].im;
        double t2 = Z[H - k].re - Z[k].re;
        out[k].re = (Z[k].re + Z[H - k].re + (t1 * p->tw2[k].re) + (t2 * p->tw2[k].im)) / 2.0;
        out[k].im = (Z[k].im - Z[H - k].im + (t2 * p->tw2[k].re) - (t1 * p->tw2[k].im)) / 2.0;
    }
    out[H].re = Z[0].re - Z[0].im; out[H].im = 0.0;
    for (int k = 1; k < H; ++k) { out[N - k].re = out[k].re; out[N - k].im = -out[k].im; }
    free(z); free(Z);
}

void orc_c2r_exec(orc_real_plan *p, const orc_cplx *in, double *out)
{
    /* src/real.c:150-193 */
    int H = p->N / 2;
    orc_cplx *z = malloc(sizeof(orc_cplx) * H), *Z = malloc(sizeof(orc_cplx) * H);
    for (int k = 0; k < H; ++k) {
        double t1 = -in[k].im - in[H - k].im;
        double t2 = -in[H - k].re + in[k].re;
        z[k].re = in[k].re + in[H - k].re + (t1 * p->tw2[k].re) - (t2 * p->tw2[k].im);
        z[k].im = in[k].im - in[H - k].im + (t2 * p->tw2[k].re) + (t1 * p->tw2[k].im);
    }
    orc_fft_exec(p->cplan, z, Z);
    for (int i = 0; i < H; ++i) { out[2 * i] = Z[i].re; out[2 * i + 1] = Z[i].im; }
    free(z); free(Z);
}

/* -------------------------------------------------------------- convolution */

int orc_next_pow2(int n)
{ /* src/convolve.c:20-25 */
    if (n <= 0) return 1;
    return (int)pow(2, ceil(log2(n)));
}

int orc_conv_fft_length(int min_len, const char *conv_type, int l1, int l2)
{ /* src/convolve.c:39-55 */
    if (strcmp(conv_type, "linear") == 0) return orc_next_pow2(min_len);
    if (strcmp(conv_type, "circular") == 0) return orc_next_pow2(l1 > l2 ? l1 : l2);
    fprintf(stderr, "Error: Invalid convolution type\n");
    exit(EXIT_FAILURE);
}

int orc_convolve(const char *type, const char *conv_type, const double *a, int la,
                 const double *b, int lb, double *out)
{
    /* src/convolve.c:74-214 */
    if (!a || !b || !out || la <= 0 || lb <= 0) return -1;
    int linear = strcmp(conv_type, "linear") == 0;
    if (!linear && strcmp(conv_type, "circular") != 0) return -1;
    int lmax = la > lb ? la : lb, lmin = la < lb ? la : lb;
    int clen = linear ? la + lb - 1 : lmax;
    int P = orc_conv_fft_length(clen, conv_type, la, lb);
    if (P < 2) P = 2; /* the reference exits for odd P=1 (real.c:28); keep the oracle total */
    orc_real_plan *f = orc_real_init(P, 1), *g = orc_real_init(P, -1);
    double *pa = calloc(P, sizeof(double)), *pb = calloc(P, sizeof(double)), *res = malloc(sizeof(double) * P);
    orc_cplx *A = malloc(sizeof(orc_cplx) * P), *B = malloc(sizeof(orc_cplx) * P), *C = malloc(sizeof(orc_cplx) * P);
    memcpy(pa, a, sizeof(double) * la);
    memcpy(pb, b, sizeof(double) * lb);
    orc_r2c_exec(f, pa, A);
    orc_r2c_exec(f, pb, B);
    for (int i = 0; i < P; ++i) {
        C[i].re = A[i].re * B[i].re - A[i].im * B[i].im;
        C[i].im = A[i].re * B[i].im + A[i].im * B[i].re;
    }
    orc_c2r_exec(g, C, res);
    for (int i = 0; i < P; ++i) res[i] /= P;
    int start = 0, n = 0;
    if (linear) {
        if (strcmp(type, "full") == 0) { start = 0; n = clen; }
        else if (strcmp(type, "same") == 0) { start = (clen - lmax) / 2; n = lmax; }
        else if (strcmp(type, "valid") == 0) { start = lmin - 1; n = lmax - lmin + 1; }
        else n = -1;
    } else { start = 0; n = P; }
    if (n > 0) memcpy(out, res + start, sizeof(double) * n);
    free(pa); free(pb); free(res); free(A); free(B); free(C);
    orc_real_free(f); orc_real_free(g);
    return n;
}

/* ------------------------------------------------------------ batch drivers */

int orc_num_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

void orc_c2c_batch(int N, int sgn, const orc_cplx *in, orc_cplx *out, long howmany, int nthreads)
{
#ifdef _OPENMP
    if (nthreads <= 0) nthreads = omp_get_max_threads();
#pragma omp parallel num_threads(nthreads)
#endif
    {
        orc_plan *p = orc_fft_init(N, sgn);
#ifdef _OPENMP
#pragma omp for schedule(static)
#endif
        for (long b = 0; b < howmany; ++b) orc_fft_exec(p, in + b * N, out + b * N);
        orc_fft_free(p);
    }
}

void orc_r2c_batch(int N, const double *in, orc_cplx *out, long howmany, int nthreads)
{
#ifdef _OPENMP
    if (nthreads <= 0) nthreads = omp_get_max_threads();
#pragma omp parallel num_threads(nthreads)
#endif
    {
        orc_real_plan *p = orc_real_init(N, 1);
#ifdef _OPENMP
#pragma omp for schedule(static)
#endif
        for (long b = 0; b < howmany; ++b) orc_r2c_exec(p, in + b * N, out + b * N);
        orc_real_free(p);
    }
}

void orc_c2r_batch(int N, const orc_cplx *in, double *out, long howmany, int nthreads)
{
    /* input rows hold N/2+1 bins each (stride N/2+1) */
#ifdef _OPENMP
    if (nthreads <= 0) nthreads = omp_get_max_threads();
#pragma omp parallel num_threads(nthreads)
#endif
    {
        orc_real_plan *p = orc_real_init(N, -1);
#ifdef _OPENMP
#pragma omp for schedule(static)
#endif
        for (long b = 0; b < howmany; ++b) orc_c2r_exec(p, in + b * (N / 2 + 1), out + b * N);
        orc_real_free(p);
    }
}

/* ------------------------------------------------------- independent checks */

void orc_naive_dft(int N, int sgn, const orc_cplx *in, orc_cplx *out)
{
    /* forward (sgn=+1): X[k] = sum x[n] exp(-2 pi i k n / N)  (src/highSpeedFFT.c:303) */
    long double *c = malloc(sizeof(long double) * N), *s = malloc(sizeof(long double) * N);
    const long double tp = 6.283185307179586476925286766559005768L;
    for (int j = 0; j < N; ++j) { c[j] = cosl(tp * j / N); s[j] = sinl(tp * j / N); }
#ifdef _OPENMP
#pragma omp parallel for schedule(static)
#endif
    for (int k = 0; k < N; ++k) {
        long double ar = 0, ai = 0;
        for (int n = 0; n < N; ++n) {
            int j = (int)(((long long)k * n) % N);
            long double wr = c[j], wi = -sgn * s[j];
            ar += in[n].re * wr - in[n].im * wi;
            ai += in[n].re * wi + in[n].im * wr;
        }
        out[k].re = (double)ar; out[k].im = (double)ai;
    }
    free(c); free(s);
}

double orc_rel_rms(const orc_cplx *y, const orc_cplx *r, long n)
{
    long double num = 0, den = 0;
    for (long i = 0; i < n; ++i) {
        long double dr = (long double)y[i].re - r[i].re, di = (long double)y[i].im - r[i].im;
        num += dr * dr + di * di;
        den += (long double)r[i].re * r[i].re + (long double)r[i].im * r[i].im;
    }
    if (den == 0) return num == 0 ? 0.0 : INFINITY;
    return (double)sqrtl(num / den);
}
