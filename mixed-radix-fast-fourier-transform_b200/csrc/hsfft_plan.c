/*
 * hsfft_plan.c -- complex plan builder and execute/destroy entry points (host side, plain C).
 *
 * Public symbols and the reference interface each replaces (paths under /root/reference):
 *   fft_init     src/highSpeedFFT.c:206-286     fft_exec   src/highSpeedFFT.c:1920-1942
 *   free_fft     src/highSpeedFFT.c:2315-2318   divideby   src/highSpeedFFT.c:1954-1968
 *   dividebyN    src/highSpeedFFT.c:1979-2025   factors    src/highSpeedFFT.c:2038-2163
 *   twiddle      src/highSpeedFFT.c:2186-2224   longvectorN src/highSpeedFFT.c:2238-2313
 *
 * The plan builder keeps the reference's routing (mixed radix vs Bluestein), its factor order and
 * its twiddle expression, so the public struct is field-for-field what the reference returns; the
 * execution plan (which sm_100a kernels run, in how many passes) is derived from those fields on
 * the first execute and cached in the registry (hs_engine.h).  No CPU fallback exists: execute
 * calls need a CUDA device and exit otherwise.
 */
#include "hs_engine.h"
#include "../../include/hsfft_b200.h"

#include <pthread.h>

/* ----------------------------------------------------------------- plan helpers (reference API) */

int divideby(int number, int divisor)
{
    if (number <= 0 || divisor <= 0) {
        fprintf(stderr, "Error: Invalid inputs for divideby - number: %d, divisor: %d\n", number, divisor);
        exit(EXIT_FAILURE);
    }
    while (number % divisor == 0) number /= divisor;
    return number == 1;
}

int dividebyN(int number)
{
    /* "is N a product of {2,3,4,5,7,8,11,13,17,23,29,31,37,41,43,47,53}?" -- 19 is not in the
     * router's list (reference :13), so multiples of 19 go to Bluestein.  The reference answers
     * from a constructor-filled table below 1024 and by trial division above; both compute this. */
    static const int radices[13] = {53, 47, 43, 41, 37, 31, 29, 23, 17, 13, 11, 7, 5};
    if (number < 1) return 0; /* the reference indexes its table out of bounds here */
    for (int i = 0; i < 13; ++i)
        while (number % radices[i] == 0) number /= radices[i];
    while (number % 3 == 0) number /= 3;
    while (number % 2 == 0) number /= 2;
    return number == 1;
}

int factors(int number, int *out)
{
    static const int order[18] = {53, 47, 43, 41, 37, 31, 29, 23, 19, 17, 13, 11, 8, 7, 5, 4, 3, 2};
    if (out == NULL) {
        fprintf(stderr, "Error: Invalid inputs for factors - number: %d, factors_array: %p\n", number, (void *)out);
        exit(EXIT_FAILURE);
    }
    if (number <= 0) { /* the reference never terminates on 0 (0 % 53 == 0); refuse cleanly */
        fprintf(stderr, "Error: Invalid inputs for factors - number: %d\n", number);
        exit(EXIT_FAILURE);
    }
    int cnt = 0, rest = number;
    for (int i = 0; i < 18; ++i)
        while (rest % order[i] == 0) { out[cnt++] = order[i]; rest /= order[i]; }
    if (rest > 31) { /* 6k+-1 sweep for what is left (reference :2140-2160) */
        for (int k = 2; rest > 1; ++k) {
            const int lo = 6 * k - 1, hi = 6 * k + 1;
            while (rest % lo == 0) { out[cnt++] = lo; rest /= lo; }
            while (rest % hi == 0) { out[cnt++] = hi; rest /= hi; }
        }
    }
    return cnt;
}

void twiddle(fft_data *sig, int N, int radix)
{
    /* Never called by the reference itself.  Its table branch reads the mis-indexed
     * twiddle_tables[]; this keeps the documented meaning:
     * sig[i] = exp(-2 pi i * i / N), i < N / radix. */
    if (N <= 0 || radix <= 0) {
        fprintf(stderr, "Error: Invalid inputs for twiddle - signal_length: %d, radix: %d\n", N, radix);
        exit(EXIT_FAILURE);
    }
    const int cnt = N / radix;
    for (int i = 0; i < cnt; ++i) {
        const fft_type angle = PI2 * i / N;
        sig[i].re = cos(angle);
        sig[i].im = -sin(angle);
    }
}

void longvectorN(fft_data *sig, int N, int *array, int M)
{
    /* Stage twiddles in the reference's layout: factors walked last to first, stage of combine
     * length L = r*Ls writes (r-1)*Ls entries cos/sin((k+1)*j*theta), theta = -PI2/L, j-major
     * (reference :2273-2289, the branch that is correct for every radix). */
    if (N <= 0 || array == NULL || M <= 0) {
        fprintf(stderr, "Error: Invalid inputs for longvectorN - signal_length: %d, prime_factors: %p, num_factors: %d\n", N,
                (void *)array, M);
        exit(EXIT_FAILURE);
    }
    int L = 1, written = 0;
    for (int s = M - 1; s >= 0; --s) {
        const int r = array[s], Ls = L;
        L *= r;
        const fft_type theta = -PI2 / L;
        for (int j = 0; j < Ls; ++j)
            for (int k = 0; k < r - 1; ++k)
                if (written < N - 1) {
                    const fft_type angle = (k + 1) * j * theta;
                    sig[written].re = cos(angle);
                    sig[written].im = sin(angle);
                    ++written;
                }
    }
    if (written != N - 1)
        fprintf(stderr, "Warning: Twiddle factor count (%d) does not match expected (%d) for N=%d\n", written, N - 1, N);
}

/* -------------------------------------------------------------------- device-side table builders */

int hs_stage_twiddle_count(const int *rx, int nst, int split_min)
{
    int m = 1, cnt = 0;
    for (int s = 0; s < nst; ++s) {
        if (s > 0) cnt += hs_stage_tw_entries(rx[s], m, split_min);
        m *= rx[s];
    }
    return cnt;
}

static void put_twiddle(fft_data *dst, int exponent, fft_type theta, int sgn)
{
    const fft_type angle = exponent * theta; /* same expression as longvectorN: (int product) * theta */
    dst->re = cos(angle);
    dst->im = (sgn == -1) ? -sin(angle) : sin(angle);
}

void hs_build_stage_twiddles(fft_data *out, const int *rx, int nst, int sgn, int split_min)
{
    /* Same values as longvectorN (same expression, so bit-identical where the stage lengths
     * coincide), laid out [i-1][k] per stage so that a warp's loads of one twiddle row coalesce.
     * Stages whose table would exceed split_min entries are stored split (hs_split_kl, hs_abi.h):
     * lo[i-1][k_lo] = W_L^(i k_lo) followed by hi[i-1][k_hi] = W_L^(i KL k_hi). */
    int m = 1, off = 0;
    for (int s = 0; s < nst; ++s) {
        const int r = rx[s];
        if (s > 0) {
            const fft_type theta = -PI2 / ((fft_type)m * r);
            const int KL = hs_split_kl(r, m, split_min);
            if (KL == 0) {
                for (int i = 1; i < r; ++i)
                    for (int k = 0; k < m; ++k) put_twiddle(&out[off + (i - 1) * m + k], i * k, theta, sgn);
            } else {
                const int KH = m / KL;
                fft_data *lo = out + off, *hi = out + off + (r - 1) * KL;
                for (int i = 1; i < r; ++i) {
                    for (int k = 0; k < KL; ++k) put_twiddle(&lo[(i - 1) * KL + k], i * k, theta, sgn);
                    for (int k = 0; k < KH; ++k) put_twiddle(&hi[(i - 1) * KH + k], i * KL * k, theta, sgn);
                }
            }
            off += hs_stage_tw_entries(r, m, split_min);
        }
        m *= r;
    }
}

void hs_build_chirp(fft_data *h, int N)
{
    /* h[n] = exp(+i pi n^2 / N), n^2 reduced mod 2N by the running sum of reference :1676-1689 */
    const fft_type PI = 3.1415926535897932384626433832795;
    const fft_type theta = PI / N;
    int l2 = 0;
    const int len2 = 2 * N;
    for (int n = 0; n < N; ++n) {
        const fft_type angle = theta * l2;
        h[n].re = cos(angle);
        h[n].im = sin(angle);
        l2 += 2 * n + 1;
        while (l2 > len2) l2 -= len2;
    }
}

int hs_bluestein_exec_pad(int N)
{
    int m = 1;
    while (m < 2 * N - 1) m *= 2; /* = 2^ceil(log2(2N-1)), reference :1750-1751 */
    return m;
}

static const hs_ct_desc *find_ct_ex(int family, int n, int allow_exact);
static void *upload(const void *h, size_t bytes);

/* The pad length the device runs with.  Any M >= 2N-1 gives the same convolution (SURVEY.md section 8c "may
 * differ"); between two powers of two the engine also offers 5/8 of the upper one where a compile-time plan with
 * full-precision radix-5 stages exists (N = 10007: 20480 instead of 32768).  HSFFT_BLU_POW2=1 keeps the power of two. */
/* Split chirp-z (hs_fused.cuh, hs_bluestein_split_kernel): the smallest compile-time plan of length M' >= N that one SM
 * holds; the convolution then runs at M = 2 M'.  Used above the lengths the plain one-kernel chirp-z covers (N > 4096:
 * M >= 16384 does not fit one SM); HSFFT_BLU_SPLIT=0 turns it off, =2 prefers it wherever a plan exists. */
static const hs_ct_desc *bluestein_split_plan(int N)
{
    /* HSFFT_BLU_SPLIT: 0 off; 1 (default) where it measured faster than the plain one-kernel chirp-z / the three-kernel
     * route (N > 2048: M' = 4096, 8192, 10240); 2 wherever a plan exists.  Two kernel families offer such plans: the
     * auto-sort pair of hs_fused.cuh (N = 4093 1.61 ms per 16384 transforms against 1.82 ms plain, N = 4099 3.28 ms against
     * 3.96 ms for the three kernels) and the in-place transforms of hs_conv_fft.cuh with their intermediates in tensor
     * memory (the only one that fits M' = 10240 without spilling: 4.33 against 8.7 ms; M' = 8192: N = 8191 3.81 against
     * 4.11 ms, but N = 4099 3.41 against 3.28 ms, hence the split of that range); HSFFT_SPLIT_FAMILY=5 / 6 forces one. */
    const int mode = getenv("HSFFT_BLU_SPLIT") ? atoi(getenv("HSFFT_BLU_SPLIT")) : 1;
    if (mode == 0 || (mode == 1 && N <= 2048)) return NULL;
    const int force = getenv("HSFFT_SPLIT_FAMILY") ? atoi(getenv("HSFFT_SPLIT_FAMILY")) : 0;
    const hs_ct_desc *best = NULL;
    const int cnt = hs_cu_ct_count();
    for (int i = 0; i < cnt; ++i) {
        const hs_ct_desc *d = hs_cu_ct_get(i);
        if ((d->family != HS_CT_BLUESTEIN_SPLIT && d->family != HS_CT_CONV_SPLIT) || d->n < N) continue;
        if (force && d->family != force) continue;
        const int want = (d->n > 8192 || (d->n == 8192 && N > 6144)) ? HS_CT_CONV_SPLIT : HS_CT_BLUESTEIN_SPLIT;
        if (!best || d->n < best->n || (d->n == best->n && best->family != want && d->family == want && !getenv("HSFFT_PREFER_CT"))) best = d;
    }
    return best;
}

/* tables of the in-place convolution transforms (hs_conv_fft.cuh): forward stage s < nst-1 works on blocks of length L,
 * m = L / r, and multiplies output j of the butterfly at in-block position p by W_L^(p j); layout [j-1][p] per stage */
static int conv_twiddle_count(const int *rx, int nst, int n)
{
    int L = n, cnt = 0;
    for (int s = 0; s + 1 < nst; ++s) { cnt += (rx[s] - 1) * (L / rx[s]); L /= rx[s]; }
    return cnt;
}
static void *upload_conv_twiddles(const int *rx, int nst, int n)
{
    const int cnt = conv_twiddle_count(rx, nst, n);
    fft_data *t = (fft_data *)malloc(sizeof(fft_data) * (cnt > 0 ? cnt : 1));
    if (!t) { fprintf(stderr, "Error: Memory allocation failed for twiddle table\n"); exit(EXIT_FAILURE); }
    int L = n, off = 0;
    for (int s = 0; s + 1 < nst; ++s) {
        const int r = rx[s], m = L / r;
        const fft_type theta = -PI2 / (fft_type)L;
        for (int j = 1; j < r; ++j)
            for (int p = 0; p < m; ++p) put_twiddle(&t[off + (j - 1) * m + p], p * j, theta, 1);
        off += (r - 1) * m;
        L = m;
    }
    void *d = upload(t, sizeof(fft_data) * cnt);
    hs_cu_stream_sync(NULL);
    free(t);
    return d;
}
/* spectrum table in the order the forward stages leave the spectrum in: position pos = sum_s j_s m_s holds bin
 * j_0 + r_0 (j_1 + r_1 (...)); stored as [i][beta] for pos = r_last beta + i (coalesced across the threads) */
static void conv_permute_hk(fft_data *dst, const fft_data *hk, const int *rx, int nst, int n)
{
    const int rl = rx[nst - 1], nb = n / rl;
    for (int pos = 0; pos < n; ++pos) {
        int m = n, k = 0, w = 1;
        for (int s = 0; s < nst; ++s) {
            m /= rx[s];
            k += ((pos / m) % rx[s]) * w;
            w *= rx[s];
        }
        dst[(size_t)(pos % rl) * nb + pos / rl] = hk[k];
    }
}

static int bluestein_device_pad(int N)
{
    const hs_ct_desc *sp = bluestein_split_plan(N);
    if (sp) return 2 * sp->n;
    const int m = hs_bluestein_exec_pad(N);
    const int alt = m / 8 * 5;
    if (m >= 8 && alt >= 2 * N - 1 && !getenv("HSFFT_BLU_POW2") && find_ct_ex(HS_CT_TWOPASS, alt, 1) &&
        find_ct_ex(HS_CT_TWOPASS, alt, 1)->exact)
        return alt;
    return m;
}

/* ------------------------------------------------------------------------------- c2c sub-plans */

#define HS_SMEM_POINTS 6400 /* complex points of one ping-pong buffer pair's half (2*6400*16 B = 200 KiB) */
#define HS_GROUP_MAX 512

/* allow_exact: also accept plans whose radix-3/5/7 stages use full-precision constants (Bluestein-internal use) */
static const hs_ct_desc *find_ct_ex(int family, int n, int allow_exact)
{
    const int cnt = hs_cu_ct_count();
    for (int i = 0; i < cnt; ++i) {
        const hs_ct_desc *d = hs_cu_ct_get(i);
        if (d->family == family && d->n == n && (allow_exact || !d->exact)) return d;
    }
    return NULL;
}
static const hs_ct_desc *find_ct(int family, int n) { return find_ct_ex(family, n, 0); }
/* the sub-plan (first- or second-pass family) that runs exactly these stages: the one whose twiddle tables a two-pass
 * descriptor's radix[] / radix2[] describe */
static const hs_ct_desc *find_sub_like(int n, const int *radix, int nst)
{
    const int cnt = hs_cu_ct_count();
    for (int i = 0; i < cnt; ++i) {
        const hs_ct_desc *d = hs_cu_ct_get(i);
        if ((d->family != HS_CT_SUB && d->family != HS_CT_SUB2) || d->n != n || d->nst != nst) continue;
        int same = 1;
        for (int s = 0; s < nst; ++s) same &= d->radix[s] == radix[s];
        if (same) return d;
    }
    return NULL;
}

static void *upload(const void *h, size_t bytes)
{
    void *d = hs_cu_malloc(bytes ? bytes : 16);
    if (bytes) hs_cu_h2d(d, h, bytes, NULL);
    return d;
}

static void *upload_stage_twiddles(const int *rx, int nst, int sgn, int split_min)
{
    const int cnt = hs_stage_twiddle_count(rx, nst, split_min);
    fft_data *t = (fft_data *)malloc(sizeof(fft_data) * (cnt > 0 ? cnt : 1));
    if (!t) { fprintf(stderr, "Error: Memory allocation failed for twiddle table\n"); exit(EXIT_FAILURE); }
    hs_build_stage_twiddles(t, rx, nst, sgn, split_min);
    void *d = upload(t, sizeof(fft_data) * cnt);
    hs_cu_stream_sync(NULL);
    free(t);
    return d;
}

static int env_int(const char *name, int dflt)
{
    const char *e = getenv(name);
    return e ? atoi(e) : dflt;
}

static hs_c2c *c2c_create(int n, int sgn, int allow_exact);
hs_c2c *hs_c2c_create(int n, int sgn) { return c2c_create(n, sgn, 0); }

static hs_c2c *c2c_create(int n, int sgn, int allow_exact)
{
    hs_c2c *p = (hs_c2c *)calloc(1, sizeof(hs_c2c));
    if (!p) { fprintf(stderr, "Error: Memory allocation failed for execution plan\n"); exit(EXIT_FAILURE); }
    p->n = n;
    p->sgn = sgn;
    p->ct_id = -1;
    p->tp_id = -1;
    p->jit_id = -1;
    p->jtp_id = -1;
    const hs_ct_desc *ct = find_ct(HS_CT_C2C, n);
    if (ct) {
        p->ct_id = ct->id;
        p->d_tw = upload_stage_twiddles(ct->radix, ct->nst, sgn, ct->split_min);
        return p;
    }
    ct = find_ct_ex(HS_CT_TWOPASS, n, allow_exact);
    if (ct) {
        /* n = n1 * n2: pass 1 = n2 strided transforms of length n1, pass 2 = n1 of length n2 with the
         * inter-pass twiddle W_n^{i k} applied on load as W_n^{k b} * (W_n^{k NB})^ii, i = b + NB ii */
        p->tp_id = ct->id;
        p->n1 = ct->n1;
        p->n2 = ct->n2;
        p->d_tw = upload_stage_twiddles(ct->radix, ct->nst, sgn, ct->split_min);
        p->d_tw2 = upload_stage_twiddles(ct->radix2, ct->nst2, sgn, ct->split_min2);
        const int NB = ct->n2 / ct->radix2[0], K = ct->n1;
        fft_data *tb = (fft_data *)malloc(sizeof(fft_data) * (size_t)NB * K), *ts = (fft_data *)malloc(sizeof(fft_data) * K);
        if (!tb || !ts) { fprintf(stderr, "Error: Memory allocation failed for twiddle table\n"); exit(EXIT_FAILURE); }
        const fft_type theta = -PI2 / (fft_type)n;
        for (int b = 0; b < NB; ++b)
            for (int k = 0; k < K; ++k) put_twiddle(&tb[(size_t)b * K + k], k * b, theta, sgn);
        for (int k = 0; k < K; ++k) put_twiddle(&ts[k], k * NB, theta, sgn);
        p->d_twb = upload(tb, sizeof(fft_data) * (size_t)NB * K);
        p->d_tws = upload(ts, sizeof(fft_data) * K);
        hs_cu_stream_sync(NULL);
        free(tb); free(ts);
        return p;
    }
    /* run-time plan: the reference's factor list reversed = execution order */
    int fac[64], rx[64];
    const int lf = factors(n, fac);
    if (lf <= 0 || lf > HS_MAX_STAGES) { fprintf(stderr, "Error: cannot plan N=%d\n", n); exit(EXIT_FAILURE); }
    for (int s = 0; s < lf; ++s) {
        rx[s] = fac[lf - 1 - s];
        if (rx[s] > 8 && (rx[s] % 2 == 0 || rx[s] >= HS_MAX_GENERIC_RADIX)) {
            fprintf(stderr, "Error: unsupported radix %d in the factor list of N=%d\n", rx[s], n);
            exit(EXIT_FAILURE);
        }
    }
    /* a length without a static plan whose stages fit one SM gets its own kernel, compiled now (hs_jit.cu) */
    {
        hs_ct_desc jd;
        const int id = hs_cu_jit_c2c_create(n, sgn, rx, lf, &jd);
        if (id >= 0) {
            p->jit_id = id;
            p->d_tw = upload_stage_twiddles(jd.radix, jd.nst, sgn, jd.split_min);
            return p;
        }
    }
    /* longer lengths whose stages are all compile-time radices (and shorter ones the one-kernel planner turned down, e.g.
     * 3^8): two tiled passes compiled now; the factor list (execution order) is cut where the two sub-transform lengths
     * are closest */
    if (n >= 1024 && lf >= 4) {
        int cand[64], nc = 0;
        for (int a = 2; a <= lf - 2; ++a) cand[nc++] = a;
        for (int i = 0; i < nc; ++i)       /* order the cuts by |log n1 - log n2| */
            for (int j = i + 1; j < nc; ++j) {
                double pi_ = 1, pj = 1;
                for (int s = 0; s < cand[i]; ++s) pi_ *= rx[s];
                for (int s = 0; s < cand[j]; ++s) pj *= rx[s];
                if (fabs(log(pi_ * pi_ / n)) > fabs(log(pj * pj / n))) { const int t = cand[i]; cand[i] = cand[j]; cand[j] = t; }
            }
        for (int i = 0; i < nc && i < 3; ++i) {
            hs_ct_desc jd;
            const int id = hs_cu_jit_twopass_create(n, sgn, rx, lf, cand[i], &jd);
            if (id < 0) continue;
            p->jtp_id = id;
            p->n1 = jd.n1; p->n2 = jd.n2;
            p->d_tw = upload_stage_twiddles(jd.radix, jd.nst, sgn, jd.split_min);
            p->d_tw2 = upload_stage_twiddles(jd.radix2, jd.nst2, sgn, jd.split_min2);
            p->d_twb = hs_upload_interpass_tables(n, sgn, jd.n2 / jd.radix2[0], jd.n1, jd.n1, &p->d_tws);
            return p;
        }
    }
    /* cos/sin half tables of the generic odd radices (reference :1533-1543) */
    double *cs = (double *)calloc((size_t)lf * HS_MAX_GENERIC_RADIX, sizeof(double));
    double *sn = (double *)calloc((size_t)lf * HS_MAX_GENERIC_RADIX, sizeof(double));
    int gen_off[64], gen_total = 0;
    for (int s = 0; s < lf; ++s) {
        gen_off[s] = 0;
        const int r = rx[s];
        if (r > 8) {
            const int h = (r - 1) / 2;
            gen_off[s] = gen_total;
            double *c = cs + gen_total, *sv = sn + gen_total;
            for (int i = 1; i <= h; ++i) { c[i - 1] = cos(i * PI2 / r); sv[i - 1] = sin(i * PI2 / r); }
            for (int i = 0; i < h; ++i) { sv[i + h] = -sv[h - 1 - i]; c[i + h] = c[h - 1 - i]; }
            gen_total += r - 1;
        }
    }
    if (gen_total) {
        p->d_gen_cs = upload(cs, sizeof(double) * gen_total);
        p->d_gen_sn = upload(sn, sizeof(double) * gen_total);
        hs_cu_stream_sync(NULL);
    }
    free(cs);
    free(sn);

    /* group consecutive stages into passes whose composite radix fits shared memory */
    int gstart[HS_MAX_PASS + 1], ng = 0;
    if (n <= HS_SMEM_POINTS) { gstart[0] = 0; gstart[1] = lf; ng = 1; }
    else {
        int s = 0;
        while (s < lf) {
            if (ng >= HS_MAX_PASS) { fprintf(stderr, "Error: N=%d needs too many passes\n", n); exit(EXIT_FAILURE); }
            gstart[ng] = s;
            long prod = rx[s++];
            if (prod > HS_SMEM_POINTS) { fprintf(stderr, "Error: radix %ld too large\n", prod); exit(EXIT_FAILURE); }
            while (s < lf && prod * rx[s] <= HS_GROUP_MAX) prod *= rx[s++];
            ++ng;
        }
        gstart[ng] = lf;
    }
    p->npass = ng;
    long long m = 1;
    for (int g = 0; g < ng; ++g) {
        hs_gen_params *q = &p->pass[g];
        const int *grx = rx + gstart[g];
        const int gst = gstart[g + 1] - gstart[g];
        long long Rg = 1;
        for (int s = 0; s < gst; ++s) Rg *= grx[s];
        memset(q, 0, sizeof(*q));
        q->n = (int)Rg;
        q->nst = gst;
        q->npad = (int)(Rg | 1);
        int mm = 1;
        int off = 0;
        for (int s = 0; s < gst; ++s) {
            q->radix[s] = grx[s];
            q->gen_off[s] = gen_off[gstart[g] + s];
            q->tw_off[s] = (s > 0) ? off : 0;
            if (s > 0) off += (grx[s] - 1) * mm;
            mm *= grx[s];
        }
        p->d_pass_tw[g] = upload_stage_twiddles(grx, gst, sgn, 0);
        q->tw = p->d_pass_tw[g];
        q->gen_cs = (const double *)p->d_gen_cs;
        q->gen_sn = (const double *)p->d_gen_sn;
        int C = HS_SMEM_POINTS / q->npad;
        if (ng == 1) { int want = 2048 / q->n; if (want < 1) want = 1; if (C > want) C = want; }
        else if (C > 16) C = 16;
        if (C < 1) C = 1;
        q->C = C;
        if (ng == 1) {
            /* whole transform in one pass: the batch is the tile dimension (set per launch) */
            q->n_o = 1; q->n_batch = 1;
            q->in_c = n; q->in_e = 1; q->out_c = n; q->out_e = 1;
            q->load_c_fast = 0; q->store_c_fast = 0;
        } else if (g == 0) {
            const long long Qp = n / Rg;
            q->n_o = 1; q->n_c = Qp;
            q->in_b = n; q->out_b = n;
            q->in_c = 1; q->in_e = Qp; q->out_c = Rg; q->out_e = 1;
            q->load_c_fast = 1; q->store_c_fast = 0;
        } else {
            const long long Qp = n / (m * Rg);
            q->n_o = Qp; q->n_c = m;
            q->in_b = n; q->out_b = n;
            q->in_o = m; q->in_c = 1; q->in_e = Qp * m;
            q->out_o = m * Rg; q->out_c = 1; q->out_e = m;
            q->load_c_fast = 1; q->store_c_fast = 1;
            /* inter-pass twiddle W_L^(e*c), L = m*Rg, split in two short tables */
            const long long L = m * Rg;
            int S = (int)ceil(sqrt((double)L));
            if (S < 1) S = 1;
            const long long nhi = (L + S - 1) / S;
            fft_data *lo = (fft_data *)malloc(sizeof(fft_data) * S), *hi = (fft_data *)malloc(sizeof(fft_data) * nhi);
            if (!lo || !hi) { fprintf(stderr, "Error: Memory allocation failed for twiddle table\n"); exit(EXIT_FAILURE); }
            const fft_type theta = -PI2 / (fft_type)L;
            for (int x = 0; x < S; ++x) {
                const fft_type a = x * theta;
                lo[x].re = cos(a); lo[x].im = (sgn == -1) ? -sin(a) : sin(a);
            }
            for (long long x = 0; x < nhi; ++x) {
                const fft_type a = (fft_type)(x * S) * theta;
                hi[x].re = cos(a); hi[x].im = (sgn == -1) ? -sin(a) : sin(a);
            }
            p->d_pass_lo[g] = upload(lo, sizeof(fft_data) * S);
            p->d_pass_hi[g] = upload(hi, sizeof(fft_data) * nhi);
            hs_cu_stream_sync(NULL);
            free(lo); free(hi);
            q->tw_L = L; q->tw_S = S; q->tw_lo = p->d_pass_lo[g]; q->tw_hi = p->d_pass_hi[g];
        }
        m *= Rg;
    }
    return p;
}

/* plan-building helpers shared with hsfft_real.c */
const hs_ct_desc *hs_find_ct(int family, int n, int allow_exact) { return find_ct_ex(family, n, allow_exact); }
/* sub-plan of length n for a second pass: one built for that role where it exists */
const hs_ct_desc *hs_find_sub_second(int n)
{
    const hs_ct_desc *d = getenv("HSFFT_NO_SUB2") ? NULL : find_ct_ex(HS_CT_SUB2, n, 0);
    return d ? d : find_ct_ex(HS_CT_SUB, n, 0);
}
void *hs_upload_stage_twiddles(const int *rx, int nst, int sgn, int split_min) { return upload_stage_twiddles(rx, nst, sgn, split_min); }
void *hs_upload_interpass_tables(long long n, int sgn, int nb, int K, int kvalid, void **d_tws)
{
    /* twb[b][k] = W_n^(k b), b < nb, and tws[k] = W_n^(k nb), k < kvalid (entries kvalid..K-1 are padding) */
    fft_data *tb = (fft_data *)calloc((size_t)nb * K, sizeof(fft_data)), *ts = (fft_data *)calloc((size_t)K, sizeof(fft_data));
    if (!tb || !ts) { fprintf(stderr, "Error: Memory allocation failed for twiddle table\n"); exit(EXIT_FAILURE); }
    const fft_type theta = -PI2 / (fft_type)n;
    for (int b = 0; b < nb; ++b)
        for (int k = 0; k < kvalid; ++k) put_twiddle(&tb[(size_t)b * K + k], k * b, theta, sgn);
    for (int k = 0; k < kvalid; ++k) put_twiddle(&ts[k], k * nb, theta, sgn);
    void *d_twb = upload(tb, sizeof(fft_data) * (size_t)nb * K);
    *d_tws = upload(ts, sizeof(fft_data) * (size_t)K);
    hs_cu_stream_sync(NULL);
    free(tb); free(ts);
    return d_twb;
}

void hs_c2c_destroy(hs_c2c *p)
{
    if (!p) return;
    hs_cu_free(p->d_tw);
    hs_cu_free(p->d_tw2); hs_cu_free(p->d_twb); hs_cu_free(p->d_tws);
    hs_cu_free(p->d_mrg_wc); hs_cu_free(p->d_mrg_wr);
    for (int g = 0; g < HS_MAX_PASS; ++g) { hs_cu_free(p->d_pass_tw[g]); hs_cu_free(p->d_pass_lo[g]); hs_cu_free(p->d_pass_hi[g]); }
    hs_cu_free(p->d_gen_cs);
    hs_cu_free(p->d_gen_sn);
    free(p);
}

size_t hs_c2c_ws_elems(const hs_c2c *p, long long howmany)
{
    if (p->tp_id >= 0 || p->jtp_id >= 0) return (size_t)howmany * p->n;
    return (p->ct_id < 0 && p->npass > 1) ? (size_t)howmany * p->n : 0;
}

void hs_c2c_exec(const hs_c2c *p, const void *d_in, void *d_out, void *d_ws, long long howmany, void *stream)
{
    if (howmany <= 0) return;
    if (p->ct_id >= 0) { hs_cu_launch_ct_c2c(p->ct_id, p->sgn, d_in, d_out, p->d_tw, howmany, stream); return; }
    if (p->jit_id >= 0) { hs_cu_jit_c2c_launch(p->jit_id, d_in, d_out, p->d_tw, howmany, stream); return; }
    if (p->jtp_id >= 0) {
        hs_cu_jit_twopass_launch(p->jtp_id, 1, d_in, d_ws, p->d_tw, NULL, NULL, howmany, stream);
        hs_cu_jit_twopass_launch(p->jtp_id, 2, d_ws, d_out, p->d_tw2, p->d_twb, p->d_tws, howmany, stream);
        return;
    }
    if (p->tp_id >= 0) {
        hs_cu_launch_ct_twopass(p->tp_id, 1, p->sgn, d_in, d_ws, p->d_tw, NULL, NULL, howmany, stream);
        hs_cu_launch_ct_twopass(p->tp_id, 2, p->sgn, d_ws, d_out, p->d_tw2, p->d_twb, p->d_tws, howmany, stream);
        return;
    }
    if (p->npass == 1) {
        hs_gen_params q = p->pass[0];
        q.n_c = howmany;
        if (q.C > howmany) q.C = (int)howmany;
        hs_cu_launch_generic(p->sgn, d_in, d_out, &q, stream);
        return;
    }
    const void *src = d_in;
    for (int g = 0; g < p->npass; ++g) {
        hs_gen_params q = p->pass[g];
        q.n_batch = howmany;
        if (q.C > q.n_c) q.C = (int)q.n_c;
        void *dst = ((p->npass - 1 - g) % 2 == 0) ? d_out : d_ws;
        hs_cu_launch_generic(p->sgn, src, dst, &q, stream);
        src = dst;
    }
}

/* ------------------------------------------------------------------------------------ registry */

static pthread_mutex_t g_lock = PTHREAD_MUTEX_INITIALIZER;
static hs_dev_obj *g_head = NULL;

static void dev_free_contents(hs_dev_obj *d)
{
    hs_scratch_free(&d->ws_pool);
    hs_c2c_destroy(d->direct);
    hs_c2c_destroy(d->fwdM);
    hs_c2c_destroy(d->invM);
    hs_cu_free(d->d_chirp);
    hs_cu_free(d->d_hk);
    hs_cu_free(d->d_blu_twf);
    hs_cu_free(d->d_blu_twi);
    hs_cu_free(d->d_big_f1); hs_cu_free(d->d_big_f2); hs_cu_free(d->d_big_i1); hs_cu_free(d->d_big_i2);
    hs_cu_free(d->d_big_fb); hs_cu_free(d->d_big_fs); hs_cu_free(d->d_big_ib); hs_cu_free(d->d_big_is);
    for (int s = 0; s < HS_NSLOT; ++s) {
        hs_cu_free(d->d_in[s]); hs_cu_free(d->d_out[s]); hs_cu_free(d->d_ws[s]); hs_cu_free(d->d_ws2[s]);
        if (d->stream[s]) hs_cu_stream_destroy(d->stream[s]);
    }
}

/* Rader's algorithm (hs_fused.cuh): a prime N whose N-1 has a shared-memory kernel of the HS_CT_RADER family becomes one
 * cyclic convolution of length L = N-1, half the points of Bluestein's padded one.  Builds perm[q] = g^q mod N and
 * Bk = FFT_L(b) / L, b[m] = exp(-sgn 2 pi i g^(-m) / N), the latter as a direct sum in long double (plan time, L^2). */
static int rader_setup(hs_dev_obj *d)
{
    const int N = d->N, L = N - 1;
    if (N < 5 || getenv("HSFFT_NO_RADER")) return 0;
    for (int f = 2; (long long)f * f <= N; ++f)
        if (N % f == 0) return 0;
    const hs_ct_desc *ct = find_ct_ex(HS_CT_RADER, L, 1);
    hs_ct_desc jd;
    int jit = 0;
    if (!ct && L >= 64 && L <= 4096 && dividebyN(L)) {
        /* no static kernel: compile one for this prime if N-1 is a product of compile-time radices (any order serves: the
         * convolution's transforms use full-precision constants).  Largest radices last keeps the hand-over wide. */
        int fac[64], rx[64];
        const int lf = factors(L, fac);
        int ok = lf >= 2 && lf <= 8;
        for (int s = 0; s < lf && ok; ++s) {
            rx[s] = fac[lf - 1 - s];
            if (rx[s] > 17 || rx[s] == 16) ok = 0;
        }
        /* merge 2s and 4s / 8s into 16s where possible: fewer passes */
        if (ok) {
            int cnt2 = 0, other[64], no = 0;
            for (int s = 0; s < lf; ++s) {
                if (rx[s] == 2) cnt2 += 1; else if (rx[s] == 4) cnt2 += 2; else if (rx[s] == 8) cnt2 += 3; else other[no++] = rx[s];
            }
            int n2 = 0, two[64];
            while (cnt2 >= 4) { two[n2++] = 16; cnt2 -= 4; }
            if (cnt2 == 3) two[n2++] = 8; else if (cnt2 == 2) two[n2++] = 4; else if (cnt2 == 1) two[n2++] = 2;
            /* small odd radices first, powers of two (the widest butterflies) last; a lone radix 2 goes first */
            int m = 0;
            for (int s = 0; s < n2; ++s) if (two[s] == 2) rx[m++] = 2;
            for (int s = 0; s < no; ++s) rx[m++] = other[s];
            for (int s = n2 - 1; s >= 0; --s) if (two[s] != 2) rx[m++] = two[s];
            /* Worth it only where the convolution has few, wide passes (measured on B200 against the chirp-z route, ms per
             * batch: N=401 0.36 vs 0.46, 769 0.56 vs 1.02, 1153 0.59 vs 1.11, 3137 0.91 vs 1.42, 2521 (5 passes) 0.85 vs 1.35;
             * but 127 0.71 vs 0.32, 251 0.47 vs 0.41, 2017 (5 passes) 0.68 vs 0.58, 4051 (7 passes) 2.5 vs 1.6).
             * HSFFT_RADER_JIT=0 never, =2 whenever a kernel can be compiled. */
            const int mode = getenv("HSFFT_RADER_JIT") ? atoi(getenv("HSFFT_RADER_JIT")) : 1;
            const int pays = (m <= 4 && L >= 384) || (m == 5 && N > 2048);
            if (m >= 2 && m <= 8 && mode != 0 && (mode == 2 || pays)) {
                const int id = hs_cu_jit_rader_create(L, rx, m, &jd);
                if (id >= 0) { ct = &jd; jit = 1; }
            }
        }
    }
    if (!ct) return 0;
    /* smallest primitive root: g^(L/p) != 1 for every prime p | L */
    int pf[16], npf = 0, m = L;
    for (int f = 2; (long long)f * f <= m; ++f)
        if (m % f == 0) { pf[npf++] = f; while (m % f == 0) m /= f; }
    if (m > 1) pf[npf++] = m;
    int g = 0;
    for (int c = 2; c < N && !g; ++c) {
        int ok = 1;
        for (int i = 0; i < npf && ok; ++i) {
            long long r = 1, b = c;
            for (int e = L / pf[i]; e; e >>= 1, b = b * b % N)
                if (e & 1) r = r * b % N;
            if (r == 1) ok = 0;
        }
        if (ok) g = c;
    }
    if (!g) return 0;
    int *perm = (int *)malloc(sizeof(int) * L);
    fft_data *bk = (fft_data *)malloc(sizeof(fft_data) * L);
    long double *wr = (long double *)malloc(sizeof(long double) * 2 * L), *br = (long double *)malloc(sizeof(long double) * 2 * L);
    if (!perm || !bk || !wr || !br) { fprintf(stderr, "Error: Memory allocation failed for Rader's tables\n"); exit(EXIT_FAILURE); }
    long long v = 1;
    for (int q = 0; q < L; ++q) { perm[q] = (int)v; v = v * g % N; }
    const long double pi2 = 6.283185307179586476925286766559005768L;
    for (int q = 0; q < L; ++q) { /* b[q] = W^(g^-q) = W^perm[(L-q) % L];  w[q] = exp(-2 pi i q / L) */
        const long double a = -(long double)d->sgn * pi2 * (long double)perm[(L - q) % L] / (long double)N;
        br[2 * q] = cosl(a); br[2 * q + 1] = sinl(a);
        const long double t = -pi2 * (long double)q / (long double)L;
        wr[2 * q] = cosl(t); wr[2 * q + 1] = sinl(t);
    }
    for (int k = 0; k < L; ++k) {
        long double sr = 0, si = 0;
        int j = 0;
        for (int q = 0; q < L; ++q) {
            sr += br[2 * q] * wr[2 * j] - br[2 * q + 1] * wr[2 * j + 1];
            si += br[2 * q] * wr[2 * j + 1] + br[2 * q + 1] * wr[2 * j];
            j += k; if (j >= L) j -= L;
        }
        bk[k].re = (fft_type)(sr / L); bk[k].im = (fft_type)(si / L);
    }
    int inv[8];
    for (int s = 0; s < ct->nst; ++s) inv[s] = ct->radix[ct->nst - 1 - s];
    d->rader_ct_id = ct->id;
    d->rader_jit = jit;
    d->M = L;
    d->d_chirp = upload(perm, sizeof(int) * L);
    d->d_hk = upload(bk, sizeof(fft_data) * L);
    d->d_blu_twf = upload_stage_twiddles(ct->radix, ct->nst, 1, ct->split_min);
    d->d_blu_twi = upload_stage_twiddles(inv, ct->nst, -1, ct->split_min);
    hs_cu_stream_sync(NULL);
    free(perm); free(bk); free(wr); free(br);
    return 1;
}

static void dev_build(hs_dev_obj *d, fft_object plan)
{
    d->key = plan;
    d->N = plan->N;
    d->sgn = plan->sgn;
    d->lt = plan->lt;
    d->blu_ct_id = -1;
    d->rader_ct_id = -1;
    d->split_ct_id = -1;
    if (d->N < 2 || (d->sgn != 1 && d->sgn != -1)) {
        fprintf(stderr, "Error: Invalid FFT object (N = %d, sgn = %d)\n", d->N, d->sgn);
        exit(EXIT_FAILURE);
    }
    for (int s = 0; s < HS_NSLOT; ++s) d->stream[s] = hs_cu_stream_create();
    if (d->lt == 0) { d->direct = hs_c2c_create(d->N, d->sgn); return; }
    if (rader_setup(d)) return;

    /* Bluestein: chirp h and Hk = FFT_M(h padded and mirrored, scaled 1/M) live in HBM */
    const int N = d->N, M = bluestein_device_pad(N);
    d->M = M;
    fft_data *h = (fft_data *)malloc(sizeof(fft_data) * N), *hp = (fft_data *)calloc(M, sizeof(fft_data));
    if (!h || !hp) { fprintf(stderr, "Error: Memory allocation failed for Bluestein's FFT arrays\n"); exit(EXIT_FAILURE); }
    hs_build_chirp(h, N);
    const fft_type scale = 1.0 / M;
    for (int n = 0; n < N; ++n) { hp[n].re = h[n].re * scale; hp[n].im = h[n].im * scale; }
    for (int n = M - N + 1; n < M; ++n) hp[n] = hp[M - n]; /* reference :1698-1703 */
    d->d_chirp = upload(h, sizeof(fft_data) * N);
    void *d_hp = upload(hp, sizeof(fft_data) * M);
    d->d_hk = hs_cu_malloc(sizeof(fft_data) * M);
    hs_c2c *f = c2c_create(M, 1, 1);
    void *ws = NULL;
    const size_t wse = hs_c2c_ws_elems(f, 1);
    if (wse) ws = hs_cu_malloc(wse * sizeof(fft_data));
    hs_c2c_exec(f, d_hp, d->d_hk, ws, 1, NULL);
    hs_cu_stream_sync(NULL);
    hs_cu_free(ws);
    hs_cu_free(d_hp);
    const hs_ct_desc *sp = bluestein_split_plan(N);
    if (sp && 2 * sp->n == M) {
        /* tables of the split kernel: h | gA = h conj(w) | gB = h w with w[n] = exp(-2 pi i n / M), and Hk de-interleaved */
        const int Mh = sp->n;
        fft_data *c3 = (fft_data *)malloc(sizeof(fft_data) * 3 * (size_t)N), *hk = (fft_data *)malloc(sizeof(fft_data) * (size_t)M);
        if (!c3 || !hk) { fprintf(stderr, "Error: Memory allocation failed for Bluestein's FFT arrays\n"); exit(EXIT_FAILURE); }
        hs_cu_d2h(hk, d->d_hk, sizeof(fft_data) * (size_t)M, NULL);
        hs_cu_stream_sync(NULL);
        for (int k = 0; k < Mh; ++k) { hp[k] = hk[2 * k]; hp[Mh + k] = hk[2 * k + 1]; }
        if (sp->family == HS_CT_CONV_SPLIT) { /* both halves in the digit-reversed order of the in-place transforms */
            memcpy(hk, hp, sizeof(fft_data) * (size_t)M);
            conv_permute_hk(hp, hk, sp->radix, sp->nst, Mh);
            conv_permute_hk(hp + Mh, hk + Mh, sp->radix, sp->nst, Mh);
        }
        hs_cu_h2d(d->d_hk, hp, sizeof(fft_data) * (size_t)M, NULL);
        for (int n = 0; n < N; ++n) {
            const fft_type a = -PI2 * (fft_type)n / (fft_type)M, wr = cos(a), wi = sin(a);
            c3[n] = h[n];
            c3[N + n].re = h[n].re * wr + h[n].im * wi;     c3[N + n].im = h[n].im * wr - h[n].re * wi;     /* h * conj(w) */
            c3[2 * N + n].re = h[n].re * wr - h[n].im * wi; c3[2 * N + n].im = h[n].im * wr + h[n].re * wi; /* h * w */
        }
        hs_cu_free(d->d_chirp);
        d->d_chirp = upload(c3, sizeof(fft_data) * 3 * (size_t)N);
        int inv[8];
        for (int s = 0; s < sp->nst; ++s) inv[s] = sp->radix[sp->nst - 1 - s];
        d->split_ct_id = sp->id;
        d->split_conv = sp->family == HS_CT_CONV_SPLIT;
        if (d->split_conv) {
            d->d_blu_twf = upload_conv_twiddles(sp->radix, sp->nst, Mh);
            /* w[p] = exp(-2 pi i p / M), p < M' / r0: the kernel that parks its intermediates in tensor memory forms every
             * w[p + (M'/r0) i] from it and compile-time constants */
            const int m0 = Mh / sp->radix[0];
            fft_data *wp = (fft_data *)malloc(sizeof(fft_data) * (size_t)m0);
            if (!wp) { fprintf(stderr, "Error: Memory allocation failed for Bluestein's FFT arrays\n"); exit(EXIT_FAILURE); }
            for (int q = 0; q < m0; ++q) { const fft_type a = -PI2 * (fft_type)q / (fft_type)M; wp[q].re = cos(a); wp[q].im = sin(a); }
            d->d_blu_twi = upload(wp, sizeof(fft_data) * (size_t)m0);
            hs_cu_stream_sync(NULL);
            free(wp);
        }
        else {
            d->d_blu_twf = upload_stage_twiddles(sp->radix, sp->nst, 1, sp->split_min);
            d->d_blu_twi = upload_stage_twiddles(inv, sp->nst, -1, sp->split_min);
        }
        hs_cu_stream_sync(NULL);
        free(c3); free(hk); free(h); free(hp);
        hs_c2c_destroy(f);
        return;
    }
    free(h);
    free(hp);

    const hs_ct_desc *tp = find_ct_ex(HS_CT_TWOPASS, M, 1);
    const hs_ct_desc *s1 = tp ? find_ct_ex(HS_CT_SUB, tp->n1, 1) : NULL, *s2 = tp ? find_ct_ex(HS_CT_SUB, tp->n2, 1) : NULL;
    if (tp && s1 && s2 && !getenv("HSFFT_NO_BIGBLU")) {
        /* long Bluestein: A = chirp/pad + forward pass 1, B = forward pass 2 * Hk + inverse pass 1, C = inverse
         * pass 2 + chirp/crop.  Forward decomposition n1 x n2, inverse n2 x n1, so B works column by column. */
        const int n1 = tp->n1, n2 = tp->n2;
        int inv2[8];
        for (int s = 0; s < s2->nst; ++s) inv2[s] = s2->radix[s2->nst - 1 - s];
        d->big = 1; d->big_n1 = n1; d->big_n2 = n2; d->big_sub1 = s1->id; d->big_sub2 = s2->id;
        d->big_tp_id = tp->id;
        d->d_big_f1 = upload_stage_twiddles(s1->radix, s1->nst, 1, s1->split_min);
        d->d_big_f2 = upload_stage_twiddles(s2->radix, s2->nst, 1, s2->split_min);
        d->d_big_i1 = upload_stage_twiddles(inv2, s2->nst, -1, s2->split_min);
        d->d_big_i2 = upload_stage_twiddles(s1->radix, s1->nst, -1, s1->split_min);
        const int NBf = n2 / s2->radix[0], NBi = n1 / s1->radix[0];
        fft_data *fb = (fft_data *)malloc(sizeof(fft_data) * (size_t)NBf * n1), *fs = (fft_data *)malloc(sizeof(fft_data) * n1);
        fft_data *ib = (fft_data *)malloc(sizeof(fft_data) * (size_t)NBi * n2), *is = (fft_data *)malloc(sizeof(fft_data) * n2);
        if (!fb || !fs || !ib || !is) { fprintf(stderr, "Error: Memory allocation failed for twiddle table\n"); exit(EXIT_FAILURE); }
        const fft_type theta = -PI2 / (fft_type)M;
        for (int b = 0; b < NBf; ++b)
            for (int k = 0; k < n1; ++k) put_twiddle(&fb[(size_t)b * n1 + k], k * b, theta, 1);
        for (int k = 0; k < n1; ++k) put_twiddle(&fs[k], k * NBf, theta, 1);
        for (int b = 0; b < NBi; ++b)
            for (int k = 0; k < n2; ++k) put_twiddle(&ib[(size_t)b * n2 + k], k * b, theta, -1);
        for (int k = 0; k < n2; ++k) put_twiddle(&is[k], k * NBi, theta, -1);
        d->d_big_fb = upload(fb, sizeof(fft_data) * (size_t)NBf * n1);
        d->d_big_fs = upload(fs, sizeof(fft_data) * n1);
        d->d_big_ib = upload(ib, sizeof(fft_data) * (size_t)NBi * n2);
        d->d_big_is = upload(is, sizeof(fft_data) * n2);
        hs_cu_stream_sync(NULL);
        free(fb); free(fs); free(ib); free(is);
        /* chunk the batch so that both M-sized workspaces of a chunk stay resident in the 126 MB L2 */
        long long mb = 1 << 20; /* measured: chunks small enough for L2 (48 MB) underfill the GPU; default = no chunking */
        if (getenv("HSFFT_BLU_CHUNK_MB")) mb = atoll(getenv("HSFFT_BLU_CHUNK_MB"));
        d->big_chunk = (mb << 20) / ((long long)2 * M * (long long)sizeof(fft_data));
        if (d->big_chunk < 1) d->big_chunk = 1;
        hs_c2c_destroy(f);
        return;
    }
    const hs_ct_desc *ct = find_ct(HS_CT_BLUESTEIN, M);
    if (ct) {
        /* fused kernel: forward plan = ct->radix, inverse plan = the same list reversed */
        int inv[8];
        for (int s = 0; s < ct->nst; ++s) inv[s] = ct->radix[ct->nst - 1 - s];
        d->blu_ct_id = ct->id;
        d->d_blu_twf = upload_stage_twiddles(ct->radix, ct->nst, 1, ct->split_min);
        d->d_blu_twi = upload_stage_twiddles(inv, ct->nst, -1, ct->split_min);
        hs_c2c_destroy(f);
    } else {
        d->fwdM = f;
        d->invM = c2c_create(M, -1, 1);
    }
}

void *hs_scratch_get(hs_scratch *s, void *stream, size_t bytes)
{
    int k = 0;
    while (k < s->used && s->stream[k] != stream) ++k;
    if (k == s->used) {
        if (s->used == HS_SCRATCH_MAX) { /* more caller streams than slots: quiesce the device and start over */
            hs_cu_device_sync();
            hs_scratch_free(s);
            k = 0;
        }
        s->stream[k] = stream; s->buf[k] = NULL; s->cap[k] = 0;
        s->used = k + 1;
    }
    if (bytes > s->cap[k]) {
        hs_cu_free_async(s->buf[k], stream);
        s->buf[k] = hs_cu_malloc_async(bytes, stream);
        s->cap[k] = bytes;
    }
    return s->buf[k];
}

void hs_scratch_free(hs_scratch *s)
{
    for (int k = 0; k < s->used; ++k) hs_cu_free(s->buf[k]);
    memset(s, 0, sizeof *s);
}

/* a plan that was free()d directly and whose address came back with other fields, or whose public struct was edited:
 * every device's state under that address is stale */
static void dev_drop_key_locked(const void *key)
{
    hs_dev_obj **pp = &g_head;
    while (*pp) {
        hs_dev_obj *d = *pp;
        if (d->key == key) {
            *pp = d->next;
            dev_free_contents(d);
            pthread_mutex_destroy(&d->lock);
            free(d);
        } else pp = &d->next;
    }
}

hs_dev_obj *hs_dev_get(fft_object plan)
{
    if (plan == NULL) { fprintf(stderr, "Error: Invalid FFT object or data pointers\n"); exit(EXIT_FAILURE); }
    hs_cu_require_device();
    const int dev = hs_cu_get_device();
    pthread_mutex_lock(&g_lock);
    hs_dev_obj *d = g_head;
    while (d && !(d->key == plan && d->device == dev)) d = d->next;
    if (d && (d->N != plan->N || d->sgn != plan->sgn || d->lt != plan->lt)) {
        dev_drop_key_locked(plan);
        d = NULL;
    }
    if (!d) {
        d = (hs_dev_obj *)calloc(1, sizeof(hs_dev_obj));
        if (!d) { fprintf(stderr, "Error: Memory allocation failed for device plan\n"); exit(EXIT_FAILURE); }
        pthread_mutexattr_t at;
        pthread_mutexattr_init(&at);
        pthread_mutexattr_settype(&at, PTHREAD_MUTEX_RECURSIVE);
        pthread_mutex_init(&d->lock, &at);
        pthread_mutexattr_destroy(&at);
        d->device = dev;
        dev_build(d, plan);
        d->next = g_head;
        g_head = d;
    }
    pthread_mutex_unlock(&g_lock);
    return d;
}

void hs_dev_lock(hs_dev_obj *d) { pthread_mutex_lock(&d->lock); }
void hs_dev_unlock(hs_dev_obj *d) { pthread_mutex_unlock(&d->lock); }

void hs_dev_release(const void *key)
{
    pthread_mutex_lock(&g_lock);
    dev_drop_key_locked(key);
    pthread_mutex_unlock(&g_lock);
}

static void grow(void **buf, size_t *cap, size_t need)
{
    if (need <= *cap) return;
    hs_cu_free(*buf);
    *buf = hs_cu_malloc(need);
    *cap = need;
}

void hs_dev_reserve(hs_dev_obj *d, int slot, size_t in_bytes, size_t out_bytes, size_t ws_bytes, size_t ws2_bytes)
{
    grow(&d->d_in[slot], &d->cap_in[slot], in_bytes);
    grow(&d->d_out[slot], &d->cap_out[slot], out_bytes);
    grow(&d->d_ws[slot], &d->cap_ws[slot], ws_bytes);
    grow(&d->d_ws2[slot], &d->cap_ws2[slot], ws2_bytes);
}

/* ----------------------------------------------------------------------------- multi-device sharding */

static int g_ndev = 1; /* hsfft_set_devices: the host-pointer batch calls shard their batch over devices 0 .. g_ndev-1 */

void hsfft_set_devices(int n)
{
    const int have = hs_cu_device_count();
    if (n < 1) n = 1;
    if (n > have) n = have > 0 ? have : 1;
    g_ndev = n;
}
int hsfft_get_devices(void) { return g_ndev; }

typedef struct shard_job { hs_shard_fn fn; void *ctx; long first, count; int dev; } shard_job;
static void *shard_thread(void *arg)
{
    shard_job *j = (shard_job *)arg;
    hs_cu_set_device(j->dev);
    j->fn(j->ctx, j->first, j->count);
    return NULL;
}

void hs_shard_run(long howmany, hs_shard_fn fn, void *ctx)
{
    const int nd = g_ndev;
    if (nd <= 1 || howmany < 2L * nd) { fn(ctx, 0, howmany); return; }
    /* contiguous batch ranges, one host thread per device (no communication) */
    shard_job job[64];
    pthread_t th[64];
    const int n = nd > 64 ? 64 : nd;
    const int home = hs_cu_get_device();
    long first = 0;
    for (int g = 0; g < n; ++g) {
        const long cnt = howmany / n + (g < howmany % n ? 1 : 0);
        job[g].fn = fn; job[g].ctx = ctx; job[g].first = first; job[g].count = cnt; job[g].dev = g;
        first += cnt;
    }
    for (int g = 1; g < n; ++g)
        if (pthread_create(&th[g], NULL, shard_thread, &job[g]) != 0) { fprintf(stderr, "Error: cannot start a device thread\n"); exit(EXIT_FAILURE); }
    shard_thread(&job[0]);
    for (int g = 1; g < n; ++g) pthread_join(th[g], NULL);
    hs_cu_set_device(home);
}

size_t hs_dev_ws_bytes(const hs_dev_obj *d, long long howmany)
{
    if (d->lt == 0) return hs_c2c_ws_elems(d->direct, howmany) * sizeof(fft_data);
    if (d->blu_ct_id >= 0 || d->rader_ct_id >= 0 || d->split_ct_id >= 0) return 0;
    if (d->big) return (size_t)2 * (howmany < d->big_chunk ? howmany : d->big_chunk) * d->M * sizeof(fft_data);
    return ((size_t)2 * howmany * d->M + hs_c2c_ws_elems(d->fwdM, howmany)) * sizeof(fft_data);
}

void hs_dev_exec(hs_dev_obj *d, const void *d_in, void *d_out, void *d_ws, long long howmany, void *stream)
{
    if (howmany <= 0) return;
    if (d->lt == 0) { hs_c2c_exec(d->direct, d_in, d_out, d_ws, howmany, stream); return; }
    if (d->rader_ct_id >= 0 && d->rader_jit) {
        hs_cu_jit_rader_launch(d->rader_ct_id, d_in, d_out, d->d_blu_twf, d->d_blu_twi, d->d_chirp, d->d_hk, howmany, stream);
        return;
    }
    if (d->rader_ct_id >= 0) {
        hs_cu_launch_ct_rader(d->rader_ct_id, d_in, d_out, d->d_blu_twf, d->d_blu_twi, d->d_chirp, d->d_hk, howmany, stream);
        return;
    }
    if (d->split_ct_id >= 0 && d->split_conv) {
        hs_cu_launch_ct_conv_split(d->split_ct_id, d->sgn, d->N, d_in, d_out, d->d_blu_twf, d->d_chirp, d->d_blu_twi, d->d_hk, howmany, stream);
        return;
    }
    if (d->split_ct_id >= 0) {
        hs_cu_launch_ct_bluestein_split(d->split_ct_id, d->sgn, d->N, d_in, d_out, d->d_blu_twf, d->d_blu_twi, d->d_chirp, d->d_hk, howmany,
                                        stream);
        return;
    }
    if (d->blu_ct_id >= 0) {
        hs_cu_launch_ct_bluestein(d->blu_ct_id, d->sgn, d->N, d_in, d_out, d->d_blu_twf, d->d_blu_twi, d->d_chirp, d->d_hk, howmany,
                                  stream);
        return;
    }
    if (d->big) {
        const long long M = d->M, N = d->N, n1 = d->big_n1, n2 = d->big_n2;
        const long long chunk = howmany < d->big_chunk ? howmany : d->big_chunk;
        fft_data *ws1 = (fft_data *)d_ws, *ws2 = ws1 + (size_t)chunk * M;
        hs_tile_geom ga, gb, gc;
        memset(&ga, 0, sizeof ga); memset(&gb, 0, sizeof gb); memset(&gc, 0, sizeof gc);
        ga.ncols = (int)n2; ga.in_xf = N; ga.in_c = 1; ga.in_e = n2; ga.out_xf = M; ga.out_c = n1; ga.out_e = 1;
        ga.chirp = d->d_chirp; ga.nlog = (int)N;
        gb.ncols = (int)n1; gb.in_xf = M; gb.in_c = 1; gb.in_e = n1; gb.out_xf = M; gb.out_c = n2; gb.out_e = 1;
        gb.twb = d->d_big_fb; gb.tws = d->d_big_fs; gb.tw_ld = (int)n1;
        gc.ncols = (int)n2; gc.in_xf = M; gc.in_c = 1; gc.in_e = n2; gc.out_xf = N; gc.out_c = 1; gc.out_e = n2;
        gc.twb = d->d_big_ib; gc.tws = d->d_big_is; gc.tw_ld = (int)n2; gc.chirp = d->d_chirp; gc.nlog = (int)N;
        for (long long first = 0; first < howmany; first += chunk) {
            const long long cnt = (first + chunk <= howmany) ? chunk : howmany - first;
            const fft_data *x = (const fft_data *)d_in + first * N;
            fft_data *o = (fft_data *)d_out + first * N;
            hs_cu_launch_tiled_io(d->big_sub1, 1, 0, 0, 1, d->sgn, x, ws1, d->d_big_f1, &ga, cnt, stream);
            hs_cu_launch_blu_mid(d->big_sub2, d->sgn, ws1, ws2, d->d_big_f2, d->d_big_i1, &gb, d->d_hk, cnt, stream);
            hs_cu_launch_tiled_io(d->big_sub1, -1, 1, 1, 2, d->sgn, ws2, o, d->d_big_i2, &gc, cnt, stream);
        }
        return;
    }
    /* unfused chirp-z: pad*chirp -> FFT_M -> *Hk -> IFFT_M -> *chirp, crop */
    fft_data *a = (fft_data *)d_ws, *b = a + (size_t)howmany * d->M, *c = b + (size_t)howmany * d->M;
    hs_cu_launch_chirp_pad(d->sgn, d_in, a, d->d_chirp, d->N, d->M, howmany, stream);
    hs_c2c_exec(d->fwdM, a, b, c, howmany, stream);
    hs_cu_launch_spectrum_mul(d->sgn, b, d->d_hk, d->M, howmany, stream);
    hs_c2c_exec(d->invM, b, a, c, howmany, stream);
    hs_cu_launch_chirp_crop(d->sgn, a, d_out, d->d_chirp, d->N, d->M, howmany, stream);
}

/* r2c with the split fused into the complex kernel (one launch, no half-length spectrum in HBM): one-kernel plans,
 * static or compiled for this plan; returns 0 when this plan has no such form (the caller runs transform + split kernel) */
int hs_dev_exec_r2c_fused(hs_dev_obj *d, const void *d_in, void *d_out, long long out_stride, int full, const void *d_tw2, long long howmany,
                          void *stream)
{
    if (d->lt != 0 || !d->direct || getenv("HSFFT_NO_R2C_FUSE")) return 0;
    const hs_c2c *p = d->direct;
    if (p->ct_id >= 0 && hs_cu_ct_has_real(p->ct_id)) {
        hs_cu_launch_ct_real(p->ct_id, 1, p->sgn, d_in, d_out, p->d_tw, d_tw2, out_stride, full, howmany, stream);
        return 1;
    }
    if (p->jit_id >= 0 && hs_cu_jit_has_real(p->jit_id)) {
        hs_cu_jit_real_launch(p->jit_id, 1, d_in, d_out, p->d_tw, d_tw2, out_stride, full, howmany, stream);
        return 1;
    }
    return 0;
}

/* twiddle2[c + n2 r] = exp(2 pi i (c + n2 r) / 2n) as a product of a column and a row factor (HsLoadColC2RPre2, HsLoadColC2RMul)
 * where the table no longer sits close to the SMs: measured on B200 per 2^28 reals, merge pass with the product / the table:
 * N = 2^20 1.11 / 1.17 ms, 2^18 0.87 / 0.97, 2^19 0.98 / 0.98, but 2^16 0.92 / 0.85 and 2^15 0.91 / 0.84 */
static void c2r_merge_factors(const hs_c2c *p, const hs_ct_desc *tp)
{
    if (p->d_mrg_wc || p->n < (1 << 17) || getenv("HSFFT_C2R_TW2_TABLE")) return;
    hs_c2c *pw = (hs_c2c *)p;
    fft_data *wc = (fft_data *)malloc(sizeof(fft_data) * tp->n2), *wr = (fft_data *)malloc(sizeof(fft_data) * tp->n1);
    if (!wc || !wr) { fprintf(stderr, "Error: Memory allocation failed for twiddle table\n"); exit(EXIT_FAILURE); }
    const fft_type theta = PI2 / (fft_type)(2 * (long long)p->n);
    for (int c = 0; c < tp->n2; ++c) { wc[c].re = cos(theta * c); wc[c].im = sin(theta * c); }
    for (int r = 0; r < tp->n1; ++r) { wr[r].re = cos(theta * ((fft_type)tp->n2 * r)); wr[r].im = sin(theta * ((fft_type)tp->n2 * r)); }
    pw->d_mrg_wc = upload(wc, sizeof(fft_data) * tp->n2);
    pw->d_mrg_wr = upload(wr, sizeof(fft_data) * tp->n1);
    hs_cu_stream_sync(NULL);
    free(wc); free(wr);
}

int hs_dev_exec_c2r_fused(hs_dev_obj *d, const void *d_spec, long long in_stride, const void *d_tw2, void *d_out, void *d_ws,
                          long long howmany, void *stream)
{
    if (d->lt != 0 || !d->direct || getenv("HSFFT_NO_C2R_FUSE")) return 0;
    if (d->direct->ct_id >= 0 && hs_cu_ct_has_real(d->direct->ct_id)) { /* one kernel: the merge rides on the first pass's loads */
        hs_cu_launch_ct_real(d->direct->ct_id, 2, d->direct->sgn, d_spec, d_out, d->direct->d_tw, d_tw2, in_stride, 0, howmany, stream);
        return 1;
    }
    if (d->direct->jit_id >= 0 && hs_cu_jit_has_real(d->direct->jit_id)) {
        hs_cu_jit_real_launch(d->direct->jit_id, 2, d_spec, d_out, d->direct->d_tw, d_tw2, in_stride, 0, howmany, stream);
        return 1;
    }
    if (d->direct->tp_id < 0) return 0;
    const hs_c2c *p = d->direct;
    const hs_ct_desc *tp = NULL;
    for (int i = 0; i < hs_cu_ct_count(); ++i) if (hs_cu_ct_get(i)->id == p->tp_id) tp = hs_cu_ct_get(i);
    const hs_ct_desc *s1 = tp ? find_sub_like(tp->n1, tp->radix, tp->nst) : NULL;
    if (!tp || !s1) return 0;
    hs_tile_geom g;
    memset(&g, 0, sizeof g);
    g.ncols = tp->n2; g.in_xf = in_stride; g.in_c = 1; g.in_e = tp->n2; g.out_xf = p->n; g.out_c = tp->n1; g.out_e = 1;
    g.chirp = d_tw2; g.nlog = p->n;
    c2r_merge_factors(p, tp);
    g.twb = p->d_mrg_wc; g.tws = p->d_mrg_wr;
    hs_cu_launch_tiled_io(s1->id, p->sgn, 0, 0, 3, 1, d_spec, d_ws, p->d_tw, &g, howmany, stream);
    hs_cu_launch_ct_twopass(p->tp_id, 2, p->sgn, d_ws, d_out, p->d_tw2, p->d_twb, p->d_tws, howmany, stream);
    return 1;
}

/* two-pass plan lookup shared by the fused real / convolution paths */
static int twopass_subs(const hs_c2c *p, const hs_ct_desc **tp, const hs_ct_desc **s1, const hs_ct_desc **s2)
{
    *tp = NULL;
    for (int i = 0; i < hs_cu_ct_count(); ++i) if (hs_cu_ct_get(i)->id == p->tp_id) *tp = hs_cu_ct_get(i);
    *s1 = *tp ? find_sub_like((*tp)->n1, (*tp)->radix, (*tp)->nst) : NULL;
    *s2 = *tp ? find_sub_like((*tp)->n2, (*tp)->radix2, (*tp)->nst2) : NULL;
    return *tp && *s1 && *s2;
}

int hs_dev_exec_padded_real(hs_dev_obj *d, const void *d_rows, int len, long long row_stride, long long row_off, long long total, void *d_z,
                            void *d_ws, long long howmany, void *stream)
{
    /* windows of len reals (window x starts at real x*row_stride + row_off of an array of `total` reals), zero padded to
     * 2*n reals on the fly, through the half-length plan: the padding rides on pass 1's loads */
    if (d->lt != 0 || !d->direct || d->direct->tp_id < 0 || getenv("HSFFT_NO_CONV_FUSE")) return 0;
    const hs_c2c *p = d->direct;
    const hs_ct_desc *tp, *s1, *s2;
    if (!twopass_subs(p, &tp, &s1, &s2)) return 0;
    hs_tile_geom g;
    memset(&g, 0, sizeof g);
    g.ncols = tp->n2; g.in_xf = 0; g.in_c = 1; g.in_e = tp->n2; g.out_xf = p->n; g.out_c = tp->n1; g.out_e = 1;
    g.nlog = len; g.pad_stride = row_stride; g.pad_off = row_off; g.pad_total = total;
    hs_cu_launch_tiled_io(s1->id, p->sgn, 0, 0, 4, 1, d_rows, d_ws, p->d_tw, &g, howmany, stream);
    hs_cu_launch_ct_twopass(p->tp_id, 2, p->sgn, d_ws, d_z, p->d_tw2, p->d_twb, p->d_tws, howmany, stream);
    return 1;
}

int hs_dev_exec_c2r_mul_slice(hs_dev_obj *d, const void *d_A, const void *d_B, long long in_stride, long long b_stride, double scale,
                              const void *d_tw2, void *d_out, long long out_stride, long long out_total, int start, int len, void *d_ws,
                              long long howmany, void *stream)
{
    /* inverse real transform of scale * A * B, keeping reals [start, start+len) of every row: the product rides on the
     * merge loads of pass 1, the slice on the stores of pass 2 */
    if (d->lt != 0 || !d->direct || d->direct->tp_id < 0 || getenv("HSFFT_NO_CONV_FUSE")) return 0;
    const hs_c2c *p = d->direct;
    const hs_ct_desc *tp, *s1, *s2;
    if (!twopass_subs(p, &tp, &s1, &s2)) return 0;
    hs_tile_geom g;
    memset(&g, 0, sizeof g);
    g.ncols = tp->n2; g.in_xf = in_stride; g.in_c = 1; g.in_e = tp->n2; g.out_xf = p->n; g.out_c = tp->n1; g.out_e = 1;
    g.chirp = d_tw2; g.nlog = p->n; g.aux = d_B; g.aux_xf = b_stride; g.scale = scale;
    c2r_merge_factors(p, tp);
    g.twb = p->d_mrg_wc; g.tws = p->d_mrg_wr;
    hs_cu_launch_tiled_io(s1->id, p->sgn, 0, 0, 5, 1, d_A, d_ws, p->d_tw, &g, howmany, stream);
    memset(&g, 0, sizeof g);
    g.ncols = tp->n1; g.in_xf = p->n; g.in_c = 1; g.in_e = tp->n1; g.out_xf = 0; g.out_c = 1; g.out_e = tp->n1;
    g.twb = p->d_twb; g.tws = p->d_tws; g.tw_ld = tp->n1;
    g.aux_xf = out_stride; g.sl_start = start; g.sl_len = len; g.sl_total = out_total;
    hs_cu_launch_tiled_io(s2->id, p->sgn, 1, 1, 6, 1, d_ws, d_out, p->d_tw2, &g, howmany, stream);
    return 1;
}

/* ---------------------------------------------------------------------------------- public API */

fft_object fft_init(int N, int sgn)
{
    if (N <= 0) { /* undefined behaviour in the reference (segfault) */
        fprintf(stderr, "Error: Signal length (%d) must be positive\n", N);
        exit(EXIT_FAILURE);
    }
    fft_object obj = NULL;
    int count;
    if (dividebyN(N)) {
        obj = (fft_object)malloc(sizeof(struct fft_set) + sizeof(fft_data) * (N - 1));
        if (obj == NULL) return NULL;
        obj->lf = factors(N, obj->factors);
        longvectorN(obj->twiddle, N, obj->factors, obj->lf); /* exits for N == 1, like the reference */
        count = N;
        obj->lt = 0;
    } else {
        /* pad length recorded in the public struct: the reference's plan-time estimate (:242-252) */
        const int K = (int)pow(2.0, ceil(log10(N) / log10(2.0)));
        int M = (K < 2 * N - 2) ? K * 2 : K;
        const int Mexec = hs_bluestein_exec_pad(N);
        if (M < Mexec) M = Mexec; /* N = 2^k+1: the reference under-allocates here (its bug 4) */
        obj = (fft_object)malloc(sizeof(struct fft_set) + sizeof(fft_data) * (M - 1));
        if (obj == NULL) return NULL;
        obj->lf = factors(M, obj->factors);
        longvectorN(obj->twiddle, M, obj->factors, obj->lf);
        count = M;
        obj->lt = 1;
    }
    obj->N = N;
    obj->sgn = sgn;
    if (sgn == -1)
        for (int i = 0; i < count; ++i) obj->twiddle[i].im = -obj->twiddle[i].im;
    /* A stale registry entry under this address (plan free()d directly by its owner) is dropped
     * lazily: hs_dev_get compares (N, sgn, lt) on the first execute. */
    return obj;
}

static long long chunk_len(long long howmany, size_t bytes_per_xf)
{
    static size_t target = 0; /* bytes per chunk of the host-pointer pipeline (HSFFT_CHUNK_MB, default 64) */
    if (!target) target = (size_t)env_int("HSFFT_CHUNK_MB", 64) << 20;
    long long c = (long long)(target / (bytes_per_xf ? bytes_per_xf : 1));
    if (c < 1) c = 1;
    if (c > howmany) c = howmany;
    return c;
}

typedef struct exec_ctx { fft_object obj; fft_data *inp, *oup; } exec_ctx;

/* one device's share of a host-pointer batch: chunked three-deep pipeline on the plan's own streams, H2D(i+1) and D2H(i-1)
 * overlap the kernels of chunk i.  The plan's staging buffers are in use until the last copy has landed, so the whole
 * call holds the (plan, device) lock: concurrent callers of one plan serialise instead of overwriting each other. */
static void exec_batch_one_device(void *vctx, long first0, long howmany)
{
    exec_ctx *c = (exec_ctx *)vctx;
    if (howmany <= 0) return;
    hs_dev_obj *d = hs_dev_get(c->obj);
    hs_dev_lock(d);
    const size_t row = sizeof(fft_data) * (size_t)d->N;
    const long long chunk = chunk_len(howmany, row);
    const long long nchunk = (howmany + chunk - 1) / chunk;
    const int nslot = nchunk < HS_NSLOT ? (int)nchunk : HS_NSLOT;
    fft_data *inp = c->inp + (size_t)first0 * d->N, *oup = c->oup + (size_t)first0 * d->N;
    for (int s = 0; s < nslot; ++s) hs_dev_reserve(d, s, row * chunk, row * chunk, hs_dev_ws_bytes(d, chunk), 0);
    for (long long i = 0; i < nchunk; ++i) {
        const int s = (int)(i % HS_NSLOT);
        const long long first = i * chunk, cnt = (first + chunk <= howmany) ? chunk : howmany - first;
        hs_cu_h2d(d->d_in[s], inp + first * d->N, row * cnt, d->stream[s]);
        hs_dev_exec(d, d->d_in[s], d->d_out[s], d->d_ws[s], cnt, d->stream[s]);
        hs_cu_d2h(oup + first * d->N, d->d_out[s], row * cnt, d->stream[s]);
    }
    for (int s = 0; s < nslot; ++s) hs_cu_stream_sync(d->stream[s]);
    hs_dev_unlock(d);
}

void fft_exec_batch(fft_object obj, fft_data *inp, fft_data *oup, long howmany)
{
    if (obj == NULL || inp == NULL || oup == NULL) {
        fprintf(stderr, "Error: Invalid FFT object or data pointers\n");
        exit(EXIT_FAILURE);
    }
    if (obj->lt != 0 && obj->lt != 1) {
        fprintf(stderr, "Error: Invalid FFT object type (lt = %d)\n", obj->lt);
        exit(EXIT_FAILURE);
    }
    if (howmany <= 0) return;
    exec_ctx c = {obj, inp, oup};
    hs_shard_run(howmany, exec_batch_one_device, &c); /* one device unless hsfft_set_devices(n > 1) */
}

void fft_exec(fft_object obj, fft_data *inp, fft_data *oup)
{
    fft_exec_batch(obj, inp, oup, 1);
}

void fft_exec_batch_dev(fft_object obj, const fft_data *d_in, fft_data *d_out, long howmany, void *stream)
{
    if (obj == NULL || d_in == NULL || d_out == NULL) {
        fprintf(stderr, "Error: Invalid FFT object or data pointers\n");
        exit(EXIT_FAILURE);
    }
    if ((const void *)d_in == (const void *)d_out) { /* like the reference, every transform here is out of place */
        fprintf(stderr, "Error: fft_exec_batch_dev is out of place: input and output must not be the same buffer\n");
        exit(EXIT_FAILURE);
    }
    if (howmany <= 0) return;
    hs_dev_obj *d = hs_dev_get(obj);
    hs_dev_lock(d);
    void *ws = hs_scratch_get(&d->ws_pool, stream, hs_dev_ws_bytes(d, howmany));
    hs_dev_exec(d, d_in, d_out, ws, howmany, stream);
    hs_dev_unlock(d);
}

/* In-place execution (additive): the transforms of d_data are replaced by their spectra.
 * Every plan whose first kernel consumes a whole transform (or the whole batch) before anything is written back is run
 * directly on the buffer; the few that do not go through a copy of the input. */
void fft_exec_batch_inplace_dev(fft_object obj, fft_data *d_data, long howmany, void *stream)
{
    if (obj == NULL || d_data == NULL) {
        fprintf(stderr, "Error: Invalid FFT object or data pointers\n");
        exit(EXIT_FAILURE);
    }
    if (howmany <= 0) return;
    hs_dev_obj *d = hs_dev_get(obj);
    hs_dev_lock(d);
    const size_t bytes = sizeof(fft_data) * (size_t)d->N * (size_t)howmany, wsb = hs_dev_ws_bytes(d, howmany);
    const hs_c2c *p = d->lt == 0 ? d->direct : NULL;
    /* unsafe on the caller's buffer: the run-time engine's multi-pass plans (pass 1 writes into the output while other CTAs
     * still read the input) and the split chirp-z kernel (parks its first half-convolution in the output row, then reads
     * the input row again) */
    const int copy = (p && p->ct_id < 0 && p->jit_id < 0 && p->tp_id < 0 && p->jtp_id < 0 && p->npass > 1) || (d->lt == 1 && d->split_ct_id >= 0);
    char *ws = (char *)hs_scratch_get(&d->ws_pool, stream, wsb + (copy ? bytes : 0));
    if (copy) {
        hs_cu_d2d(ws + wsb, d_data, bytes, stream);
        hs_dev_exec(d, ws + wsb, d_data, ws, howmany, stream);
    } else hs_dev_exec(d, d_data, d_data, ws, howmany, stream);
    hs_dev_unlock(d);
}

void free_fft(fft_object object)
{
    if (object) hs_dev_release(object);
    free(object);
}

/* ----------------------------------------------------------------------------- device helpers */

int hsfft_device_count(void) { return hs_cu_device_count(); }
void hsfft_set_device(int dev) { hs_cu_set_device(dev); }
int hsfft_get_device(void) { return hs_cu_get_device(); }
void *hsfft_malloc_device(size_t bytes) { hs_cu_require_device(); return hs_cu_malloc(bytes); }
void hsfft_free_device(void *d) { hs_cu_free(d); }
void *hsfft_malloc_host(size_t bytes) { hs_cu_require_device(); return hs_cu_malloc_host(bytes); }
void hsfft_free_host(void *h) { hs_cu_free_host(h); }
void hsfft_copy_h2d(void *d, const void *h, size_t bytes) { hs_cu_h2d(d, h, bytes, NULL); hs_cu_stream_sync(NULL); }
void hsfft_copy_d2h(void *h, const void *d, size_t bytes) { hs_cu_d2h(h, d, bytes, NULL); hs_cu_stream_sync(NULL); }
void hsfft_sync(void) { hs_cu_stream_sync(NULL); }
long long hsfft_launch_count(void) { return hs_cu_launch_count(); }
void hsfft_timing_begin(void *stream) { hs_cu_timing_begin(stream); }
int hsfft_timing_report(char *out, int cap) { return hs_cu_timing_report(out, cap); }
const char *hsfft_version(void) { return "hsfft-b200 0.1 (sm_100a)"; }

const char *hsfft_plan_describe(fft_object obj)
{
    static __thread char buf[512]; /* one buffer per calling thread */
    if (!obj) return "null plan";
    hs_dev_obj *d = hs_dev_get(obj);
    int n = 0;
    if (d->lt == 0) {
        const hs_c2c *p = d->direct;
        if (p->ct_id >= 0) n = snprintf(buf, sizeof buf, "mixed-radix N=%d: compile-time smem kernel #%d", d->N, p->ct_id);
        else if (p->jit_id >= 0)
            n = snprintf(buf, sizeof buf, "mixed-radix N=%d: smem kernel compiled for this plan (NVRTC) #%d%s", d->N, p->jit_id,
                         hs_cu_jit_is_staged(p->jit_id) ? ", persistent with staged input for large batches" : "");
        else if (p->jtp_id >= 0)
            n = snprintf(buf, sizeof buf, "mixed-radix N=%d = %d x %d: two tiled passes compiled for this plan (NVRTC) #%d", d->N, p->n1, p->n2, p->jtp_id);
        else if (p->tp_id >= 0)
            n = snprintf(buf, sizeof buf, "mixed-radix N=%d: compile-time two-pass kernels #%d", d->N, p->tp_id);
        else {
            n = snprintf(buf, sizeof buf, "mixed-radix N=%d: %d run-time pass(es):", d->N, p->npass);
            for (int g = 0; g < p->npass && n < (int)sizeof buf - 32; ++g)
                n += snprintf(buf + n, sizeof buf - n, " [n=%d C=%d]", p->pass[g].n, p->pass[g].C);
        }
    } else if (d->big)
        n = snprintf(buf, sizeof buf, "bluestein N=%d M=%d=%dx%d: 3 fused tiled kernels", d->N, d->M, d->big_n1, d->big_n2);
    else if (d->rader_ct_id >= 0)
        n = snprintf(buf, sizeof buf, "rader N=%d (prime), cyclic convolution of length %d: fused %s kernel #%d", d->N, d->M,
                     d->rader_jit ? "plan-time (NVRTC)" : "compile-time", d->rader_ct_id);
    else if (d->split_ct_id >= 0)
        n = snprintf(buf, sizeof buf, "bluestein N=%d M=%d: one kernel, two convolutions of length %d (split chirp-z, %s transforms) #%d", d->N, d->M,
                     d->M / 2, d->split_conv ? "in-place" : "auto-sort", d->split_ct_id);
    else if (d->blu_ct_id >= 0)
        n = snprintf(buf, sizeof buf, "bluestein N=%d M=%d: fused compile-time kernel #%d", d->N, d->M, d->blu_ct_id);
    else
        n = snprintf(buf, sizeof buf, "bluestein N=%d M=%d: unfused (chirp, FFT_M x%d pass, mul, IFFT_M, chirp)", d->N, d->M,
                     d->fwdM->ct_id >= 0 ? 1 : (d->fwdM->tp_id >= 0 ? 2 : d->fwdM->npass));
    (void)n;
    return buf;
}
