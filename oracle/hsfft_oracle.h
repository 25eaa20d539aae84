/*
 * hsfft_oracle.h -- CPU restatement of the reference FFT hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is linked into, imported by
 * or executed from the product library (libhsfft_b200.so).  It is used by
 * tests/, by __graft_entry__.smoke() and by bench.py's cpu_baseline /
 * --impl reference legs, and only as the checker or as the timed CPU baseline.
 *
 * Parity status: PINNED.  tests/test_oracle.py checks this restatement against
 * (a) the reference's own C sources compiled in place into oracle/_ref/
 * (libhsfft_ref_asis.so / libhsfft_ref_repaired.so, see oracle/Makefile),
 * (b) the known-answer vectors the reference's gtest file holds for this path
 * (SURVEY.md section 4), and (c) an independent long-double DFT.
 *
 * Each function cites the reference file:line it restates
 * (paths relative to /root/reference).
 */
#ifndef HSFFT_ORACLE_H_
#define HSFFT_ORACLE_H_

#ifdef __cplusplus
extern "C" {
#endif

typedef struct { double re, im; } orc_cplx;

/* plan: restates struct fft_set (src/highspeedFFT.h:36-43) without the
 * flexible-array trick; `len` is the transform length the twiddle table was
 * built for (N for mixed radix, the padded M for Bluestein). */
typedef struct orc_plan {
    int N;
    int sgn;
    int factors[64];
    int lf;
    int lt;        /* 0 = mixed radix, 1 = Bluestein */
    int len;       /* N or M */
    orc_cplx *tw;  /* len-1 (+1 spare) stage twiddles, conjugated when sgn == -1 */
} orc_plan;

typedef struct orc_real_plan {
    orc_plan *cplan;   /* half-length complex plan */
    int N;             /* real length (even) */
    orc_cplx *tw2;     /* exp(+2 pi i k / N), k < N/2 */
} orc_real_plan;

/* plan helpers (src/highSpeedFFT.c:1979-2025, :2038-2163, :2238-2313 dynamic branch) */
int  orc_dividebyN(int n);
int  orc_factors(int n, int *out);
void orc_stage_twiddles(orc_cplx *tw, int n, const int *fac, int lf);

/* complex transform (src/highSpeedFFT.c:206-286, :318-1629, :1645-1907, :1920-1942) */
orc_plan *orc_fft_init(int N, int sgn);
void      orc_fft_exec(orc_plan *p, const orc_cplx *in, orc_cplx *out);
void      orc_fft_free(orc_plan *p);
int       orc_bluestein_pad(int N);   /* M used at exec time, :1750-1751 */

/* real transforms (src/real.c:26-64, :78-136, :150-193) */
orc_real_plan *orc_real_init(int N, int sgn);
void orc_r2c_exec(orc_real_plan *p, const double *in, orc_cplx *out /* N entries */);
void orc_c2r_exec(orc_real_plan *p, const orc_cplx *in /* N/2+1 read */, double *out);
void orc_real_free(orc_real_plan *p);

/* convolution (src/convolve.c:20-25, :39-55, :74-214) */
int orc_next_pow2(int n);
int orc_conv_fft_length(int min_len, const char *conv_type, int l1, int l2);
int orc_convolve(const char *type, const char *conv_type, const double *a, int la,
                 const double *b, int lb, double *out);

/* batch drivers used by tests and by the CPU-baseline timing leg.
 * nthreads <= 0 -> all OpenMP threads.  One plan per thread (the reference's
 * Bluestein exec mutates its plan, src/highSpeedFFT.c:1759-1760). */
void orc_c2c_batch(int N, int sgn, const orc_cplx *in, orc_cplx *out, long howmany, int nthreads);
void orc_r2c_batch(int N, const double *in, orc_cplx *out, long howmany, int nthreads);
void orc_c2r_batch(int N, const orc_cplx *in, double *out, long howmany, int nthreads);
int  orc_num_threads(void);

/* independent cross-check: O(N^2) long-double DFT with (k*n mod N) phase reduction */
void orc_naive_dft(int N, int sgn, const orc_cplx *in, orc_cplx *out);

/* relative RMS error sqrt(sum|y-r|^2 / sum|r|^2) */
double orc_rel_rms(const orc_cplx *y, const orc_cplx *r, long n);

#ifdef __cplusplus
}
#endif
#endif
