/*
 * ref_driver.c -- batch/timing driver linked WITH the reference's own C sources
 * (compiled in place from /root/reference/src by oracle/Makefile) into
 * oracle/_ref/libhsfft_ref_{asis,repaired}.so.
 *
 * TEST INFRASTRUCTURE ONLY.  It calls the reference's public API exactly as a
 * user would (src/highspeedFFT.h:34-59, src/real.h:46-84, src/convolve.c:74) and
 * adds nothing but an OpenMP loop over the batch with one plan per thread
 * (the reference's Bluestein exec mutates its plan, src/highSpeedFFT.c:1759-1760).
 */
#include "real.h"
#ifdef _OPENMP
#include <omp.h>
#endif

int fft_convolve(const char *type, const char *conv_type, fft_type *input1, int length1,
                 fft_type *input2, int length2, fft_type *output);

int refdrv_num_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

void refdrv_c2c_batch(int N, int sgn, fft_data *in, fft_data *out, long howmany, int nthreads)
{
#ifdef _OPENMP
    if (nthreads <= 0) nthreads = omp_get_max_threads();
#pragma omp parallel num_threads(nthreads)
#endif
    {
        fft_object p = fft_init(N, sgn);
#ifdef _OPENMP
#pragma omp for schedule(static)
#endif
        for (long b = 0; b < howmany; ++b) fft_exec(p, in + b * N, out + b * N);
        free_fft(p);
    }
}

void refdrv_r2c_batch(int N, fft_type *in, fft_data *out, long howmany, int nthreads)
{
#ifdef _OPENMP
    if (nthreads <= 0) nthreads = omp_get_max_threads();
#pragma omp parallel num_threads(nthreads)
#endif
    {
        fft_real_object p = fft_real_init(N, 1);
#ifdef _OPENMP
#pragma omp for schedule(static)
#endif
        for (long b = 0; b < howmany; ++b) fft_r2c_exec(p, in + b * N, out + b * N);
        free_real_fft(p);
    }
}

/* input rows hold N/2+1 bins (row stride N/2+1), as fft_c2r_exec reads in[0..N/2] */
void refdrv_c2r_batch(int N, fft_data *in, fft_type *out, long howmany, int nthreads)
{
#ifdef _OPENMP
    if (nthreads <= 0) nthreads = omp_get_max_threads();
#pragma omp parallel num_threads(nthreads)
#endif
    {
        fft_real_object p = fft_real_init(N, -1);
#ifdef _OPENMP
#pragma omp for schedule(static)
#endif
        for (long b = 0; b < howmany; ++b) fft_c2r_exec(p, in + b * (N / 2 + 1), out + b * N);
        free_real_fft(p);
    }
}

/* howmany independent linear "full" convolutions of la- and lb-sample rows */
void refdrv_convolve_batch(const char *type, const char *conv_type, fft_type *a, int la, fft_type *b, int lb,
                           fft_type *out, int out_stride, long howmany, int nthreads)
{
#ifdef _OPENMP
    if (nthreads <= 0) nthreads = omp_get_max_threads();
#pragma omp parallel for schedule(static) num_threads(nthreads)
#endif
    for (long i = 0; i < howmany; ++i)
        fft_convolve(type, conv_type, a + i * la, la, b + i * lb, lb, out + i * (long)out_stride);
}
