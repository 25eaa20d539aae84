/*
 * highspeedFFT.h -- complex plan-create / execute / destroy API of hsfft-b200.
 *
 * Drop-in replacement for the reference's header of the same name
 * (reference src/highspeedFFT.h:1-68): the same symbols, the same argument
 * meaning, the same PUBLIC plan layout, so code compiled against the reference
 * links against libhsfft_b200.so unchanged.  What changed is what runs behind
 * fft_exec: hand-written sm_100a kernels on a B200 instead of the recursive
 * CPU DIT.  There is no CPU fallback; without a CUDA device every exec call
 * prints an error and exits, like every other failure in this API.
 *
 * Entry points and the reference interface each one replaces:
 *   fft_init     reference src/highspeedFFT.h:34  (src/highSpeedFFT.c:206-286)
 *   fft_exec     reference src/highspeedFFT.h:45  (src/highSpeedFFT.c:1920-1942)
 *   free_fft     reference src/highspeedFFT.h:59  (src/highSpeedFFT.c:2315-2318)
 *   divideby     reference src/highspeedFFT.h:47  (src/highSpeedFFT.c:1954-1968)
 *   dividebyN    reference src/highspeedFFT.h:49  (src/highSpeedFFT.c:1979-2025)
 *   factors      reference src/highspeedFFT.h:53  (src/highSpeedFFT.c:2038-2163)
 *   twiddle      reference src/highspeedFFT.h:55  (src/highSpeedFFT.c:2186-2224)
 *   longvectorN  reference src/highspeedFFT.h:57  (src/highSpeedFFT.c:2238-2313)
 *
 * Conventions kept from the reference: transforms are out of place and
 * unnormalised in both directions; sgn = +1 means exp(-2 pi i k n / N);
 * argument errors print to stderr and exit(EXIT_FAILURE).
 * Batched, device-resident and multi-GPU entry points are additive and live in
 * hsfft_b200.h.
 */
#ifndef HSFFT_H_
#define HSFFT_H_

#include <stdlib.h>
#include <stdio.h>
#include <math.h>
#include <string.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PI2 6.28318530717958647692528676655900577

#ifndef fft_type
#define fft_type double
#endif

/* interleaved complex sample, 16 bytes */
typedef struct fft_t {
  fft_type re;
  fft_type im;
} fft_data;

typedef struct fft_set* fft_object;

/*
 * Public plan.  Field offsets are part of the ABI (N@0 sgn@4 factors@8 lf@264
 * lt@268 twiddle@272, sizeof 288); callers and the reference's tests read them.
 * The block is one malloc() and stays free()-able; GPU-side state (HBM-resident
 * twiddle / chirp tables, workspace, stream) lives in an internal registry keyed
 * by the plan address and is released by free_fft.
 *   lt == 0: mixed radix, twiddle[] holds the N-1 stage twiddles
 *   lt == 1: Bluestein,   twiddle[] holds the M-1 stage twiddles of the padded length
 */
struct fft_set{
	int N;
	int sgn;
	int factors[64];
	int lf;
	int lt;
	fft_data twiddle[1];
};

fft_object fft_init(int N, int sgn);

void fft_exec(fft_object obj,fft_data *inp,fft_data *oup);

int divideby(int M,int d);

int dividebyN(int N);

int factors(int M, int* arr);

void twiddle(fft_data *sig,int N, int radix);

void longvectorN(fft_data *sig,int N, int *array, int M);

void free_fft(fft_object object);

#ifdef __cplusplus
}
#endif

#endif /* HSFFT_H_ */
