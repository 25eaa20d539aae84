/*
 * real.h -- real-input plan-create / execute / destroy API of hsfft-b200.
 *
 * Drop-in replacement for the reference's src/real.h:1-90.
 *   fft_real_init   reference src/real.h:46  (src/real.c:26-64)
 *   fft_r2c_exec    reference src/real.h:60  (src/real.c:78-136)
 *   fft_c2r_exec    reference src/real.h:74  (src/real.c:150-193)
 *   free_real_fft   reference src/real.h:84  (src/real.c:259-267)
 *
 * Buffer contract, as the reference's CODE has it (its header comment says
 * N/2+1, its loop at src/real.c:128-132 writes the mirror):
 *   fft_r2c_exec writes ALL N bins (out[N-k] = conj(out[k])): out must hold N fft_data.
 *   fft_c2r_exec reads bins 0..N/2 and returns N unnormalised reals (N*x for a round trip).
 * Where the half-length complex plan is a one-kernel plan (N/2 <= 8192, static or compiled at plan time) or N = 2^15..2^20,
 * the split (r2c) and the merge (c2r) are fused into the transform kernels themselves: r2c splits in shared memory after
 * the last pass (one kernel) or on pass 1 of a real four-step (two kernels), c2r merges on the first pass's loads.  Only
 * half-lengths that run as Bluestein / Rader / the run-time engine, and r2c on an inverse plan, use the separate
 * element-wise split / merge kernels (csrc/hs_generic.cuh).
 */
#ifndef REAL_H_
#define REAL_H_

#include "highspeedFFT.h"

#ifdef __cplusplus
extern "C"
{
#endif

    typedef struct fft_real_set *fft_real_object;

    /* Public real plan: half-length complex plan + exp(+2 pi i k / N), k < N/2. */
    struct fft_real_set
    {
        fft_object cobj;
        fft_data twiddle2[1];
    };

    fft_real_object fft_real_init(int signal_length, int transform_direction);

    void fft_r2c_exec(fft_real_object real_obj, fft_type *input_data, fft_data *output_data);

    void fft_c2r_exec(fft_real_object real_obj, fft_data *input_data, fft_type *output_data);

    void free_real_fft(fft_real_object real_obj);

    /* Helper exports the reference's real.c also defines (src/real.c:198-257); host-only. */
    void generate_signal(fft_data *signal, int length, double freq, double amplitude);
    double compute_mse(fft_data *original, fft_data *reconstructed, int length);
    void print_complex(fft_data *data, int length, const char *label);
    void generate_real_signal(fft_type *signal, int length, double freq, double amplitude);
    double compute_mse_real(fft_type *original, fft_type *reconstructed, int length);
    void print_real(fft_type *data, int length, const char *label);

#ifdef __cplusplus
}
#endif

#endif /* REAL_H_ */
