/*
 * hsfft_b200.h -- additive API of hsfft-b200 (nothing here exists in the reference; nothing here
 * changes a reference symbol).  Everything is plain C ABI: pointers and sizes.
 *
 *   1. fft_convolve / next_power_of_two / find_optimal_fft_length: the reference DEFINES these in
 *      src/convolve.c:20,39,74 but declares them in no header; they are declared here.
 *   2. Batched entry points with HOST pointers (fft_exec_batch ...): what a caller looping
 *      fft_exec over a batch should call instead; copies are chunked and pipelined.
 *   3. Batched entry points with DEVICE pointers (*_dev): the path the roofline is measured on.
 *   4. Small device helpers so a C caller needs no CUDA headers.
 *
 * Layouts: complex data is interleaved {re, im} doubles (fft_data); a batch is `howmany`
 * transforms stored back to back (row stride = transform length unless a stride is given).
 */
#ifndef HSFFT_B200_H_
#define HSFFT_B200_H_

#include "real.h"

#ifdef __cplusplus
extern "C" {
#endif

/* ---- 1. convolution (reference src/convolve.c:20-25, :39-55, :74-214) ---- */
int next_power_of_two(int n);
int find_optimal_fft_length(int min_length, const char *conv_type, int length1, int length2);
/* type: "full" | "same" | "valid"; conv_type: "linear" | "circular".  Returns the number of
 * samples written to output, or -1 on invalid arguments. */
int fft_convolve(const char *type, const char *conv_type, fft_type *input1, int length1,
                 fft_type *input2, int length2, fft_type *output);

/* ---- 2. batched, host pointers ---- */
void fft_exec_batch(fft_object obj, fft_data *inp, fft_data *oup, long howmany);
/* r2c: rows of N reals in, rows of N bins out (full mirrored spectrum, like fft_r2c_exec) */
void fft_r2c_exec_batch(fft_real_object obj, fft_type *inp, fft_data *oup, long howmany);
/* c2r: rows of N bins in (bins 0..N/2 are read), rows of N reals out */
void fft_c2r_exec_batch(fft_real_object obj, fft_data *inp, fft_type *oup, long howmany);
/* howmany independent convolutions: rows of length1 / length2 samples, output rows of out_stride */
int fft_convolve_batch(const char *type, const char *conv_type, fft_type *input1, int length1,
                       fft_type *input2, int length2, fft_type *output, long out_stride, long howmany);

/* ---- 3. batched, device pointers (stream: a cudaStream_t passed as void*, NULL = default).  Out of place, like every
 *          transform of the reference: input and output must not overlap.  A plan may be executed from several host
 *          threads and on several streams at once: calls on one (plan, device) serialise on an internal lock while they
 *          enqueue, and every caller stream has its own intermediates. ---- */
void fft_exec_batch_dev(fft_object obj, const fft_data *d_in, fft_data *d_out, long howmany, void *stream);
/* in place: every transform of d_data is replaced by its spectrum (the one call that may be given a single buffer) */
void fft_exec_batch_inplace_dev(fft_object obj, fft_data *d_data, long howmany, void *stream);
/* out_stride: complex elements between output rows (>= N/2+1); full != 0 also writes the mirror
 * bins N/2+1..N-1 (needs out_stride >= N), full == 0 writes bins 0..N/2 only. */
void fft_r2c_exec_batch_dev(fft_real_object obj, const fft_type *d_in, fft_data *d_out, long out_stride, int full,
                            long howmany, void *stream);
/* in_stride: complex elements between input rows (>= N/2+1); bins 0..N/2 are read */
void fft_c2r_exec_batch_dev(fft_real_object obj, const fft_data *d_in, long in_stride, fft_type *d_out, long howmany,
                            void *stream);
int fft_convolve_batch_dev(const char *type, const char *conv_type, const fft_type *d_in1, int length1,
                           const fft_type *d_in2, int length2, fft_type *d_out, long out_stride, long howmany,
                           void *stream);

/* ---- 3a. overlap-save streaming convolution: y = x (*) h for a long signal and a short filter ("full" linear
 *          convolution, Lx + Lh - 1 samples).  Windows of P samples (a power of two 2^15..2^21, at least 2*Lh; P <= 0
 *          picks one) are transformed in batches; they are read in place, never copied.  Returns the number of
 *          samples written or -1. ---- */
long fft_convolve_stream(fft_type *x, long Lx, fft_type *h, int Lh, fft_type *y, int P);
long fft_convolve_stream_dev(const fft_type *d_x, long Lx, const fft_type *d_h, int Lh, fft_type *d_y, int P, void *stream);

/* ---- 3c. two-dimensional complex transforms (row-major n0 x n1, unnormalised, same sign convention as fft_exec):
 *          rows as batched 1-D transforms, columns by a strided pass (n0 = 128, 256, 512, 1024 and n1 a multiple of 16)
 *          or through a transpose.  Device-pointer calls are out of place; in and out must not overlap. ---- */
typedef struct fft2_set *fft2_object;
fft2_object fft2_init(int n0, int n1, int sgn);
void fft2_exec(fft2_object p, fft_data *inp, fft_data *oup);
void fft2_exec_batch(fft2_object p, fft_data *inp, fft_data *oup, long howmany);
void fft2_exec_batch_dev(fft2_object p, const fft_data *d_in, fft_data *d_out, long howmany, void *stream);
const char *fft2_describe(fft2_object p);
void free_fft2(fft2_object p);

/* ---- 3b. distributed four-step transform: local phases (the exchange between them is the caller's
 *          all-to-all, e.g. torch.distributed / NCCL; see csrc/hsfft_dist.c for the data layout) ----
 * N = N1*N2 (powers of two), `nranks` devices; rank g holds a[N1][N2/nranks] = x[N2 n1 + g N2/nranks + c]. */
typedef struct hsfft_dist_set *hsfft_dist_object;
hsfft_dist_object hsfft_dist_init(int N1, int N2, int nranks, int rank, int sgn);
/* phase A: b[k1][c] = W_N^(n2 k1) * FFT_N1 over n1 of a[n1][c]; ws holds N1 * N2/nranks elements */
void hsfft_dist_cols_dev(hsfft_dist_object d, const fft_data *d_a, fft_data *d_b, fft_data *d_ws, void *stream);
/* phase B: out[k1l][k2] = FFT_N2 over n2 of r[n2 / Cw][k1l][n2 % Cw] (r as received from the all-to-all) */
void hsfft_dist_rows_dev(hsfft_dist_object d, const fft_data *d_r, fft_data *d_out, fft_data *d_ws, void *stream);
long long hsfft_dist_slab_elems(hsfft_dist_object d);
/* Exchange fused into phase A (no all-to-all): every rank allocates its receive buffer with hsfft_malloc_device,
 * publishes hsfft_ipc_get_handle()'s 64 bytes to the others (any transport), and passes all ranks' handles, ordered by
 * rank, to hsfft_dist_set_peers.  hsfft_dist_cols_exchange_dev then stores each output row, already multiplied by
 * the exchange twiddle, straight into the receive buffer of the rank that owns it (NVLink peer stores issued by the
 * transform kernel itself).  The caller barriers the ranks before hsfft_dist_rows_dev reads the buffer, and alternates
 * two receive buffers (set_peers once per buffer, on two plans or between calls) if transforms follow back to back.
 * hsfft_dist_set_peer_pointers: same for buffers already addressable from this process. */
void hsfft_ipc_get_handle(void *d_buf, void *handle64);
void hsfft_dist_set_peers(hsfft_dist_object d, void *d_own_r, const void *handles64, int nhandles);
void hsfft_dist_set_peer_pointers(hsfft_dist_object d, void *const *peer_r, int npeers);
void hsfft_dist_cols_exchange_dev(hsfft_dist_object d, const fft_data *d_a, fft_data *d_ws, void *stream);
void hsfft_dist_free(hsfft_dist_object d);
/* the same four-step from ONE process driving devices 0 .. ndev-1 (peer access instead of IPC handles): host pointers,
 * natural order in and out, N = N1*N2, unnormalised.  Returns 0, -1 when the devices cannot reach each other's memory. */
int fft_exec_dist(int N1, int N2, int sgn, const fft_data *inp, fft_data *oup, int ndev);

/* ---- 4. device helpers ---- */
int   hsfft_device_count(void);          /* 0 when no usable CUDA device */
void  hsfft_set_device(int dev);         /* device state of a plan is kept per (plan, device): a plan may execute on any device */
/* The host-pointer batch calls (fft_exec_batch, fft_r2c/c2r_exec_batch, fft_convolve_batch) shard their batch over devices
 * 0 .. n-1 by batch index, one host thread per device, no communication (default n = 1: the current device only). */
void  hsfft_set_devices(int n);
int   hsfft_get_devices(void);
int   hsfft_get_device(void);
void *hsfft_malloc_device(size_t bytes);
void  hsfft_free_device(void *d);
void *hsfft_malloc_host(size_t bytes);   /* pinned host memory for the host-pointer batch calls */
void  hsfft_free_host(void *h);
void  hsfft_copy_h2d(void *d, const void *h, size_t bytes);
void  hsfft_copy_d2h(void *h, const void *d, size_t bytes);
void  hsfft_sync(void);
long long hsfft_launch_count(void);      /* kernels this library has launched in this process */
/* per-launch device timing (CUDA events on the launching stream, for roofline records): hsfft_timing_begin(stream), run
 * any calls, then hsfft_timing_report writes "kernel tag:launches:total ms;..." and returns its length (-1: buffer too small) */
void  hsfft_timing_begin(void *stream);
int   hsfft_timing_report(char *out, int cap);
/* human-readable description of how a plan will execute (kernel family, passes); static buffer */
const char *hsfft_plan_describe(fft_object obj);
const char *hsfft_version(void);

#ifdef __cplusplus
}
#endif
#endif /* HSFFT_B200_H_ */
