/*
 * hsfft_convolve.c -- FFT convolution of real signals (host side, plain C).
 *
 * Replaces reference src/convolve.c: next_power_of_two :20-25, find_optimal_fft_length :39-55,
 * fft_convolve :74-214.  Same argument meaning, same return values (-1 on bad arguments, exit on
 * an unknown conv_type inside find_optimal_fft_length), same "full" / "same" / "valid" slicing.
 *
 * The reference builds and destroys two real plans and six P-sized buffers on every call
 * (:104-122, :204-211) and multiplies all P mirrored bins (:147-151).  Here the pair of plans for a
 * given P is cached, the spectra are produced half-length only (bins 0..P/2 -- all c2r reads),
 * the product carries the 1/P scale (exact: P is a power of two), and a batch entry point runs
 * many independent convolutions per launch.
 */
#include "hs_engine.h"
#include "../../include/hsfft_b200.h"

#include <pthread.h>

#ifndef MIN
#define MIN(a, b) ((a) < (b) ? (a) : (b))
#endif
#ifndef MAX
#define MAX(a, b) ((a) > (b) ? (a) : (b))
#endif

int next_power_of_two(int n)
{
    if (n <= 0) return 1;
    return (int)pow(2, ceil(log2(n)));
}

int find_optimal_fft_length(int min_length, const char *conv_type, int length1, int length2)
{
    if (strcmp(conv_type, "linear") == 0) return next_power_of_two(min_length);
    if (strcmp(conv_type, "circular") == 0) return next_power_of_two(MAX(length1, length2));
    fprintf(stderr, "Error: Invalid convolution type\n");
    exit(EXIT_FAILURE);
}

/* plan cache: one forward and one inverse real plan per (padded length, device).  Scratch of the device-pointer calls is
 * kept per caller stream (hs_scratch); the host-pointer batch call stages through three slots on the cache entry's own
 * streams and holds the entry's lock for the whole call. */
typedef struct conv_plans {
    int P, device;
    pthread_mutex_t lock;
    fft_real_object fwd, inv;
    hs_scratch a, b, A, B, res;
    void *stream[HS_NSLOT], *d_1[HS_NSLOT], *d_2[HS_NSLOT], *d_o[HS_NSLOT];
    size_t cap_1[HS_NSLOT], cap_2[HS_NSLOT], cap_o[HS_NSLOT];
    struct conv_plans *next;
} conv_plans;

static pthread_mutex_t c_lock = PTHREAD_MUTEX_INITIALIZER;
static conv_plans *c_head = NULL;

static conv_plans *plans_for(int P)
{
    hs_cu_require_device();
    const int dev = hs_cu_get_device();
    pthread_mutex_lock(&c_lock);
    conv_plans *c = c_head;
    while (c && !(c->P == P && c->device == dev)) c = c->next;
    if (!c) {
        c = (conv_plans *)calloc(1, sizeof(*c));
        if (!c) { fprintf(stderr, "Error: Memory allocation failed\n"); exit(EXIT_FAILURE); }
        pthread_mutexattr_t at;
        pthread_mutexattr_init(&at);
        pthread_mutexattr_settype(&at, PTHREAD_MUTEX_RECURSIVE);
        pthread_mutex_init(&c->lock, &at);
        pthread_mutexattr_destroy(&at);
        c->P = P;
        c->device = dev;
        c->fwd = fft_real_init(P, 1);
        c->inv = fft_real_init(P, -1);
        for (int s = 0; s < HS_NSLOT; ++s) c->stream[s] = hs_cu_stream_create();
        c->next = c_head;
        c_head = c;
    }
    pthread_mutex_unlock(&c_lock);
    return c;
}

static void grow(void **buf, size_t *cap, size_t need)
{
    if (need <= *cap) return;
    hs_cu_free(*buf);
    *buf = hs_cu_malloc(need);
    *cap = need;
}

/* slicing rules of reference :163-195; returns the output length (or -1), sets *start */
static int conv_slice(const char *type, int linear, int length1, int length2, int conv_length, int P, int *start)
{
    *start = 0;
    if (!linear) return P;
    if (type == NULL || strcmp(type, "full") == 0) return conv_length;
    if (strcmp(type, "same") == 0) {
        const int larger = MAX(length1, length2);
        *start = (conv_length - larger) / 2;
        return larger;
    }
    if (strcmp(type, "valid") == 0) {
        const int smaller = MIN(length1, length2);
        *start = smaller - 1;
        return MAX(length1, length2) - smaller + 1;
    }
    fprintf(stderr, "Error: Invalid output type. Use 'full', 'same', or 'valid'.\n");
    return -1;
}

static int conv_setup(const char *type, const char *conv_type, const void *in1, int length1, const void *in2, int length2, const void *out,
                      int *P, int *start, int *linear)
{
    if (type == NULL) type = "full"; /* the reference's `|| type == NULL` intent (:165) */
    if (in1 == NULL || in2 == NULL || out == NULL || length1 <= 0 || length2 <= 0 || conv_type == NULL) {
        fprintf(stderr, "Error: Invalid inputs for fft_convolve\n");
        return -1;
    }
    int conv_length;
    if (strcmp(conv_type, "linear") == 0) { conv_length = length1 + length2 - 1; *linear = 1; }
    else if (strcmp(conv_type, "circular") == 0) { conv_length = MAX(length1, length2); *linear = 0; }
    else {
        fprintf(stderr, "Error: Invalid convolution type. Use 'linear' or 'circular'.\n");
        return -1;
    }
    *P = find_optimal_fft_length(conv_length, conv_type, length1, length2);
    if (*P < 2) { /* P == 1 is odd: the reference exits inside fft_real_init (src/real.c:28-32) */
        fprintf(stderr, "Error: Signal length (%d) must be positive and even\n", *P);
        exit(EXIT_FAILURE);
    }
    return conv_slice(type, *linear, length1, length2, conv_length, *P, start);
}

int fft_convolve_batch_dev(const char *type, const char *conv_type, const fft_type *d_in1, int length1, const fft_type *d_in2, int length2,
                           fft_type *d_out, long out_stride, long howmany, void *stream)
{
    int P, start, linear;
    const int n = conv_setup(type, conv_type, d_in1, length1, d_in2, length2, d_out, &P, &start, &linear);
    if (n <= 0 || howmany <= 0) return n;
    conv_plans *c = plans_for(P);
    const size_t H1 = (size_t)P / 2 + 1;
    pthread_mutex_lock(&c->lock);
    fft_data *dA = (fft_data *)hs_scratch_get(&c->A, stream, sizeof(fft_data) * H1 * howmany);
    fft_data *dB = (fft_data *)hs_scratch_get(&c->B, stream, sizeof(fft_data) * H1 * howmany);
    /* long convolutions: padding, product, scale and slice ride on the transforms' own loads and stores (6 launches,
     * about two thirds of the HBM traffic of the unfused sequence below) */
    if (hs_real_r2c_padded_dev(c->fwd, d_in1, length1, length1, 0, (long long)length1 * howmany, dA, (long)H1, howmany, stream)) {
        if (!hs_real_r2c_padded_dev(c->fwd, d_in2, length2, length2, 0, (long long)length2 * howmany, dB, (long)H1, howmany, stream) ||
            !hs_real_c2r_mul_slice_dev(c->inv, dA, dB, (long)H1, (long)H1, 1.0 / P, d_out, out_stride,
                                       (long long)out_stride * (howmany - 1) + n, start, n, howmany, stream)) {
            fprintf(stderr, "Error: fused convolution path became unavailable\n");
            exit(EXIT_FAILURE);
        }
        pthread_mutex_unlock(&c->lock);
        return n;
    }
    fft_type *da = (fft_type *)hs_scratch_get(&c->a, stream, sizeof(fft_type) * (size_t)P * howmany);
    fft_type *db = (fft_type *)hs_scratch_get(&c->b, stream, sizeof(fft_type) * (size_t)P * howmany);
    fft_type *dres = (fft_type *)hs_scratch_get(&c->res, stream, sizeof(fft_type) * (size_t)P * howmany);
    /* zero-pad -> r2c x2 (half spectra) -> product * 1/P -> c2r -> slice */
    hs_cu_launch_pad_rows(d_in1, length1, da, P, howmany, stream);
    hs_cu_launch_pad_rows(d_in2, length2, db, P, howmany, stream);
    fft_r2c_exec_batch_dev(c->fwd, da, dA, (long)H1, 0, howmany, stream);
    fft_r2c_exec_batch_dev(c->fwd, db, dB, (long)H1, 0, howmany, stream);
    hs_cu_launch_conv_mul(dA, dB, dA, (long long)(H1 * howmany), 1.0 / P, stream);
    fft_c2r_exec_batch_dev(c->inv, dA, (long)H1, dres, howmany, stream);
    hs_cu_launch_slice_rows(dres, P, start, n, d_out, out_stride, howmany, stream);
    pthread_mutex_unlock(&c->lock);
    return n;
}

typedef struct conv_ctx {
    const char *type, *conv_type;
    fft_type *in1, *in2, *out;
    int length1, length2, P, n;
    long out_stride;
} conv_ctx;

/* one device's share of a host-pointer batch: three slots on the cache entry's own streams, so that the copies of chunk
 * i+1 and i-1 overlap the kernels of chunk i; staging buffers persist across calls (no allocation per call) */
static void convolve_batch_one_device(void *vctx, long first0, long howmany)
{
    conv_ctx *x = (conv_ctx *)vctx;
    if (howmany <= 0) return;
    conv_plans *c = plans_for(x->P);
    pthread_mutex_lock(&c->lock);
    const int l1 = x->length1, l2 = x->length2, n = x->n;
    /* chunk so that the device scratch (about 5 P-sized rows per convolution) stays bounded */
    long chunk = (long)(((size_t)128 << 20) / (sizeof(fft_type) * (size_t)x->P * 5));
    if (chunk < 1) chunk = 1;
    if (chunk > howmany) chunk = howmany;
    const long nchunk = (howmany + chunk - 1) / chunk;
    fft_type *in1 = x->in1 + (size_t)first0 * l1, *in2 = x->in2 + (size_t)first0 * l2, *out = x->out + (size_t)first0 * x->out_stride;
    for (long i = 0; i < nchunk; ++i) {
        const int s = (int)(i % HS_NSLOT);
        const long first = i * chunk, cnt = (first + chunk <= howmany) ? chunk : howmany - first;
        if (i < HS_NSLOT) {
            grow(&c->d_1[s], &c->cap_1[s], sizeof(fft_type) * (size_t)l1 * chunk);
            grow(&c->d_2[s], &c->cap_2[s], sizeof(fft_type) * (size_t)l2 * chunk);
            grow(&c->d_o[s], &c->cap_o[s], sizeof(fft_type) * (size_t)n * chunk);
        }
        hs_cu_h2d(c->d_1[s], in1 + first * l1, sizeof(fft_type) * (size_t)l1 * cnt, c->stream[s]);
        hs_cu_h2d(c->d_2[s], in2 + first * l2, sizeof(fft_type) * (size_t)l2 * cnt, c->stream[s]);
        fft_convolve_batch_dev(x->type, x->conv_type, (const fft_type *)c->d_1[s], l1, (const fft_type *)c->d_2[s], l2, (fft_type *)c->d_o[s], n,
                               cnt, c->stream[s]);
        hs_cu_d2h_2d(out + first * x->out_stride, sizeof(fft_type) * (size_t)x->out_stride, c->d_o[s], sizeof(fft_type) * (size_t)n,
                     sizeof(fft_type) * (size_t)n, (size_t)cnt, c->stream[s]);
    }
    for (int s = 0; s < HS_NSLOT; ++s) hs_cu_stream_sync(c->stream[s]);
    pthread_mutex_unlock(&c->lock);
}

int fft_convolve_batch(const char *type, const char *conv_type, fft_type *input1, int length1, fft_type *input2, int length2,
                       fft_type *output, long out_stride, long howmany)
{
    int P, start, linear;
    const int n = conv_setup(type, conv_type, input1, length1, input2, length2, output, &P, &start, &linear);
    if (n <= 0 || howmany <= 0) return n;
    hs_cu_require_device();
    conv_ctx x = {type, conv_type, input1, input2, output, length1, length2, P, n, out_stride};
    hs_shard_run(howmany, convolve_batch_one_device, &x);
    return n;
}

int fft_convolve(const char *type, const char *conv_type, fft_type *input1, int length1, fft_type *input2, int length2, fft_type *output)
{
    int P, start, linear;
    const int n = conv_setup(type, conv_type, input1, length1, input2, length2, output, &P, &start, &linear);
    if (n <= 0) return n;
    return fft_convolve_batch(type, conv_type, input1, length1, input2, length2, output, n, 1);
}


/* ------------------------------------------------------------ overlap-save streaming convolution (additive)
 * y = x (*) h, a long signal and a short filter, "full" linear convolution of Lx + Lh - 1 samples.  The signal is cut into
 * windows of P samples that overlap by Lh - 1 (hop S = P - Lh + 1); every window is circularly convolved with the filter
 * through the batched real transforms and only its last S samples -- the ones free of wrap-around -- are kept.  The
 * windows are never copied: window b is read in place by the forward transform's first pass (zero outside the signal),
 * the filter's spectrum is one shared row of the product, and each block's S samples go straight to y. */
long fft_convolve_stream_dev(const fft_type *d_x, long Lx, const fft_type *d_h, int Lh, fft_type *d_y, int P, void *stream)
{
    if (!d_x || !d_h || !d_y || Lx <= 0 || Lh <= 0) {
        fprintf(stderr, "Error: Invalid inputs for fft_convolve_stream\n");
        return -1;
    }
    if (P <= 0) { /* 8 filter lengths per window keeps the overlap at 1/8; at least 2^15 (the fused two-pass path), at most 2^20 */
        P = 1 << 15;
        while (P < 8 * Lh && P < (1 << 20)) P *= 2;
    }
    if (P < (1 << 15) || P > (1 << 21) || (P & (P - 1)) || Lh > P / 2) {
        fprintf(stderr, "Error: fft_convolve_stream needs a power-of-two window 2^15..2^21 of at least twice the filter (P = %d, filter %d)\n", P, Lh);
        return -1;
    }
    const long total = Lx + Lh - 1, S = P - Lh + 1, nblocks = (total + S - 1) / S;
    conv_plans *c = plans_for(P);
    const size_t H1 = (size_t)P / 2 + 1;
    long chunk = (long)(((size_t)512 << 20) / (sizeof(fft_data) * H1)); /* spectra of a chunk of windows: at most 512 MiB */
    if (chunk < 1) chunk = 1;
    if (chunk > nblocks) chunk = nblocks;
    pthread_mutex_lock(&c->lock);
    fft_data *dA = (fft_data *)hs_scratch_get(&c->A, stream, sizeof(fft_data) * H1 * (size_t)chunk);
    fft_data *dB = (fft_data *)hs_scratch_get(&c->B, stream, sizeof(fft_data) * H1);
    if (!hs_real_r2c_padded_dev(c->fwd, d_h, Lh, Lh, 0, Lh, dB, (long)H1, 1, stream)) {
        fprintf(stderr, "Error: fft_convolve_stream: no fused transform for P = %d\n", P);
        pthread_mutex_unlock(&c->lock);
        return -1;
    }
    for (long b0 = 0; b0 < nblocks; b0 += chunk) {
        const long cnt = (b0 + chunk <= nblocks) ? chunk : nblocks - b0;
        /* window b0 + i starts at signal index (b0 + i) * S - (Lh - 1) */
        hs_real_r2c_padded_dev(c->fwd, d_x, P, S, b0 * S - (Lh - 1), Lx, dA, (long)H1, cnt, stream);
        hs_real_c2r_mul_slice_dev(c->inv, dA, dB, (long)H1, 0, 1.0 / P, d_y + b0 * S, S, total - b0 * S, Lh - 1, (int)S, cnt, stream);
    }
    pthread_mutex_unlock(&c->lock);
    return total;
}

long fft_convolve_stream(fft_type *x, long Lx, fft_type *h, int Lh, fft_type *y, int P)
{
    if (!x || !h || !y || Lx <= 0 || Lh <= 0) {
        fprintf(stderr, "Error: Invalid inputs for fft_convolve_stream\n");
        return -1;
    }
    hs_cu_require_device();
    const long total = Lx + Lh - 1;
    void *dx = hs_cu_malloc(sizeof(fft_type) * (size_t)Lx), *dh = hs_cu_malloc(sizeof(fft_type) * (size_t)Lh);
    void *dy = hs_cu_malloc(sizeof(fft_type) * (size_t)total);
    hs_cu_h2d(dx, x, sizeof(fft_type) * (size_t)Lx, NULL);
    hs_cu_h2d(dh, h, sizeof(fft_type) * (size_t)Lh, NULL);
    const long n = fft_convolve_stream_dev((const fft_type *)dx, Lx, (const fft_type *)dh, Lh, (fft_type *)dy, P, NULL);
    if (n > 0) hs_cu_d2h(y, dy, sizeof(fft_type) * (size_t)n, NULL);
    hs_cu_stream_sync(NULL);
    hs_cu_free(dx); hs_cu_free(dh); hs_cu_free(dy);
    return n;
}
