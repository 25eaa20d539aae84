/*
 * hsfft_real.c -- real-input transforms (host side, plain C).
 *
 * Replaces reference src/real.c: fft_real_init :26-64, fft_r2c_exec :78-136, fft_c2r_exec :150-193,
 * free_real_fft :259-267 and the helper exports :198-257.
 *
 * The reference packs N reals as N/2 complex samples (a pure reinterpretation of the same bytes,
 * :99-103), runs the half-length complex plan, and splits / merges with the table
 * twiddle2[k] = exp(+2 pi i k / N).  Here the pack is free (the device reads the real rows as
 * complex rows), the half-length transform runs on the GPU through the complex plan's own
 * execution plan, and the split (r2c) / merge (c2r) arithmetic keeps the reference's expressions.
 */
#include "hs_engine.h"
#include "../../include/hsfft_b200.h"

#include <pthread.h>

/* device state of a real plan: twiddle2 in HBM, keyed by the real plan's address */
typedef struct hs_real_dev {
    const void *key;
    int device;            /* one entry per (real plan, device) */
    pthread_mutex_t lock;  /* execute calls on one (plan, device) serialise on it */
    hs_scratch z_pool, ws_pool; /* device-pointer calls: intermediates per caller stream */
    int H;
    void *d_tw2;
    /* real four-step (N = n1 n2, both with a tiled sub-plan; forward plans): see hs_tiled.cuh */
    int r4, r4_n1, r4_n2, r4_sub1, r4_sub2, r4_ld;
    void *d_r4_tw1, *d_r4_tw2, *d_r4_twb, *d_r4_tws;
    void *stream[HS_NSLOT];
    void *d_in[HS_NSLOT], *d_out[HS_NSLOT], *d_z[HS_NSLOT], *d_ws[HS_NSLOT];
    size_t cap_in[HS_NSLOT], cap_out[HS_NSLOT], cap_z[HS_NSLOT], cap_ws[HS_NSLOT];
    struct hs_real_dev *next;
} hs_real_dev;

static pthread_mutex_t r_lock = PTHREAD_MUTEX_INITIALIZER;
static hs_real_dev *r_head = NULL;

static void real_dev_clear(hs_real_dev *r)
{
    hs_scratch_free(&r->z_pool); hs_scratch_free(&r->ws_pool);
    hs_cu_free(r->d_tw2);
    hs_cu_free(r->d_r4_tw1); hs_cu_free(r->d_r4_tw2); hs_cu_free(r->d_r4_twb); hs_cu_free(r->d_r4_tws);
    for (int s = 0; s < HS_NSLOT; ++s) {
        hs_cu_free(r->d_in[s]); hs_cu_free(r->d_out[s]); hs_cu_free(r->d_z[s]); hs_cu_free(r->d_ws[s]);
        if (r->stream[s]) hs_cu_stream_destroy(r->stream[s]);
    }
}

/* Real four-step for long forward transforms: N = n1 * n2 with a tiled sub-plan for both lengths (2^15 .. 2^20).
 * HSFFT_NO_R2C_FUSE=1 keeps the half-length complex transform followed by the split kernel. */
static void real_fourstep_setup(fft_real_object obj, hs_real_dev *r)
{
    r->r4 = 0;
    const long long N = 2LL * r->H;
    if (obj->cobj->sgn != 1 || obj->cobj->lt != 0 || N < (1 << 15) || getenv("HSFFT_NO_R2C_FUSE")) return;
    /* N = n1 * n2, both sub-lengths from 128..1024.  The second pass (mirrored stores, twice the bytes of the first) decides:
     * measured on B200 per 2^28 reals (profiles/r2_subplan_second_pass_and_r4_splits.txt, r2_r4_mirror_tile_width.txt) it
     * takes 1.09 ms with n2 = 256 and 1.10-1.14 with 128 (16 columns per tile: 256-byte runs), 1.35 with 512 (the second-pass
     * plan, 128-byte runs) and 1.57 with 1024 (64-byte runs), while the first pass takes 0.76-0.89 ms whatever n1 is (512 and
     * 1024 staged by the copy engine).  So: the best n2 whose partner exists -- 2^15 128x256, 2^16 256x256, 2^17 512x256,
     * 2^18 1024x256 1.88 ms, 2^19 1024x512 2.17, 2^20 1024x1024 2.39 (round 1's choices: 2.31 / 2.41 / 2.62). */
    static const int order2[4] = {256, 128, 512, 1024};
    int n1 = 0, n2 = 0;
    for (int i = 0; i < 4 && !n1; ++i)
        if (N % order2[i] == 0 && N / order2[i] >= 128 && N / order2[i] <= 1024) { n2 = order2[i]; n1 = (int)(N / n2); }
    if (!n1) return;
    if (getenv("HSFFT_R4_SWAP")) { const int t = n1; n1 = n2; n2 = t; }
    if (getenv("HSFFT_R4_N1")) { /* experiments: force the first sub-length */
        const int f = atoi(getenv("HSFFT_R4_N1"));
        if (f >= 128 && f <= 1024 && N % f == 0 && N / f >= 128 && N / f <= 1024) { n1 = f; n2 = (int)(N / f); }
    }
    const hs_ct_desc *s1 = hs_find_ct(HS_CT_SUB, n1, 0), *s2 = hs_find_sub_second(n2);
    if (!s1 || !s2) return;
    r->r4_n1 = n1; r->r4_n2 = n2; r->r4_sub1 = s1->id; r->r4_sub2 = s2->id;
    r->r4_ld = n1 / 2 + 2; /* rows k1 = 0..n1/2 of every real column, padded to an even count (32-byte aligned rows) */
    r->d_r4_tw1 = hs_upload_stage_twiddles(s1->radix, s1->nst, 1, s1->split_min);
    r->d_r4_tw2 = hs_upload_stage_twiddles(s2->radix, s2->nst, 1, s2->split_min);
    r->d_r4_twb = hs_upload_interpass_tables(N, 1, n2 / s2->radix[0], r->r4_ld, n1 / 2 + 1, &r->d_r4_tws);
    r->r4 = 1;
}

/* complex elements of the per-transform intermediate (half-length spectrum, or the four-step's padded rows) */
static size_t real_z_elems(const hs_real_dev *r) { return r->r4 ? (size_t)r->r4_n2 * r->r4_ld : (size_t)r->H; }

static void real_drop_key_locked(const void *key)
{
    hs_real_dev **pp = &r_head;
    while (*pp) {
        hs_real_dev *r = *pp;
        if (r->key == key) {
            *pp = r->next;
            real_dev_clear(r);
            pthread_mutex_destroy(&r->lock);
            free(r);
        } else pp = &r->next;
    }
}

static hs_real_dev *real_dev_get(fft_real_object obj)
{
    hs_cu_require_device();
    const int H = obj->cobj->N, dev = hs_cu_get_device();
    pthread_mutex_lock(&r_lock);
    hs_real_dev *r = r_head;
    while (r && !(r->key == obj && r->device == dev)) r = r->next;
    if (r && r->H != H) { /* address re-used by a different plan */
        real_drop_key_locked(obj);
        r = NULL;
    }
    if (!r) {
        r = (hs_real_dev *)calloc(1, sizeof(*r));
        if (!r) { fprintf(stderr, "Error: Memory allocation failed for real FFT object\n"); exit(EXIT_FAILURE); }
        pthread_mutexattr_t at;
        pthread_mutexattr_init(&at);
        pthread_mutexattr_settype(&at, PTHREAD_MUTEX_RECURSIVE);
        pthread_mutex_init(&r->lock, &at);
        pthread_mutexattr_destroy(&at);
        r->key = obj;
        r->device = dev;
        r->H = H;
        r->d_tw2 = hs_cu_malloc(sizeof(fft_data) * (size_t)H);
        hs_cu_h2d(r->d_tw2, obj->twiddle2, sizeof(fft_data) * (size_t)H, NULL);
        for (int s = 0; s < HS_NSLOT; ++s) r->stream[s] = hs_cu_stream_create();
        real_fourstep_setup(obj, r);
        r->next = r_head;
        r_head = r;
    }
    pthread_mutex_unlock(&r_lock);
    return r;
}

static void real_dev_release(const void *key)
{
    pthread_mutex_lock(&r_lock);
    real_drop_key_locked(key);
    pthread_mutex_unlock(&r_lock);
}

static void grow(void **buf, size_t *cap, size_t need)
{
    if (need <= *cap) return;
    hs_cu_free(*buf);
    *buf = hs_cu_malloc(need);
    *cap = need;
}

/* ---------------------------------------------------------------------------------- public API */

fft_real_object fft_real_init(int signal_length, int transform_direction)
{
    if (signal_length <= 0 || signal_length % 2 != 0) {
        fprintf(stderr, "Error: Signal length (%d) must be positive and even\n", signal_length);
        exit(EXIT_FAILURE);
    }
    const int half = signal_length / 2;
    fft_real_object obj = (fft_real_object)malloc(sizeof(struct fft_real_set) + sizeof(fft_data) * half);
    if (obj == NULL) {
        fprintf(stderr, "Error: Memory allocation failed for real FFT object\n");
        exit(EXIT_FAILURE);
    }
    obj->cobj = fft_init(half, transform_direction);
    if (obj->cobj == NULL) {
        free(obj);
        fprintf(stderr, "Error: Failed to initialize complex FFT object\n");
        exit(EXIT_FAILURE);
    }
    for (int k = 0; k < half; ++k) {
        const fft_type angle = PI2 * k / signal_length;
        obj->twiddle2[k].re = cos(angle);
        obj->twiddle2[k].im = sin(angle);
    }
    return obj;
}

static void check_real_args(fft_real_object obj, const void *a, const void *b)
{
    if (obj == NULL || a == NULL || b == NULL || obj->cobj == NULL) {
        fprintf(stderr, "Error: Invalid real FFT object or data pointers\n");
        exit(EXIT_FAILURE);
    }
}

/* device-pointer cores; d_z holds howmany*H complex, d_ws the complex plan's workspace */
static void r2c_dev(fft_real_object obj, hs_real_dev *r, const fft_type *d_in, fft_data *d_out, long out_stride, int full, void *d_z,
                    void *d_ws, long howmany, void *stream)
{
    hs_dev_obj *c = hs_dev_get(obj->cobj);
    if (r->r4) { /* long transforms: real four-step, the split rides on pass 1, the mirror on pass 2's stores (d_z = workspace) */
        const long long n1 = r->r4_n1, n2 = r->r4_n2, ld = r->r4_ld;
        hs_tile_geom g;
        memset(&g, 0, sizeof g);
        g.ncols = (int)(n2 / 2); g.in_xf = r->H; g.in_c = 1; g.in_e = n2 / 2; g.out_xf = n2 * ld; g.out_c = 2 * ld; g.out_e = 1;
        hs_cu_launch_tiled_io(r->r4_sub1, 1, 0, 0, 8, 1, d_in, d_z, r->d_r4_tw1, &g, howmany, stream);
        memset(&g, 0, sizeof g);
        g.ncols = (int)(n1 / 2 + 1); g.in_xf = n2 * ld; g.in_c = 1; g.in_e = ld; g.out_xf = out_stride;
        g.twb = r->d_r4_twb; g.tws = r->d_r4_tws; g.tw_ld = (int)ld; g.nlog = (int)n1; g.sl_start = full;
        hs_cu_launch_tiled_io(r->r4_sub2, 1, 1, 1, 9, 1, d_z, d_out, r->d_r4_tw2, &g, howmany, stream);
        return;
    }
    /* short transforms: the split rides on the one-kernel complex transform (the spectrum is split in shared memory) */
    if (hs_dev_exec_r2c_fused(c, d_in, d_out, out_stride, full, r->d_tw2, howmany, stream)) return;
    hs_dev_exec(c, d_in, d_z, d_ws, howmany, stream); /* rows of N reals == rows of H complex */
    hs_cu_launch_r2c_post(d_z, d_out, r->d_tw2, r->H, out_stride, full, howmany, stream);
}

static void c2r_dev(fft_real_object obj, hs_real_dev *r, const fft_data *d_in, long in_stride, fft_type *d_out, void *d_z, void *d_ws,
                    long howmany, void *stream)
{
    hs_dev_obj *c = hs_dev_get(obj->cobj);
    /* long transforms: the merge rides on pass 1's loads (d_z then serves as the pass workspace) */
    if (hs_dev_exec_c2r_fused(c, d_in, in_stride, r->d_tw2, d_out, d_z, howmany, stream)) return;
    hs_cu_launch_c2r_pre(d_in, d_z, r->d_tw2, r->H, in_stride, howmany, stream);
    hs_dev_exec(c, d_z, d_out, d_ws, howmany, stream);
}

void fft_r2c_exec_batch_dev(fft_real_object obj, const fft_type *d_in, fft_data *d_out, long out_stride, int full, long howmany,
                            void *stream)
{
    check_real_args(obj, d_in, d_out);
    if (howmany <= 0) return;
    hs_real_dev *r = real_dev_get(obj);
    if (out_stride < r->H + 1 || (full && out_stride < 2 * r->H)) {
        fprintf(stderr, "Error: r2c output stride %ld too small for N = %d\n", out_stride, 2 * r->H);
        exit(EXIT_FAILURE);
    }
    if ((const void *)d_in == (const void *)d_out) { fprintf(stderr, "Error: fft_r2c_exec_batch_dev is out of place\n"); exit(EXIT_FAILURE); }
    hs_dev_obj *c = hs_dev_get(obj->cobj);
    pthread_mutex_lock(&r->lock);
    void *z = hs_scratch_get(&r->z_pool, stream, sizeof(fft_data) * (size_t)howmany * real_z_elems(r));
    void *ws = hs_scratch_get(&r->ws_pool, stream, hs_dev_ws_bytes(c, howmany));
    r2c_dev(obj, r, d_in, d_out, out_stride, full, z, ws, howmany, stream);
    pthread_mutex_unlock(&r->lock);
}

void fft_c2r_exec_batch_dev(fft_real_object obj, const fft_data *d_in, long in_stride, fft_type *d_out, long howmany, void *stream)
{
    check_real_args(obj, d_in, d_out);
    if (howmany <= 0) return;
    hs_real_dev *r = real_dev_get(obj);
    if (in_stride < r->H + 1) {
        fprintf(stderr, "Error: c2r input stride %ld too small for N = %d\n", in_stride, 2 * r->H);
        exit(EXIT_FAILURE);
    }
    if ((const void *)d_in == (const void *)d_out) { fprintf(stderr, "Error: fft_c2r_exec_batch_dev is out of place\n"); exit(EXIT_FAILURE); }
    hs_dev_obj *c = hs_dev_get(obj->cobj);
    pthread_mutex_lock(&r->lock);
    void *z = hs_scratch_get(&r->z_pool, stream, sizeof(fft_data) * (size_t)howmany * real_z_elems(r));
    void *ws = hs_scratch_get(&r->ws_pool, stream, hs_dev_ws_bytes(c, howmany));
    c2r_dev(obj, r, d_in, in_stride, d_out, z, ws, howmany, stream);
    pthread_mutex_unlock(&r->lock);
}

int hs_real_r2c_padded_dev(fft_real_object obj, const fft_type *d_rows, int len, long long row_stride, long long row_off, long long total,
                           fft_data *d_out, long out_stride, long howmany, void *stream)
{
    check_real_args(obj, d_rows, d_out);
    hs_real_dev *r = real_dev_get(obj);
    hs_dev_obj *c = hs_dev_get(obj->cobj);
    if (c->lt != 0 || !c->direct || c->direct->tp_id < 0) return 0;
    pthread_mutex_lock(&r->lock);
    void *z = hs_scratch_get(&r->z_pool, stream, sizeof(fft_data) * (size_t)howmany * real_z_elems(r));
    if (r->r4 && !getenv("HSFFT_NO_CONV_FUSE")) { /* real four-step with the zero padding on pass 1's loads: two launches, no split kernel */
        const long long n1 = r->r4_n1, n2 = r->r4_n2, ld = r->r4_ld;
        hs_tile_geom g;
        memset(&g, 0, sizeof g);
        g.ncols = (int)(n2 / 2); g.in_xf = 0; g.in_c = 1; g.in_e = n2 / 2; g.out_xf = n2 * ld; g.out_c = 2 * ld; g.out_e = 1;
        g.nlog = len; g.pad_stride = row_stride; g.pad_off = row_off; g.pad_total = total;
        hs_cu_launch_tiled_io(r->r4_sub1, 1, 0, 0, 10, 1, d_rows, z, r->d_r4_tw1, &g, howmany, stream);
        memset(&g, 0, sizeof g);
        g.ncols = (int)(n1 / 2 + 1); g.in_xf = n2 * ld; g.in_c = 1; g.in_e = ld; g.out_xf = out_stride;
        g.twb = r->d_r4_twb; g.tws = r->d_r4_tws; g.tw_ld = (int)ld; g.nlog = (int)n1; g.sl_start = 0;
        hs_cu_launch_tiled_io(r->r4_sub2, 1, 1, 1, 9, 1, z, d_out, r->d_r4_tw2, &g, howmany, stream);
        pthread_mutex_unlock(&r->lock);
        return 1;
    }
    void *ws = hs_scratch_get(&r->ws_pool, stream, sizeof(fft_data) * (size_t)howmany * real_z_elems(r));
    const int ok = hs_dev_exec_padded_real(c, d_rows, len, row_stride, row_off, total, z, ws, howmany, stream);
    if (ok) hs_cu_launch_r2c_post(z, d_out, r->d_tw2, r->H, out_stride, 0, howmany, stream);
    pthread_mutex_unlock(&r->lock);
    return ok;
}

int hs_real_c2r_mul_slice_dev(fft_real_object obj, const fft_data *d_A, const fft_data *d_B, long in_stride, long b_stride, double scale,
                              fft_type *d_out, long long out_stride, long long out_total, int start, int len, long howmany, void *stream)
{
    check_real_args(obj, d_A, d_out);
    hs_real_dev *r = real_dev_get(obj);
    hs_dev_obj *c = hs_dev_get(obj->cobj);
    if (c->lt != 0 || !c->direct || c->direct->tp_id < 0) return 0;
    pthread_mutex_lock(&r->lock);
    void *ws = hs_scratch_get(&r->ws_pool, stream, sizeof(fft_data) * (size_t)howmany * real_z_elems(r));
    const int ok = hs_dev_exec_c2r_mul_slice(c, d_A, d_B, in_stride, b_stride, scale, r->d_tw2, d_out, out_stride, out_total, start, len, ws,
                                             howmany, stream);
    pthread_mutex_unlock(&r->lock);
    return ok;
}

static long real_chunk(long howmany, size_t row_bytes)
{
    long c = (long)(((size_t)64 << 20) / row_bytes);
    if (c < 1) c = 1;
    return c > howmany ? howmany : c;
}

typedef struct real_ctx { fft_real_object obj; void *inp, *oup; long total; } real_ctx;

static void r2c_batch_one_device(void *vctx, long first0, long howmany)
{
    real_ctx *x = (real_ctx *)vctx;
    fft_real_object obj = x->obj;
    if (howmany <= 0) return;
    hs_real_dev *r = real_dev_get(obj);
    hs_dev_obj *c = hs_dev_get(obj->cobj);
    pthread_mutex_lock(&r->lock);
    const size_t N = 2 * (size_t)r->H, in_row = sizeof(fft_type) * N, out_row = sizeof(fft_data) * N;
    fft_type *inp = (fft_type *)x->inp + (size_t)first0 * N;
    fft_data *oup = (fft_data *)x->oup + (size_t)first0 * N;
    const long chunk = real_chunk(howmany, out_row), nchunk = (howmany + chunk - 1) / chunk;
    for (long i = 0; i < nchunk; ++i) {
        const int s = (int)(i % HS_NSLOT);
        const long first = i * chunk, cnt = (first + chunk <= howmany) ? chunk : howmany - first;
        if (i < HS_NSLOT) {
            grow(&r->d_in[s], &r->cap_in[s], in_row * chunk);
            grow(&r->d_out[s], &r->cap_out[s], out_row * chunk);
            grow(&r->d_z[s], &r->cap_z[s], sizeof(fft_data) * (size_t)chunk * real_z_elems(r));
            grow(&r->d_ws[s], &r->cap_ws[s], hs_dev_ws_bytes(c, chunk));
        }
        hs_cu_h2d(r->d_in[s], inp + first * N, in_row * cnt, r->stream[s]);
        r2c_dev(obj, r, (const fft_type *)r->d_in[s], (fft_data *)r->d_out[s], (long)N, 1, r->d_z[s], r->d_ws[s], cnt, r->stream[s]);
        hs_cu_d2h(oup + first * N, r->d_out[s], out_row * cnt, r->stream[s]);
    }
    for (int s = 0; s < HS_NSLOT; ++s) hs_cu_stream_sync(r->stream[s]);
    pthread_mutex_unlock(&r->lock);
}

void fft_r2c_exec_batch(fft_real_object obj, fft_type *inp, fft_data *oup, long howmany)
{
    check_real_args(obj, inp, oup);
    if (howmany <= 0) return;
    real_ctx x = {obj, inp, oup, howmany};
    hs_shard_run(howmany, r2c_batch_one_device, &x);
}

static void c2r_batch_one_device(void *vctx, long first0, long howmany)
{
    real_ctx *x = (real_ctx *)vctx;
    fft_real_object obj = x->obj;
    if (howmany <= 0) return;
    hs_real_dev *r = real_dev_get(obj);
    hs_dev_obj *c = hs_dev_get(obj->cobj);
    pthread_mutex_lock(&r->lock);
    const size_t N = 2 * (size_t)r->H, in_row = sizeof(fft_data) * N, out_row = sizeof(fft_type) * N;
    fft_data *inp = (fft_data *)x->inp + (size_t)first0 * N;
    fft_type *oup = (fft_type *)x->oup + (size_t)first0 * N;
    const long chunk = real_chunk(howmany, in_row), nchunk = (howmany + chunk - 1) / chunk;
    for (long i = 0; i < nchunk; ++i) {
        const int s = (int)(i % HS_NSLOT);
        const long first = i * chunk, cnt = (first + chunk <= howmany) ? chunk : howmany - first;
        if (i < HS_NSLOT) {
            grow(&r->d_in[s], &r->cap_in[s], in_row * chunk);
            grow(&r->d_out[s], &r->cap_out[s], out_row * chunk);
            grow(&r->d_z[s], &r->cap_z[s], sizeof(fft_data) * (size_t)chunk * real_z_elems(r));
            grow(&r->d_ws[s], &r->cap_ws[s], hs_dev_ws_bytes(c, chunk));
        }
        /* a single transform is given as N/2+1 bins by reference-style callers; rows of a batch are N bins apart */
        const size_t bytes = (x->total == 1) ? sizeof(fft_data) * (size_t)(r->H + 1) : in_row * cnt;
        hs_cu_h2d(r->d_in[s], inp + first * N, bytes, r->stream[s]);
        c2r_dev(obj, r, (const fft_data *)r->d_in[s], (long)N, (fft_type *)r->d_out[s], r->d_z[s], r->d_ws[s], cnt, r->stream[s]);
        hs_cu_d2h(oup + first * N, r->d_out[s], out_row * cnt, r->stream[s]);
    }
    for (int s = 0; s < HS_NSLOT; ++s) hs_cu_stream_sync(r->stream[s]);
    pthread_mutex_unlock(&r->lock);
}

void fft_c2r_exec_batch(fft_real_object obj, fft_data *inp, fft_type *oup, long howmany)
{
    check_real_args(obj, inp, oup);
    if (howmany <= 0) return;
    real_ctx x = {obj, inp, oup, howmany};
    hs_shard_run(howmany, c2r_batch_one_device, &x);
}

void fft_r2c_exec(fft_real_object real_obj, fft_type *input_data, fft_data *output_data)
{
    fft_r2c_exec_batch(real_obj, input_data, output_data, 1);
}

void fft_c2r_exec(fft_real_object real_obj, fft_data *input_data, fft_type *output_data)
{
    fft_c2r_exec_batch(real_obj, input_data, output_data, 1);
}

void free_real_fft(fft_real_object real_obj)
{
    if (real_obj != NULL) {
        real_dev_release(real_obj);
        free_fft(real_obj->cobj);
        free(real_obj);
    }
}

/* ------------------------------------------------- helper exports of the reference's real.c */
#define HS_M_PI 3.14159265358979323846

void generate_signal(fft_data *signal, int length, double freq, double amplitude)
{
    for (int i = 0; i < length; i++) {
        signal[i].re = amplitude * sin(2.0 * HS_M_PI * freq * i / length);
        signal[i].im = 0.0;
    }
}

double compute_mse(fft_data *original, fft_data *reconstructed, int length)
{
    double acc = 0.0;
    for (int i = 0; i < length; i++) {
        const double dr = original[i].re - reconstructed[i].re, di = original[i].im - reconstructed[i].im;
        acc += dr * dr + di * di;
    }
    return acc / length;
}

void print_complex(fft_data *data, int length, const char *label)
{
    printf("%s:\n", label);
    for (int i = 0; i < length; i++) printf("  [%d] %.6f + %.6fi\n", i, data[i].re, data[i].im);
    printf("\n");
}

void generate_real_signal(fft_type *signal, int length, double freq, double amplitude)
{
    for (int i = 0; i < length; i++) signal[i] = amplitude * sin(2.0 * HS_M_PI * freq * i / length);
}

double compute_mse_real(fft_type *original, fft_type *reconstructed, int length)
{
    double acc = 0.0;
    for (int i = 0; i < length; i++) {
        const double d = original[i] - reconstructed[i];
        acc += d * d;
    }
    return acc / length;
}

void print_real(fft_type *data, int length, const char *label)
{
    printf("%s:\n", label);
    for (int i = 0; i < length; i++) printf("  [%d] %.6f\n", i, data[i]);
    printf("\n");
}
