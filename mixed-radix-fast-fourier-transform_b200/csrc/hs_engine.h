/*
 * hs_engine.h -- host-side (plain C) execution planner shared by hsfft_plan.c, hsfft_real.c and
 * hsfft_convolve.c.  Internal; the public surface is the three headers under include.
 *
 * A public plan (struct fft_set, a bare malloc block the caller may even free() directly --
 * reference src/real.c:264, test/test_mixedRadixFFT.cpp:126,305) carries no GPU state.  The GPU
 * state (HBM-resident stage twiddles, chirp tables, pass list, staging buffers, streams) lives in
 * an hs_dev_obj kept in a registry keyed by the plan's address and validated against the plan's
 * (N, sgn, lt) on every call; it is created on the first execute and released by free_fft.
 */
#ifndef HS_ENGINE_H_
#define HS_ENGINE_H_

#include "../../include/real.h"
#include "hs_abi.h"

#ifdef __cplusplus
extern "C" {
#endif

#include <pthread.h>

#define HS_MAX_PASS 6
#define HS_NSLOT 3 /* pipeline depth of the host-pointer batch path */

/* Scratch of the device-pointer entry points, one buffer per caller stream: two streams that execute the same plan
 * never share an intermediate, and a buffer grows with stream-ordered allocation (no device-wide synchronisation, no
 * race with the kernels already queued on that stream).  Guarded by the owning object's lock. */
#define HS_SCRATCH_MAX 8
typedef struct hs_scratch {
    void *stream[HS_SCRATCH_MAX], *buf[HS_SCRATCH_MAX];
    size_t cap[HS_SCRATCH_MAX];
    int used;
} hs_scratch;
void *hs_scratch_get(hs_scratch *s, void *stream, size_t bytes);
void  hs_scratch_free(hs_scratch *s);

/* batched, contiguous, out-of-place c2c of one length and direction */
typedef struct hs_c2c {
    int n, sgn;
    int ct_id;          /* >= 0: one compile-time kernel does the whole transform */
    void *d_tw;         /* its stage twiddles in HBM */
    int jit_id;         /* >= 0: a kernel compiled for this length when the plan was created (hs_jit.cu) */
    int tp_id;          /* >= 0: compile-time two-pass transform (n = n1 * n2) */
    int jtp_id;         /* >= 0: the two passes compiled for this length when the plan was created (hs_jit.cu) */
    int n1, n2;
    void *d_tw2, *d_twb, *d_tws; /* pass-2 stage twiddles and the inter-pass tables W_n^{k b}, W_n^{k NB} */
    void *d_mrg_wc, *d_mrg_wr; /* c2r merge on pass 1: exp(2 pi i c / 2n), c < n2, and exp(2 pi i n2 r / 2n), r < n1 (made on first use) */
    int npass;          /* otherwise: run-time planned passes */
    hs_gen_params pass[HS_MAX_PASS];
    void *d_pass_tw[HS_MAX_PASS], *d_pass_lo[HS_MAX_PASS], *d_pass_hi[HS_MAX_PASS];
    void *d_gen_cs, *d_gen_sn;
} hs_c2c;

typedef struct hs_dev_obj {
    const void *key;    /* plan address */
    int device;         /* the CUDA device this state lives on: one entry per (plan, device) */
    pthread_mutex_t lock; /* execute calls on one (plan, device) serialise on it (staging buffers, scratch look-up, launches) */
    hs_scratch ws_pool; /* device-pointer calls: workspace per caller stream */
    int N, sgn, lt;
    hs_c2c *direct;     /* lt == 0 */
    /* Bluestein (lt == 1) */
    int M;
    hs_c2c *fwdM, *invM;
    void *d_chirp, *d_hk;
    int blu_ct_id;
    int rader_jit;   /* rader_ct_id is a handle of hs_cu_jit_rader_create (kernel compiled for this prime) instead of a static kernel id */
    int rader_ct_id; /* >= 0: prime N run as a cyclic convolution of length N-1 (d_chirp = int perm[N-1], d_hk = Bk) */
    int split_conv;  /* the split kernel runs on the in-place transforms of hs_conv_fft.cuh (d_blu_twf = their stage table) */
    int split_ct_id; /* >= 0: split chirp-z, two convolutions of length M/2 >= N in one kernel (d_chirp = h | h conj(w) | h w,
                      * d_hk = Hk[even] | Hk[odd]) */
    void *d_blu_twf, *d_blu_twi;
    /* long Bluestein (M = n1*n2 two-pass): three fused kernels, see hs_tiled.cuh */
    int big, big_n1, big_n2, big_sub1, big_sub2;
    int big_tp_id;
    void *d_big_f1, *d_big_f2, *d_big_i1, *d_big_i2, *d_big_fb, *d_big_fs, *d_big_ib, *d_big_is;
    long long big_chunk;
    /* pipeline slots for host-pointer calls */
    void *stream[HS_NSLOT];
    void *d_in[HS_NSLOT], *d_out[HS_NSLOT], *d_ws[HS_NSLOT], *d_ws2[HS_NSLOT];
    size_t cap_in[HS_NSLOT], cap_out[HS_NSLOT], cap_ws[HS_NSLOT], cap_ws2[HS_NSLOT];
    struct hs_dev_obj *next;
} hs_dev_obj;

/* plan tables built on the host with the reference's own expressions */
void hs_build_stage_twiddles(fft_data *out, const int *radix_exec, int nst, int sgn, int split_min); /* [i-1][k] per stage */
int  hs_stage_twiddle_count(const int *radix_exec, int nst, int split_min);
void hs_build_chirp(fft_data *h, int N);                                              /* exp(+i pi n^2/N) */
int  hs_bluestein_exec_pad(int N);                                                    /* 2^ceil(log2(2N-1)) */

/* plan-building helpers (hsfft_plan.c) */
const hs_ct_desc *hs_find_ct(int family, int n, int allow_exact);
const hs_ct_desc *hs_find_sub_second(int n);
void *hs_upload_stage_twiddles(const int *radix_exec, int nst, int sgn, int split_min);
void *hs_upload_interpass_tables(long long n, int sgn, int nb, int K, int kvalid, void **d_tws);

/* c2c sub-plans */
hs_c2c *hs_c2c_create(int n, int sgn);
void    hs_c2c_destroy(hs_c2c *p);
size_t  hs_c2c_ws_elems(const hs_c2c *p, long long howmany); /* workspace (complex elements) exec needs */
void    hs_c2c_exec(const hs_c2c *p, const void *d_in, void *d_out, void *d_ws, long long howmany, void *stream);

/* r2c with the split fused into the (one-kernel) complex transform; returns 0 when this plan has no such form */
int     hs_dev_exec_r2c_fused(hs_dev_obj *d, const void *d_in, void *d_out, long long out_stride, int full, const void *d_tw2, long long howmany,
                              void *stream);
/* c2r with the merge pass fused into the first pass of the half-length transform; returns 0 when this plan
 * has no such fusion (the caller then runs the separate merge kernel) */
int     hs_dev_exec_c2r_fused(hs_dev_obj *d, const void *d_spec, long long in_stride, const void *d_tw2, void *d_out, void *d_ws,
                              long long howmany, void *stream);

/* convolution fusions over a two-pass half-length plan (return 0 when the plan has none):
 * rows of len reals zero padded on pass 1's loads -> half-length spectrum d_z (the caller splits it);
 * inverse real transform of scale*A*B with the product on pass 1's merge loads and the output slice on pass 2's stores */
int     hs_dev_exec_padded_real(hs_dev_obj *d, const void *d_rows, int len, long long row_stride, long long row_off, long long total, void *d_z,
                                void *d_ws, long long howmany, void *stream);
int     hs_dev_exec_c2r_mul_slice(hs_dev_obj *d, const void *d_A, const void *d_B, long long in_stride, long long b_stride, double scale,
                                  const void *d_tw2, void *d_out, long long out_stride, long long out_total, int start, int len, void *d_ws,
                                  long long howmany, void *stream);
/* hsfft_real.c internals used by hsfft_convolve.c: r2c of zero-padded windows of a real array to half spectra (window x
 * = len reals from index x*row_stride + row_off, zero outside [0, total)); c2r of scale*A*B (B row stride b_stride, 0 = one
 * shared spectrum) keeping reals [start, start+len) of every row at output index x*out_stride.., clipped to out_total */
int     hs_real_r2c_padded_dev(fft_real_object obj, const fft_type *d_rows, int len, long long row_stride, long long row_off, long long total,
                               fft_data *d_out, long out_stride, long howmany, void *stream);
int     hs_real_c2r_mul_slice_dev(fft_real_object obj, const fft_data *d_A, const fft_data *d_B, long in_stride, long b_stride, double scale,
                                  fft_type *d_out, long long out_stride, long long out_total, int start, int len, long howmany, void *stream);

/* registry, keyed by (plan address, current device) */
hs_dev_obj *hs_dev_get(fft_object plan);      /* find or build on the current device; exits when no CUDA device */
void        hs_dev_lock(hs_dev_obj *d);
void        hs_dev_unlock(hs_dev_obj *d);
/* batch sharding over the devices chosen with hsfft_set_devices (hsfft_plan.c): fn(ctx, first, count) runs once per
 * device, each on its own host thread with that device current; a single device (the default) calls fn directly */
typedef void (*hs_shard_fn)(void *ctx, long first, long count);
void        hs_shard_run(long howmany, hs_shard_fn fn, void *ctx);
void        hs_dev_release(const void *key);  /* no-op when absent */
void        hs_dev_reserve(hs_dev_obj *d, int slot, size_t in_bytes, size_t out_bytes, size_t ws_bytes, size_t ws2_bytes);

/* device-pointer execution of a full plan (mixed radix or Bluestein) */
size_t hs_dev_ws_bytes(const hs_dev_obj *d, long long howmany);
void   hs_dev_exec(hs_dev_obj *d, const void *d_in, void *d_out, void *d_ws, long long howmany, void *stream);

#ifdef __cplusplus
}
#endif
#endif
