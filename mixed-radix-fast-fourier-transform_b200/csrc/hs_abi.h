/*
 * hs_abi.h -- the thin C ABI between the host side (plain C: plan builder, public API,
 * hsfft_plan.c / hsfft_real.c / hsfft_convolve.c) and the CUDA side (hs_device.cu: memory,
 * streams, kernel launches).  Plain pointers and sizes only.  Not installed; the public
 * headers are include/highspeedFFT.h, include/real.h and include/hsfft_b200.h.
 *
 * Error convention (reference SURVEY.md section 5 "Metrics / logging"): every CUDA failure
 * prints one line to stderr and exit(EXIT_FAILURE)s, like the reference's own argument checks.
 */
#ifndef HS_ABI_H_
#define HS_ABI_H_
#ifndef __CUDACC_RTC__
#include <stddef.h>
#endif

#ifdef __cplusplus
extern "C" {
#endif

#define HS_MAX_STAGES 32
#define HS_MAX_GENERIC_RADIX 128

/* Split stage twiddles.  A stage of radix r on sub-transforms of length m needs W_L^(i k), L = r m,
 * i < r, k < m: (r-1) m table entries, each used once per transform -- for the last stage that is a
 * table as large as the signal itself, streamed from L2 with its latency exposed in every CTA.
 * When (r-1) m exceeds split_min entries the compile-time plans use k = k_lo + KL k_hi and
 * W_L^(i k) = W_L^(i k_lo) * W_L^(i KL k_hi): two tables of (r-1) KL and (r-1) m/KL entries that stay
 * L1 resident, at the price of one extra complex multiply per twiddle.  KL = largest divisor of m
 * not above sqrt(m).  Shared by the kernels (constexpr) and the host table builder (plain C). */
#ifdef __cplusplus
#define HS_RULE constexpr
#else
#define HS_RULE static inline
#endif
HS_RULE int hs_split_kl(int r, int m, int split_min)
{
    if (split_min <= 0 || (r - 1) * m <= split_min) return 0;
    int best = 1;
    for (int d = 1; d * d <= m; ++d)
        if (m % d == 0) best = d;
    return best > 1 ? best : 0;
}
HS_RULE int hs_stage_tw_entries(int r, int m, int split_min)
{
    return hs_split_kl(r, m, split_min) ? (r - 1) * (hs_split_kl(r, m, split_min) + m / hs_split_kl(r, m, split_min)) : (r - 1) * m;
}

/* compile-time kernel families */
enum { HS_CT_C2C = 0, HS_CT_BLUESTEIN = 1, HS_CT_TWOPASS = 2, HS_CT_SUB = 3, HS_CT_RADER = 4, HS_CT_BLUESTEIN_SPLIT = 5, HS_CT_CONV_SPLIT = 6,
       HS_CT_SUB2 = 7 /* sub-plans built for second passes only (strided stores, inter-pass twiddle on load) */ };

/* one compile-time specialised shared-memory transform the CUDA side offers */
typedef struct hs_ct_desc {
    int id;
    int family;     /* HS_CT_* */
    int n;          /* transform length held in shared memory */
    int nst;        /* stages */
    int radix[8];   /* execution order (leaf first) */
    int split_min;  /* twiddle-table split threshold of this plan (0 = never split), see hs_split_kl */
    int threads, smem_bytes;
    /* HS_CT_TWOPASS: n = n1 * n2; radix[]/nst/split_min describe the pass-1 sub-plan of length n1
     * (columns of stride n2), the fields below the pass-2 sub-plan of length n2 (columns of stride n1) */
    int n1, n2, nst2, radix2[8], split_min2;
    /* 1: the radix-3/5/7 stages use full-precision constants -- usable inside Bluestein only (a transform the
     * reference runs itself must reproduce its 11-digit literals) */
    int exact;
} hs_ct_desc;

/* parameters of one run-time planned pass (see hs_generic.cuh) */
typedef struct hs_gen_params {
    int n, nst, C, npad;
    int radix[HS_MAX_STAGES];
    int tw_off[HS_MAX_STAGES];
    int gen_off[HS_MAX_STAGES];
    long long n_c, n_o, n_batch;
    long long in_b, in_o, in_c, in_e;
    long long out_b, out_o, out_c, out_e;
    int load_c_fast, store_c_fast;
    long long tw_L;
    int tw_S;
    const void *tw, *tw_lo, *tw_hi;
    const double *gen_cs, *gen_sn;
} hs_gen_params;

/* geometry of one tiled pass (hs_tiled.cuh): `howmany` outer items x `ncols` columns, each column a
 * strided sub-transform; element e of column c of outer item x sits at
 *     x*in_xf + c*in_c + remap(e*in_e)          remap(o) = (o >> blk_log)*blk_stride + (o & (2^blk_log-1))
 * (blk_log == 0: remap is the identity; it un-blocks the layout an all-to-all leaves behind), outputs
 * likewise without the remap.
 * tw_mode: 0 none; 1 inter-pass twiddle W^(k b) * (W^(k NB))^ii with k = column; 2 same with k = outer item. */
typedef struct hs_tile_geom {
    long long in_xf, out_xf;
    long long in_c, in_e, out_c, out_e;
    int ncols;
    int blk_log;
    long long blk_stride;
    const void *twb, *tws;
    int tw_ld; /* row length of twb: entries per b */
    /* Bluestein fusions (io_mode of hs_cu_launch_tiled_io): the chirp h[n] and the logical length */
    const void *chirp;
    int nlog;
    /* convolution fusions (in_xf / out_xf must be 0 where a row stride in REALS applies):
     * io_mode 4 = row x is the nlog reals starting at real index x*pad_stride + pad_off of one long array of pad_total reals;
     *             reals outside [0, pad_total) and beyond the row's nlog read as zero (plain zero padding: pad_stride = nlog,
     *             pad_off = 0, pad_total = rows*nlog; overlap-save blocks: pad_stride = hop, pad_off = -(filter-1));
     * io_mode 5 = c2r merge of the product aux[k]*in[k]*scale, aux = second spectrum with row stride aux_xf (0: one shared row);
     * io_mode 6 = store only reals [sl_start, sl_start+sl_len) of every row, to index x*aux_xf + (r - sl_start) of one
     *             output array of sl_total reals (later indices are dropped) */
    const void *aux;
    long long aux_xf;
    double scale;
    int sl_start, sl_len;
    long long pad_stride, pad_off, pad_total, sl_total;
} hs_tile_geom;

/* Four-step exchange fused into phase A's last pass (io_mode 7 of hs_cu_launch_tiled_io): every output row k1 of the
 * rank's column slab is multiplied by the exchange twiddle W_N^(n2 k1) and stored straight into the receive buffer of
 * the rank that owns row k1 -- peer memory mapped through CUDA IPC, the stores travel over NVLink while the
 * pass computes.  One such struct per plan lives in device memory (hs_tile_geom.chirp points to it). */
#define HS_DIST_MAX_RANKS 16
typedef struct hs_dist_store {
    void *peer[HS_DIST_MAX_RANKS]; /* peer[g']: rank g''s receive buffer r[g][k1l][c] (own buffer for g' = rank) */
    const void *hi, *lo;           /* W_N^e = hi[e >> logS] * lo[e & (2^logS - 1)] */
    long long Cw, Kw;              /* columns / rows per rank */
    long long row_mul;             /* global row of output element k of outer item x: x + row_mul * k */
    int logKw, logN, logS, rank;
} hs_dist_store;

/* ---- device / memory / streams ---- */
int   hs_cu_device_count(void);
void  hs_cu_require_device(void);            /* exits with a message when no CUDA device is usable */
int   hs_cu_get_device(void);
void  hs_cu_set_device(int dev);
void *hs_cu_malloc(size_t bytes);
void  hs_cu_free(void *d);
/* stream-ordered allocation (cudaMallocAsync / cudaFreeAsync): scratch that grows on a caller's stream without a device-wide
 * synchronisation and without racing the kernels already queued there */
void *hs_cu_malloc_async(size_t bytes, void *stream);
void  hs_cu_free_async(void *d, void *stream);
void  hs_cu_device_sync(void);
int   hs_cu_can_access_peer(int dev, int peer);
void  hs_cu_enable_peer_access(int peer);    /* from the current device; already-enabled is not an error */
/* rows x width_bytes between host and device with pitches (cudaMemcpy2DAsync) */
void  hs_cu_h2d_2d(void *d, size_t dpitch, const void *h, size_t hpitch, size_t width_bytes, size_t rows, void *stream);
void  hs_cu_d2h_2d(void *h, size_t hpitch, const void *d, size_t dpitch, size_t width_bytes, size_t rows, void *stream);
void *hs_cu_malloc_host(size_t bytes);       /* pinned */
void  hs_cu_free_host(void *h);
void  hs_cu_h2d(void *d, const void *h, size_t bytes, void *stream);
void  hs_cu_d2h(void *h, const void *d, size_t bytes, void *stream);
void  hs_cu_d2d(void *dst, const void *src, size_t bytes, void *stream);
void  hs_cu_memset0(void *d, size_t bytes, void *stream);
void *hs_cu_stream_create(void);
void  hs_cu_stream_destroy(void *stream);
void  hs_cu_stream_sync(void *stream);
void *hs_cu_event_create(void);
void  hs_cu_event_destroy(void *ev);
void  hs_cu_event_record(void *ev, void *stream);
void  hs_cu_stream_wait_event(void *stream, void *ev);
float hs_cu_event_elapsed_ms(void *start, void *stop);
void  hs_cu_event_sync(void *ev);
int   hs_cu_sm_count(void);
long long hs_cu_launch_count(void);          /* kernels launched by this library so far */
/* per-launch device timing (CUDA events on the launching stream): begin, run calls, then report "tag:count:total_ms;..." */
void  hs_cu_timing_begin(void *stream);
int   hs_cu_timing_report(char *out, int cap);
/* CUDA IPC: a 64-byte handle of a hs_cu_malloc'ed buffer, and its mapping in another process on the same node */
void  hs_cu_ipc_get(void *d, void *handle64);
void *hs_cu_ipc_open(const void *handle64);
void  hs_cu_ipc_close(void *mapped);

/* ---- compile-time kernels ---- */
int   hs_cu_ct_count(void);
const hs_ct_desc *hs_cu_ct_get(int idx);
/* batched contiguous c2c, transform b at in + b*n */
void  hs_cu_launch_ct_c2c(int id, int sgn, const void *d_in, void *d_out, const void *d_tw, long long howmany, void *stream);
/* real transforms riding on the one-kernel complex plan `id` of length H = N/2 (hs_smem_fft.cuh):
 * mode 1 = r2c: d_in rows of N reals (contiguous), d_out rows of `stride` bins, bins 0..H and, when full, the mirrored N-k;
 * mode 2 = c2r: d_in rows of `stride` bins (0..H are read), d_out rows of N reals.  d_tw2 = exp(+2 pi i k / N), k < H. */
void  hs_cu_launch_ct_real(int id, int mode, int sgn, const void *d_in, void *d_out, const void *d_tw, const void *d_tw2, long long stride,
                           int full, long long howmany, void *stream);
int   hs_cu_ct_has_real(int id);
/* one pass (1 or 2) of a two-pass transform; d_tw: that pass's sub-plan twiddles; d_twb/d_tws: the
 * inter-pass twiddle tables (pass 2 only) */
void  hs_cu_launch_ct_twopass(int id, int pass, int sgn, const void *d_in, void *d_out, const void *d_tw, const void *d_twb,
                              const void *d_tws, long long howmany, void *stream);
/* building blocks for strided / distributed transforms: sub-plans (family HS_CT_SUB, n <= 1024) run by
 * the tiled kernel with explicit geometry */
void  hs_cu_launch_tiled(int sub_id, int sgn, int tw_mode, int stcol, const void *d_in, void *d_out, const void *d_tw,
                         const hs_tile_geom *g, long long howmany, void *stream);
/* same with a fused end (see hs_tile_geom for io_mode 3..7): io_mode 1 = load x[n]*conj(h[n]) (chirp_sgn=+1; *h[n] for -1) for n < nlog
 * and zero beyond (input rows are nlog long); io_mode 2 = store y[n]*conj(h[n]) (resp. *h[n]) for n < nlog only */
void  hs_cu_launch_tiled_io(int sub_id, int sgn, int tw_mode, int stcol, int io_mode, int chirp_sgn, const void *d_in, void *d_out,
                            const void *d_tw, const hs_tile_geom *g, long long howmany, void *stream);
/* middle kernel of the long Bluestein: per column k of the padded length M = n1*n2 (n2 = the sub-plan's
 * length): FFT_n2 (inter-pass twiddle on load) -> *Hk (conj for chirp_sgn=-1) -> inverse FFT_n2, written as
 * contiguous rows ws2[k][j'].  g describes the loads like pass 2 of a two-pass transform. */
void  hs_cu_launch_blu_mid(int sub_id, int chirp_sgn, const void *d_in, void *d_out, const void *d_twf, const void *d_twi,
                           const hs_tile_geom *g, const void *d_hk, long long howmany, void *stream);
/* b[k1][c] *= W_N^((c0 + c) * k1), N = 2^logN, via the split tables hi[e >> logS], lo[e & (2^logS - 1)] */
void  hs_cu_launch_dist_twiddle(void *b, const void *hi, const void *lo, long long rows, long long cols, long long c0, int logN,
                                int logS, void *stream);
/* fused Bluestein: d_tw_f / d_tw_i are the forward-order / inverse-order stage tables of length M */
/* Rader kernel for prime N = n + 1 (descriptor family HS_CT_RADER, length n): d_perm[q] = g^q mod N (n ints),
 * d_bk = FFT_n(b) / n with b[m] = exp(-sgn 2 pi i g^-m / N); d_tw_i belongs to the reversed radix list */
void  hs_cu_launch_ct_rader(int id, const void *d_in, void *d_out, const void *d_tw_f, const void *d_tw_i, const void *d_perm,
                            const void *d_bk, long long howmany, void *stream);
void  hs_cu_launch_ct_bluestein(int id, int sgn, int N, const void *d_in, void *d_out, const void *d_tw_f, const void *d_tw_i,
                                const void *d_chirp, const void *d_hk, long long howmany, void *stream);

/* split chirp-z (descriptor family HS_CT_BLUESTEIN_SPLIT, length M' >= N): two cyclic convolutions of length M' per transform in
 * one kernel.  d_chirp3 = h | h*conj(w) | h*w (N entries each, w[n] = exp(-2 pi i n / 2M')), d_hk2 = Hk[even] | Hk[odd] of
 * Hk = FFT_2M'(padded mirrored chirp / 2M') (M' entries each) */
void  hs_cu_launch_ct_bluestein_split(int id, int sgn, int N, const void *d_in, void *d_out, const void *d_tw_f, const void *d_tw_i,
                                      const void *d_chirp3, const void *d_hk2, long long howmany, void *stream);

/* the same on the in-place transforms of hs_conv_fft.cuh (descriptor family HS_CT_CONV_SPLIT): d_tw = the DIF stage table
 * of hs_build_conv_twiddles, d_hkt2 = Hk[even] | Hk[odd] each permuted by hs_conv_permute_hk, d_wp[p] = exp(-i pi p / M'),
 * p < M' / radix[0] (the variant that parks its intermediates in tensor memory rebuilds w from it) */
void  hs_cu_launch_ct_conv_split(int id, int sgn, int N, const void *d_in, void *d_out, const void *d_tw, const void *d_chirp3,
                                 const void *d_wp, const void *d_hkt2, long long howmany, void *stream);

/* ---- plan-time specialisation (hs_jit.cu): a compile-time kernel for any length whose stages fit one SM, compiled with
 * NVRTC when the plan is created.  radix_exec: stage radices in execution order (2,3,4,5,7,8,16 only).  create returns a
 * handle >= 0 and fills desc (radices, twiddle mode) or -1 (NVRTC missing, HSFFT_JIT=0, length unsuitable): the caller
 * then falls back to the run-time planned engine.  hs_cu_jit_source: the generated translation unit (tests). ---- */
int   hs_cu_jit_c2c_create(int n, int sgn, const int *radix_exec, int nst, hs_ct_desc *desc);
void  hs_cu_jit_c2c_launch(int handle, const void *d_in, void *d_out, const void *d_tw, long long howmany, void *stream);
int   hs_cu_jit_source(int n, int sgn, const int *radix_exec, int nst, char *out, int cap);
/* the real transforms of hs_cu_launch_ct_real on a plan-time kernel (same module, entry points hs_jit_r2c / hs_jit_c2r) */
int   hs_cu_jit_has_real(int handle);
int   hs_cu_jit_is_staged(int handle); /* large batches run as persistent CTAs with the next input staged by the copy engine */
/* Rader kernel (hs_fused.cuh) compiled for a prime N = L + 1 whose L is a product of compile-time radices, L <= 4096:
 * radix_exec = any ordering of L's radices (full-precision constants); launch arguments as hs_cu_launch_ct_rader */
int   hs_cu_jit_rader_create(int L, const int *radix_exec, int nst, hs_ct_desc *desc);
void  hs_cu_jit_rader_launch(int handle, const void *d_in, void *d_out, const void *d_tw_f, const void *d_tw_i, const void *d_perm,
                             const void *d_bk, long long howmany, void *stream);
/* two tiled passes compiled for this plan (N > 8192, all stages compile-time radices): radix_exec cut after `split` stages;
 * fills desc like a static HS_CT_TWOPASS descriptor; launch runs pass 1 or 2 over a contiguous batch */
int   hs_cu_jit_twopass_create(long long n, int sgn, const int *radix_exec, int nst, int split, hs_ct_desc *desc);
void  hs_cu_jit_twopass_launch(int handle, int pass, const void *d_in, void *d_out, const void *d_tw, const void *d_twb, const void *d_tws,
                               long long howmany, void *stream);
void  hs_cu_jit_real_launch(int handle, int mode, const void *d_in, void *d_out, const void *d_tw, const void *d_tw2, long long stride,
                            int full, long long howmany, void *stream);

/* ---- run-time planned pass and element-wise helpers ---- */
void  hs_cu_launch_generic(int sgn, const void *d_in, void *d_out, const hs_gen_params *p, void *stream);
void  hs_cu_launch_chirp_pad(int sgn, const void *x, void *a, const void *h, int N, int M, long long howmany, void *stream);
void  hs_cu_launch_spectrum_mul(int sgn, void *A, const void *Hk, int M, long long howmany, void *stream);
void  hs_cu_launch_chirp_crop(int sgn, const void *y, void *out, const void *h, int N, int M, long long howmany, void *stream);
void  hs_cu_launch_r2c_post(const void *Z, void *out, const void *tw2, int H, long long out_stride, int full, long long howmany,
                            void *stream);
void  hs_cu_launch_c2r_pre(const void *in, void *z, const void *tw2, int H, long long in_stride, long long howmany, void *stream);
void  hs_cu_launch_conv_mul(const void *A, const void *B, void *C, long long total, double scale, void *stream);
/* rows of len reals -> rows of P reals, zero padded; rows of P reals -> rows [start, start+len) */
void  hs_cu_launch_pad_rows(const void *src, int len, void *dst, int P, long long howmany, void *stream);
/* out[img][c][r] = in[img][r][c] for howmany rows x cols complex matrices */
void  hs_cu_launch_transpose(const void *in, void *out, int rows, int cols, long long howmany, void *stream);
void  hs_cu_launch_slice_rows(const void *src, int P, int start, int len, void *dst, long long dst_stride, long long howmany,
                              void *stream);

#ifdef __cplusplus
}
#endif
#endif
