/*
 * hsfft_dist.c -- local phases of the distributed four-step transform (host side, plain C).
 *
 * No counterpart in the reference (it recurses on one core: N = 2^27 takes 19 s there, SURVEY.md
 * section 5 "long-context" row); this is north_star's "single very large 1D transform ... four-step
 * decomposition whose transpose is an NCCL all-to-all".
 *
 *   N = N1 * N2,  n = N2 n1 + n2,  k = k1 + N1 k2,  G ranks, Cw = N2 / G columns and Kw = N1 / G rows per rank
 *   X[k1 + N1 k2] = sum_n2 W_N2^(n2 k2) * [ W_N^(n2 k1) * sum_n1 x[N2 n1 + n2] W_N1^(n1 k1) ]
 *
 *   rank g holds   a[n1][c]      = x[N2 n1 + g Cw + c]              (its column slab of the N1 x N2 matrix)
 *   phase A (hsfft_dist_cols_dev):  b[k1][c] = W_N^((g Cw + c) k1) * FFT_N1 over n1          -- local
 *   exchange (caller, NCCL all-to-all): rows [g' Kw, (g'+1) Kw) of b go to rank g'; a rank receives
 *                                   r[g][k1l][c] -- contiguous blocks on both sides, no pack kernel
 *   phase B (hsfft_dist_rows_dev):  out[k1l][k2] = FFT_N2 over n2 of r[n2 / Cw][k1l][n2 % Cw]  -- local,
 *                                   the un-blocking of the received layout is fused into the loads
 *   rank g ends with out[k1l][k2] = X[(g Kw + k1l) + N1 k2]         (its row slab, transposed order)
 *
 * The length-N1 / length-N2 transforms are the same compile-time sub-plans as everywhere else, run
 * by the tiled kernel (hs_tiled.cuh) in one pass (128..1024) or two (2^14..2^20, e.g. 2^15 = 128 x 256
 * for N = 2^30).
 */
#include "hs_engine.h"
#include "../../include/hsfft_b200.h"

#ifndef HS_DIST_BULK_DEFAULT
#define HS_DIST_BULK_DEFAULT 0
#endif

typedef struct hs_axis { /* transform of length L along one axis, as one or two tiled passes */
    int L, two, sub1, sub2, n1, n2, nb2;
    void *d_tw1, *d_tw2, *d_twb, *d_tws;
} hs_axis;

struct hsfft_dist_set {
    int N1, N2, nranks, rank, sgn;
    long long Cw, Kw;
    int logN, logS, logCw;
    hs_axis ax1, ax2;
    void *d_hi, *d_lo;
    void *d_store;          /* hs_dist_store in device memory once the peers' receive buffers are known */
    void *s1, *s2, *ev_fork, *ev_join1, *ev_join2, *ev_chunk[16]; /* two internal streams: pass 1 of chunk i+1 runs under pass 2 of chunk i */
    void *mapped[HS_DIST_MAX_RANKS]; /* IPC mappings this plan opened (closed by hsfft_dist_free) */
};

static int ilog2_exact(long long v)
{
    int l = 0;
    while ((1LL << l) < v) ++l;
    return ((1LL << l) == v) ? l : -1;
}

static const hs_ct_desc *find_family(int family, int n)
{
    const int cnt = hs_cu_ct_count();
    for (int i = 0; i < cnt; ++i) {
        const hs_ct_desc *d = hs_cu_ct_get(i);
        if (d->family == family && d->n == n) return d;
    }
    return NULL;
}

static void *upload_tw(const int *rx, int nst, int sgn, int split_min)
{
    const int cnt = hs_stage_twiddle_count(rx, nst, split_min);
    fft_data *t = (fft_data *)malloc(sizeof(fft_data) * (cnt > 0 ? cnt : 1));
    if (!t) { fprintf(stderr, "Error: Memory allocation failed for twiddle table\n"); exit(EXIT_FAILURE); }
    hs_build_stage_twiddles(t, rx, nst, sgn, split_min);
    void *d = hs_cu_malloc(sizeof(fft_data) * (cnt > 0 ? cnt : 1));
    hs_cu_h2d(d, t, sizeof(fft_data) * cnt, NULL);
    free(t);
    return d;
}

static void put_tw(fft_data *dst, double e, double theta, int sgn)
{
    const double a = e * theta;
    dst->re = cos(a);
    dst->im = (sgn == -1) ? -sin(a) : sin(a);
}

static void axis_init(hs_axis *ax, int L, int sgn)
{
    memset(ax, 0, sizeof(*ax));
    ax->L = L;
    const hs_ct_desc *s = find_family(HS_CT_SUB, L);
    if (s) {
        ax->sub1 = s->id;
        ax->d_tw1 = upload_tw(s->radix, s->nst, sgn, s->split_min);
        return;
    }
    const hs_ct_desc *t = find_family(HS_CT_TWOPASS, L);
    if (!t) {
        fprintf(stderr, "Error: distributed transform needs axis lengths in {128,256,512,1024} or 2^14..2^20 (got %d)\n", L);
        exit(EXIT_FAILURE);
    }
    const hs_ct_desc *s1 = find_family(HS_CT_SUB, t->n1), *s2 = find_family(HS_CT_SUB, t->n2);
    if (!s1 || !s2) {
        fprintf(stderr, "Error: distributed transform: no tiled sub-plan for the factors %d x %d of axis length %d\n", t->n1, t->n2, L);
        exit(EXIT_FAILURE);
    }
    ax->two = 1;
    ax->n1 = t->n1; ax->n2 = t->n2;
    ax->sub1 = s1->id; ax->sub2 = s2->id;
    ax->d_tw1 = upload_tw(s1->radix, s1->nst, sgn, s1->split_min);
    ax->d_tw2 = upload_tw(s2->radix, s2->nst, sgn, s2->split_min);
    ax->nb2 = t->n2 / s2->radix[0];
    fft_data *tb = (fft_data *)malloc(sizeof(fft_data) * (size_t)ax->nb2 * t->n1), *ts = (fft_data *)malloc(sizeof(fft_data) * t->n1);
    if (!tb || !ts) { fprintf(stderr, "Error: Memory allocation failed for twiddle table\n"); exit(EXIT_FAILURE); }
    const double theta = -PI2 / (double)L;
    for (int b = 0; b < ax->nb2; ++b)
        for (int k = 0; k < t->n1; ++k) put_tw(&tb[(size_t)b * t->n1 + k], (double)(k * b), theta, sgn);
    for (int k = 0; k < t->n1; ++k) put_tw(&ts[k], (double)(k * ax->nb2), theta, sgn);
    ax->d_twb = hs_cu_malloc(sizeof(fft_data) * (size_t)ax->nb2 * t->n1);
    ax->d_tws = hs_cu_malloc(sizeof(fft_data) * t->n1);
    hs_cu_h2d(ax->d_twb, tb, sizeof(fft_data) * (size_t)ax->nb2 * t->n1, NULL);
    hs_cu_h2d(ax->d_tws, ts, sizeof(fft_data) * t->n1, NULL);
    hs_cu_stream_sync(NULL);
    free(tb); free(ts);
}

static void axis_free(hs_axis *ax)
{
    hs_cu_free(ax->d_tw1); hs_cu_free(ax->d_tw2); hs_cu_free(ax->d_twb); hs_cu_free(ax->d_tws);
}

hsfft_dist_object hsfft_dist_init(int N1, int N2, int nranks, int rank, int sgn)
{
    hs_cu_require_device();
    const int l1 = ilog2_exact(N1), l2 = ilog2_exact(N2), lg = ilog2_exact(nranks);
    if (l1 < 0 || l2 < 0 || lg < 0 || rank < 0 || rank >= nranks || N1 % nranks || N2 % nranks || (sgn != 1 && sgn != -1)) {
        fprintf(stderr, "Error: Invalid inputs for hsfft_dist_init - N1: %d, N2: %d, ranks: %d, rank: %d, sgn: %d\n", N1, N2, nranks, rank, sgn);
        exit(EXIT_FAILURE);
    }
    hsfft_dist_object d = (hsfft_dist_object)calloc(1, sizeof(*d));
    if (!d) { fprintf(stderr, "Error: Memory allocation failed for distributed plan\n"); exit(EXIT_FAILURE); }
    d->N1 = N1; d->N2 = N2; d->nranks = nranks; d->rank = rank; d->sgn = sgn;
    d->Cw = N2 / nranks; d->Kw = N1 / nranks;
    d->logN = l1 + l2; d->logS = (d->logN + 1) / 2; d->logCw = ilog2_exact(d->Cw);
    if (d->Cw < 16 || d->Kw < 16) { fprintf(stderr, "Error: distributed transform needs at least 16 columns and rows per rank\n"); exit(EXIT_FAILURE); }
    axis_init(&d->ax1, N1, sgn);
    axis_init(&d->ax2, N2, sgn);
    if (d->ax2.two && d->ax2.n2 > d->Cw) {
        fprintf(stderr, "Error: N2 = %d over %d ranks leaves slabs narrower than its second factor %d\n", N2, nranks, d->ax2.n2);
        exit(EXIT_FAILURE);
    }
    /* exchange twiddle W_N^e, e = n2 k1 mod N, as hi[e >> logS] * lo[e mod 2^logS] */
    const long long S = 1LL << d->logS, H = 1LL << (d->logN - d->logS);
    fft_data *hi = (fft_data *)malloc(sizeof(fft_data) * H), *lo = (fft_data *)malloc(sizeof(fft_data) * S);
    if (!hi || !lo) { fprintf(stderr, "Error: Memory allocation failed for twiddle table\n"); exit(EXIT_FAILURE); }
    const double theta = -PI2 / ldexp(1.0, d->logN);
    for (long long e = 0; e < S; ++e) put_tw(&lo[e], (double)e, theta, sgn);
    for (long long e = 0; e < H; ++e) put_tw(&hi[e], ldexp((double)e, d->logS), theta, sgn);
    d->d_hi = hs_cu_malloc(sizeof(fft_data) * H);
    d->d_lo = hs_cu_malloc(sizeof(fft_data) * S);
    hs_cu_h2d(d->d_hi, hi, sizeof(fft_data) * H, NULL);
    hs_cu_h2d(d->d_lo, lo, sizeof(fft_data) * S, NULL);
    hs_cu_stream_sync(NULL);
    free(hi); free(lo);
    return d;
}

void hsfft_dist_free(hsfft_dist_object d)
{
    if (!d) return;
    axis_free(&d->ax1); axis_free(&d->ax2);
    hs_cu_free(d->d_hi); hs_cu_free(d->d_lo); hs_cu_free(d->d_store);
    for (int g = 0; g < HS_DIST_MAX_RANKS; ++g) hs_cu_ipc_close(d->mapped[g]);
    if (d->s1) {
        hs_cu_stream_destroy(d->s1); hs_cu_stream_destroy(d->s2);
        hs_cu_event_destroy(d->ev_fork); hs_cu_event_destroy(d->ev_join1); hs_cu_event_destroy(d->ev_join2);
        for (int i = 0; i < 16; ++i) hs_cu_event_destroy(d->ev_chunk[i]);
    }
    free(d);
}

/* ---- exchange fused into phase A: every rank's receive buffer is mapped into every other rank (CUDA IPC over
 * NVLink peer access) and phase A's last pass stores each twiddled row straight into its owner's buffer ---- */

void hsfft_ipc_get_handle(void *d_buf, void *handle64) { hs_cu_require_device(); hs_cu_ipc_get(d_buf, handle64); }

void hsfft_dist_set_peers(hsfft_dist_object d, void *d_own_r, const void *handles64, int nhandles)
{
    /* handles64: nranks handles of 64 bytes, ordered by rank (the own slot is ignored: d_own_r is used).  Buffers that are
     * already addressable from this process go through hsfft_dist_set_peer_pointers instead. */
    if (!d || !d_own_r || !handles64 || nhandles != d->nranks || d->nranks > HS_DIST_MAX_RANKS) {
        fprintf(stderr, "Error: Invalid inputs for hsfft_dist_set_peers\n");
        exit(EXIT_FAILURE);
    }
    hs_dist_store st;
    memset(&st, 0, sizeof st);
    for (int g = 0; g < d->nranks; ++g) {
        if (g == d->rank) { st.peer[g] = d_own_r; continue; }
        hs_cu_ipc_close(d->mapped[g]);
        d->mapped[g] = hs_cu_ipc_open((const char *)handles64 + 64 * (size_t)g);
        st.peer[g] = d->mapped[g];
    }
    st.hi = d->d_hi; st.lo = d->d_lo;
    st.Cw = d->Cw; st.Kw = d->Kw;
    st.row_mul = d->ax1.two ? d->ax1.n1 : 1;
    st.logKw = ilog2_exact(d->Kw); st.logN = d->logN; st.logS = d->logS; st.rank = d->rank;
    if (!d->d_store) d->d_store = hs_cu_malloc(sizeof st);
    hs_cu_h2d(d->d_store, &st, sizeof st, NULL);
    hs_cu_stream_sync(NULL);
}

void hsfft_dist_set_peer_pointers(hsfft_dist_object d, void *const *peer_r, int npeers)
{
    /* same, for peers whose buffers are already addressable from this process (one process driving several plans) */
    if (!d || !peer_r || npeers != d->nranks || d->nranks > HS_DIST_MAX_RANKS) {
        fprintf(stderr, "Error: Invalid inputs for hsfft_dist_set_peer_pointers\n");
        exit(EXIT_FAILURE);
    }
    hs_dist_store st;
    memset(&st, 0, sizeof st);
    for (int g = 0; g < d->nranks; ++g) st.peer[g] = peer_r[g];
    st.hi = d->d_hi; st.lo = d->d_lo;
    st.Cw = d->Cw; st.Kw = d->Kw;
    st.row_mul = d->ax1.two ? d->ax1.n1 : 1;
    st.logKw = ilog2_exact(d->Kw); st.logN = d->logN; st.logS = d->logS; st.rank = d->rank;
    if (!d->d_store) d->d_store = hs_cu_malloc(sizeof st);
    hs_cu_h2d(d->d_store, &st, sizeof st, NULL);
    hs_cu_stream_sync(NULL);
}

/* how phase A's last pass sends its rows: 7 = 16-byte stores from the registers, 11 = rows staged in shared memory and handed
 * to the copy engine (cp.async.bulk shared -> peer global); HSFFT_DIST_BULK=0 / 1 */
static int dist_store_mode(void)
{
    static int mode = -1;
    if (mode < 0) { const char *e = getenv("HSFFT_DIST_BULK"); mode = (e ? atoi(e) : HS_DIST_BULK_DEFAULT) ? 11 : 7; }
    return mode;
}

void hsfft_dist_cols_exchange_dev(hsfft_dist_object d, const fft_data *d_a, fft_data *d_ws, void *stream)
{
    /* phase A and the exchange in one: b is never materialised, its rows go to their owners as they are produced.
     * The caller synchronises the ranks afterwards (all stores have landed when every rank's kernel has ended). */
    if (!d || !d_a || !d_ws || !d->d_store) { fprintf(stderr, "Error: Invalid distributed plan, data pointers or peers not set\n"); exit(EXIT_FAILURE); }
    const hs_axis *ax = &d->ax1;
    const long long W = d->Cw;
    hs_tile_geom g;
    memset(&g, 0, sizeof g);
    g.ncols = (int)W;
    g.in_c = 1; g.out_c = 1;
    g.chirp = d->d_store;
    if (!ax->two) {
        g.in_e = W; g.out_e = W;
        hs_cu_launch_tiled_io(ax->sub1, d->sgn, 0, 1, dist_store_mode(), 1, d_a, NULL, ax->d_tw1, &g, 1, stream);
        return;
    }
    /* Two passes.  Only the second one talks to the peers; HSFFT_DIST_CHUNKS=k cuts the slab into k column chunks and
     * runs pass 1 of chunk i+1 (on a second stream) under pass 2 of chunk i so that the NVLink stores start earlier.
     * Measured at 2^30 on 8 GPUs: 4.65 ms with 1 chunk, 4.68 with 2, 4.73 with 4 -- pass 2's CTAs, stalled on their
     * peer stores, hold the SMs and the next chunk's pass 1 does not get in; so the default is one chunk. */
    const long long L1 = ax->n1, L2 = ax->n2;
    int chunks = 1;
    if (getenv("HSFFT_DIST_CHUNKS")) chunks = atoi(getenv("HSFFT_DIST_CHUNKS"));
    if (chunks > 16) chunks = 16;
    while (chunks > 1 && (W % chunks != 0 || (W / chunks) % 64 != 0)) chunks /= 2;
    if (chunks < 1) chunks = 1;
    const long long Wc = W / chunks;
    if (chunks > 1 && !d->s1) {
        d->s1 = hs_cu_stream_create(); d->s2 = hs_cu_stream_create();
        d->ev_fork = hs_cu_event_create(); d->ev_join1 = hs_cu_event_create(); d->ev_join2 = hs_cu_event_create();
        for (int i = 0; i < 16; ++i) d->ev_chunk[i] = hs_cu_event_create();
    }
    void *sa = chunks > 1 ? d->s1 : stream, *sb = chunks > 1 ? d->s2 : stream;
    if (chunks > 1) {
        hs_cu_event_record(d->ev_fork, stream);
        hs_cu_stream_wait_event(sa, d->ev_fork);
        hs_cu_stream_wait_event(sb, d->ev_fork);
    }
    for (int i = 0; i < chunks; ++i) {
        const fft_data *src = d_a + i * Wc;
        fft_data *mid = d_ws + i * Wc;
        memset(&g, 0, sizeof g);
        g.ncols = (int)Wc; g.in_c = 1; g.out_c = 1;
        g.in_xf = W; g.in_e = W * L2; g.out_xf = L1 * W; g.out_e = W;
        hs_cu_launch_tiled(ax->sub1, d->sgn, 0, 1, src, mid, ax->d_tw1, &g, L2, sa);
        if (chunks > 1) { hs_cu_event_record(d->ev_chunk[i], sa); hs_cu_stream_wait_event(sb, d->ev_chunk[i]); }
        g.in_xf = W; g.in_e = L1 * W; g.out_xf = W; g.out_e = L1 * W;
        g.twb = ax->d_twb; g.tws = ax->d_tws; g.tw_ld = (int)L1;
        g.chirp = d->d_store; g.nlog = (int)(i * Wc);
        hs_cu_launch_tiled_io(ax->sub2, d->sgn, 2, 1, dist_store_mode(), 1, mid, NULL, ax->d_tw2, &g, L1, sb);
    }
    if (chunks > 1) {
        hs_cu_event_record(d->ev_join1, sa); hs_cu_event_record(d->ev_join2, sb);
        hs_cu_stream_wait_event(stream, d->ev_join1);
        hs_cu_stream_wait_event(stream, d->ev_join2);
    }
}

long long hsfft_dist_slab_elems(hsfft_dist_object d) { return (long long)d->N1 * d->Cw; }

/* transform of length ax->L along the rows of src[L][W] (column stride 1), result dst[L][W]; ws: L*W */
static void axis_cols(const hs_axis *ax, int sgn, const void *src, void *dst, void *ws, long long W, void *stream)
{
    hs_tile_geom g;
    memset(&g, 0, sizeof g);
    g.ncols = (int)W;
    g.in_c = 1; g.out_c = 1;
    if (!ax->two) {
        g.in_e = W; g.out_e = W;
        hs_cu_launch_tiled(ax->sub1, sgn, 0, 1, src, dst, ax->d_tw1, &g, 1, stream);
        return;
    }
    const long long L1 = ax->n1, L2 = ax->n2;
    g.in_xf = W; g.in_e = W * L2; g.out_xf = L1 * W; g.out_e = W;
    hs_cu_launch_tiled(ax->sub1, sgn, 0, 1, src, ws, ax->d_tw1, &g, L2, stream);
    g.in_xf = W; g.in_e = L1 * W; g.out_xf = W; g.out_e = L1 * W;
    g.twb = ax->d_twb; g.tws = ax->d_tws; g.tw_ld = (int)L1;
    hs_cu_launch_tiled(ax->sub2, sgn, 2, 1, ws, dst, ax->d_tw2, &g, L1, stream);
}

void hsfft_dist_cols_dev(hsfft_dist_object d, const fft_data *d_a, fft_data *d_b, fft_data *d_ws, void *stream)
{
    if (!d || !d_a || !d_b || !d_ws) { fprintf(stderr, "Error: Invalid distributed plan or data pointers\n"); exit(EXIT_FAILURE); }
    axis_cols(&d->ax1, d->sgn, d_a, d_b, d_ws, d->Cw, stream);
    hs_cu_launch_dist_twiddle(d_b, d->d_hi, d->d_lo, d->N1, d->Cw, (long long)d->rank * d->Cw, d->logN, d->logS, stream);
}

void hsfft_dist_rows_dev(hsfft_dist_object d, const fft_data *d_r, fft_data *d_out, fft_data *d_ws, void *stream)
{
    if (!d || !d_r || !d_out || !d_ws) { fprintf(stderr, "Error: Invalid distributed plan or data pointers\n"); exit(EXIT_FAILURE); }
    const hs_axis *ax = &d->ax2;
    const long long Cw = d->Cw, Kw = d->Kw, N2 = d->N2;
    hs_tile_geom g;
    memset(&g, 0, sizeof g);
    g.blk_log = d->logCw; g.blk_stride = Kw * Cw; /* element n2 of row r: r*Cw + (n2 / Cw) * Kw*Cw + n2 % Cw */
    if (!ax->two) {
        /* lanes = rows (strided loads; only the small test sizes come here), contiguous stores */
        g.ncols = (int)Kw; g.in_c = Cw; g.in_e = 1; g.out_c = N2; g.out_e = 1;
        hs_cu_launch_tiled(ax->sub1, d->sgn, 0, 0, d_r, d_out, ax->d_tw1, &g, 1, stream);
        return;
    }
    const long long L1 = ax->n1, L2 = ax->n2;
    /* pass 1: per row r, columns q < L2: FFT_L1 over i of row[q + L2 i]  ->  ws[r][q*L1 + j] */
    g.ncols = (int)L2; g.in_xf = Cw; g.in_c = 1; g.in_e = L2; g.out_xf = N2; g.out_c = L1; g.out_e = 1;
    hs_cu_launch_tiled(ax->sub1, d->sgn, 0, 0, d_r, d_ws, ax->d_tw1, &g, Kw, stream);
    /* pass 2: per row r, columns k < L1: FFT_L2 over i of W^(i k) ws[r][i*L1 + k]  ->  out[r][k + L1 j] */
    memset(&g, 0, sizeof g);
    g.ncols = (int)L1; g.in_xf = N2; g.in_c = 1; g.in_e = L1; g.out_xf = N2; g.out_c = 1; g.out_e = L1;
    g.twb = ax->d_twb; g.tws = ax->d_tws; g.tw_ld = (int)L1;
    hs_cu_launch_tiled(ax->sub2, d->sgn, 1, 1, d_ws, d_out, ax->d_tw2, &g, Kw, stream);
}

/* ---- the whole four-step from ONE process driving `ndev` GPUs (devices 0 .. ndev-1): host pointers, natural order in
 * and out.  The self-contained C entry for a single huge transform (north_star config 5) -- no launcher, no handle
 * exchange: the devices map each other's receive buffers through peer access, phase A's transform kernel stores every
 * twiddled row straight into its owner over NVLink, one synchronisation of the devices is the barrier, phase B runs
 * the rows.  inp[N1*N2] -> oup[N1*N2], unnormalised, sgn as in fft_init.  Returns 0, or -1 when the devices cannot
 * reach each other's memory.  (Many processes: hsfft_dist_init .. hsfft_dist_rows_dev above, with the caller's transport.) */
int fft_exec_dist(int N1, int N2, int sgn, const fft_data *inp, fft_data *oup, int ndev)
{
    hs_cu_require_device();
    if (!inp || !oup || ndev < 1 || ndev > HS_DIST_MAX_RANKS || ndev > hs_cu_device_count()) {
        fprintf(stderr, "Error: Invalid inputs for fft_exec_dist - devices: %d (of %d)\n", ndev, hs_cu_device_count());
        exit(EXIT_FAILURE);
    }
    for (int g = 0; g < ndev; ++g)
        for (int p = 0; p < ndev; ++p)
            if (g != p && !hs_cu_can_access_peer(g, p)) {
                fprintf(stderr, "Error: fft_exec_dist: device %d cannot access device %d's memory\n", g, p);
                return -1;
            }
    const int home = hs_cu_get_device();
    hsfft_dist_object plan[HS_DIST_MAX_RANKS];
    void *a[HS_DIST_MAX_RANKS], *r[HS_DIST_MAX_RANKS], *ws[HS_DIST_MAX_RANKS], *o[HS_DIST_MAX_RANKS], *st[HS_DIST_MAX_RANKS];
    const long long Cw = N2 / ndev, Kw = N1 / ndev;
    const size_t slab = sizeof(fft_data) * (size_t)N1 * (size_t)Cw;
    for (int g = 0; g < ndev; ++g) {
        hs_cu_set_device(g);
        for (int p = 0; p < ndev; ++p)
            if (p != g) hs_cu_enable_peer_access(p);
        plan[g] = hsfft_dist_init(N1, N2, ndev, g, sgn);
        a[g] = hs_cu_malloc(slab); r[g] = hs_cu_malloc(slab); ws[g] = hs_cu_malloc(slab); o[g] = hs_cu_malloc(slab);
        st[g] = hs_cu_stream_create();
    }
    for (int g = 0; g < ndev; ++g) {
        hs_cu_set_device(g);
        hsfft_dist_set_peer_pointers(plan[g], r, ndev);
        /* column slab a[n1][c] = x[N2 n1 + g Cw + c] */
        hs_cu_h2d_2d(a[g], sizeof(fft_data) * (size_t)Cw, inp + (size_t)g * Cw, sizeof(fft_data) * (size_t)N2, sizeof(fft_data) * (size_t)Cw,
                     (size_t)N1, st[g]);
        hsfft_dist_cols_exchange_dev(plan[g], (const fft_data *)a[g], (fft_data *)ws[g], st[g]);
    }
    for (int g = 0; g < ndev; ++g) { hs_cu_set_device(g); hs_cu_stream_sync(st[g]); } /* every rank's peer stores have landed */
    for (int g = 0; g < ndev; ++g) {
        hs_cu_set_device(g);
        hsfft_dist_rows_dev(plan[g], (const fft_data *)r[g], (fft_data *)o[g], (fft_data *)ws[g], st[g]);
        /* out[k1l][k2] = X[(g Kw + k1l) + N1 k2]  ->  [k2][k1l], then rows of Kw bins into the natural-order array */
        hs_cu_launch_transpose(o[g], a[g], (int)Kw, N2, 1, st[g]);
        hs_cu_d2h_2d(oup + (size_t)g * Kw, sizeof(fft_data) * (size_t)N1, a[g], sizeof(fft_data) * (size_t)Kw, sizeof(fft_data) * (size_t)Kw,
                     (size_t)N2, st[g]);
    }
    for (int g = 0; g < ndev; ++g) {
        hs_cu_set_device(g);
        hs_cu_stream_sync(st[g]);
        hsfft_dist_free(plan[g]);
        hs_cu_free(a[g]); hs_cu_free(r[g]); hs_cu_free(ws[g]); hs_cu_free(o[g]);
        hs_cu_stream_destroy(st[g]);
    }
    hs_cu_set_device(home);
    return 0;
}
