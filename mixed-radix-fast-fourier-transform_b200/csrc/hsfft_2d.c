/*
 * hsfft_2d.c -- two-dimensional complex transforms on top of the batched 1-D engine (host side, plain C).
 *
 * Additive: the reference has 1-D transforms only.  X[k0][k1] = sum over j0, j1 of
 * x[j0][j1] exp(-sgn 2 pi i (j0 k0 / n0 + j1 k1 / n1)), row-major n0 x n1, unnormalised like fft_exec.
 *
 *   rows     every row is a contiguous 1-D transform of length n1: one batched call of the 1-D plan (static, plan-time
 *            compiled or Bluestein -- whatever fft_init(n1) gives);
 *   columns  n0 = 128, 256, 512 or 1024 with n1 a multiple of 16 / 16 / 8 / 4: the strided tiled pass (hs_tiled.cuh) transforms the
 *            columns in place of a transpose, 16-byte x C contiguous runs, one launch for the whole batch;
 *            any other n0: transpose, batched 1-D transforms of length n0, transpose back.
 * Butterflies, literals and twiddles are those of the 1-D paths, so every row and column transform is within rounding
 * of what the reference's fft_exec computes for it.
 */
#include "hs_engine.h"
#include "../../include/hsfft_b200.h"

struct fft2_set {
    int n0, n1, sgn;
    fft_object row, col;   /* 1-D plans of length n1 and (transpose path) n0 */
    int col_sub;           /* >= 0: compile-time sub-plan of length n0 run by the strided tiled pass */
    void *d_col_tw;
    void *d_tmp, *d_tmp2;  /* device scratch, grown on demand */
    size_t cap, cap2;
};

fft2_object fft2_init(int n0, int n1, int sgn)
{
    if (n0 < 2 || n1 < 2 || (sgn != 1 && sgn != -1)) {
        fprintf(stderr, "Error: Invalid inputs for fft2_init - n0: %d, n1: %d, sgn: %d\n", n0, n1, sgn);
        exit(EXIT_FAILURE);
    }
    fft2_object p = (fft2_object)calloc(1, sizeof(*p));
    if (!p) { fprintf(stderr, "Error: Memory allocation failed for 2-D plan\n"); exit(EXIT_FAILURE); }
    p->n0 = n0; p->n1 = n1; p->sgn = sgn;
    p->col_sub = -1;
    p->row = fft_init(n1, sgn);
    if (!p->row) { free(p); return NULL; }
    return p;
}

static void plan_columns(fft2_object p)
{
    if (p->col_sub >= 0 || p->col) return;
    hs_cu_require_device();
    const int need = p->n0 == 1024 ? 4 : (p->n0 == 512 ? 8 : 16); /* columns per tile of the sub-plan (HsTileCols, hs_tiled.cuh) */
    if (p->n1 % need == 0 && !getenv("HSFFT_2D_TRANSPOSE")) {
        const int cnt = hs_cu_ct_count();
        /* the column pass stores strided runs like a second pass: a sub-plan built for that role first (HS_CT_SUB2) */
        for (int fam = getenv("HSFFT_NO_SUB2") ? 1 : 0; fam < 2; ++fam)
        for (int i = 0; i < cnt; ++i) {
            const hs_ct_desc *d = hs_cu_ct_get(i);
            if (d->family == (fam == 0 ? HS_CT_SUB2 : HS_CT_SUB) && d->n == p->n0 && !d->exact) {
                const int nt = hs_stage_twiddle_count(d->radix, d->nst, d->split_min);
                fft_data *t = (fft_data *)malloc(sizeof(fft_data) * (nt > 0 ? nt : 1));
                if (!t) { fprintf(stderr, "Error: Memory allocation failed for twiddle table\n"); exit(EXIT_FAILURE); }
                hs_build_stage_twiddles(t, d->radix, d->nst, p->sgn, d->split_min);
                p->d_col_tw = hs_cu_malloc(sizeof(fft_data) * (nt > 0 ? nt : 1));
                hs_cu_h2d(p->d_col_tw, t, sizeof(fft_data) * nt, NULL);
                hs_cu_stream_sync(NULL);
                free(t);
                p->col_sub = d->id;
                return;
            }
        }
    }
    p->col = fft_init(p->n0, p->sgn);
    if (!p->col) { fprintf(stderr, "Error: Memory allocation failed for 2-D plan\n"); exit(EXIT_FAILURE); }
}

static void grow2(void **buf, size_t *cap, size_t need)
{
    if (need <= *cap) return;
    hs_cu_free(*buf);
    *buf = hs_cu_malloc(need);
    *cap = need;
}

void fft2_exec_batch_dev(fft2_object p, const fft_data *d_in, fft_data *d_out, long howmany, void *stream)
{
    if (!p || !d_in || !d_out) { fprintf(stderr, "Error: Invalid 2-D FFT object or data pointers\n"); exit(EXIT_FAILURE); }
    if (howmany <= 0) return;
    plan_columns(p);
    const size_t img = (size_t)p->n0 * p->n1, bytes = sizeof(fft_data) * img * (size_t)howmany;
    grow2(&p->d_tmp, &p->cap, bytes);
    fft_exec_batch_dev(p->row, d_in, (fft_data *)p->d_tmp, howmany * p->n0, stream);
    if (p->col_sub >= 0) {
        hs_tile_geom g;
        memset(&g, 0, sizeof g);
        g.ncols = p->n1; g.in_c = 1; g.out_c = 1; g.in_e = p->n1; g.out_e = p->n1;
        g.in_xf = (long long)img; g.out_xf = (long long)img;
        hs_cu_launch_tiled(p->col_sub, p->sgn, 0, 1, p->d_tmp, d_out, p->d_col_tw, &g, howmany, stream);
        return;
    }
    grow2(&p->d_tmp2, &p->cap2, bytes);
    hs_cu_launch_transpose(p->d_tmp, p->d_tmp2, p->n0, p->n1, howmany, stream);
    fft_exec_batch_dev(p->col, (const fft_data *)p->d_tmp2, (fft_data *)p->d_tmp, howmany * p->n1, stream);
    hs_cu_launch_transpose(p->d_tmp, d_out, p->n1, p->n0, howmany, stream);
}

void fft2_exec_batch(fft2_object p, fft_data *inp, fft_data *oup, long howmany)
{
    if (!p || !inp || !oup) { fprintf(stderr, "Error: Invalid 2-D FFT object or data pointers\n"); exit(EXIT_FAILURE); }
    if (howmany <= 0) return;
    hs_cu_require_device();
    const size_t bytes = sizeof(fft_data) * (size_t)p->n0 * p->n1 * (size_t)howmany;
    void *din = hs_cu_malloc(bytes), *dout = hs_cu_malloc(bytes);
    hs_cu_h2d(din, inp, bytes, NULL);
    fft2_exec_batch_dev(p, (const fft_data *)din, (fft_data *)dout, howmany, NULL);
    hs_cu_d2h(oup, dout, bytes, NULL);
    hs_cu_stream_sync(NULL);
    hs_cu_free(din); hs_cu_free(dout);
}

void fft2_exec(fft2_object p, fft_data *inp, fft_data *oup) { fft2_exec_batch(p, inp, oup, 1); }

const char *fft2_describe(fft2_object p)
{
    static char buf[256];
    if (!p) return "null plan";
    plan_columns(p);
    snprintf(buf, sizeof buf, "2-D %d x %d: rows as batched 1-D transforms, columns by %s", p->n0, p->n1,
             p->col_sub >= 0 ? "the strided tiled pass (no transpose)" : "transpose, batched 1-D transforms, transpose");
    return buf;
}

void free_fft2(fft2_object p)
{
    if (!p) return;
    free_fft(p->row);
    free_fft(p->col);
    hs_cu_free(p->d_col_tw); hs_cu_free(p->d_tmp); hs_cu_free(p->d_tmp2);
    free(p);
}
