// hs_generic.cuh -- run-time planned, strided, tiled shared-memory FFT pass.
//
// This is the "every N works" engine behind the drop-in API: any factor list the reference's
// factors() can emit (reference src/highSpeedFFT.c:2038-2163), including the O(r^2) generic odd
// radices (:1475-1628), and -- through the stride / inter-pass-twiddle parameters -- one level of a
// multi-pass decomposition of a transform too long for shared memory.  Hot sizes do not come here;
// they use the compile-time plans of hs_smem_fft.cuh.
//
// One pass computes, for every (batch, o, c):
//     out[bt*out_b + o*out_o + c*out_c + j*out_e] =
//         FFT_n{ in[bt*in_b + o*in_o + c*in_c + e*in_e] * W_L^{e*c} }[j]            (W_L only if tw_L > 0)
// A CTA owns a tile of C consecutive c (so that global accesses are C*16-byte contiguous runs when
// the c stride is 1) and keeps the C transforms in two ping-pong shared-memory buffers.
#pragma once
#include "hs_butterflies.cuh"

#include "hs_abi.h"
typedef hs_gen_params HsGenParams; // pointers travel as void*, see hs_abi.h

template <int R, int SGN>
HS_DEVICE void hs_gen_stage(const HsGenParams &p, int s, const cplx *src, cplx *dst, int m)
{
    const int n = p.n, qp = n / (m * R), nb = n / R;
    const int total = p.C * nb;
    for (int idx = HS_TID; idx < total; idx += HS_NTHREADS) {
        const int b = idx % nb, c = idx / nb;
        const int q = b % qp, k = b / qp;
        const cplx *sc = src + c * p.npad;
        cplx v[R];
        HS_UNROLL
        for (int i = 0; i < R; ++i) v[i] = sc[q + qp * i + qp * R * k];
        if (s > 0) {
            const cplx *w = (const cplx *)p.tw + p.tw_off[s] + k;
            HS_UNROLL
            for (int i = 1; i < R; ++i) v[i] = hs_mul(v[i], hs_ld_tab(w + (i - 1) * m));
        }
        hs_bfly<R, SGN>(v);
        cplx *dc = dst + c * p.npad;
        HS_UNROLL
        for (int j = 0; j < R; ++j) dc[b + nb * j] = v[j];
    }
}

template <int SGN>
HS_DEVICE void hs_gen_stage_odd(const HsGenParams &p, int s, const cplx *src, cplx *dst, int m, int r)
{
    const int n = p.n, qp = n / (m * r), nb = n / r;
    const int total = p.C * nb;
    const double *cs = p.gen_cs + p.gen_off[s], *sn = p.gen_sn + p.gen_off[s];
    for (int idx = HS_TID; idx < total; idx += HS_NTHREADS) {
        const int b = idx % nb, c = idx / nb;
        const int q = b % qp, k = b / qp;
        const cplx *sc = src + c * p.npad;
        cplx v[HS_MAX_GENERIC_RADIX];
        for (int i = 0; i < r; ++i) v[i] = sc[q + qp * i + qp * r * k];
        if (s > 0) {
            const cplx *w = (const cplx *)p.tw + p.tw_off[s] + k;
            for (int i = 1; i < r; ++i) v[i] = hs_mul(v[i], hs_ld_tab(w + (i - 1) * m));
        }
        hs_bfly_generic<SGN>(v, r, cs, sn);
        cplx *dc = dst + c * p.npad;
        for (int j = 0; j < r; ++j) dc[b + nb * j] = v[j];
    }
}

template <int SGN>
HS_GLOBAL void HS_LAUNCH_BOUNDS(256, 1) hs_generic_pass_kernel(const cplx *HS_RESTRICT in, cplx *HS_RESTRICT out, const HsGenParams p)
{
    HS_DYN_SMEM(cplx, smem);
    cplx *bufA = smem, *bufB = smem + (size_t)p.C * p.npad;
    const long long ntile = (p.n_c + p.C - 1) / p.C;
    const long long bid = HS_BID;
    const long long tile = bid % ntile, o = (bid / ntile) % p.n_o, bt = bid / (ntile * p.n_o);
    const long long c0 = tile * p.C;
    const int n = p.n;
    const cplx *gin = in + bt * p.in_b + o * p.in_o;
    cplx *gout = out + bt * p.out_b + o * p.out_o;

    for (int idx = HS_TID; idx < p.C * n; idx += HS_NTHREADS) {
        int c, e;
        if (p.load_c_fast) { c = idx % p.C; e = idx / p.C; }
        else               { e = idx % n;   c = idx / n; }
        if (c0 + c < p.n_c) {
            cplx val = hs_ld_stream(gin + (c0 + c) * p.in_c + (long long)e * p.in_e);
            if (p.tw_L > 0) {
                const long long x = (long long)e * (c0 + c); // < tw_L by construction
                const cplx w = hs_mul(hs_ld_tab((const cplx *)p.tw_lo + (x % p.tw_S)), hs_ld_tab((const cplx *)p.tw_hi + (x / p.tw_S)));
                val = hs_mul(val, w);
            }
            bufA[c * p.npad + e] = val;
        } else {
            bufA[c * p.npad + e] = hs_c(0.0, 0.0);
        }
    }
    HS_SYNC();

    int m = 1;
    for (int s = 0; s < p.nst; ++s) {
        const int r = p.radix[s];
        switch (r) {
        case 2: hs_gen_stage<2, SGN>(p, s, bufA, bufB, m); break;
        case 3: hs_gen_stage<3, SGN>(p, s, bufA, bufB, m); break;
        case 4: hs_gen_stage<4, SGN>(p, s, bufA, bufB, m); break;
        case 5: hs_gen_stage<5, SGN>(p, s, bufA, bufB, m); break;
        case 7: hs_gen_stage<7, SGN>(p, s, bufA, bufB, m); break;
        case 8: hs_gen_stage<8, SGN>(p, s, bufA, bufB, m); break;
        case 16: hs_gen_stage<16, SGN>(p, s, bufA, bufB, m); break;
        default: hs_gen_stage_odd<SGN>(p, s, bufA, bufB, m, r); break;
        }
        m *= r;
        HS_SYNC();
        cplx *t = bufA; bufA = bufB; bufB = t;
    }

    for (int idx = HS_TID; idx < p.C * n; idx += HS_NTHREADS) {
        int c, k;
        if (p.store_c_fast) { c = idx % p.C; k = idx / p.C; }
        else                { k = idx % n;   c = idx / n; }
        if (c0 + c < p.n_c) hs_st_stream(gout + (c0 + c) * p.out_c + (long long)k * p.out_e, bufA[c * p.npad + k]);
    }
}

// ------------------------------------------------------------------ element-wise helper kernels
// (unfused forms; the hot sizes use the fused kernels in hs_fused.cuh)

// Bluestein step 1 (reference src/highSpeedFFT.c:1803-1827): a[n] = x[n]*conj(h[n]) (sgn=+1) or
// x[n]*h[n] (sgn=-1) for n < N, zero for N <= n < M.
template <int SGN>
HS_GLOBAL void hs_chirp_pad_kernel(const cplx *HS_RESTRICT x, cplx *HS_RESTRICT a, const cplx *HS_RESTRICT h, int N, int M,
                                   long long howmany)
{
    const long long total = howmany * M;
    for (long long idx = HS_BID * HS_NTHREADS + HS_TID; idx < total; idx += HS_NBLOCKS * HS_NTHREADS) {
        const long long b = idx / M;
        const int n = (int)(idx % M);
        cplx v = hs_c(0.0, 0.0);
        if (n < N) {
            const cplx xv = hs_ld_stream(x + b * N + n), hv = hs_ld_tab(h + n);
            v = (SGN == 1) ? hs_mulc(xv, hv) : hs_mul(xv, hv);
        }
        a[idx] = v;
    }
}

// Bluestein step 3 (:1838-1855): A[k] *= Hk[k] (sgn=+1) or conj(Hk[k]) (sgn=-1), in place.
template <int SGN>
HS_GLOBAL void hs_spectrum_mul_kernel(cplx *HS_RESTRICT A, const cplx *HS_RESTRICT Hk, int M, long long howmany)
{
    const long long total = howmany * M;
    for (long long idx = HS_BID * HS_NTHREADS + HS_TID; idx < total; idx += HS_NBLOCKS * HS_NTHREADS) {
        const cplx hv = hs_ld_tab(Hk + (idx % M));
        const cplx av = A[idx];
        A[idx] = (SGN == 1) ? hs_mul(av, hv) : hs_mulc(av, hv);
    }
}

// Bluestein step 5 (:1871-1886): out[k] = y[k]*conj(h[k]) (sgn=+1) or y[k]*h[k] (sgn=-1), k < N.
template <int SGN>
HS_GLOBAL void hs_chirp_crop_kernel(const cplx *HS_RESTRICT y, cplx *HS_RESTRICT out, const cplx *HS_RESTRICT h, int N, int M,
                                    long long howmany)
{
    const long long total = howmany * N;
    for (long long idx = HS_BID * HS_NTHREADS + HS_TID; idx < total; idx += HS_NBLOCKS * HS_NTHREADS) {
        const long long b = idx / N;
        const int k = (int)(idx % N);
        const cplx yv = y[b * M + k], hv = hs_ld_tab(h + k);
        hs_st_stream(out + idx, (SGN == 1) ? hs_mulc(yv, hv) : hs_mul(yv, hv));
    }
}

// r2c split (reference src/real.c:109-132).  Z: half-length spectrum (H = N/2 per row);
// out row stride out_stride, bins 0..H always written, mirror bins H+1..N-1 when full != 0.
// One thread per pair (k, H-k), k = 0..H/2: bins k and H-k need the same two inputs Z[k] and Z[H-k], so every
// input is read once (the first version read it twice: 32 N bytes of traffic for 24 N algorithmic).
HS_GLOBAL void hs_r2c_post_kernel(const cplx *HS_RESTRICT Z, cplx *HS_RESTRICT out, const cplx *HS_RESTRICT tw2, int H, long long out_stride,
                                  int full, long long howmany)
{
    const int P = H / 2 + 1; // pairs per row
    const long long total = howmany * P;
    for (long long idx = HS_BID * HS_NTHREADS + HS_TID; idx < total; idx += HS_NBLOCKS * HS_NTHREADS) {
        const long long b = idx / P;
        const int k = (int)(idx % P);
        const cplx *z = Z + b * H;
        cplx *o = out + b * out_stride;
        if (k == 0) {
            const cplx z0 = hs_ld_stream(z);
            hs_st_stream(o, hs_c(z0.x + z0.y, 0.0));
            hs_st_stream(o + H, hs_c(z0.x - z0.y, 0.0));
        } else {
            const cplx a = hs_ld_stream(z + k), c = (2 * k == H) ? a : hs_ld_stream(z + (H - k));
            const cplx r = hs_r2c_bin(a, c, hs_ld_tab(tw2 + k));
            hs_st_stream(o + k, r);
            if (full) hs_st_stream(o + (2 * H - k), hs_c(r.x, -r.y));
            if (2 * k != H) {
                const cplx r2 = hs_r2c_bin(c, a, hs_ld_tab(tw2 + (H - k)));
                hs_st_stream(o + (H - k), r2);
                if (full) hs_st_stream(o + (H + k), hs_c(r2.x, -r2.y));
            }
        }
    }
}

// c2r merge (reference src/real.c:171-179).  in: rows of bins 0..H (row stride in_stride); z: H per row.
HS_GLOBAL void hs_c2r_pre_kernel(const cplx *HS_RESTRICT in, cplx *HS_RESTRICT z, const cplx *HS_RESTRICT tw2, int H, long long in_stride,
                                 long long howmany)
{
    const long long total = howmany * H;
    for (long long idx = HS_BID * HS_NTHREADS + HS_TID; idx < total; idx += HS_NBLOCKS * HS_NTHREADS) {
        const long long b = idx / H;
        const int k = (int)(idx % H);
        const cplx *x = in + b * in_stride;
        const cplx a = x[k], c = x[H - k], w = hs_ld_tab(tw2 + k);
        const double t1 = -a.y - c.y, t2 = -c.x + a.x;
        z[idx] = hs_c(a.x + c.x + (t1 * w.x) - (t2 * w.y), a.y - c.y + (t2 * w.x) + (t1 * w.y));
    }
}

// convolution product (reference src/convolve.c:147-160): C[k] = A[k]*B[k] * scale, bins 0..H
HS_GLOBAL void hs_conv_mul_kernel(const cplx *A, const cplx *HS_RESTRICT B, cplx *Cc, long long total, double scale) // Cc may be A (in-place product)
{
    for (long long idx = HS_BID * HS_NTHREADS + HS_TID; idx < total; idx += HS_NBLOCKS * HS_NTHREADS) {
        const cplx p = hs_mul(A[idx], B[idx]);
        Cc[idx] = hs_c(p.x * scale, p.y * scale);
    }
}

// convolution staging (reference src/convolve.c:117-140 zero padding, :163-200 slicing)
HS_GLOBAL void hs_pad_rows_kernel(const double *HS_RESTRICT src, int len, double *HS_RESTRICT dst, int P, long long howmany)
{
    const long long total = howmany * P;
    for (long long idx = HS_BID * HS_NTHREADS + HS_TID; idx < total; idx += HS_NBLOCKS * HS_NTHREADS) {
        const long long b = idx / P;
        const int i = (int)(idx % P);
        dst[idx] = (i < len) ? hs_ld_stream_d(src + b * len + i) : 0.0;
    }
}
HS_GLOBAL void hs_slice_rows_kernel(const double *HS_RESTRICT src, int P, int start, int len, double *HS_RESTRICT dst, long long dst_stride,
                                    long long howmany)
{
    const long long total = howmany * len;
    for (long long idx = HS_BID * HS_NTHREADS + HS_TID; idx < total; idx += HS_NBLOCKS * HS_NTHREADS) {
        const long long b = idx / len;
        const int i = (int)(idx % len);
        hs_st_stream_d(dst + b * dst_stride + i, src[b * P + start + i]);
    }
}

// batched out-of-place transpose of rows x cols complex matrices (2-D transforms whose column length has no strided
// tiled plan): 32 x 32 tiles through shared memory, rows of 33 slots so that the transposed reads hit different banks
HS_GLOBAL void hs_transpose_kernel(const cplx *HS_RESTRICT in, cplx *HS_RESTRICT out, int rows, int cols, long long howmany)
{
    HS_DYN_SMEM(cplx, tile);
    const int tx = HS_TID % 32, ty = HS_TID / 32, ny = HS_NTHREADS / 32;
    const long long tiles_c = (cols + 31) / 32, tiles_r = (rows + 31) / 32, per = tiles_c * tiles_r;
    for (long long b = HS_BID; b < howmany * per; b += HS_NBLOCKS) {
        const long long img = b / per, t = b % per;
        const int r0 = (int)(t / tiles_c) * 32, c0 = (int)(t % tiles_c) * 32;
        const cplx *src = in + img * rows * cols;
        cplx *dst = out + img * rows * cols;
        for (int j = ty; j < 32; j += ny)
            if (r0 + j < rows && c0 + tx < cols) tile[j * 33 + tx] = hs_ld_stream(src + (long long)(r0 + j) * cols + c0 + tx);
        HS_SYNC();
        for (int j = ty; j < 32; j += ny)
            if (c0 + j < cols && r0 + tx < rows) hs_st_stream(dst + (long long)(c0 + j) * rows + r0 + tx, tile[tx * 33 + j]);
        HS_SYNC();
    }
}
