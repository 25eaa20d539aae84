// hs_tiled.cuh -- one pass of a transform too long for one SM: C adjacent strided sub-transforms per CTA.
//
// A transform of length N = R1*R2 whose states do not fit shared memory is run as two of the
// reference's "super stages" (the recursion of reference
// src/highSpeedFFT.c:318 unrolled two levels at a time):
//   pass 1:  for every q < R2:   A[q][j] = FFT_R1{ x[q + R2 i] }            in stride R2, out contiguous
//   pass 2:  for every k < R1:   X[k + R1 j] = FFT_R2{ W_N^{i k} A[i][k] }  in stride R1, out stride R1
// Each sub-transform is a compile-time plan of hs_smem_fft.cuh run by a team of threads.  A CTA
// owns C adjacent columns (q or k), so every global access is a run of C*16 contiguous bytes:
//   * first pass of the plan: teams across lanes (lane = column) -> the strided loads coalesce;
//   * later passes: pass 1 switches to threads-of-a-team across lanes so that its contiguous output
//     rows coalesce; pass 2 keeps lane = column because its stores are strided like its loads.
// The inter-pass twiddle W_N^{i k} is applied in registers right after the load.  Element
// i = b + NB*ii of column k needs W_N^{k b} * (W_N^{k NB})^ii: one table load per butterfly plus one
// per thread, the powers are formed by multiplication (the L1 data pipe, not FP64, limits these
// kernels -- see DESIGN.md section 3).
#pragma once
#include "hs_smem_fft.cuh"

typedef hs_tile_geom HsTileGeom; // see hs_abi.h

// Distance (in 16-byte slots) between the buffers of adjacent columns of a tile.  In the passes that map
// lane -> column, a quarter-warp (one LDS.128 / STS.128 wavefront) holds min(C, 8) columns x 8/min(C, 8)
// consecutive threads of a team, i.e. slots  col*stride + t  (mod 8 they must all differ): the stride has to be
// odd for C >= 8, 2 mod 4 for C = 4, 4 mod 8 for C = 2.  (ncu on the first pipeline build, C = 4 with an odd
// stride: 44 M bank conflicts against 0.5 M, shared-memory wavefronts x 1.7.)
constexpr int hs_col_stride(int npad, int C)
{
    const int want = C >= 8 ? 1 : (C == 4 ? 2 : (C == 2 ? 4 : 0));
    int s = npad;
    while (C >= 8 ? (s % 2 == 0) : (want ? (s % (2 * want) != want) : false)) ++s;
    return s;
}

// columns per CTA: as many as shared memory and the register file allow (longer contiguous runs on the
// strided side), 128..512 threads per CTA
// 512-point sub-plans: 4 columns (two 256-thread CTAs per SM) for the one-CTA-per-tile kernel, 8 for the persistent TMA-staged
// one (measured after the bank-conflict fix, fraction of the HBM peak at 2^18 / 2^19: plain kernel 29.9 / 32.8 % with 8
// columns, 41.2 / 40.8 % with 4; TMA kernel 41.1 / 39.7 % with 8, 41.6 / 38.2 % with 4; c2r 2^20 28.7 -> 31.4 %)
// columns per tile of the plain tiled kernels (CTA = T * C threads).  Measured on B200 (profiles/r1b_tile_width_ab.txt):
// 8 instead of 16 columns for the 160- and 256-point plans (320 / 256 -> 160 / 128 threads, more CTAs per SM) takes
// Bluestein N = 10007 from 6.21 to 5.57 ms and 2^16..2^17 up by 1-3 %; the 128-point plan (T = 8) loses 4 % with 64-thread
// CTAs and keeps 16.
#ifndef HS_TILE_C_SMALL
#define HS_TILE_C_SMALL (P::T <= 8 ? 16 : 8)
#endif
// (2 columns for the 1024-point plan -- four 128-thread CTAs per SM -- measured slower: 2^20 31.3 vs 40.9 %; 32-byte rows)
#ifndef HS_TILE_C_BIG
#define HS_TILE_C_BIG 4
#endif
// passes whose column count need not be a multiple of the tile width: the real four-step's second pass (n1/2 + 1 rows);
// the plan-time (NVRTC) passes define it as 1 for every pass
#ifndef HS_TILE_PARTIAL
#define HS_TILE_PARTIAL(IO) ((IO) == 9)
#endif
template <class P> struct HsTileCols { static constexpr int C = P::T >= 64 ? HS_TILE_C_BIG : HS_TILE_C_SMALL; };
// the real four-step's mirrored second pass (IO 9) stores every run twice, the mirror image 16 bytes off the run's natural
// alignment: wider tiles pay there.  256 points: 16 instead of 8 columns takes that pass from 1.50 to 1.09 ms per 2^28 reals,
// while the plain second pass of a complex transform does not care (1.23 vs 1.24-1.26 ms)
#ifndef HS_TILE_C9_256
#define HS_TILE_C9_256 16
#endif
#ifndef HS_TILE_C9_512
#define HS_TILE_C9_512 8
#endif
#ifndef HS_TILE_C9_1024
#define HS_TILE_C9_1024 4
#endif
template <class P> struct HsTileColsMirror {
    static constexpr int C = (P::N == 256 && P::T == 16) ? HS_TILE_C9_256 : (P::N == 512 && P::T == 32) ? HS_TILE_C9_512 : (P::N == 1024 && P::T == 64) ? HS_TILE_C9_1024 : HsTileCols<P>::C;
};
template <class P, int C> struct HsTileColsTma { static constexpr int value = (P::N == 512 && C == 4) ? 8 : C; };

// LDP / STP: residency class of the loads / stores (hs_ld_mp / hs_st_mp in hs_common.cuh): 0 for the stand-alone
// passes, 1 / 2 inside the persistent pipeline kernels of hs_pipe.cuh.
template <int LDP = 0>
struct HsLoadCol {
    static constexpr bool kPost = false;
    const cplx *HS_RESTRICT base; // outer item's base
    long long off0, es;           // column offset, element stride
    int blk_log;
    long long blk_stride;
    HS_DEVICE long long addr(int n) const
    {
        const long long e = n * es;
        return off0 + (blk_log ? (e >> blk_log) * blk_stride + (e & ((1LL << blk_log) - 1)) : e);
    }
    HS_DEVICE cplx operator()(int, int, int n) const { return hs_ld_mp<LDP>(base + addr(n)); }
};
// inter-pass twiddle of element i = b + NB*ii of column k: v[ii] *= t[ii], t[ii] = W^(k b) * (W^(k NB))^ii.
// t[] costs R + 1 complex products for R = 16 (the first version took 30): the step's powers 2 and 4 by squaring,
// t[1], t[2], t[4] from w0, t[8] = t[4]*s[4], and every other t[ii] from the entry without its lowest set bit,
// t[ii] = t[ii - low] * s[low].  The longest chain is 5 products (ii = 15), the same depth as before.
constexpr int hs_pow2_floor(int x) { return x < 2 ? 1 : 2 * hs_pow2_floor(x / 2); }
template <int R>
HS_DEVICE void hs_interpass_twiddle(const cplx w0, const cplx step, cplx *v)
{
    constexpr int SMAX = hs_pow2_floor((R - 1) / 2); // largest power of the step that is needed (any radix, not only 2^k)
    cplx s[SMAX + 1], t[R];
    s[1] = step;
    HS_UNROLL
    for (int p = 2; p <= SMAX; p *= 2) s[p] = hs_mul(s[p / 2], s[p / 2]);
    t[0] = w0;
    HS_UNROLL
    for (int p = 1; p < R; p *= 2) t[p] = (p <= SMAX) ? hs_mul(w0, s[p]) : hs_mul(t[p / 2], s[p / 2]);
    HS_UNROLL
    for (int i = 3; i < R; ++i)
        if (i & (i - 1)) {
            const int low = i & -i;
            t[i] = hs_mul(t[i - low], s[low]);
        }
    HS_UNROLL
    for (int i = 0; i < R; ++i) v[i] = hs_mul(v[i], t[i]);
}
template <int LDP = 0>
struct HsLoadColTw : HsLoadCol<LDP> {
    static constexpr bool kPost = true;
    const cplx *HS_RESTRICT twb; // already offset by k
    int tw_ld;
    cplx step;
    template <int R>
    HS_DEVICE void post(int b, cplx *v) const
    {
        hs_interpass_twiddle<R>(hs_ld_tab(twb + (long long)b * tw_ld), step, v);
    }
};
template <int STP = 0>
struct HsStoreCol {
    cplx *HS_RESTRICT base;
    long long es;
    HS_DEVICE void operator()(int, int, int k, cplx v) const { hs_st_mp<STP>(base + k * es, v); }
};

// Bluestein ends fused into the tiled passes (reference src/highSpeedFFT.c:1803-1827 and :1871-1886):
// pass 1 of the forward transform loads x[n]*conj(h[n]) (or *h[n]) and the zero padding directly,
// the last pass of the inverse transform multiplies by the chirp and stores only the N logical outputs.
template <int CS, int LDP = 0>
struct HsLoadColChirp {
    static constexpr bool kPost = false;
    const cplx *HS_RESTRICT x; // this transform's N logical samples
    const cplx *HS_RESTRICT h;
    int N;
    long long off0, es;
    HS_DEVICE cplx operator()(int, int, int n_) const
    {
        const long long n = off0 + n_ * es;
        if (n >= N) return hs_c(0.0, 0.0);
        const cplx xv = hs_ld_mp<LDP>(x + n), hv = hs_ld_tab(h + n);
        return (CS == 1) ? hs_mulc(xv, hv) : hs_mul(xv, hv);
    }
};
template <int CS, int STP = 0>
struct HsStoreColChirpCrop {
    cplx *HS_RESTRICT out;
    const cplx *HS_RESTRICT h;
    int N;
    long long off0, es;
    HS_DEVICE void operator()(int, int, int k, cplx v) const
    {
        const long long n = off0 + k * es;
        if (n < N) {
            const cplx hv = hs_ld_tab(h + n);
            hs_st_mp<STP>(out + n, (CS == 1) ? hs_mulc(v, hv) : hs_mul(v, hv));
        }
    }
};

// Four-step exchange fused into the last pass of phase A (north_star config 5; see hs_dist_store in hs_abi.h): output
// element k of column c of outer item x is row k1 = x + row_mul k of the rank's slab b[k1][c]; it is multiplied by
// W_N^((rank Cw + c) k1) and stored into the receive buffer of rank k1 / Kw at r[rank][k1 mod Kw][c].  With lane =
// column a warp writes runs of C*16 contiguous bytes per row, which is what NVLink wants.
struct HsStoreColDist {
    const hs_dist_store *HS_RESTRICT ds;
    long long x; // outer item
    int c;       // column inside the slab
    HS_DEVICE void operator()(int, int, int k, cplx v) const
    {
        const unsigned long long k1 = (unsigned long long)(x + ds->row_mul * k);
        const unsigned long long n2 = (unsigned long long)(ds->rank * ds->Cw + c);
        const unsigned long long e = (k1 * n2) & ((1ULL << ds->logN) - 1);
        const cplx w = hs_mul(hs_ld_tab((const cplx *)ds->hi + (e >> ds->logS)), hs_ld_tab((const cplx *)ds->lo + (e & ((1ULL << ds->logS) - 1))));
        cplx *dst = (cplx *)ds->peer[k1 >> ds->logKw] + ((ds->rank * ds->Kw + (long long)(k1 & (unsigned long long)(ds->Kw - 1))) * ds->Cw + c);
        *dst = hs_mul(v, w); // plain store: peer memory (or the own buffer), visible system-wide when the kernel ends
    }
};

// The same exchange with the rows staged in shared memory: the last pass leaves its twiddled outputs as stage[k][c]
// (C adjacent columns of output row k = C*16 contiguous bytes of the owner's receive buffer), and after one barrier one
// thread per row hands the run to the copy engine (cp.async.bulk shared -> peer global): the SM's load/store unit issues
// one instruction per row instead of C stores, and the warps do not sit on NVLink back-pressure.
struct HsStoreColDistStage {
    static constexpr bool kToSmem = true;
    const hs_dist_store *HS_RESTRICT ds;
    long long x;  // outer item
    int c;        // column inside the slab
    cplx *stage;  // [k][C]
    int cs, C;    // column slot inside the tile
    HS_DEVICE void operator()(int, int, int k, cplx v) const
    {
        const unsigned long long k1 = (unsigned long long)(x + ds->row_mul * k);
        const unsigned long long n2 = (unsigned long long)(ds->rank * ds->Cw + c);
        const unsigned long long e = (k1 * n2) & ((1ULL << ds->logN) - 1);
        const cplx w = hs_mul(hs_ld_tab((const cplx *)ds->hi + (e >> ds->logS)), hs_ld_tab((const cplx *)ds->lo + (e & ((1ULL << ds->logS) - 1))));
        stage[k * C + cs] = hs_mul(v, w);
    }
};

// c2r merge (reference src/real.c:171-179) fused into pass 1 of the half-length transform: element k of the
// packed signal is built from bins k and H-k of the Hermitian input on the fly, so the merged signal is
// never written to HBM.
template <int LDP = 0>
struct HsLoadColC2RPre {
    static constexpr bool kPost = false;
    const cplx *HS_RESTRICT x;   // this transform's bins 0..H
    const cplx *HS_RESTRICT tw2; // exp(+2 pi i k / N)
    int H;
    long long off0, es;
    HS_DEVICE cplx operator()(int, int, int n_) const
    {
        const long long k = off0 + n_ * es;
        const cplx a = hs_ld_mp<LDP>(x + k), c = hs_ld_mp<LDP>(x + (H - k)), w = hs_ld_tab(tw2 + k);
        const double t1 = -a.y - c.y, t2 = -c.x + a.x;
        return hs_c(a.x + c.x + (t1 * w.x) - (t2 * w.y), a.y - c.y + (t2 * w.x) + (t1 * w.y));
    }
};
// the same with twiddle2[column + n2 row] formed as the product of a per-column factor (held in a register) and a per-row
// factor (n1 entries, L1 resident) instead of read from the H-entry table: at N = 2^20 that table is 8 MB, read once per
// transform -- a third of the pass's L2 traffic
template <int LDP = 0>
struct HsLoadColC2RPre2 {
    static constexpr bool kPost = false;
    const cplx *HS_RESTRICT x;  // this transform's bins 0..H
    const cplx *HS_RESTRICT wr; // exp(+2 pi i n2 row / N)
    cplx wc;                    // exp(+2 pi i column / N)
    int H;
    long long off0, es;
    HS_DEVICE cplx operator()(int, int, int n_) const
    {
        const long long k = off0 + n_ * es;
        const cplx a = hs_ld_mp<LDP>(x + k), c = hs_ld_mp<LDP>(x + (H - k)), w = hs_mul(wc, hs_ld_tab(wr + n_));
        const double t1 = -a.y - c.y, t2 = -c.x + a.x;
        return hs_c(a.x + c.x + (t1 * w.x) - (t2 * w.y), a.y - c.y + (t2 * w.x) + (t1 * w.y));
    }
};

// loaders for a tile the copy engine staged in shared memory, layout [row][C] (dense): with lane = column a
// quarter-warp reads adjacent 16-byte slots -- conflict free
struct HsLoadTile {
    static constexpr bool kPost = false;
    const cplx *tile; // staged [row][C], already offset by the column
    int C;
    HS_DEVICE cplx operator()(int, int, int n) const { return tile[n * C]; }
};
struct HsLoadTileTw : HsLoadTile {
    static constexpr bool kPost = true;
    const cplx *HS_RESTRICT twb;
    int tw_ld;
    cplx step;
    template <int R>
    HS_DEVICE void post(int b, cplx *v) const
    {
        hs_interpass_twiddle<R>(hs_ld_tab(twb + (long long)b * tw_ld), step, v);
    }
};

// Convolution fusions (reference src/convolve.c:117-200): the zero padding rides on the first pass's loads of the
// forward transforms, the spectrum product and the 1/P scale on the merge loads of the inverse one, the output
// slice on its last pass's stores -- the padded inputs, the product and the full-length result never exist in HBM.
struct HsLoadColPadReal { // a window of a long real array read as a row of complex pairs, zero outside the row's len reals and
    static constexpr bool kPost = false; // outside the array (8-byte loads: windows may be 8-byte aligned only)
    const double *HS_RESTRICT x;
    long long first, total; // the window's first real (may be negative), reals in the array
    int len;
    long long off0, es;
    HS_DEVICE double at(long long r) const
    {
        const long long gidx = first + r;
        return (r < len && gidx >= 0 && gidx < total) ? hs_ld_stream_d(x + gidx) : 0.0;
    }
    HS_DEVICE cplx operator()(int, int, int n_) const
    {
        const long long r = 2 * (off0 + n_ * es);
        return hs_c(at(r), at(r + 1));
    }
};
struct HsLoadColC2RMul { // element k of the packed signal from bins k and H-k of the product of two spectra
    static constexpr bool kPost = false;
    const cplx *HS_RESTRICT xa, *HS_RESTRICT xb; // bins 0..H of the two spectra
    const cplx *HS_RESTRICT tw2;                 // twiddle2, or (wr != nullptr) unused: twiddle2[k] = wc * wr[row]
    const cplx *HS_RESTRICT wr;
    cplx wc;
    int H;
    double scale;
    long long off0, es;
    HS_DEVICE cplx operator()(int, int, int n_) const
    {
        const long long k = off0 + n_ * es;
        const cplx pa = hs_mul(hs_ld_stream(xa + k), hs_ld_stream(xb + k)), pc = hs_mul(hs_ld_stream(xa + (H - k)), hs_ld_stream(xb + (H - k)));
        const cplx a = hs_c(pa.x * scale, pa.y * scale), c = hs_c(pc.x * scale, pc.y * scale);
        const cplx w = wr ? hs_mul(wc, hs_ld_tab(wr + n_)) : hs_ld_tab(tw2 + k); // reference :147-160
        const double t1 = -a.y - c.y, t2 = -c.x + a.x;
        return hs_c(a.x + c.x + (t1 * w.x) - (t2 * w.y), a.y - c.y + (t2 * w.x) + (t1 * w.y));
    }
};
struct HsStoreColSlice { // complex element k of the row = reals 2k, 2k+1; keep [start, start+len), clipped to the output array
    double *HS_RESTRICT y;
    long long first, total; // output index of the row's first kept real, reals in the output array
    int start, len;
    long long off0, es;
    HS_DEVICE void put(long long r, double v) const
    {
        if (r >= 0 && r < len && first + r < total) hs_st_stream_d(y + first + r, v);
    }
    HS_DEVICE void operator()(int, int, int k_, cplx v) const
    {
        const long long r = 2 * (off0 + k_ * es) - start;
        put(r, v.x);
        put(r + 1, v.y);
    }
};

// Real four-step (r2c of N = n1 n2 reals, reference src/real.c:78-136 computes the same bins through a half-length complex
// transform and a split pass).  The packed signal z[m] = x[2m] + i x[2m+1] is an n1 x n2/2 complex matrix whose complex
// column q' carries the REAL columns 2q' and 2q'+1 of the n1 x n2 real matrix.  Pass 1 transforms the complex columns and
// separates the two real columns' spectra where both partners Z[k1], Z[n1-k1] sit in the same column buffer:
//     A_2q'[k1] = (Z[k1] + conj Z[n1-k1]) / 2,     A_2q'+1[k1] = (Z[k1] - conj Z[n1-k1]) / 2i,      k1 = 0 .. n1/2
// (Hermitian: the other rows are their conjugates and are never stored).  Pass 2 runs the n1/2+1 rows:
//     X[k1 + n1 k2] = sum_q W_N^(q k1) A_q[k1] W_n2^(q k2)
// and writes every bin together with its conjugate mirror X[N-k].  The split costs one shared-memory round trip in the
// pass that has HBM bandwidth to spare; no separate split kernel, no half-length spectrum in HBM.
template <class P>
struct HsStoreSmemCol { // last pass of a column transform: leave the spectrum in the column's buffer, natural order
    static constexpr bool kToSmem = true;
    cplx *sm;
    HS_DEVICE void operator()(int, int, int k, cplx v) const { sm[P::pad(k)] = v; }
};
template <int STP = 0>
struct HsStoreColR2CMirror { // element k2 of row k1: bin k = k1 + n1 k2 and its mirror N - k
    cplx *HS_RESTRICT out; // this transform's output row
    int k1, n1, full;
    long long H;
    HS_DEVICE void operator()(int, int, int k2, cplx v) const
    {
        const long long k = k1 + (long long)n1 * k2;
        const bool selfrow = (k1 == 0) || (2 * k1 == n1); // rows whose mirror bins lie in the row itself
        if (k <= H) {
            hs_st_mp<STP>(out + k, v);
            if (full && k != 0 && k != H) hs_st_mp<STP>(out + (2 * H - k), hs_conj(v));
        } else if (!selfrow) {
            hs_st_mp<STP>(out + (2 * H - k), hs_conj(v));
            if (full) hs_st_mp<STP>(out + k, v);
        }
    }
};

// One tile: C adjacent columns c0..c0+C-1 of one outer item, whose input / output start at in_x / out_x.
// P: plan of the sub-transform (P::X is ignored); C: columns per CTA; TW: inter-pass twiddle on load
// (0 none, 1 keyed by the column, 2 keyed by the outer item xf); STCOL: stores keep lane = column
// instead of switching to lane = thread of a team (contiguous output rows).
// IO: 0 plain; 1 chirp-and-pad load (CS = chirp direction); 2 chirp-and-crop store; 3 c2r merge on load;
// 4 zero padding on load, 5 spectrum product on merge load, 6 slice on store (convolution);
// 7 four-step exchange on store (g.chirp = the plan's hs_dist_store in device memory);
// 8 / 9 the two passes of the real four-step (g.out_c = two workspace rows; g.nlog = n1, g.sl_start = full spectrum);
// 10 = 8 with the zero padding of 4 on its loads; 11 = 7 with the rows staged in shared memory and sent as bulk copies.
// tid: thread index in the tile's team of P::T*C threads; on: false for threads (or whole CTAs) that only
// take part in the barriers.
// STG: the input tile was staged in shared memory by the copy engine ([row][C] at in_x); IO 0 and 2 only.
template <class P, int SGN, int C, int TW, bool STCOL, int IO, int CS, int LDP, int STP, bool STG = false, class HK>
HS_DEVICE void hs_tile_run(const cplx *HS_RESTRICT in_x, cplx *HS_RESTRICT out_x, const cplx *HS_RESTRICT tw, const HsTileGeom &g,
                           const long long xf, const int c0, cplx *smem, const int tid, const bool on, HK &hk)
{
    constexpr int NPADC = hs_col_stride(P::NPAD, C);
    const int colA = tid % C, tA = tid / C;       // lane = column
    const int colB = tid / P::T, tB = tid % P::T; // lane = thread of a team
    // a partial last tile (column count not a multiple of C): columns beyond the end only take part in the barriers
    // (only where the column count need not be a multiple of C: a run-time `active` costs the 512-point passes 8-14 %)
    constexpr bool PART = HS_TILE_PARTIAL(IO);
    const HsLane first{smem + colA * NPADC, tA, PART ? (on && c0 + colA < g.ncols) : on};
    const HsLane rest = STCOL ? first : HsLane{smem + colB * NPADC, tB, PART ? (on && c0 + colB < g.ncols) : on};
    const int colS = STCOL ? colA : colB;
    static_assert(!STG || IO == 0 || IO == 2, "only plain tile loads can be staged");
    if constexpr (IO == 1) { // forward pass 1 of the long Bluestein
        HsStoreCol<STP> st{out_x + (long long)(c0 + colS) * g.out_c, g.out_e};
        HsLoadColChirp<CS, LDP> ld{in_x, (const cplx *)g.chirp, g.nlog, (long long)(c0 + colA) * g.in_c, g.in_e};
        hs_fft_team2<P, SGN>(ld, st, first, rest, tw, hk);
    } else if constexpr (IO == 3) { // c2r: merge fused into pass 1 (g.chirp = twiddle2, g.nlog = H)
        HsStoreCol<STP> st{out_x + (long long)(c0 + colS) * g.out_c, g.out_e};
        if (g.twb) { // column / row factors of twiddle2 (in_c == 1: the column index is the offset)
            HsLoadColC2RPre2<LDP> ld{in_x, (const cplx *)g.tws, hs_ld_tab((const cplx *)g.twb + (c0 + colA)), g.nlog, (long long)(c0 + colA) * g.in_c, g.in_e};
            hs_fft_team2<P, SGN>(ld, st, first, rest, tw, hk);
        } else {
            HsLoadColC2RPre<LDP> ld{in_x, (const cplx *)g.chirp, g.nlog, (long long)(c0 + colA) * g.in_c, g.in_e};
            hs_fft_team2<P, SGN>(ld, st, first, rest, tw, hk);
        }
    } else if constexpr (IO == 4) { // convolution: zero padding on load (windows of g.nlog reals of one long array)
        HsStoreCol<STP> st{out_x + (long long)(c0 + colS) * g.out_c, g.out_e};
        HsLoadColPadReal ld{(const double *)in_x, xf * g.pad_stride + g.pad_off, g.pad_total, g.nlog, (long long)(c0 + colA) * g.in_c, g.in_e};
        hs_fft_team2<P, SGN>(ld, st, first, rest, tw, hk);
    } else if constexpr (IO == 5) { // convolution: product of two spectra, scaled, merged (g.chirp = twiddle2, g.nlog = H)
        HsStoreCol<STP> st{out_x + (long long)(c0 + colS) * g.out_c, g.out_e};
        HsLoadColC2RMul ld{in_x, (const cplx *)g.aux + xf * g.aux_xf, (const cplx *)g.chirp, (const cplx *)g.tws,
                           g.twb ? hs_ld_tab((const cplx *)g.twb + (c0 + colA)) : hs_c(1.0, 0.0), g.nlog, g.scale, (long long)(c0 + colA) * g.in_c, g.in_e};
        hs_fft_team2<P, SGN>(ld, st, first, rest, tw, hk);
    } else if constexpr (IO == 6) { // convolution: slice on store (rows of g.aux_xf reals)
        static_assert(TW == 1 && STCOL, "the slice store belongs to a second pass");
        HsStoreColSlice st{(double *)out_x, xf * g.aux_xf, g.sl_total, g.sl_start, g.sl_len, (long long)(c0 + colS) * g.out_c, g.out_e};
        HsLoadColTw<LDP> ld;
        ld.base = in_x; ld.off0 = (long long)(c0 + colA) * g.in_c; ld.es = g.in_e;
        ld.blk_log = g.blk_log; ld.blk_stride = g.blk_stride;
        ld.twb = (const cplx *)g.twb + (c0 + colA); ld.tw_ld = g.tw_ld; ld.step = first.active ? hs_ld_tab((const cplx *)g.tws + (c0 + colA)) : hs_c(0.0, 0.0);
        hs_fft_team2<P, SGN>(ld, st, first, rest, tw, hk);
    } else if constexpr (IO == 7) { // last pass of phase A of the distributed transform: twiddle and store to the owning rank
        static_assert(STCOL && TW != 1, "the exchange store keeps lane = column");
        HsStoreColDist st{(const hs_dist_store *)g.chirp, xf, g.nlog + c0 + colS}; // g.nlog: first slab column of this launch's chunk
        if constexpr (TW == 2) {
            HsLoadColTw<LDP> ld;
            ld.base = in_x; ld.off0 = (long long)(c0 + colA) * g.in_c; ld.es = g.in_e;
            ld.blk_log = g.blk_log; ld.blk_stride = g.blk_stride;
            ld.twb = (const cplx *)g.twb + xf; ld.tw_ld = g.tw_ld; ld.step = first.active ? hs_ld_tab((const cplx *)g.tws + xf) : hs_c(0.0, 0.0);
            hs_fft_team2<P, SGN>(ld, st, first, rest, tw, hk);
        } else {
            HsLoadCol<LDP> ld;
            ld.base = in_x; ld.off0 = (long long)(c0 + colA) * g.in_c; ld.es = g.in_e;
            ld.blk_log = g.blk_log; ld.blk_stride = g.blk_stride;
            hs_fft_team2<P, SGN>(ld, st, first, rest, tw, hk);
        }
    } else if constexpr (IO == 11) { // as 7, rows staged in shared memory and sent with bulk copies
        static_assert(STCOL && TW != 1 && !STG, "the exchange store keeps lane = column");
        const hs_dist_store *ds = (const hs_dist_store *)g.chirp;
        HsStoreColDistStage st{ds, xf, g.nlog + c0 + colS, smem, colS, C};
        if constexpr (TW == 2) {
            HsLoadColTw<LDP> ld;
            ld.base = in_x; ld.off0 = (long long)(c0 + colA) * g.in_c; ld.es = g.in_e;
            ld.blk_log = g.blk_log; ld.blk_stride = g.blk_stride;
            ld.twb = (const cplx *)g.twb + xf; ld.tw_ld = g.tw_ld; ld.step = first.active ? hs_ld_tab((const cplx *)g.tws + xf) : hs_c(0.0, 0.0);
            hs_fft_team2<P, SGN>(ld, st, first, rest, tw, hk);
        } else {
            HsLoadCol<LDP> ld;
            ld.base = in_x; ld.off0 = (long long)(c0 + colA) * g.in_c; ld.es = g.in_e;
            ld.blk_log = g.blk_log; ld.blk_stride = g.blk_stride;
            hs_fft_team2<P, SGN>(ld, st, first, rest, tw, hk);
        }
        hs_fence_async_proxy(); // this thread's staged outputs become visible to the copy engine ...
        hk.sync();              // ... and after the barrier so are everybody's
        for (int k = tid; k < P::N; k += P::T * C) {
            const unsigned long long k1 = (unsigned long long)(xf + ds->row_mul * k);
            cplx *dst = (cplx *)ds->peer[k1 >> ds->logKw] +
                        ((ds->rank * ds->Kw + (long long)(k1 & (unsigned long long)(ds->Kw - 1))) * ds->Cw + (g.nlog + c0));
            hs_bulk_s2g(dst, smem + k * C, (unsigned)(C * sizeof(cplx)));
        }
        hs_bulk_commit_wait_read();
    } else if constexpr (IO == 8 || IO == 10) { // real four-step, pass 1: transform the complex columns, split into the two real columns' rows 0..n1/2
        static_assert(TW == 0 && !STCOL && !STG, "the split belongs to a first pass with contiguous output rows");
        HsStoreSmemCol<P> st{rest.sm};
        if constexpr (IO == 10) { // convolution: the rows are zero-padded windows of a real array (as IO 4)
            HsLoadColPadReal ld{(const double *)in_x, xf * g.pad_stride + g.pad_off, g.pad_total, g.nlog, (long long)(c0 + colA) * g.in_c, g.in_e};
            hs_fft_team2<P, SGN>(ld, st, first, rest, tw, hk);
        } else {
            HsLoadCol<LDP> ld;
            ld.base = in_x; ld.off0 = (long long)(c0 + colA) * g.in_c; ld.es = g.in_e;
            ld.blk_log = g.blk_log; ld.blk_stride = g.blk_stride;
            hs_fft_team2<P, SGN>(ld, st, first, rest, tw, hk);
        }
        hk.sync();
        if (rest.active) {
            const cplx *z = rest.sm;
            cplx *oe = out_x + (long long)(c0 + colB) * g.out_c, *oo = oe + g.out_c / 2; // g.out_c = two rows of the workspace
            for (int k1 = tB; k1 <= P::N / 2; k1 += P::T) {
                const cplx a = z[P::pad(k1)], c = z[P::pad(k1 ? P::N - k1 : 0)];
                hs_st_mp<STP>(oe + k1, hs_c((a.x + c.x) * 0.5, (a.y - c.y) * 0.5));
                hs_st_mp<STP>(oo + k1, hs_c((a.y + c.y) * 0.5, (c.x - a.x) * 0.5));
            }
        }
    } else if constexpr (IO == 9) { // real four-step, pass 2: rows k1 = 0..n1/2, every bin stored with its mirror
        static_assert(TW == 1 && STCOL && !STG, "the mirrored store belongs to a second pass");
        HsStoreColR2CMirror<STP> st{out_x, c0 + colS, g.nlog, g.sl_start, (long long)g.nlog * P::N / 2};
        HsLoadColTw<LDP> ld;
        ld.base = in_x; ld.off0 = (long long)(c0 + colA) * g.in_c; ld.es = g.in_e;
        ld.blk_log = g.blk_log; ld.blk_stride = g.blk_stride;
        ld.twb = (const cplx *)g.twb + (c0 + colA); ld.tw_ld = g.tw_ld; ld.step = first.active ? hs_ld_tab((const cplx *)g.tws + (c0 + colA)) : hs_c(0.0, 0.0);
        hs_fft_team2<P, SGN>(ld, st, first, rest, tw, hk);
    } else if constexpr (IO == 2) { // inverse last pass of the long Bluestein
        static_assert(TW != 0, "the chirp-and-crop store belongs to a second pass");
        HsStoreColChirpCrop<CS, STP> st{out_x, (const cplx *)g.chirp, g.nlog, (long long)(c0 + colS) * g.out_c, g.out_e};
        const long long k = (TW == 1) ? (long long)(c0 + colA) : xf;
        if constexpr (STG) {
            HsLoadTileTw ld;
            ld.tile = in_x + colA; ld.C = C;
            ld.twb = (const cplx *)g.twb + k; ld.tw_ld = g.tw_ld; ld.step = first.active ? hs_ld_tab((const cplx *)g.tws + k) : hs_c(0.0, 0.0);
            hs_fft_team2<P, SGN>(ld, st, first, rest, tw, hk);
        } else {
            HsLoadColTw<LDP> ld;
            ld.base = in_x; ld.off0 = (long long)(c0 + colA) * g.in_c; ld.es = g.in_e;
            ld.blk_log = g.blk_log; ld.blk_stride = g.blk_stride;
            ld.twb = (const cplx *)g.twb + k; ld.tw_ld = g.tw_ld; ld.step = first.active ? hs_ld_tab((const cplx *)g.tws + k) : hs_c(0.0, 0.0);
            hs_fft_team2<P, SGN>(ld, st, first, rest, tw, hk);
        }
    } else if constexpr (TW != 0) {
        HsStoreCol<STP> st{out_x + (long long)(c0 + colS) * g.out_c, g.out_e};
        const long long k = (TW == 1) ? (long long)(c0 + colA) : xf;
        if constexpr (STG) {
            HsLoadTileTw ld;
            ld.tile = in_x + colA; ld.C = C;
            ld.twb = (const cplx *)g.twb + k; ld.tw_ld = g.tw_ld; ld.step = first.active ? hs_ld_tab((const cplx *)g.tws + k) : hs_c(0.0, 0.0);
            hs_fft_team2<P, SGN>(ld, st, first, rest, tw, hk);
        } else {
            HsLoadColTw<LDP> ld;
            ld.base = in_x; ld.off0 = (long long)(c0 + colA) * g.in_c; ld.es = g.in_e;
            ld.blk_log = g.blk_log; ld.blk_stride = g.blk_stride;
            ld.twb = (const cplx *)g.twb + k; ld.tw_ld = g.tw_ld; ld.step = first.active ? hs_ld_tab((const cplx *)g.tws + k) : hs_c(0.0, 0.0);
            hs_fft_team2<P, SGN>(ld, st, first, rest, tw, hk);
        }
    } else {
        HsStoreCol<STP> st{out_x + (long long)(c0 + colS) * g.out_c, g.out_e};
        if constexpr (STG) {
            HsLoadTile ld;
            ld.tile = in_x + colA; ld.C = C;
            hs_fft_team2<P, SGN>(ld, st, first, rest, tw, hk);
        } else {
            HsLoadCol<LDP> ld;
            ld.base = in_x; ld.off0 = (long long)(c0 + colA) * g.in_c; ld.es = g.in_e;
            ld.blk_log = g.blk_log; ld.blk_stride = g.blk_stride;
            hs_fft_team2<P, SGN>(ld, st, first, rest, tw, hk);
        }
    }
}

// Stand-alone pass: one CTA per tile.
template <class P, int SGN, int C, int TW, bool STCOL, int MINB, int IO = 0, int CS = 1>
HS_GLOBAL void HS_LAUNCH_BOUNDS(P::T *C, MINB)
hs_tiled_kernel(const cplx *HS_RESTRICT in, cplx *HS_RESTRICT out, const cplx *HS_RESTRICT tw, const HsTileGeom g, long long howmany)
{
    HS_DYN_SMEM(cplx, smem);
    const int tiles = (g.ncols + C - 1) / C; // the last tile may be partial (hs_tile_run masks its missing columns)
    const long long xf = HS_BID / tiles;
    const int c0 = (int)(HS_BID % tiles) * C;
    const bool on = true;
    HsNoHooks hk;
    hs_tile_run<P, SGN, C, TW, STCOL, IO, CS, 0, 0>(in + xf * g.in_xf, out + xf * g.out_xf, tw, g, xf, c0, smem, HS_TID, on, hk);
}

// ----------------------------------------------------------------------------------------------
// Persistent variant with TMA staging of the strided tile (north_star: "TMA staging for the strided
// passes").  One CTA per SM walks the tiles; while the teams transform tile j out of one staging
// buffer, the bulk-copy engine fills the other with tile j+1: R rows of C*16 contiguous bytes, one
// cp.async.bulk per row, completion counted by an mbarrier.  The load latency that the plain kernel
// exposes once per CTA (ncu: 25-30 % of its stall samples) disappears behind the previous tile's work.
// Staging layout [row][column] (dense): with lane = column, a quarter-warp reads 8 adjacent 16-byte
// slots -- conflict free.  Requires in_c == 1 and no remap; other geometries use hs_tiled_kernel.
template <class P, int C>
struct HsTmaTileSmem {
    static constexpr int NPADC = hs_col_stride(P::NPAD, C);
    static constexpr int tile_elems = P::N * C;
    static constexpr int bytes = (2 * tile_elems + C * NPADC) * 16 + 64;
};

// IO: 0 plain stores; 8 the real four-step's first pass (split of the two real columns' spectra after the transform, see IO 8 of hs_tile_run)
template <class P, int SGN, int C, int TW, bool STCOL, int IO = 0>
HS_GLOBAL void HS_LAUNCH_BOUNDS(P::T *C, 1)
hs_tiled_tma_kernel(const HS_GRID_CONSTANT HsTmap tmap, cplx *HS_RESTRICT out, const cplx *HS_RESTRICT tw, const HsTileGeom g, long long howmany)
{
    static_assert(IO == 0 || (IO == 8 && TW == 0 && !STCOL), "staged tiles: plain stores, or the real four-step's first pass");
    constexpr int BOXR = P::N < 256 ? P::N : 256; // rows per TMA box (hardware limit 256 per dimension)
    constexpr int NPADC = hs_col_stride(P::NPAD, C);
    constexpr int TE = P::N * C;             // elements of one staged tile
    constexpr unsigned ROWB = C * 16;        // bytes of one row
    HS_DYN_SMEM(cplx, smem);
    cplx *stage0 = smem, *stage1 = smem + TE, *exch = smem + 2 * TE; // TMA destinations: 128-byte aligned
    HsMbar *bar = reinterpret_cast<HsMbar *>(exch + C * NPADC); // two barriers, 16-byte aligned slot
    const int tiles = g.ncols / C;
    const long long total = howmany * tiles;
    const int colA = HS_TID % C, tA = HS_TID / C;
    const int colB = HS_TID / P::T, tB = HS_TID % P::T;
    const HsLane first{exch + colA * NPADC, tA, true};
    const HsLane rest = STCOL ? first : HsLane{exch + colB * NPADC, tB, true};
    const int colS = STCOL ? colA : colB;

    if (HS_TID == 0) { hs_mbar_init(&bar[0]); hs_mbar_init(&bar[1]); }
    hs_fence_async_proxy();
    HS_SYNC();

    auto issue = [&](long long tile_id, int buf) {
        if (HS_TID != 0) return; // one thread drives the copy engine
        const long long xf = tile_id / tiles;
        const int c0 = (int)(tile_id % tiles) * C;
        cplx *dst = buf ? stage1 : stage0;
        hs_mbar_expect(&bar[buf], (unsigned)(TE * 16));
        for (int r0 = 0; r0 < P::N; r0 += BOXR) hs_tma_load_3d(dst + r0 * C, &tmap, 2 * c0, r0, xf, BOXR, (int)ROWB, &bar[buf]);
    };

    long long j = 0;
    if (HS_BID < total) issue(HS_BID, 0);
    for (long long tile_id = HS_BID; tile_id < total; tile_id += HS_NBLOCKS, ++j) {
        const int buf = (int)(j & 1);
        const long long next = tile_id + HS_NBLOCKS;
        if (next < total) issue(next, buf ^ 1); // the other buffer was drained before the last barrier of tile j-1
        hs_mbar_wait(&bar[buf], (unsigned)((j >> 1) & 1));
        const long long xf = tile_id / tiles;
        const int c0 = (int)(tile_id % tiles) * C;
        const cplx *stage = buf ? stage1 : stage0;
        if constexpr (IO == 8) {
            HsStoreSmemCol<P> st{rest.sm};
            HsLoadTile ld;
            ld.tile = stage + colA; ld.C = C;
            hs_fft_team2<P, SGN>(ld, st, first, rest, tw);
            HS_SYNC();
            const cplx *z = rest.sm;
            cplx *oe = out + xf * g.out_xf + (long long)(c0 + colB) * g.out_c, *oo = oe + g.out_c / 2;
            for (int k1 = tB; k1 <= P::N / 2; k1 += P::T) {
                const cplx a = z[P::pad(k1)], c = z[P::pad(k1 ? P::N - k1 : 0)];
                hs_st_stream(oe + k1, hs_c((a.x + c.x) * 0.5, (a.y - c.y) * 0.5));
                hs_st_stream(oo + k1, hs_c((a.y + c.y) * 0.5, (c.x - a.x) * 0.5));
            }
        } else {
        HsStoreCol<0> st{out + xf * g.out_xf + (long long)(c0 + colS) * g.out_c, g.out_e};
        if constexpr (TW != 0) {
            const long long k = (TW == 1) ? (long long)(c0 + colA) : xf;
            HsLoadTileTw ld;
            ld.tile = stage + colA; ld.C = C;
            ld.twb = (const cplx *)g.twb + k; ld.tw_ld = g.tw_ld; ld.step = hs_ld_tab((const cplx *)g.tws + k);
            hs_fft_team2<P, SGN>(ld, st, first, rest, tw);
        } else {
            HsLoadTile ld;
            ld.tile = stage + colA; ld.C = C;
            hs_fft_team2<P, SGN>(ld, st, first, rest, tw);
        }
        }
        hs_fence_async_proxy(); // this tile's reads of the staging buffer precede the copy that refills it
        HS_SYNC();              // ... and the exchange buffer is free for the next tile
    }
}

// Middle kernel of the long Bluestein (padded length M = n1 * P::N): for column k < n1
//   Y[k + n1 j] = FFT_{P::N} over i of ( W_M^{i k} ws1[i n1 + k] )      forward pass 2
//   Y *= Hk (conj(Hk) for the inverse plan direction)                   reference :1838-1855
//   ws2[k][j'] = inverse FFT_{P::N} over j of Y[k + n1 j]               inverse pass 1 (decomposition n2 x n1)
// The spectrum Y never leaves the registers between the two transforms (same radix, same map).
template <int CS, int RL>
struct HsStoreRegsMulCol {
    cplx (*y)[RL];
    const cplx *HS_RESTRICT hk; // already offset by the column
    long long es;
    HS_DEVICE void operator()(int u, int j, int k, cplx v) const
    {
        const cplx hv = hs_ld_tab(hk + k * es);
        y[u][j] = (CS == 1) ? hs_mul(v, hv) : hs_mulc(v, hv);
    }
};
// hook adapters for two transforms back to back: only the first one's loads and the second one's last pass count
template <class HK>
struct HsFirstHook {
    HK &h;
    HS_DEVICE void sync() { h.sync(); }
    HS_DEVICE void after_first_loads() { h.after_first_loads(); }
    HS_DEVICE void after_first_pass() { h.after_first_pass(); }
    HS_DEVICE void before_last_pass() {}
};
template <class HK>
struct HsLastHook {
    HK &h;
    HS_DEVICE void sync() { h.sync(); }
    HS_DEVICE void after_first_loads() {}
    HS_DEVICE void after_first_pass() {}
    HS_DEVICE void before_last_pass() { h.before_last_pass(); }
};
template <int RL>
struct HsLoadRegsCol {
    static constexpr bool kPost = false;
    cplx (*y)[RL];
    HS_DEVICE cplx operator()(int u, int i, int) const { return y[u][i]; }
};

// STG: the tile's input was staged in shared memory by the copy engine ([row][C] at in_x) instead of being read from global memory
template <class PF, class PI, int CS, int C, int LDP, int STP, bool STG = false, class HK>
HS_DEVICE void hs_blu_mid_run(const cplx *HS_RESTRICT in_x, cplx *HS_RESTRICT out_x, const cplx *HS_RESTRICT twf, const cplx *HS_RESTRICT twi,
                              const HsTileGeom &g, const cplx *HS_RESTRICT hk, const int c0, cplx *smem, const int tid, const bool on, HK &hooks)
{
    static_assert(PF::N == PI::N && PF::T == PI::T && PF::NPAD == PI::NPAD, "plans must pair");
    constexpr int RL = PF::radix(PF::NST - 1);
    static_assert(PI::radix(0) == RL && PF::pass_size(PF::NPASS - 1) == 1 && PI::pass_size(0) == 1, "register hand-over needs matching radices");
    constexpr int NB = PF::N / RL, B = (NB + PF::T - 1) / PF::T;
    constexpr int NPADC = hs_col_stride(PF::NPAD, C);
    const int colA = tid % C, tA = tid / C;         // lane = column
    const int colB = tid / PF::T, tB = tid % PF::T; // lane = thread of a team
    const HsLane laneA{smem + colA * NPADC, tA, on}, laneB{smem + colB * NPADC, tB, on};
    cplx y[B][RL];

    const long long k = c0 + colA;
    HsStoreRegsMulCol<CS, RL> st{y, hk + k, (long long)g.ncols}; // Hk[k + n1 j]
    HsFirstHook<HK> h1{hooks}; // the forward transform loads from memory, the inverse one stores to it
    if constexpr (STG) {
        HsLoadTileTw ld;
        ld.tile = in_x + colA; ld.C = C;
        ld.twb = (const cplx *)g.twb + k; ld.tw_ld = g.tw_ld; ld.step = on ? hs_ld_tab((const cplx *)g.tws + k) : hs_c(0.0, 0.0);
        hs_fft_team2<PF, 1>(ld, st, laneA, laneA, twf, h1);
    } else {
        HsLoadColTw<LDP> ld;
        ld.base = in_x; ld.off0 = k * g.in_c; ld.es = g.in_e; ld.blk_log = 0; ld.blk_stride = 0;
        ld.twb = (const cplx *)g.twb + k; ld.tw_ld = g.tw_ld; ld.step = on ? hs_ld_tab((const cplx *)g.tws + k) : hs_c(0.0, 0.0);
        hs_fft_team2<PF, 1>(ld, st, laneA, laneA, twf, h1);
    }
    hooks.sync();
    HsLoadRegsCol<RL> ld2{y};
    HsStoreCol<STP> st2{out_x + (long long)(c0 + colB) * g.out_c, g.out_e};
    HsLastHook<HK> h2{hooks};
    hs_fft_team2<PI, -1>(ld2, st2, laneA, laneB, twi, h2);
}

template <class PF, class PI, int CS, int C, int MINB>
HS_GLOBAL void HS_LAUNCH_BOUNDS(PF::T *C, MINB)
hs_blu_mid_kernel(const cplx *HS_RESTRICT in, cplx *HS_RESTRICT out, const cplx *HS_RESTRICT twf, const cplx *HS_RESTRICT twi,
                  const HsTileGeom g, const cplx *HS_RESTRICT hk, long long howmany)
{
    HS_DYN_SMEM(cplx, smem);
    const int tiles = g.ncols / C;
    const long long xf = HS_BID / tiles;
    const int c0 = (int)(HS_BID % tiles) * C;
    HsNoHooks hooks;
    hs_blu_mid_run<PF, PI, CS, C, 0, 0>(in + xf * g.in_xf, out + xf * g.out_xf, twf, twi, g, hk, c0, smem, HS_TID, true, hooks);
}

// Four-step exchange twiddle (north_star config 5): b[k1][c] *= W_N^{(c0 + c) k1}, N = 2^logN, from two
// short tables (W_N^e = hi[e >> logS] * lo[e mod 2^logS]) instead of an N-entry one.
template <int SGN_UNUSED>
HS_GLOBAL void hs_dist_twiddle_kernel(cplx *HS_RESTRICT b, const cplx *HS_RESTRICT hi, const cplx *HS_RESTRICT lo, long long rows, long long cols,
                                      long long c0, int logN, int logS)
{
    const long long total = rows * cols;
    const unsigned long long maskN = (1ULL << logN) - 1, maskS = (1ULL << logS) - 1;
    for (long long idx = HS_BID * HS_NTHREADS + HS_TID; idx < total; idx += HS_NBLOCKS * HS_NTHREADS) {
        const unsigned long long k1 = (unsigned long long)(idx / cols), n2 = (unsigned long long)(c0 + idx % cols);
        const unsigned long long e = (k1 * n2) & maskN;
        const cplx w = hs_mul(hs_ld_tab(hi + (e >> logS)), hs_ld_tab(lo + (e & maskS)));
        b[idx] = hs_mul(b[idx], w);
    }
}
