// hs_pipe.cuh -- persistent, warp-specialised dataflow pipeline: the passes of a long transform as ONE kernel
// whose intermediates never leave the L2.
//
// A transform of 2^14..2^20 points needs two tiled passes (hs_tiled.cuh); run as two launches over the whole
// batch the intermediate makes a round trip through HBM (r2c: 56 N bytes of traffic for 24 N algorithmic, long
// Bluestein: 4 M + 2 N points for 2 N).  Here the passes are *phases* of one persistent kernel:
//   * work is cut into tiles (C columns of one transform in one phase) and handed out through a global
//     ticket counter in an order in which every tile's inputs are produced by tiles with smaller tickets:
//     round s = [phase 0 of group s | phase 1 of group s-lag | phase 2 of group s-2 lag], a group being G
//     consecutive transforms;
//   * phase j writes its tiles into slot (group mod nslot) of a small ring; with nslot*G*N*16 bytes well below
//     the 126 MB L2 and the accesses tagged evict_last (hs_ld_mp / hs_st_mp), the ring is overwritten in place
//     and its dirty lines never reach HBM.  HBM sees the signal once on the way in and once on the way out;
//   * a tile of phase j+1 needs the completion counter of phase j of its group to be full; tiles of phase j need
//     phase j+1 to have drained the ring slot they overwrite.  Dependencies only point to smaller tickets and
//     nobody blocks while it holds unpublished work, so the scheme cannot deadlock whatever the number of
//     resident CTAs;
//   * every CTA has two PRODUCER warps next to its compute warps.  The issuer takes the tickets, peeks the two
//     counters of the next tile, and -- when they are full -- has the TMA engine stage the tile's C x n strided
//     input (cp.async.bulk.tensor, UTMALDG) into one of two staging buffers while the compute warps still work
//     on the previous tile; the publisher makes finished tiles public (fence + atomic).  None of the protocol's
//     latencies is paid by the warps that do the arithmetic.  Compute warps wait on the staging buffer's
//     mbarrier, transform, store, bump two shared-memory counters; they synchronise among themselves with a
//     named barrier (hooks of hs_smem_fft.cuh).
// The tile bodies are the same device functions the stand-alone pass kernels use (hs_tile_run,
// hs_blu_mid_run), so results are bit-identical to the multi-launch path.
#pragma once
#include "hs_tiled.cuh"


// ----------------------------------------------------------------------------------------- phases
// NT: compute threads; SMEM: exchange buffer bytes; STAGED: the input tile is ROWS x C and the producer stages it
struct HsPhNone {
    static constexpr int NT = 0, SMEM = 0, C = 1, ROWS = 0, STAGE_ELEMS = 0;
    static constexpr bool STAGED = false;
    template <class HK>
    static HS_DEVICE void run(const hs_phase_args &, const cplx *, cplx *, long long, int, cplx *, HK &) {}
};

// one tiled pass; NT_ threads = P::T threads per column x C columns per tile
template <class P, int SGN, int NT_, int TW, bool STCOL, int IO, int CS, int LDP, int STP>
struct HsPhTiled {
    static constexpr int NT = NT_, C = NT_ / P::T, ROWS = P::N;
    static_assert(C * P::T == NT_ && C >= 1, "thread count must be a multiple of the team size");
    static constexpr int SMEM = C * hs_col_stride(P::NPAD, C) * 16;
    static constexpr bool STAGED = (IO == 0 || IO == 2); // plain strided loads; the fused loads (chirp, c2r merge) read global memory themselves
    static constexpr int STAGE_ELEMS = STAGED ? P::N * C : 0;
    template <class HK>
    static HS_DEVICE void run(const hs_phase_args &pa, const cplx *src, cplx *out_x, long long x, int tile, cplx *exch, HK &hk)
    {
        hs_tile_run<P, SGN, C, TW, STCOL, IO, CS, LDP, STP, STAGED>(src, out_x, (const cplx *)pa.tw, pa.g, x, tile * C, exch, HS_TID, true, hk);
    }
};

// middle of the long Bluestein: forward pass 2, spectrum product, inverse pass 1
template <class PF, class PI, int CS, int NT_, int LDP, int STP>
struct HsPhBluMid {
    static constexpr int NT = NT_, C = NT_ / PF::T, ROWS = PF::N;
    static_assert(C * PF::T == NT_ && C >= 1, "thread count must be a multiple of the team size");
    static constexpr int SMEM = C * hs_col_stride(PF::NPAD, C) * 16;
    static constexpr bool STAGED = true;
    static constexpr int STAGE_ELEMS = PF::N * C;
    template <class HK>
    static HS_DEVICE void run(const hs_phase_args &pa, const cplx *src, cplx *out_x, long long, int tile, cplx *exch, HK &hk)
    {
        hs_blu_mid_run<PF, PI, CS, C, LDP, STP, true>(src, out_x, (const cplx *)pa.tw, (const cplx *)pa.tw_b, pa.g, (const cplx *)pa.aux, tile * C, exch,
                                                      HS_TID, true, hk);
    }
};

// r2c split (reference src/real.c:109-132) as a tile phase: a tile owns KPT*NT values of k in 0..H/2 and
// produces bins k and H-k (and their mirrors N-k, H+k) from Z[k] and Z[H-k], so Z is read once.
// pa.g: chirp = twiddle2, nlog = H; pa.flags & 1: write the mirrored upper half.
template <int NT_, int KPT, int LDP, int STP>
struct HsPhR2CSplit {
    static constexpr int NT = NT_, C = 1, SMEM = 0, ROWS = 0, STAGE_ELEMS = 0;
    static constexpr bool STAGED = false;
    static constexpr int KT = NT_ * KPT; // k values per tile
    template <class HK>
    static HS_DEVICE void run(const hs_phase_args &pa, const cplx *z, cplx *o, long long, int tile, cplx *, HK &)
    {
        const int H = pa.g.nlog;
        const cplx *HS_RESTRICT tw2 = (const cplx *)pa.g.chirp;
        const bool full = pa.flags & 1;
        HS_UNROLL
        for (int u = 0; u < KPT; ++u) {
            const int k = tile * KT + u * NT + HS_TID;
            if (2 * k > H) continue;
            if (k == 0) {
                const cplx z0 = hs_ld_mp<LDP>(z);
                hs_st_mp<STP>(o, hs_c(z0.x + z0.y, 0.0));
                hs_st_mp<STP>(o + H, hs_c(z0.x - z0.y, 0.0));
            } else {
                const cplx a = hs_ld_mp<LDP>(z + k), c = (2 * k == H) ? a : hs_ld_mp<LDP>(z + (H - k));
                const cplx r = hs_r2c_bin(a, c, hs_ld_tab(tw2 + k));
                hs_st_mp<STP>(o + k, r);
                if (full) hs_st_mp<STP>(o + (2 * H - k), hs_c(r.x, -r.y));
                if (2 * k != H) {
                    const cplx r2 = hs_r2c_bin(c, a, hs_ld_tab(tw2 + (H - k)));
                    hs_st_mp<STP>(o + (H - k), r2);
                    if (full) hs_st_mp<STP>(o + (H + k), hs_c(r2.x, -r2.y));
                }
            }
        }
    }
};

// ----------------------------------------------------------------------------------------- kernel
// per-CTA queue between the producer warp and the compute warps
struct HsPipeQ {
    HsMbar full[2];            // producer -> compute: the slot's ticket is valid and its staged tile has landed
    unsigned long long tk[2];  // ticket in the slot (>= total: no more work)
    unsigned consumed;         // compute -> issuer: tiles whose staging buffer has been read (monotonic)
    unsigned finished;         // compute -> publisher: tiles whose stores are issued (monotonic)
    unsigned issued;           // issuer -> publisher: tiles handed to the compute warps (monotonic); bit 31: no more will follow
    unsigned pad_;
    unsigned *sig[8];          // issuer -> publisher: completion counter of tile (index mod 8)
};

template <class PH0, class PH1, class PH2>
struct HsPipeShape {
    static constexpr int NPH = PH2::NT ? 3 : 2;
    static constexpr int cmax(int a, int b) { return a > b ? a : b; }
    static constexpr int NT = cmax(PH0::NT, cmax(PH1::NT, PH2::NT)); // compute threads
    static_assert(PH0::NT == NT && PH1::NT == NT && (PH2::NT == NT || PH2::NT == 0), "every phase uses all compute warps");
    static constexpr int NTHREADS = NT + 64;                         // + the producer warps (issuer, publisher)
    static constexpr int EXCH = cmax(PH0::SMEM, cmax(PH1::SMEM, PH2::SMEM));
    static constexpr int STAGE_ELEMS = cmax(PH0::STAGE_ELEMS, cmax(PH1::STAGE_ELEMS, PH2::STAGE_ELEMS));
    static constexpr int SMEM = 2 * STAGE_ELEMS * 16 + EXCH + 128;   // staging buffers first (TMA destinations), the queue last
    // two CTAs per SM where they fit: their load / exchange / arithmetic phases then overlap on the SM's pipes
    static constexpr int MINB = (2 * (SMEM + 1024) <= 227 * 1024 && NTHREADS <= 256) ? 2 : 1;
    // Registers are handed out per warp in blocks of 512: 168 per thread (what ptxas picks for 2 x 192 threads) rounds
    // up to 176 and the second CTA no longer fits.  Declaring a slightly larger CTA caps ptxas at 160.
    static constexpr int BOUND = (MINB == 2 && NTHREADS == 192) ? 204 : NTHREADS;
};

struct HsPipeItem {
    int j;          // phase
    long long grp;  // group
    unsigned q;     // item inside the group's phase
    bool valid;
};
HS_DEVICE HsPipeItem hs_pipe_decode(const hs_pipe_args &a, const unsigned long long tk)
{
    HsPipeItem it;
    const long long s = (long long)(tk / a.round_items);
    unsigned q = (unsigned)(tk % a.round_items);
    it.j = 0;
    if (q >= a.items[0] + a.items[1]) { it.j = 2; q -= a.items[0] + a.items[1]; }
    else if (q >= a.items[0]) { it.j = 1; q -= a.items[0]; }
    it.q = q;
    it.grp = s - (long long)it.j * a.lag;
    it.valid = it.grp >= 0 && it.grp < a.ngroups;
    return it;
}
// where a tile sits: transform x, ring index r, tile number inside the transform
struct HsPipeWhere { long long x, r; int tile; };
HS_DEVICE HsPipeWhere hs_pipe_where(const hs_pipe_args &a, const HsPipeItem &it)
{
    const unsigned tiles = (unsigned)a.ph[it.j].tiles;
    const int xl = (int)(it.q / tiles);
    return HsPipeWhere{it.grp * a.G + xl, (it.grp % a.nslot) * a.G + xl, (int)(it.q % tiles)};
}
// the two counters a ticket depends on: its producer phase (same group) and the consumer phase of the group
// that held its ring slot before.  need == 0: no such dependency.
template <int NPH>
HS_DEVICE bool hs_pipe_deps_ready(const hs_pipe_args &a, const HsPipeItem &it)
{
    const unsigned *done = (const unsigned *)((const char *)a.ctl + 128);
    const bool up = it.j > 0, dn = it.j < NPH - 1 && it.grp >= a.nslot;
    // both loads are issued before either is looked at (one L2 round trip, not two)
    const unsigned vu = up ? hs_peek(done + (long long)(it.j - 1) * a.ngroups + it.grp) : 0u;
    const unsigned vd = dn ? hs_peek(done + (long long)(it.j + 1) * a.ngroups + (it.grp - a.nslot)) : 0u;
    return (!up || vu >= a.items[it.j - 1]) && (!dn || vd >= a.items[it.j + 1]);
}

// hooks of the compute warps: barriers among the compute threads only, staging buffer handed back after the first pass
template <int NT>
struct HsComputeHooks {
    HsPipeQ *q;
    bool released;
    HS_DEVICE void sync() const { HS_SYNC_COMPUTE(NT); }
    HS_DEVICE void after_first_loads() const { hs_fence_async_proxy(); } // my reads of the staging buffer precede its asynchronous refill
    HS_DEVICE void after_first_pass()
    {
        released = true;
        if (HS_TID == 0) hs_flag_add(&q->consumed); // the pass's trailing barrier is behind us: everybody has read the buffer
    }
    HS_DEVICE void before_last_pass() const {}
};

template <class PH, int J, int NPH, class HK>
HS_DEVICE void hs_pipe_tile(const hs_pipe_args &a, const HsPipeItem &it, const cplx *stage, cplx *exch, HK &hk)
{
    const hs_phase_args &pa = a.ph[J];
    const HsPipeWhere w = hs_pipe_where(a, it);
    const cplx *src = PH::STAGED ? stage
                                 : ((J == 0) ? (const cplx *)a.in + w.x * pa.g.in_xf : (const cplx *)a.ring[J ? J - 1 : 0] + w.r * pa.g.in_xf);
    cplx *out_x = (J == NPH - 1) ? (cplx *)a.out + w.x * pa.g.out_xf : (cplx *)a.ring[J < NPH - 1 ? J : 0] + w.r * pa.g.out_xf;
    PH::run(pa, src, out_x, w.x, w.tile, exch, hk);
}

// producer: make the tile of `it` available in slot b (stage it through the TMA engine, or just release the slot)
template <class PH, int J>
HS_DEVICE void hs_pipe_issue(const hs_pipe_args &a, const HsTmap *tm, const HsPipeItem &it, cplx *stage, HsMbar *bar)
{
    if constexpr (PH::STAGED) {
        constexpr int BOXR = PH::ROWS < 256 ? PH::ROWS : 256; // rows per TMA box (hardware limit 256 per dimension)
        const HsPipeWhere w = hs_pipe_where(a, it);
        hs_fence_proxy_async_all(); // the counters were read through the generic proxy; the tile is read through the asynchronous one
        hs_mbar_expect(bar, (unsigned)(PH::STAGE_ELEMS * 16));
        for (int r0 = 0; r0 < PH::ROWS; r0 += BOXR)
            hs_tma_load_3d_mp<(J == 0 ? 2 : 1)>(stage + r0 * PH::C, tm, 2 * w.tile * PH::C, r0, J == 0 ? w.x : w.r, BOXR, PH::C * 16, bar);
    } else {
        hs_mbar_arrive(bar);
    }
}

template <class PH0, class PH1, class PH2>
HS_GLOBAL void HS_LAUNCH_BOUNDS((HsPipeShape<PH0, PH1, PH2>::BOUND), (HsPipeShape<PH0, PH1, PH2>::MINB))
hs_pipe_kernel(const HS_GRID_CONSTANT hs_pipe_args a, const HS_GRID_CONSTANT HsTmap tm0, const HS_GRID_CONSTANT HsTmap tm1,
               const HS_GRID_CONSTANT HsTmap tm2)
{
    using SH = HsPipeShape<PH0, PH1, PH2>;
    constexpr int NPH = SH::NPH, NT = SH::NT;
    HS_DYN_SMEM(cplx, smem);
    cplx *const stage0 = smem, *const exch = smem + 2 * SH::STAGE_ELEMS;
    HsPipeQ *const q = reinterpret_cast<HsPipeQ *>(exch + SH::EXCH / 16);

    if (HS_TID == 0) {
        hs_mbar_init(&q->full[0]); hs_mbar_init(&q->full[1]);
        q->consumed = 0; q->finished = 0; q->issued = 0;
    }
    hs_fence_async_proxy();
    HS_SYNC();

    if (HS_TID < NT) {
        // ------------------------------------------------------------------ compute warps
        for (unsigned i = 0;; ++i) {
            const int b = (int)(i & 1u);
            hs_mbar_wait(&q->full[b], (i >> 1) & 1u);
            const unsigned long long tk = q->tk[b];
            if (tk >= a.total) break;
            const HsPipeItem it = hs_pipe_decode(a, tk);
            const cplx *stage = stage0 + b * SH::STAGE_ELEMS;
            HsComputeHooks<NT> hk{q, false};
            if (it.j == 0) hs_pipe_tile<PH0, 0, NPH>(a, it, stage, exch, hk);
            else if (it.j == 1) hs_pipe_tile<PH1, 1, NPH>(a, it, stage, exch, hk);
            else if constexpr (NPH == 3) hs_pipe_tile<PH2, 2, NPH>(a, it, stage, exch, hk);
            if (!hk.released) { hs_fence_async_proxy(); HS_SYNC_COMPUTE(NT); hk.after_first_pass(); } // a phase without passes
            HS_SYNC_COMPUTE(NT); // every compute thread's stores are issued; the exchange buffer is free
            if (HS_TID == 0) hs_flag_add(&q->finished);
        }
    } else if (HS_TID == NT) {
        // ------------------------------------------------------------------ issuer (one lane of the first producer warp)
        unsigned long long *const ticket = (unsigned long long *)a.ctl;
        unsigned *const done = (unsigned *)((char *)a.ctl + 128);
        unsigned issued = 0;
        unsigned long long tk = hs_ticket_take(ticket);
        for (unsigned idle = 0;;) {
            bool progress = false;
            // slot free: its previous tile has left the staging buffer, and the publisher's ring of 8 has room
            if ((issued < 2 || hs_flag_read(&q->consumed) + 1 >= issued) && issued < hs_flag_read(&q->finished) + 6) {
                const int b = (int)(issued & 1u);
                if (tk >= a.total) {
                    q->tk[b] = tk;
                    hs_mbar_arrive(&q->full[b]);
                    hs_flag_set(&q->issued, issued | 0x80000000u);
                    break;
                }
                const HsPipeItem it = hs_pipe_decode(a, tk);
                bool skip = !it.valid;
                if (it.valid && hs_pipe_where(a, it).x >= a.howmany) { // padding of the last group: nothing to compute, still counted
                    hs_signal_add(done + (long long)it.j * a.ngroups + it.grp);
                    skip = true;
                }
                if (skip) {
                    tk = hs_ticket_take(ticket);
                    progress = true;
                } else if (hs_pipe_deps_ready<NPH>(a, it)) {
                    cplx *stage = stage0 + b * SH::STAGE_ELEMS;
                    q->tk[b] = tk;
                    q->sig[issued & 7u] = done + (long long)it.j * a.ngroups + it.grp;
                    if (it.j == 0) hs_pipe_issue<PH0, 0>(a, &tm0, it, stage, &q->full[b]);
                    else if (it.j == 1) hs_pipe_issue<PH1, 1>(a, &tm1, it, stage, &q->full[b]);
                    else hs_pipe_issue<PH2, 2>(a, &tm2, it, stage, &q->full[b]);
                    ++issued;
                    hs_flag_set(&q->issued, issued);
                    tk = hs_ticket_take(ticket);
                    progress = true;
                }
            }
            if (progress) idle = 0;
            else { hs_backoff(); if (++idle > (1u << 25)) HS_TRAP(); } // seconds without progress: fail loudly instead of hanging the GPU
        }
    } else if (HS_TID == NT + 32) {
        // ------------------------------------------------------------------ publisher (one lane of the second producer warp)
        unsigned published = 0;
        for (unsigned idle = 0;;) {
            const unsigned iss = hs_flag_read(&q->issued);
            if (published < (iss & 0x7fffffffu) && hs_flag_read(&q->finished) > published) {
                hs_signal_add(q->sig[published & 7u]); // fence (cumulative over the compute warps' stores), then count the tile
                ++published;
                idle = 0;
            } else if ((iss & 0x80000000u) && published == (iss & 0x7fffffffu)) {
                break;
            } else {
                hs_backoff();
                if (++idle > (1u << 25)) HS_TRAP();
            }
        }
    }
}
