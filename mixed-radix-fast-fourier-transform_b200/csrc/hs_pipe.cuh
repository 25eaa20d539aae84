// hs_pipe.cuh -- persistent dataflow pipeline: the passes of a long transform as ONE kernel whose
// intermediates never leave the L2.
//
// A transform of 2^14..2^20 points needs two tiled passes (hs_tiled.cuh); run as two launches over the whole
// batch the intermediate makes a round trip through HBM, which caps the pair at 50 % of the one-pass roofline
// (r2c: 43 %, long Bluestein: 13 %).  Here the passes are *phases* of one persistent kernel:
//   * work is cut into tiles (C columns of one transform in one phase) and handed out through a global
//     ticket counter in an order in which every tile's inputs are produced by tiles with smaller tickets:
//     round s = [phase 0 of group s | phase 1 of group s-lag | phase 2 of group s-2 lag], a group being G
//     consecutive transforms;
//   * a tile of phase j+1 needs the completion counter of phase j of its group to be full; tiles of phase j need
//     phase j+1 to have drained the ring slot they overwrite.  Both conditions were met long ago when lag and
//     the slot count are chosen as in hsfft_plan.c -- and because dependencies only point to smaller tickets,
//     which are held by CTAs that are already running, the scheme cannot deadlock whatever the number of
//     resident CTAs;
//   * none of the protocol's latencies is exposed: the next ticket is taken when a tile starts, its two
//     counters are peeked before the tile's last pass, and the tile's own completion is published (fence +
//     atomic) by one thread after the NEXT tile has issued its loads (hooks of hs_smem_fft.cuh); a tile whose
//     counters were full at the peek starts without any extra barrier;
//   * phase j writes its tiles into slot (group mod nslot) of a small ring; with nslot*G*N*16 bytes well below
//     the 126 MB L2 and the accesses tagged evict_last (hs_ld_mp / hs_st_mp), the ring is overwritten in place
//     and its dirty lines never reach HBM.  HBM sees the signal once on the way in and once on the way out.
// The tile bodies are the same device functions the stand-alone pass kernels use (hs_tile_run,
// hs_blu_mid_run), so results are bit-identical to the multi-launch path.
#pragma once
#include "hs_tiled.cuh"

#ifndef HS_PIPE_TILE_FN
#define HS_PIPE_TILE_FN HS_DEVICE
#endif

// ----------------------------------------------------------------------------------------- phases
struct HsPhNone {
    static constexpr int NT = 0, SMEM = 0, C = 1;
    template <int KNT, class HK>
    static HS_DEVICE void run(const hs_phase_args &, const cplx *, cplx *, long long, int, cplx *, HK &) {}
};

// one tiled pass; NT_ threads = P::T threads per column x C columns per tile
template <class P, int SGN, int NT_, int TW, bool STCOL, int IO, int CS, int LDP, int STP>
struct HsPhTiled {
    static constexpr int NT = NT_, C = NT_ / P::T;
    static_assert(C * P::T == NT_ && C >= 1, "thread count must be a multiple of the team size");
    static constexpr int SMEM = C * hs_col_stride(P::NPAD, C) * 16;
    template <int KNT, class HK> // KNT: threads of the kernel's CTA (>= NT; the surplus only takes part in the barriers)
    static HS_DEVICE void run(const hs_phase_args &pa, const cplx *in_x, cplx *out_x, long long x, int tile, cplx *smem, HK &hk)
    {
        hs_tile_run<P, SGN, C, TW, STCOL, IO, CS, LDP, STP>(in_x, out_x, (const cplx *)pa.tw, pa.g, x, tile * C, smem, HS_TID,
                                                            KNT == NT ? true : HS_TID < NT, hk);
    }
};

// middle of the long Bluestein: forward pass 2, spectrum product, inverse pass 1
template <class PF, class PI, int CS, int NT_, int LDP, int STP>
struct HsPhBluMid {
    static constexpr int NT = NT_, C = NT_ / PF::T;
    static_assert(C * PF::T == NT_ && C >= 1, "thread count must be a multiple of the team size");
    static constexpr int SMEM = C * hs_col_stride(PF::NPAD, C) * 16;
    template <int KNT, class HK>
    static HS_DEVICE void run(const hs_phase_args &pa, const cplx *in_x, cplx *out_x, long long, int tile, cplx *smem, HK &hk)
    {
        hs_blu_mid_run<PF, PI, CS, C, LDP, STP>(in_x, out_x, (const cplx *)pa.tw, (const cplx *)pa.tw_b, pa.g, (const cplx *)pa.aux, tile * C, smem,
                                                HS_TID, KNT == NT ? true : HS_TID < NT, hk);
    }
};

// r2c split (reference src/real.c:109-132) as a tile phase: a tile owns KPT*NT values of k in 0..H/2 and
// produces bins k and H-k (and their mirrors N-k, H+k) from Z[k] and Z[H-k], so Z is read once.
// pa.g: chirp = twiddle2, nlog = H; pa.flags & 1: write the mirrored upper half.
template <int NT_, int KPT, int LDP, int STP>
struct HsPhR2CSplit {
    static constexpr int NT = NT_, C = 1, SMEM = 0;
    static constexpr int KT = NT_ * KPT; // k values per tile
    static HS_DEVICE cplx bin(cplx a, cplx c, cplx w) { return hs_r2c_bin(a, c, w); }
    template <int KNT, class HK>
    static HS_DEVICE void run(const hs_phase_args &pa, const cplx *z, cplx *o, long long, int tile, cplx *, HK &)
    {
        const int H = pa.g.nlog;
        const cplx *HS_RESTRICT tw2 = (const cplx *)pa.g.chirp;
        const bool full = pa.flags & 1;
        HS_UNROLL
        for (int u = 0; u < KPT; ++u) {
            const int k = tile * KT + u * NT + HS_TID;
            if (HS_TID >= NT || 2 * k > H) continue;
            if (k == 0) {
                const cplx z0 = hs_ld_mp<LDP>(z);
                hs_st_mp<STP>(o, hs_c(z0.x + z0.y, 0.0));
                hs_st_mp<STP>(o + H, hs_c(z0.x - z0.y, 0.0));
            } else {
                const cplx a = hs_ld_mp<LDP>(z + k), c = (2 * k == H) ? a : hs_ld_mp<LDP>(z + (H - k));
                const cplx r = bin(a, c, hs_ld_tab(tw2 + k));
                hs_st_mp<STP>(o + k, r);
                if (full) hs_st_mp<STP>(o + (2 * H - k), hs_c(r.x, -r.y));
                if (2 * k != H) {
                    const cplx r2 = bin(c, a, hs_ld_tab(tw2 + (H - k)));
                    hs_st_mp<STP>(o + (H - k), r2);
                    if (full) hs_st_mp<STP>(o + (H + k), hs_c(r2.x, -r2.y));
                }
            }
        }
    }
};

// ----------------------------------------------------------------------------------------- kernel
template <class PH0, class PH1, class PH2>
struct HsPipeShape {
    static constexpr int NPH = PH2::NT ? 3 : 2;
    static constexpr int cmax(int a, int b) { return a > b ? a : b; }
    static constexpr int NT = cmax(PH0::NT, cmax(PH1::NT, PH2::NT));
    static constexpr int SMEM_TILES = cmax(PH0::SMEM, cmax(PH1::SMEM, PH2::SMEM));
    static constexpr int SMEM = SMEM_TILES + 64; // + the control block below
    static constexpr int SIG = NT >= 64 ? 32 : 0; // the thread that publishes completions (not in the poller's warp)
};

// per-CTA control block in shared memory, double buffered by iteration parity
struct HsPipeCtl {
    unsigned long long tk[2];  // ticket of the iteration
    unsigned *sig[2];          // counter of the previous iteration's tile, still to be published (or null)
    unsigned rdy[2];           // the ticket's dependencies were seen complete when peeked
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
// the two counters a ticket depends on: its producer phase (same group) and the consumer phase of the group
// that held its ring slot before.  need == 0: no such dependency.
struct HsPipeDeps {
    const unsigned *up, *dn;
    unsigned need_up, need_dn;
};
template <int NPH>
HS_DEVICE HsPipeDeps hs_pipe_deps(const hs_pipe_args &a, const HsPipeItem &it)
{
    const unsigned *done = (const unsigned *)((const char *)a.ctl + 128);
    HsPipeDeps d{done, done, 0u, 0u};
    if (!it.valid) return d;
    if (it.j > 0) { d.up = done + (long long)(it.j - 1) * a.ngroups + it.grp; d.need_up = a.items[it.j - 1]; }
    if (it.j < NPH - 1 && it.grp >= a.nslot) { d.dn = done + (long long)(it.j + 1) * a.ngroups + (it.grp - a.nslot); d.need_dn = a.items[it.j + 1]; }
    return d;
}

template <int NPH, int SIG>
struct HsPipeHooks {
    const hs_pipe_args &a;
    unsigned *sigp;                // SIG thread: completion still to be published
    unsigned long long nxt;        // thread 0: the next ticket (its atomic may still be in flight)
    bool next_ready;               // thread 0: set by before_last_pass
    bool fired1, fired2;
    HS_DEVICE void after_first_loads()
    {
        fired1 = true;
        if (HS_TID == SIG && sigp) hs_signal_add(sigp);
        sigp = nullptr;
    }
    HS_DEVICE void after_first_pass() const {}
    HS_DEVICE void before_last_pass()
    {
        fired2 = true;
        if (HS_TID == 0) {
            bool ok = true;
            if (nxt < a.total) {
                const HsPipeDeps d = hs_pipe_deps<NPH>(a, hs_pipe_decode(a, nxt));
                const unsigned vu = d.need_up ? hs_peek(d.up) : 0u, vd = d.need_dn ? hs_peek(d.dn) : 0u;
                ok = vu >= d.need_up && vd >= d.need_dn;
            }
            next_ready = ok;
        }
    }
    HS_DEVICE void flush() // a ticket without a transform inside (padding), or a phase without passes
    {
        if (!fired1) after_first_loads();
        if (!fired2) before_last_pass();
    }
};

template <class PH, int J, int NPH, int KNT, class HK>
HS_PIPE_TILE_FN void hs_pipe_tile(const hs_pipe_args &a, const long long grp, const unsigned q, cplx *smem, HK &hk)
{
    const hs_phase_args &pa = a.ph[J];
    const int xl = (int)(q / (unsigned)pa.tiles), tile = (int)(q % (unsigned)pa.tiles);
    const long long x = grp * a.G + xl;
    if (x >= a.howmany) return; // padding of the last group
    const long long r = (grp % a.nslot) * a.G + xl;
    const cplx *in_x = (J == 0) ? (const cplx *)a.in + x * pa.g.in_xf : (const cplx *)a.ring[J ? J - 1 : 0] + r * pa.g.in_xf;
    cplx *out_x = (J == NPH - 1) ? (cplx *)a.out + x * pa.g.out_xf : (cplx *)a.ring[J < NPH - 1 ? J : 0] + r * pa.g.out_xf;
    PH::template run<KNT>(pa, in_x, out_x, x, tile, smem, hk);
}

template <class PH0, class PH1, class PH2, int MINB>
HS_GLOBAL void HS_LAUNCH_BOUNDS((HsPipeShape<PH0, PH1, PH2>::NT), MINB)
hs_pipe_kernel(const HS_GRID_CONSTANT hs_pipe_args a)
{
    using SH = HsPipeShape<PH0, PH1, PH2>;
    constexpr int NPH = SH::NPH;
    HS_DYN_SMEM(cplx, smem);
    HsPipeCtl *const ctl = reinterpret_cast<HsPipeCtl *>(smem + SH::SMEM_TILES / 16);
    unsigned long long *const ticket = (unsigned long long *)a.ctl;
    unsigned *const done = (unsigned *)((char *)a.ctl + 128);

    if (HS_TID == 0) { ctl->tk[0] = hs_ticket_take(ticket); ctl->sig[0] = nullptr; ctl->rdy[0] = 0u; }
    HS_SYNC();
    for (int it = 0;; ++it) {
        const int cur = it & 1;
        const unsigned long long tk = ctl->tk[cur];
        HsPipeHooks<NPH, SH::SIG> hk{a, ctl->sig[cur], 0ULL, true, false, false};
        if (tk >= a.total) { hk.after_first_loads(); break; }
        if (HS_TID == 0) hk.nxt = hs_ticket_take(ticket); // in flight while this tile runs
        const HsPipeItem item = hs_pipe_decode(a, tk);
        if (item.valid) {
            if (!ctl->rdy[cur]) { // rare: the peek found a counter short.  Publish our own completion first (it may be the
                hk.after_first_loads(); // one we are waiting for), then poll.
                hk.fired1 = false;
                if (HS_TID == 0) {
                    const HsPipeDeps d = hs_pipe_deps<NPH>(a, item);
                    if (d.need_up) hs_wait_geq(d.up, d.need_up);
                    if (d.need_dn) hs_wait_geq(d.dn, d.need_dn);
                }
                HS_SYNC();
            }
            if (item.j == 0) hs_pipe_tile<PH0, 0, NPH, SH::NT>(a, item.grp, item.q, smem, hk);
            else if (item.j == 1) hs_pipe_tile<PH1, 1, NPH, SH::NT>(a, item.grp, item.q, smem, hk);
            else if constexpr (NPH == 3) hs_pipe_tile<PH2, 2, NPH, SH::NT>(a, item.grp, item.q, smem, hk);
        }
        hk.flush();
        if (HS_TID == 0) {
            ctl->tk[cur ^ 1] = hk.nxt;
            ctl->rdy[cur ^ 1] = hk.next_ready ? 1u : 0u;
            ctl->sig[cur ^ 1] = item.valid ? done + (long long)item.j * a.ngroups + item.grp : nullptr;
        }
        HS_SYNC(); // every thread's stores of this tile are issued; the tile buffers and the control block are handed over
    }
}
