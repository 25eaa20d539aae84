"""Multi-GPU plumbing of hsfft-b200: one process per GPU, torch.distributed for the exchange.

* Batched transforms shard by batch index with no communication (bench.py does exactly that).
* A single very large 1-D transform runs as a four-step decomposition N = N1*N2: two local phases
  (sm_100a kernels behind the C ABI, csrc/hsfft_dist.c) with ONE all-to-all between them
  (NCCL over NVLink on GPUs; gloo on CPU for the tests).

Layouts (rank g of G, Cw = N2/G, Kw = N1/G):
    input   a[n1][c]      = x[N2*n1 + g*Cw + c]           column slab of the N1 x N2 matrix
    output  out[k1l][k2]  = X[(g*Kw + k1l) + N1*k2]       row slab, transposed order
`scatter_input` / `gather_output` convert from / to natural order for tests.

Two ways to run the exchange:
* `forward`      -- phase A, `torch.distributed.all_to_all_single` (NCCL), phase B;
* `forward_p2p`  -- phase A's last pass multiplies by the exchange twiddle and stores every row straight into the
                    receive buffer of the rank that owns it (peer memory mapped with CUDA IPC, stores issued by the
                    transform kernel over NVLink while it computes); one tiny all-reduce as a barrier, then phase B.
"""
import ctypes as C

import numpy as np


class DistPlan:
    def __init__(self, hs, N1, N2, sgn, rank=0, world=1, lib=None, device=None):
        import torch
        self.torch = torch
        self.L = lib or hs.load()
        self.N1, self.N2, self.sgn, self.rank, self.world = N1, N2, sgn, rank, world
        self.Cw, self.Kw = N2 // world, N1 // world
        self.device = device if device is not None else torch.device("cpu")
        self.p = self.L.hsfft_dist_init(N1, N2, world, rank, sgn)
        n = N1 * self.Cw
        mk = lambda: torch.empty((n, 2), dtype=torch.float64, device=self.device)
        self.b, self.ws, self.r = mk(), mk(), mk()

    def _stream(self):
        if self.device.type == "cuda":
            return C.c_void_p(self.torch.cuda.current_stream().cuda_stream)
        return None

    def cols(self, a):
        self.L.hsfft_dist_cols_dev(self.p, C.c_void_p(a.data_ptr()), C.c_void_p(self.b.data_ptr()), C.c_void_p(self.ws.data_ptr()),
                                   self._stream())
        return self.b

    def exchange(self, group=None):
        """rows [g'*Kw, (g'+1)*Kw) of b -> rank g'; received blocks land as r[g][k1l][c]."""
        if self.world == 1:
            self.r.copy_(self.b)
        else:
            import torch.distributed as dist
            dist.all_to_all_single(self.r, self.b, group=group)
        return self.r

    def rows(self, out):
        self.L.hsfft_dist_rows_dev(self.p, C.c_void_p(self.r.data_ptr()), C.c_void_p(out.data_ptr()), C.c_void_p(self.ws.data_ptr()),
                                   self._stream())
        return out

    def forward(self, a, out, group=None):
        self.cols(a)
        self.exchange(group)
        return self.rows(out)

    # ---- exchange fused into phase A (peer stores instead of an all-to-all)
    def enable_p2p(self, group=None):
        """Allocate two receive buffers per rank (alternated call by call), exchange their IPC handles, map the peers."""
        torch, L = self.torch, self.L
        nbytes = self.N1 * self.Cw * 16
        self.pp, self.rbuf, self.calls = [], [], 0
        for _ in range(2):
            plan = self.p if not self.pp else L.hsfft_dist_init(self.N1, self.N2, self.world, self.rank, self.sgn)
            r = L.hsfft_malloc_device(nbytes)
            if self.world == 1:
                arr = (C.c_void_p * 1)(r)
                L.hsfft_dist_set_peer_pointers(plan, arr, 1)
            else:
                import torch.distributed as dist
                h = (C.c_ubyte * 64)()
                L.hsfft_ipc_get_handle(C.c_void_p(r), h)
                mine = torch.tensor(list(h), dtype=torch.uint8, device=self.device)
                allh = [torch.empty_like(mine) for _ in range(self.world)]
                dist.all_gather(allh, mine, group=group)
                blob = bytes(torch.cat(allh).cpu().tolist())
                L.hsfft_dist_set_peers(plan, C.c_void_p(r), blob, self.world)
            self.pp.append(plan)
            self.rbuf.append(r)
        self._flag = torch.zeros(1, device=self.device)
        return self

    def forward_p2p(self, a, out, group=None):
        par = self.calls & 1
        self.calls += 1
        L = self.L
        L.hsfft_dist_cols_exchange_dev(self.pp[par], C.c_void_p(a.data_ptr()), C.c_void_p(self.ws.data_ptr()), self._stream())
        if self.world > 1:
            import torch.distributed as dist
            dist.all_reduce(self._flag, group=group)  # stream-ordered barrier: every rank's peer stores have landed
        L.hsfft_dist_rows_dev(self.pp[par], C.c_void_p(self.rbuf[par]), C.c_void_p(out.data_ptr()), C.c_void_p(self.ws.data_ptr()),
                              self._stream())
        return out

    def close(self):
        for plan in getattr(self, "pp", [])[1:]:
            self.L.hsfft_dist_free(plan)
        for r in getattr(self, "rbuf", []):
            self.L.hsfft_free_device(r)
        self.pp, self.rbuf = [], []
        if self.p:
            self.L.hsfft_dist_free(self.p)
            self.p = None


def local_group_forward(hs, x, N1, N2, sgn, world, lib=None, device=None):
    """All `world` ranks of the peer-store path driven by ONE process (buffers addressable from every plan): the
    addressing of the fused exchange without a second process.  x: natural-order signal; returns natural-order X."""
    import torch
    L = lib or hs.load()
    dev = device if device is not None else torch.device("cpu")
    Cw, Kw = N2 // world, N1 // world
    mk = lambda n: torch.empty((n, 2), dtype=torch.float64, device=dev)
    plans = [L.hsfft_dist_init(N1, N2, world, g, sgn) for g in range(world)]
    rb = [mk(N1 * Cw) for _ in range(world)]
    ptrs = (C.c_void_p * world)(*[r.data_ptr() for r in rb])
    for pl in plans:
        L.hsfft_dist_set_peer_pointers(pl, ptrs, world)
    ws, outs = mk(N1 * Cw), []
    for g, pl in enumerate(plans):
        a = torch.view_as_real(torch.from_numpy(scatter_input(x, N1, N2, g, world))).reshape(-1, 2).contiguous().to(dev)
        L.hsfft_dist_cols_exchange_dev(pl, C.c_void_p(a.data_ptr()), C.c_void_p(ws.data_ptr()), None)
    if dev.type == "cuda":
        torch.cuda.synchronize()
    for g, pl in enumerate(plans):
        out = mk(Kw * N2)
        L.hsfft_dist_rows_dev(pl, C.c_void_p(rb[g].data_ptr()), C.c_void_p(out.data_ptr()), C.c_void_p(ws.data_ptr()), None)
        if dev.type == "cuda":
            torch.cuda.synchronize()
        outs.append(torch.view_as_complex(out).reshape(Kw, N2).cpu().numpy())
    for pl in plans:
        L.hsfft_dist_free(pl)
    return gather_output(outs, N1, N2)


def scatter_input(x, N1, N2, rank, world):
    """Natural-order x[N] -> the rank's column slab a[N1][Cw] (numpy)."""
    Cw = N2 // world
    return np.ascontiguousarray(np.asarray(x).reshape(N1, N2)[:, rank * Cw:(rank + 1) * Cw])


def gather_output(parts, N1, N2):
    """Per-rank out[Kw][N2] (list ordered by rank) -> natural-order X[k1 + N1*k2] (numpy)."""
    full = np.concatenate([np.asarray(p) for p in parts], axis=0)  # [k1][k2]
    return np.ascontiguousarray(full.T).reshape(-1)               # index k2*N1 + k1


def bench_four_step(hs, log2n, steps, p2p, rank, world, dev, bins=64):
    """Device-timed phases (max over ranks) of one c2c of N = 2^log2n points over `world` ranks (torch.distributed already
    initialised when world > 1), verified on `bins` sampled output bins against direct O(N) DFT sums (fp64, exact integer
    phase reduction) plus Parseval -- SURVEY.md section 8d's check for sizes no CPU oracle can hold.  Returns a dict."""
    import math
    import torch
    import torch.distributed as dist
    l1 = log2n // 2
    N1, N2 = 1 << l1, 1 << (log2n - l1)
    N = N1 * N2
    Cw, Kw = N2 // world, N1 // world
    plan = DistPlan(hs, N1, N2, 1, rank=rank, world=world, device=dev)
    g = torch.Generator(device=dev)
    g.manual_seed(1234 + rank)
    a = torch.rand((N1 * Cw, 2), generator=g, device=dev, dtype=torch.float64) * 2 - 1
    out = torch.empty((Kw * N2, 2), device=dev, dtype=torch.float64)
    ptr = lambda t: C.c_void_p(t.data_ptr())

    def sync():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    if p2p:
        plan.enable_p2p()
    fwd = plan.forward_p2p if p2p else plan.forward
    for _ in range(3):
        fwd(a, out)  # warm-up (NCCL channels, kernels, both receive buffers)
    sync()
    tA = tX = tB = 0.0
    for _ in range(steps):
        e = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        sync()
        if p2p:  # phase A includes the exchange; what is left of "exchange" is the barrier
            par = plan.calls & 1
            plan.calls += 1
            e[0].record()
            plan.L.hsfft_dist_cols_exchange_dev(plan.pp[par], ptr(a), ptr(plan.ws), plan._stream())
            e[1].record()
            if world > 1:
                dist.all_reduce(plan._flag)
            e[2].record()
            plan.L.hsfft_dist_rows_dev(plan.pp[par], plan.rbuf[par], ptr(out), ptr(plan.ws), plan._stream())
            e[3].record()
        else:
            e[0].record(); plan.cols(a); e[1].record(); plan.exchange(); e[2].record(); plan.rows(out); e[3].record()
        sync()
        t = torch.tensor([e[0].elapsed_time(e[1]), e[1].elapsed_time(e[2]), e[2].elapsed_time(e[3])], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        tA += t[0].item(); tX += t[1].item(); tB += t[2].item()
    tA, tX, tB = tA / steps, tX / steps, tB / steps

    # verification: sampled bins + Parseval
    xa = torch.view_as_complex(a).reshape(N1, Cw)
    n1 = torch.arange(N1, device=dev, dtype=torch.int64).reshape(N1, 1)
    c = torch.arange(Cw, device=dev, dtype=torch.int64).reshape(1, Cw)
    nidx = N2 * n1 + (rank * Cw + c)                       # natural index of every local sample
    gen = torch.Generator()
    gen.manual_seed(7)
    ks = torch.randint(0, N, (bins,), generator=gen, dtype=torch.int64).tolist()
    ks[0], ks[1] = 0, N - 1
    want = torch.zeros(bins, dtype=torch.complex128, device=dev)
    for i, k in enumerate(ks):
        ph = (nidx * k) % N                                # exact: < 2^60
        ang = ph.to(torch.float64) * (-2.0 * math.pi / N)
        want[i] = (xa * torch.complex(torch.cos(ang), torch.sin(ang))).sum()
    energy_in = (xa.abs() ** 2).sum()
    X = torch.view_as_complex(out).reshape(Kw, N2)
    energy_out = (X.abs() ** 2).sum()
    got = torch.zeros(bins, dtype=torch.complex128, device=dev)
    for i, k in enumerate(ks):
        k1, k2 = k % N1, k // N1
        if k1 // Kw == rank:
            got[i] = X[k1 % Kw, k2]
    tot = torch.stack([energy_in, energy_out]).to(torch.float64)
    if world > 1:
        wr, wi, gr, gi = want.real.contiguous(), want.imag.contiguous(), got.real.contiguous(), got.imag.contiguous()
        for t in (wr, wi, gr, gi, tot):
            dist.all_reduce(t)
        want, got = torch.complex(wr, wi), torch.complex(gr, gi)
    err = float(((got - want).abs().pow(2).sum() / want.abs().pow(2).sum()).sqrt())
    parseval = float(abs(tot[1] / N - tot[0]) / tot[0])
    total = tA + tX + tB
    slab = N1 * Cw * 16.0
    passes_a = 2 * (2 if N1 > 1024 else 1) + (0 if p2p else 2)  # tiled passes (read + write each) + the twiddle pass of the NCCL variant
    passes_b = 2 * (2 if N2 > 1024 else 1)
    sent = slab * (world - 1) / world
    rec = {"four_step": "%dx%d" % (N1, N2),
           "exchange": "fused into phase A: NVLink peer stores from the transform kernel, all-reduce barrier" if p2p
                       else "NCCL all_to_all_single between the phases",
           "ms_total": total, "ms_cols": tA, "ms_exchange": tX, "ms_rows": tB,
           "gflops": 5.0 * N * log2n / (total * 1e-3) / 1e9,
           "local_hbm_gbs_per_gpu": {"cols": passes_a * slab / (tA * 1e-3) / 1e9, "rows": passes_b * slab / (tB * 1e-3) / 1e9},
           "exchange_gbs_per_gpu_per_direction": (sent / ((tA if p2p else tX) * 1e-3) / 1e9) if world > 1 else None,
           "exchange_bytes_sent_per_gpu": sent, "nvlink_peak_gbs": 900, "nvlink_measured_gbs": 770,
           "sampled_bins": bins, "rel_rms_err_sampled": err, "tolerance": 1e-12 * log2n, "parseval_rel": parseval,
           "ok": bool(err <= 1e-12 * log2n and parseval < 1e-12)}
    if world > 1 and rec["exchange_gbs_per_gpu_per_direction"]:
        rec["exchange_frac_of_measured_nvlink"] = rec["exchange_gbs_per_gpu_per_direction"] / 770.0
    plan.close()
    del a, out
    torch.cuda.empty_cache()
    return rec

