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
