"""Multi-GPU plumbing of hsfft-b200: one process per GPU, torch.distributed for the exchange.

* Batched transforms shard by batch index with no communication (bench.py does exactly that).
* A single very large 1-D transform runs as a four-step decomposition N = N1*N2: two local phases
  (sm_100a kernels behind the C ABI, csrc/hsfft_dist.c) with ONE all-to-all between them
  (NCCL over NVLink on GPUs; gloo on CPU for the tests).

Layouts (rank g of G, Cw = N2/G, Kw = N1/G):
    input   a[n1][c]      = x[N2*n1 + g*Cw + c]           column slab of the N1 x N2 matrix
    output  out[k1l][k2]  = X[(g*Kw + k1l) + N1*k2]       row slab, transposed order
`scatter_input` / `gather_output` convert from / to natural order for tests.
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

    def close(self):
        if self.p:
            self.L.hsfft_dist_free(self.p)
            self.p = None


def scatter_input(x, N1, N2, rank, world):
    """Natural-order x[N] -> the rank's column slab a[N1][Cw] (numpy)."""
    Cw = N2 // world
    return np.ascontiguousarray(np.asarray(x).reshape(N1, N2)[:, rank * Cw:(rank + 1) * Cw])


def gather_output(parts, N1, N2):
    """Per-rank out[Kw][N2] (list ordered by rank) -> natural-order X[k1 + N1*k2] (numpy)."""
    full = np.concatenate([np.asarray(p) for p in parts], axis=0)  # [k1][k2]
    return np.ascontiguousarray(full.T).reshape(-1)               # index k2*N1 + k1
