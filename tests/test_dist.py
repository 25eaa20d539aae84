"""CPU tests of the multi-rank paths (gloo, world_size 2) -- the host-side logic of
 * the distributed four-step transform (local phases run the real kernel source through tests/emu,
   the exchange is torch.distributed.all_to_all_single over gloo), and
 * batch sharding (no collective on the data path, ranks own contiguous batch ranges).
"""
import importlib
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import importlib, os, sys
sys.path.insert(0, %(root)r); sys.path.insert(0, os.path.join(%(root)r, "tests"))
import numpy as np, torch, torch.distributed as dist
import oracle_lib as ol
hs = importlib.import_module("mixed-radix-fast-fourier-transform_b200")
dmod = importlib.import_module("mixed-radix-fast-fourier-transform_b200.dist")
L = hs.load(os.path.join(%(root)r, "tests", "emu", "_build", "libhsfft_emu.so"))
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
orc = ol.Oracle()
ok = True
# ---- four-step single transform
for (N1, N2, sgn) in ((128, 256, 1), (256, 128, -1), (512, 128, 1)):
    N = N1 * N2
    x = ol.rand_cplx(1, N, seed=N1 + 3 * N2)[0]          # same seeded signal on every rank
    plan = dmod.DistPlan(hs, N1, N2, sgn, rank=rank, world=world, lib=L)
    a = torch.view_as_real(torch.from_numpy(dmod.scatter_input(x, N1, N2, rank, world))).reshape(-1, 2).contiguous()
    out = torch.empty((N1 // world * N2, 2), dtype=torch.float64)
    plan.forward(a, out)
    mine = torch.view_as_complex(out).reshape(N1 // world, N2)
    parts = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(parts, mine)
    X = dmod.gather_output([p.numpy() for p in parts], N1, N2)
    err = ol.rel_rms(X, orc.c2c(x, sgn))
    if rank == 0:
        print("four-step %%dx%%d sgn=%%+d relRMS=%%.3e" %% (N1, N2, sgn, err))
    ok = ok and err <= ol.tol(N)
    plan.close()
# ---- batch sharding: contiguous batch ranges, no data-path collective
n, batch = 1024, 6
x = ol.rand_cplx(batch, n, seed=99)
lo, hi = rank * batch // world, (rank + 1) * batch // world
with hs.Plan(n, 1, lib=L) as p:
    mine = p.exec(x[lo:hi])
ok = ok and ol.rel_rms(mine, orc.c2c(x[lo:hi], 1)) <= ol.tol(n)
flag = torch.tensor([1 if ok else 0])
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
if rank == 0:
    print("ALL_OK" if int(flag) == 1 else "FAILED")
dist.destroy_process_group()
'''


def test_four_step_and_batch_sharding_world2(emu_lib, tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER % {"root": ROOT})
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", OMP_NUM_THREADS="1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", "29533", str(script)], capture_output=True, text=True, env=env, timeout=600)
    assert "ALL_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]


def test_four_step_layout_helpers():
    dmod = importlib.import_module("mixed-radix-fast-fourier-transform_b200.dist")
    N1, N2, G = 4, 8, 2
    x = np.arange(N1 * N2)
    a0 = dmod.scatter_input(x, N1, N2, 0, G)
    assert a0.shape == (N1, N2 // G) and a0[1, 2] == N2 * 1 + 2
    a1 = dmod.scatter_input(x, N1, N2, 1, G)
    assert a1[3, 0] == N2 * 3 + 4
    # out[k1l][k2] on rank g is X[(g Kw + k1l) + N1 k2]
    parts = [np.array([[(g * 2 + k1l) + N1 * k2 for k2 in range(N2)] for k1l in range(2)]) for g in range(G)]
    assert np.array_equal(dmod.gather_output(parts, N1, N2), np.arange(N1 * N2))


@pytest.mark.parametrize("N1,N2,world,sgn", [(128, 256, 2, 1), (256, 128, 4, -1), (16384, 128, 2, 1)])
def test_four_step_exchange_fused_into_phase_a(hs, emu_lib, oracle, N1, N2, world, sgn):
    """Peer-store path: phase A's last pass (one-pass and two-pass axis) twiddles and stores every row into the
    owning rank's receive buffer; all ranks are driven by one process here (same address space)."""
    import oracle_lib as ol
    dmod = importlib.import_module("mixed-radix-fast-fourier-transform_b200.dist")
    x = ol.rand_cplx(1, N1 * N2, seed=N1 + N2 + world)[0]
    X = dmod.local_group_forward(hs, x, N1, N2, sgn, world, lib=emu_lib)
    assert ol.rel_rms(X, oracle.c2c(x, sgn)) <= ol.tol(N1 * N2)
