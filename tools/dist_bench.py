#!/usr/bin/env python
"""tools/dist_bench.py -- single large 1-D complex FFT over all ranks (BASELINE.json config 5).

    torchrun --nproc-per-node G tools/dist_bench.py --log2n 30 [--steps 3]

Four-step N = N1*N2 with one all-to-all (NCCL over NVLink).  Reports per-phase device times (max
over ranks), GFLOP/s (5 N log2 N), local-phase HBM GB/s and exchange GB/s per GPU, and verifies the
result on 64 sampled bins against direct O(N) DFT sums (fp64, exact integer phase reduction) plus
Parseval, as SURVEY.md section 8d prescribes for sizes no CPU oracle can hold.
"""
import argparse
import importlib
import json
import math
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def plan_ptr(t):
    import ctypes
    return ctypes.c_void_p(t.data_ptr())


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--log2n", type=int, default=30)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--bins", type=int, default=64)
    ap.add_argument("--p2p", action="store_true", help="exchange fused into phase A (peer stores over NVLink) instead of an NCCL all-to-all")
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    hs = importlib.import_module("mixed-radix-fast-fourier-transform_b200")
    dmod = importlib.import_module("mixed-radix-fast-fourier-transform_b200.dist")
    rank, local, world = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    hs.load().hsfft_set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    l1 = args.log2n // 2
    N1, N2 = 1 << l1, 1 << (args.log2n - l1)
    N = N1 * N2
    Cw, Kw = N2 // world, N1 // world
    plan = dmod.DistPlan(hs, N1, N2, 1, rank=rank, world=world, device=dev)
    g = torch.Generator(device=dev); g.manual_seed(1234 + rank)
    a = torch.rand((N1 * Cw, 2), generator=g, device=dev, dtype=torch.float64) * 2 - 1
    out = torch.empty((Kw * N2, 2), device=dev, dtype=torch.float64)

    def sync():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    if args.p2p:
        plan.enable_p2p()
    fwd = plan.forward_p2p if args.p2p else plan.forward
    fwd(a, out)  # warm-up (NCCL channels, kernels)
    fwd(a, out)
    sync()
    tA = tX = tB = 0.0
    for _ in range(args.steps):
        e = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        sync()
        if args.p2p:  # phase A includes the exchange; "exchange" is what is left of it: the barrier
            par = plan.calls & 1
            plan.calls += 1
            e[0].record()
            plan.L.hsfft_dist_cols_exchange_dev(plan.pp[par], plan_ptr(a), plan_ptr(plan.ws), plan._stream())
            e[1].record()
            if world > 1:
                dist.all_reduce(plan._flag)
            e[2].record()
            plan.L.hsfft_dist_rows_dev(plan.pp[par], plan.rbuf[par], plan_ptr(out), plan_ptr(plan.ws), plan._stream())
            e[3].record()
        else:
            e[0].record(); plan.cols(a); e[1].record(); plan.exchange(); e[2].record(); plan.rows(out); e[3].record()
        sync()
        t = torch.tensor([e[0].elapsed_time(e[1]), e[1].elapsed_time(e[2]), e[2].elapsed_time(e[3])], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        tA += t[0].item(); tX += t[1].item(); tB += t[2].item()
    tA, tX, tB = tA / args.steps, tX / args.steps, tB / args.steps

    # ---- verification: sampled bins + Parseval
    xa = torch.view_as_complex(a).reshape(N1, Cw)
    n1 = torch.arange(N1, device=dev, dtype=torch.int64).reshape(N1, 1)
    c = torch.arange(Cw, device=dev, dtype=torch.int64).reshape(1, Cw)
    nidx = N2 * n1 + (rank * Cw + c)                       # natural index of every local sample
    gen = torch.Generator(); gen.manual_seed(7)
    ks = torch.randint(0, N, (args.bins,), generator=gen, dtype=torch.int64).tolist()
    ks[0], ks[1] = 0, N - 1
    want = torch.zeros(args.bins, dtype=torch.complex128, device=dev)
    for i, k in enumerate(ks):
        ph = (nidx * k) % N                                # exact: < 2^60
        ang = ph.to(torch.float64) * (-2.0 * math.pi / N)
        want[i] = (xa * torch.complex(torch.cos(ang), torch.sin(ang))).sum()
    energy_in = (xa.abs() ** 2).sum()
    X = torch.view_as_complex(out).reshape(Kw, N2)
    energy_out = (X.abs() ** 2).sum()
    got = torch.zeros(args.bins, dtype=torch.complex128, device=dev)
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
    if rank == 0:
        total = tA + tX + tB
        slab = N1 * Cw * 16.0
        passes_a = 2 * (2 if N1 > 1024 else 1) + 2  # tiled passes (read+write each) + twiddle pass
        passes_b = 2 * (2 if N2 > 1024 else 1)
        line = {"workload": "single c2c N=2^%d fp64, four-step %dx%d on %d GPUs" % (args.log2n, N1, N2, world),
                "exchange": "fused into phase A: peer stores over NVLink from the transform kernel, all-reduce barrier" if args.p2p
                            else "NCCL all_to_all_single between the phases",
                "ms_total": total, "ms_cols": tA, "ms_exchange": tX, "ms_rows": tB,
                "gflops": 5.0 * N * args.log2n / (total * 1e-3) / 1e9,
                "local_hbm_gbs_per_gpu": {"cols": passes_a * slab / (tA * 1e-3) / 1e9, "rows": passes_b * slab / (tB * 1e-3) / 1e9},
                "exchange_gbs_per_gpu_per_direction": (slab * (world - 1) / world / ((tA if args.p2p else tX) * 1e-3) / 1e9) if world > 1 else None,
                "nvlink_peak_gbs": 900, "nvlink_measured_gbs": 770,
                "sampled_bins": args.bins, "rel_rms_err_sampled": err, "tolerance": 1e-12 * args.log2n, "parseval_rel": parseval,
                "ok": bool(err <= 1e-12 * args.log2n and parseval < 1e-12)}
        print(json.dumps(line), flush=True)
    plan.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
