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
    rec = dmod.bench_four_step(hs, args.log2n, args.steps, args.p2p, rank, world, dev, bins=args.bins)
    if rank == 0:
        rec["workload"] = "single c2c N=2^%d fp64, four-step %s on %d GPUs" % (args.log2n, rec["four_step"], world)
        print(json.dumps(rec), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
