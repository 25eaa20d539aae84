#!/usr/bin/env python
"""tools/sweep.py -- device-resident throughput sweep over transform lengths (development aid).

    python tools/sweep.py 4096:65536 3000:65536 1009:16384 ...   [--steps 10]

Prints one line per case: ms/step, GB/s (32 N batch bytes), fraction of the measured HBM peak,
GFLOP/s (5 N log2 N) and the plan description.  Set HSFFT_PREFER_CT / HSFFT_NO_CT to compare plans.
"""
import importlib
import json
import math
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    hs = importlib.import_module("mixed-radix-fast-fourier-transform_b200")
    steps = 10
    cases = []
    args = sys.argv[1:]
    while args:
        a = args.pop(0)
        if a == "--steps":
            steps = int(args.pop(0))
        else:
            n, b = a.split(":")
            cases.append((int(n), int(b)))
    try:
        peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        peak = 6650.0
    dev = torch.device("cuda", 0)
    for n, batch in cases:
        x = torch.rand((batch, n, 2), device=dev, dtype=torch.float64) * 2 - 1
        y = torch.empty_like(x)
        for sgn in (1,):
            with hs.Plan(n, sgn) as p:
                desc = p.describe()
                for _ in range(3):
                    p.exec_dev(x, y, batch)
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(steps):
                    p.exec_dev(x, y, batch)
                e1.record()
                torch.cuda.synchronize()
                ms = e0.elapsed_time(e1) / steps
            gbs = 32.0 * n * batch / (ms * 1e-3) / 1e9
            gfl = 5.0 * n * math.log2(n) * batch / (ms * 1e-3) / 1e9
            print("N=%-7d batch=%-7d %8.3f ms  %8.1f GB/s  %5.1f%% of %.0f  %9.1f GFLOP/s  [%s]" %
                  (n, batch, ms, gbs, 100 * gbs / peak, peak, gfl, desc), flush=True)
        del x, y
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
