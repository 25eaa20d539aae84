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
    kernels = False
    cases = []
    args = sys.argv[1:]
    while args:
        a = args.pop(0)
        if a == "--steps":
            steps = int(args.pop(0))
        elif a == "--kernels":       # per-kernel times of one extra pass (CUDA events around every launch)
            kernels = True
        else:
            f = a.split(":")
            if f[0] == "fft2":           # fft2:<n0>x<n1>:<batch>
                n0, n1 = (int(v) for v in f[1].split("x"))
                cases.append(("fft2", (n0, n1), int(f[2])))
            elif len(f) == 2:
                cases.append(("c2c", int(f[0]), int(f[1])))
            else:
                cases.append((f[0], int(f[1]), int(f[2])))
    try:
        peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        peak = 6650.0
    dev = torch.device("cuda", 0)
    for kind, n, batch in cases:
        if kind == "fft2":
            n0, n1 = n
            x = torch.rand((batch, n0, n1, 2), device=dev, dtype=torch.float64) * 2 - 1
            y = torch.empty_like(x)
            p2 = hs.Plan2D(n0, n1, 1)
            p = None
            run = lambda: p2.exec_dev(x, y, batch)
            n = n0 * n1
            nbytes, flops = 32.0 * n * batch, 5.0 * n * math.log2(n) * batch
            kind = "fft2 " + p2.describe()[4:40]
        elif kind == "c2c":
            x = torch.rand((batch, n, 2), device=dev, dtype=torch.float64) * 2 - 1
            y = torch.empty_like(x)
            p = hs.Plan(n, 1)
            run = lambda: p.exec_dev(x, y, batch)
            nbytes, flops = 32.0 * n * batch, 5.0 * n * math.log2(n) * batch
        elif kind == "r2c":      # full mirrored spectrum, like the reference
            x = torch.rand((batch, n), device=dev, dtype=torch.float64) * 2 - 1
            y = torch.empty((batch, n, 2), device=dev, dtype=torch.float64)
            p = hs.RealPlan(n, 1)
            run = lambda: p.r2c_dev(x, y, n, 1, batch)
            nbytes, flops = 24.0 * n * batch, 2.5 * n * math.log2(n) * batch
        elif kind == "r2ch":     # half spectrum (bins 0..N/2)
            x = torch.rand((batch, n), device=dev, dtype=torch.float64) * 2 - 1
            y = torch.empty((batch, n // 2 + 1, 2), device=dev, dtype=torch.float64)
            p = hs.RealPlan(n, 1)
            run = lambda: p.r2c_dev(x, y, n // 2 + 1, 0, batch)
            nbytes, flops = (8.0 * n + 16.0 * (n // 2 + 1)) * batch, 2.5 * n * math.log2(n) * batch
        elif kind == "c2r":
            x = torch.rand((batch, n // 2 + 1, 2), device=dev, dtype=torch.float64) * 2 - 1
            y = torch.empty((batch, n), device=dev, dtype=torch.float64)
            p = hs.RealPlan(n, -1)
            run = lambda: p.c2r_dev(x, n // 2 + 1, y, batch)
            nbytes, flops = (8.0 * n + 16.0 * (n // 2 + 1)) * batch, 2.5 * n * math.log2(n) * batch
        elif kind == "conv":     # n-sample rows convolved pairwise, linear / full
            import ctypes as C
            x = torch.rand((batch, n), device=dev, dtype=torch.float64) * 2 - 1
            x2 = torch.rand((batch, n), device=dev, dtype=torch.float64) * 2 - 1
            y = torch.empty((batch, 2 * n), device=dev, dtype=torch.float64)
            L = hs.load()
            p = None
            run = lambda: L.fft_convolve_batch_dev(b"full", b"linear", C.c_void_p(x.data_ptr()), n, C.c_void_p(x2.data_ptr()), n,
                                                   C.c_void_p(y.data_ptr()), 2 * n, batch, None)
            P = 2 * n
            nbytes, flops = (16.0 * n + 8.0 * (2 * n - 1)) * batch, 3 * 2.5 * P * math.log2(P) * batch
        elif kind == "convs":    # overlap-save: one signal of n samples, filter of `batch` taps ("convs:<Lx>:<Lh>")
            import ctypes as C
            lh = batch
            x = torch.rand((n,), device=dev, dtype=torch.float64) * 2 - 1
            x2 = torch.rand((lh,), device=dev, dtype=torch.float64) * 2 - 1
            y = torch.empty((n + lh,), device=dev, dtype=torch.float64)
            L = hs.load()
            p = None
            run = lambda: L.fft_convolve_stream_dev(C.c_void_p(x.data_ptr()), n, C.c_void_p(x2.data_ptr()), lh, C.c_void_p(y.data_ptr()), 0, None)
            Pw = 1 << 15
            while Pw < 8 * lh and Pw < (1 << 20):
                Pw *= 2
            nblk = (n + lh - 1 + (Pw - lh)) // (Pw - lh + 1)
            nbytes, flops = 16.0 * n, 2 * 2.5 * Pw * math.log2(Pw) * nblk
        desc = p.describe() if kind == "c2c" else kind
        for _ in range(3):
            run()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            run()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        gbs = nbytes / (ms * 1e-3) / 1e9
        gfl = flops / (ms * 1e-3) / 1e9
        print("%-4s N=%-8d batch=%-7d %8.3f ms  %8.1f GB/s  %5.1f%% of %.0f  %9.1f GFLOP/s  [%s]" %
              (kind, n, batch, ms, gbs, 100 * gbs / peak, peak, gfl, desc), flush=True)
        if kernels:
            with hs.launch_timing() as t:
                run()
                torch.cuda.synchronize()
            for tag, cnt, kms in t.kernels:
                print("       %-40s x%-3d %8.3f ms" % (tag, cnt, kms), flush=True)
        if p is not None:
            p.close()
        del x, y
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
