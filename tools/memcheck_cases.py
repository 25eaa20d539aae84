#!/usr/bin/env python
"""tools/memcheck_cases.py -- one small call of every kernel family, for `compute-sanitizer --tool memcheck`.

    compute-sanitizer --tool memcheck --error-exitcode 1 python tools/memcheck_cases.py

Small batches so that the instrumented run stays short; results are checked against numpy so that a wrong
answer is not mistaken for a clean run.
"""
import importlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    hs = importlib.import_module("mixed-radix-fast-fourier-transform_b200")
    dmod = importlib.import_module("mixed-radix-fast-fourier-transform_b200.dist")
    rng = np.random.default_rng(3)
    worst = 0.0

    def rel(a, b):
        return float(np.sqrt((np.abs(a - b) ** 2).sum() / (np.abs(b) ** 2).sum()))

    for n, batch in ((4096, 3), (3000, 3), (1024, 5), (8192, 2), (1000, 3), (96, 40), (1009, 3), (257, 5), (2039, 2),
                     (10007, 2), (4099, 2), (16384, 2), (1 << 17, 1), (1 << 19, 1), (24576, 1), (13 * 64, 3)):
        x = rng.uniform(-1, 1, (batch, n)) + 1j * rng.uniform(-1, 1, (batch, n))
        for sgn in (1, -1):
            with hs.Plan(n, sgn) as p:
                y = p.exec(x)
                d = p.describe()
            ref = np.fft.fft(x, axis=1) if sgn == 1 else np.fft.ifft(x, axis=1) * n
            e = rel(y, ref)
            worst = max(worst, e)
            print("c2c %7d %+d %.2e  %s" % (n, sgn, e, d[:70]), flush=True)
            assert e < 1e-9
    for n in (2048, 32768, 1 << 18):
        x = rng.uniform(-1, 1, (2, n))
        with hs.RealPlan(n, 1) as f, hs.RealPlan(n, -1) as g:
            X = f.r2c(x)
            back = g.c2r(X) / n
        e = max(rel(X, np.fft.fft(x, axis=1)), rel(back, x))
        print("real %7d %.2e" % (n, e), flush=True)
        assert e < 1e-9
    for la, lb in ((100, 37), (20001, 12999), (70001, 33333)):
        a, b = rng.uniform(-1, 1, (2, la)), rng.uniform(-1, 1, (2, lb))
        n, y = hs.convolve_batch(a, b)
        e = max(rel(y[i], np.convolve(a[i], b[i])) for i in range(2))
        print("conv %d x %d %.2e" % (la, lb, e), flush=True)
        assert e < 1e-9
    x, h = rng.uniform(-1, 1, 300001), rng.uniform(-1, 1, 1001)
    n, y = hs.convolve_stream(x, h, 0)
    e = rel(y, np.convolve(x, h))
    print("overlap-save %.2e" % e, flush=True)
    assert e < 1e-9
    dev = torch.device("cuda", 0)
    for N1, N2, world in ((256, 512, 2), (16384, 256, 4)):
        x = rng.uniform(-1, 1, N1 * N2) + 1j * rng.uniform(-1, 1, N1 * N2)
        X = dmod.local_group_forward(hs, x, N1, N2, 1, world, device=dev)
        e = rel(X, np.fft.fft(x))
        print("four-step fused exchange %dx%d / %d ranks %.2e" % (N1, N2, world, e), flush=True)
        assert e < 1e-9
    print("ALL OK, worst c2c error %.2e" % worst)


if __name__ == "__main__":
    main()
