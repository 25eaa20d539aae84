#!/usr/bin/env python
"""Generate tests/golden/ref_vectors.npz from the REFERENCE'S OWN C code (oracle/_ref, built by
oracle/Makefile from /root/reference/src in place).  Run in the build container only:

    python tests/golden/make_golden.py

The GPU box has no /root/reference; the committed .npz lets `pytest -m gpu` and the CPU suite pin
the oracle and the CUDA path to outputs of the unmodified ("asis") and repaired reference.
Each entry stores the seeded input, the full output for small N and, for large N, 64 sampled bins
plus the sum of all bins (a checksum).  Which flavour is authoritative per size follows SURVEY.md 8c.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import oracle_lib as ol  # noqa: E402


def main():
    orc = ol.Oracle()
    asis, rep = ol.RefLib("asis"), ol.RefLib("repaired")
    out = {}
    meta = []
    # complex transforms: reference self-test sizes (src/real.c:278), BASELINE sizes, odd radices, Bluestein
    for n in [2, 3, 4, 5, 7, 8, 15, 16, 20, 64, 11, 13, 26, 143, 19, 59, 257, 1009, 1024, 3000, 4096, 10007]:
        flavour = "asis" if ol.asis_is_authoritative(n, orc) and n % 13 else "repaired"
        lib = asis if flavour == "asis" else rep
        x = ol.rand_cplx(1, n, seed=424242 + n)[0]
        for sgn in (1, -1):
            y = lib.c2c(x, sgn)
            key = "c2c_%d_%s" % (n, "f" if sgn == 1 else "i")
            if n <= 64:
                out[key] = y
            else:
                idx = np.linspace(0, n - 1, 64).astype(np.int64)
                out[key + "_idx"] = idx
                out[key] = y[idx]
                out[key + "_sum"] = np.array([y.sum()])
            meta.append("%s flavour=%s" % (key, flavour))
    # real transforms (repaired flavour: half lengths ending in a factor 2 hit reference bug 2)
    for n in [4, 8, 16, 32, 64, 2018, 4096]:
        x = ol.xorshift_uniform(n, seed=777 + n)
        X = rep.r2c(x)
        y = rep.c2r(X, n)
        if n <= 64:
            out["r2c_%d" % n], out["c2r_%d" % n] = X, y
        else:
            idx = np.linspace(0, n - 1, 64).astype(np.int64)
            out["r2c_%d_idx" % n] = idx
            out["r2c_%d" % n], out["c2r_%d" % n] = X[idx], y[idx]
            out["r2c_%d_sum" % n], out["c2r_%d_sum" % n] = np.array([X.sum()]), np.array([y.sum()])
    # convolution
    for la, lb in [(5, 3), (100, 37), (512, 512)]:
        a, b = ol.xorshift_uniform(la, seed=la), ol.xorshift_uniform(lb, seed=lb + 1)
        for t in ("full", "same", "valid"):
            n, y = rep.convolve(a, b, t, "linear")
            out["conv_%d_%d_%s" % (la, lb, t)] = y
    # plan fields the reference's tests read (test/test_mixedRadixFFT.cpp:103-123,134-189)
    for n in [15, 17, 64, 1024, 3000, 4096, 1009]:
        f = rep.plan_fields(n, 1)
        out["plan_%d_factors" % n] = np.array(f["factors"], dtype=np.int32)
        out["plan_%d_lt" % n] = np.array([f["lt"]], dtype=np.int32)
        if "twiddle" in f and n <= 4096:
            out["plan_%d_twiddle" % n] = f["twiddle"]
    out["meta"] = np.array(meta)
    path = os.path.join(HERE, "ref_vectors.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes,", len(out), "arrays")


if __name__ == "__main__":
    main()
