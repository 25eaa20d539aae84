#!/bin/bash
# 1024-point second pass by pairs of CTAs (cluster of two, 128-byte store runs) against the plain tiled kernel
mkdir -p gpurun_out
cat > /tmp/chk.py <<'PY'
import importlib, sys, os
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import numpy as np
import oracle_lib as ol
hs = importlib.import_module("mixed-radix-fast-fourier-transform_b200")
orc = ol.Oracle()
n = 1 << 20
for full in (1,):
    rng = np.random.default_rng(n); x = rng.standard_normal((3, n))
    with hs.RealPlan(n, 1) as p:
        y = p.r2c(x)
    r = orc.r2c(x)
    e = max(ol.rel_rms(y[i], r[i]) for i in range(3))
    print("r2c", n, "err %.3e" % e, "OK" if e <= ol.tol(n) else "FAIL", "mirror exact", np.array_equal(y[:, n - 1: n // 2: -1], np.conj(y[:, 1: n // 2])), flush=True)
for sgn in (1, -1):
    z = ol.rand_cplx(3, n, seed=n)
    with hs.Plan(n, sgn) as p:
        y = p.exec(z)
    r = orc.c2c(z, sgn)
    e = max(ol.rel_rms(y[i], r[i]) for i in range(3))
    print("c2c", n, sgn, "err %.3e" % e, "OK" if e <= ol.tol(n) else "FAIL", flush=True)
PY
for v in 1 0; do echo "== parity pair=$v"; HSFFT_PAIR=$v timeout 300 python /tmp/chk.py; done
for v in 1 0; do echo "== pair=$v"; HSFFT_PAIR=$v timeout 300 python tools/sweep.py --kernels --steps 20 1048576:256 r2c:1048576:256 r2ch:1048576:256 conv:524288:64; done
