#!/bin/bash
# c2r merge from TMA-staged tiles against the plain tiled kernel
mkdir -p gpurun_out
cat > /tmp/chk.py <<'PY'
import importlib, sys, os
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import numpy as np
import oracle_lib as ol
hs = importlib.import_module("mixed-radix-fast-fourier-transform_b200")
orc = ol.Oracle()
for n in (1 << 18, 1 << 19, 1 << 20):
    X = ol.rand_cplx(5, n, seed=n)
    with hs.RealPlan(n, -1) as p:
        y = p.c2r(X)
    r = orc.c2r(X, n)
    e = max(ol.rel_rms(y[i], r[i]) for i in range(5))
    print("c2r", n, "err %.3e" % e, "OK" if e <= ol.tol(n) else "FAIL", flush=True)
PY
for v in 1 0; do echo "== parity c2r_tma=$v"; HSFFT_C2R_TMA=$v timeout 300 python /tmp/chk.py; done
for v in 1 0; do echo "== c2r_tma=$v"; HSFFT_C2R_TMA=$v timeout 300 python tools/sweep.py --kernels --steps 20 c2r:1048576:256 c2r:524288:512 c2r:262144:1024; done
timeout 900 python -m pytest tests -m gpu -x -q -k "real or c2r or half" 2>&1 | tail -3
