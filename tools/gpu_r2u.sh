#!/bin/bash
# real four-step: first pass on the TMA-staged kernel (HSFFT_R4_TMA=1) against the plain tiled kernel
mkdir -p gpurun_out
cat > /tmp/chk.py <<'PY'
import importlib, sys, os
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import numpy as np
import oracle_lib as ol
hs = importlib.import_module("mixed-radix-fast-fourier-transform_b200")
orc = ol.Oracle()
for n in (1 << 17, 1 << 18, 1 << 20, 1 << 19):
    rng = np.random.default_rng(n); x = rng.standard_normal((3, n))
    with hs.RealPlan(n, 1) as p:
        y = p.r2c(x)
    r = orc.r2c(x)
    e = max(ol.rel_rms(y[i], r[i]) for i in range(3))
    print(n, "err %.3e" % e, "OK" if e <= ol.tol(n) else "FAIL", flush=True)
PY
for v in 0 1; do echo "== parity r4tma=$v"; HSFFT_R4_TMA=$v timeout 200 python /tmp/chk.py; done
S="timeout 300 python tools/sweep.py --steps 20 r2c:1048576:256 r2c:524288:512 r2c:262144:1024 r2c:131072:2048"
for v in 0 1; do echo "== r4tma=$v"; HSFFT_R4_TMA=$v $S; HSFFT_R4_TMA=$v HSFFT_R4_SWAP=1 $S; done
