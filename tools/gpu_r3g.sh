#!/bin/bash
# plan-time one-kernel plans: persistent CTAs with the next input staged by cp.async.bulk (HSFFT_JIT_STAGED) vs one-shot CTAs
mkdir -p gpurun_out
cat > /tmp/chk.py <<'PY'
import importlib, sys, os
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import numpy as np
import oracle_lib as ol
hs = importlib.import_module("mixed-radix-fast-fourier-transform_b200")
orc = ol.Oracle()
for n in (6000, 6561, 5000, 3125, 2000):
    b = 1500 if n > 4000 else 2500
    for sgn in (1, -1):
        x = ol.rand_cplx(b, n, seed=n)
        with hs.Plan(n, sgn) as p:
            y = p.exec(x)
        idx = [0, 1, 147, 148, 149, b // 2, b - 2, b - 1]
        r = orc.c2c(x[idx], sgn)
        e = max(ol.rel_rms(y[i], r[k]) for k, i in enumerate(idx))
        print(n, sgn, "err %.3e" % e, "OK" if e <= ol.tol(n) else "FAIL", flush=True)
PY
for v in 2 1; do echo "== parity staged=$v"; HSFFT_JIT_STAGED=$v timeout 300 python /tmp/chk.py; done
CASES="2000:32768 2187:32768 2400:32768 2500:32768 3125:16384 3200:16384 3600:16384 4000:16384 4500:16384 5000:16384 5120:16384 5625:16384 6000:16384 6144:16384 6250:16384 6400:16384 6561:16384 7000:16384 7500:16384 8000:16384"
for v in 0 1 2; do echo "== staged=$v"; HSFFT_JIT_STAGED=$v timeout 600 python tools/sweep.py --steps 20 $CASES; done
