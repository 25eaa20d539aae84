#!/bin/bash
mkdir -p gpurun_out
S="timeout 200 python tools/sweep.py --steps 10 8191:16384 4099:16384 4093:16384 2053:32768"
echo "== default (auto-sort pair for M' <= 8192)"; $S
echo "== in-place + TMEM"; HSFFT_SPLIT_FAMILY=6 $S
echo "== in-place, L2 parking"; HSFFT_SPLIT_FAMILY=6 HSFFT_SPLIT_TMEM=0 $S
cat > /tmp/chk.py <<'PY'
import importlib, sys, os
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import oracle_lib as ol
hs = importlib.import_module("mixed-radix-fast-fourier-transform_b200")
orc = ol.Oracle()
for n in (8191, 4099, 4093, 2053):
    for sgn in (1, -1):
        x = ol.rand_cplx(5, n, seed=n)
        with hs.Plan(n, sgn) as p:
            y = p.exec(x); d = p.describe()
        r = orc.c2c(x, sgn)
        e = max(ol.rel_rms(y[i], r[i]) for i in range(5))
        print(n, sgn, "err %.3e" % e, "OK" if e <= ol.tol(n) else "FAIL", d[-60:], flush=True)
PY
HSFFT_SPLIT_FAMILY=6 timeout 120 python /tmp/chk.py
