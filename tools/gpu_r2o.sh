#!/bin/bash
mkdir -p gpurun_out
cat > /tmp/tmem_check.py <<'PY'
import importlib, sys, os
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import oracle_lib as ol
hs = importlib.import_module("mixed-radix-fast-fourier-transform_b200")
orc = ol.Oracle()
for n in (10007, 8209, 10239, 8193):
    for sgn in (1, -1):
        x = ol.rand_cplx(5, n, seed=n)
        with hs.Plan(n, sgn) as p:
            y = p.exec(x); d = p.describe()
        r = orc.c2c(x, sgn)
        e = max(ol.rel_rms(y[i], r[i]) for i in range(5))
        print(n, sgn, "err %.3e tol %.3e" % (e, ol.tol(n)), "OK" if e <= ol.tol(n) else "FAIL", flush=True)
PY
echo "== TMEM variant parity"; HSFFT_SPLIT_TMEM=1 timeout 120 python /tmp/tmem_check.py; echo "exit $?"
S="timeout 200 python tools/sweep.py --steps 10"
echo "== L2 parking"; HSFFT_SPLIT_TMEM=0 $S 10007:16384 8209:16384
echo "== TMEM parking"; HSFFT_SPLIT_TMEM=1 $S 10007:16384 8209:16384
