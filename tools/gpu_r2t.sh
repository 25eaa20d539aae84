#!/bin/bash
mkdir -p gpurun_out
cat > /tmp/chk.py <<'PY'
import importlib, sys, os
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import oracle_lib as ol
hs = importlib.import_module("mixed-radix-fast-fourier-transform_b200")
orc = ol.Oracle()
for n, b in ((1009, 1000), (257, 777)):
    for sgn in (1, -1):
        x = ol.rand_cplx(b, n, seed=n)
        with hs.Plan(n, sgn) as p:
            y = p.exec(x); d = p.describe()
        r = orc.c2c(x, sgn)
        e = max(ol.rel_rms(y[i], r[i]) for i in range(b))
        print(n, sgn, "err %.3e" % e, "OK" if e <= ol.tol(n) else "FAIL", flush=True)
PY
for p in 4 5 6; do echo "== parity pipe=$p"; HSFFT_RADER_PIPE=$p timeout 120 python /tmp/chk.py; done
S="timeout 200 python tools/sweep.py --steps 20 1009:16384 1009:65536 257:65536"
echo "== one-shot kernel"; $S
for p in 4 5 6; do echo "== pipe, $p CTAs per SM"; HSFFT_RADER_PIPE=$p $S; done
