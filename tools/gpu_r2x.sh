#!/bin/bash
# sub-plan variants (1024 points with 128 threads, 512 points with 32 threads) against the default, real four-step with the staged first pass
mkdir -p gpurun_out
cat > /tmp/chk.py <<'PY'
import importlib, sys, os
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import numpy as np
import oracle_lib as ol
hs = importlib.import_module("mixed-radix-fast-fourier-transform_b200")
orc = ol.Oracle()
for n in (1 << 19, 1 << 20, 1 << 18):
    rng = np.random.default_rng(n); x = rng.standard_normal((2, n))
    with hs.RealPlan(n, 1) as p:
        y = p.r2c(x)
    r = orc.r2c(x)
    e = max(ol.rel_rms(y[i], r[i]) for i in range(2))
    print("r2c", n, "err %.3e" % e, "OK" if e <= ol.tol(n) else "FAIL", flush=True)
    for sgn in (1, -1):
        z = ol.rand_cplx(2, n, seed=n)
        with hs.Plan(n, sgn) as p:
            y = p.exec(z)
        r = orc.c2c(z, sgn)
        e = max(ol.rel_rms(y[i], r[i]) for i in range(2))
        print("c2c", n, sgn, "err %.3e" % e, "OK" if e <= ol.tol(n) else "FAIL", flush=True)
PY
CASES="262144:1024 524288:512 1048576:256 r2c:1048576:256 r2c:524288:512 r2c:262144:1024 c2r:1048576:256"
for v in "" _s1024t128 _s512t32; do
  export HSFFT_LIB=$PWD/mixed-radix-fast-fourier-transform_b200/libhsfft_b200$v.so
  echo "== lib '$v' parity"; HSFFT_R4_TMA=2 timeout 300 python /tmp/chk.py
  echo "== lib '$v' r4tma=1"; HSFFT_R4_TMA=1 timeout 300 python tools/sweep.py --kernels --steps 10 $CASES
  echo "== lib '$v' r4tma=2 (512-point first pass staged too)"; HSFFT_R4_TMA=2 timeout 300 python tools/sweep.py --kernels --steps 10 r2c:524288:512 r2c:262144:1024 r2c:65536:4096
  echo "== lib '$v' r4tma=2 n1=512"; HSFFT_R4_N1=512 HSFFT_R4_TMA=2 timeout 300 python tools/sweep.py --kernels --steps 10 r2c:524288:512 r2c:262144:1024 r2c:65536:4096
  echo "== lib '$v' all passes staged (HSFFT_TMA=1)"; HSFFT_TMA=1 HSFFT_R4_TMA=1 timeout 300 python tools/sweep.py --kernels --steps 10 262144:1024 524288:512 1048576:256
done
