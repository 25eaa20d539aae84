#!/bin/bash
# plan-time kernels: the fat-thread rule (HSFFT_JIT_FAT=1, default) against the thin-thread rule (=0) over many lengths;
# real transforms that ride on the changed plans; correctness against torch.fft
mkdir -p gpurun_out
exec > gpurun_out/r4c.log 2>&1
NS="1536 1600 1800 2000 2400 2500 2560 2880 3072 3125 3200 3456 3600 3840 4000 4374 4500 4608 4800 5000 5120 5400 5625 6000 6144 6250 6400 6750 6912 7000 7200 7680 7776 8000"
CASES=""; for n in $NS; do CASES="$CASES $n:$((67108864 / n))"; done
REAL="r2c:6000:16384 c2r:6000:16384 r2c:16384:8192 c2r:16384:8192 r2c:10000:8192 c2r:10000:8192 r2c:4800:16384 c2r:4800:16384 r2c:12000:8192 c2r:12000:8192 r2c:5000:16384 c2r:5000:16384"
python - <<'PY'
import importlib, torch
hs = importlib.import_module("mixed-radix-fast-fourier-transform_b200")
bad = 0
for n in (1536, 1800, 2000, 2400, 2560, 2880, 3125, 3200, 3600, 4000, 4500, 5000, 5120, 6000, 6250, 6912, 7000, 7776, 8000):
    b = 700
    x = torch.rand((b, n, 2), device="cuda", dtype=torch.float64) * 2 - 1
    y = torch.empty_like(x)
    for sgn in (1, -1):
        p = hs.Plan(n, sgn); p.exec_dev(x, y, b); torch.cuda.synchronize()
        xc = torch.view_as_complex(x)
        r = torch.fft.fft(xc) if sgn == 1 else torch.fft.ifft(xc) * n
        e = ((torch.view_as_complex(y) - r).norm() / r.norm()).item()
        bad += e > 2e-11
        print("check c2c N=%d sgn %d rel err %.2e %s" % (n, sgn, e, p.describe()[-60:])); p.close()
for n in (4000, 4800, 5000, 6000, 10000, 12000, 16384):
    b = 500
    x = torch.rand((b, n), device="cuda", dtype=torch.float64) * 2 - 1
    y = torch.empty((b, n, 2), device="cuda", dtype=torch.float64)
    p = hs.RealPlan(n, 1); p.r2c_dev(x, y, n, 1, b); torch.cuda.synchronize()
    r = torch.fft.fft(x)
    e = ((torch.view_as_complex(y) - r).norm() / r.norm()).item(); bad += e > 2e-11
    h = torch.view_as_real(torch.fft.rfft(x)).contiguous(); z = torch.empty_like(x)
    q = hs.RealPlan(n, -1); q.c2r_dev(h, n // 2 + 1, z, b); torch.cuda.synchronize()
    e2 = ((z - x * n).norm() / (x * n).norm()).item(); bad += e2 > 2e-11
    print("check real N=%d r2c %.2e c2r %.2e" % (n, e, e2)); p.close(); q.close()
print("CHECK", "FAILED" if bad else "ok", bad)
PY
echo "== fat (default)"; timeout 900 python tools/sweep.py --steps 10 $CASES $REAL 2>&1 | cut -c1-200
echo "== thin (HSFFT_JIT_FAT=0)"; HSFFT_JIT_FAT=0 timeout 900 python tools/sweep.py --steps 10 $CASES r2c:10000:8192 c2r:10000:8192 r2c:4800:16384 c2r:4800:16384 r2c:12000:8192 c2r:12000:8192 r2c:5000:16384 c2r:5000:16384 2>&1 | cut -c1-200
