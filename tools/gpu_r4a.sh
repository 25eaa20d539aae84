#!/bin/bash
# A/B: fewer, fatter threads for the one-kernel plans (N = 3000 at 125 threads, N = 8192 in three passes at 256 threads)
mkdir -p gpurun_out
exec > gpurun_out/r4a.log 2>&1
chk() { python - "$1" <<'PY'
import importlib, sys, torch
hs = importlib.import_module("mixed-radix-fast-fourier-transform_b200")
n = int(sys.argv[1]); b = 300
x = torch.rand((b, n, 2), device="cuda", dtype=torch.float64) * 2 - 1
y = torch.empty_like(x)
for sgn in (1, -1):
    p = hs.Plan(n, sgn); p.exec_dev(x, y, b); torch.cuda.synchronize()
    xc = torch.view_as_complex(x)
    r = torch.fft.fft(xc) if sgn == 1 else torch.fft.ifft(xc) * n
    e = (torch.view_as_complex(y) - r).norm() / r.norm()
    print("   check", p.describe(), "sgn", sgn, "rel err vs torch.fft %.3e" % e.item()); p.close()
PY
}
for rep in 1 2; do
echo "== N=3000 default";           python tools/sweep.py --steps 20 3000:65536
echo "== N=3000 T=125, 3 CTAs/SM";  HSFFT_PREFER_CT=46 python tools/sweep.py --steps 20 3000:65536
echo "== N=3000 T=125, 4 CTAs/SM";  HSFFT_PREFER_CT=47 python tools/sweep.py --steps 20 3000:65536
echo "== N=8192 default";           python tools/sweep.py --steps 20 8192:32768
echo "== N=8192 {4,8}{16}{16}";     HSFFT_PREFER_CT=48 python tools/sweep.py --steps 20 8192:32768
echo "== N=8192 {16}{16}{8,4}";     HSFFT_PREFER_CT=49 python tools/sweep.py --steps 20 8192:32768
done
HSFFT_PREFER_CT=46 chk 3000; HSFFT_PREFER_CT=47 chk 3000; HSFFT_PREFER_CT=48 chk 8192; HSFFT_PREFER_CT=49 chk 8192
