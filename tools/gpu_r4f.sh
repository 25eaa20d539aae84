#!/bin/bash
# A/B: N = 3000 team sizes 125 (default) / 128 (whole warps) / 120 (25-point pass exact)
mkdir -p gpurun_out
exec > gpurun_out/r4f.log 2>&1
for rep in 1 2; do
echo "== N=3000 T=125 (default)"; python tools/sweep.py --steps 20 3000:65536
echo "== N=3000 T=128";           HSFFT_PREFER_CT=47 python tools/sweep.py --steps 20 3000:65536
echo "== N=3000 T=120";           HSFFT_PREFER_CT=50 python tools/sweep.py --steps 20 3000:65536
done
