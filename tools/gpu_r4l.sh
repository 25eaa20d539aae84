#!/bin/bash
# plan-time kernels too long for the staged second buffer (N >= 6900): fat / thin rule, and N = 7000 at the team size the exploration found
mkdir -p gpurun_out
exec > gpurun_out/r4l.log 2>&1
C="7000:9586 7168:9362 7290:9205 7840:8559 8000:8388"
echo "== fat (default)"; timeout 200 python tools/sweep.py --steps 10 $C 2>&1 | cut -c1-150
echo "== thin (HSFFT_JIT_FAT=0)"; HSFFT_JIT_FAT=0 timeout 200 python tools/sweep.py --steps 10 $C 2>&1 | cut -c1-150
echo "== N=7000 FUSE=16 T=350"; HSFFT_JIT_FUSE=16 HSFFT_JIT_T=350 HSFFT_JIT_MINB=1 timeout 100 python tools/sweep.py --steps 10 7000:9586 2>&1 | cut -c1-150
