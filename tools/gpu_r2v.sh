#!/bin/bash
# real four-step with the TMA-staged first pass: which split per length
mkdir -p gpurun_out
export HSFFT_R4_TMA=1
for c in "r2c:262144:1024 512 1024 256" "r2c:65536:4096 256 512 128" "r2c:32768:8192 256 128" "r2c:131072:2048 1024 512" "r2c:524288:512 1024 512"; do
  set -- $c; w=$1; shift
  for n1 in "$@"; do echo "== $w n1=$n1"; HSFFT_R4_N1=$n1 timeout 200 python tools/sweep.py --steps 20 $w; done
done
