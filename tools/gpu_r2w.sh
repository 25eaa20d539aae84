#!/bin/bash
mkdir -p gpurun_out
export HSFFT_R4_TMA=1
for c in "r2c:1048576:256 1024" "r2c:524288:512 1024 512" "r2c:262144:1024 512 1024 256" "r2c:131072:2048 1024 512 256 128" "r2c:65536:4096 256 512 128"; do
  set -- $c; w=$1; shift
  for n1 in "$@"; do echo "== $w n1=$n1"; HSFFT_R4_N1=$n1 timeout 200 python tools/sweep.py --kernels --steps 10 $w; done
done
