#!/bin/bash
mkdir -p gpurun_out
S="timeout 200 python tools/sweep.py --steps 10 10007:16384"
P=mixed-radix-fast-fourier-transform_b200
export HSFFT_SPLIT_TMEM=1
for v in u4 u2tw2 u2tw1 u2tw0; do echo "== $v"; HSFFT_LIB=$P/libhsfft_b200_$v.so $S; done
