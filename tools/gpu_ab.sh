#!/bin/bash
# A/B: default library vs a build variant (HSFFT_LIB) on the same cases
mkdir -p gpurun_out
CASES="${CASES:-16384:16384 65536:4096 131072:2048 262144:1024 524288:512 1048576:256 r2c:1048576:256 c2r:1048576:256 10007:16384 1009:16384}"
echo "== default"; timeout 600 python tools/sweep.py $CASES 2>&1 | tee gpurun_out/ab_default.log
for v in "$@"; do echo "== variant $v"; HSFFT_LIB=mixed-radix-fast-fourier-transform_b200/libhsfft_b200_$v.so timeout 600 python tools/sweep.py $CASES 2>&1 | tee gpurun_out/ab_$v.log; done
