#!/bin/bash
# after the partial-tile fix and the second-pass 512-point plan: two-pass sizes, real transforms, convolution
mkdir -p gpurun_out
export HSFFT_R4_TMA=2
CASES="131072:2048 262144:1024 524288:512 1048576:256 c2r:1048576:256 c2r:262144:1024 conv:524288:64 conv:131072:256 40009:4096"
echo "== default"; timeout 300 python tools/sweep.py --kernels --steps 10 $CASES
echo "== HSFFT_TMA=3 (512-point first passes staged)"; HSFFT_TMA=3 timeout 300 python tools/sweep.py --kernels --steps 10 $CASES
for c in "r2c:1048576:256 1024" "r2c:524288:512 1024 512" "r2c:262144:1024 512 1024 256" "r2c:131072:2048 1024 512 256" "r2c:65536:4096 256 512 128" "r2c:32768:8192 256 128"; do
  set -- $c; w=$1; shift
  for n1 in "$@"; do echo "== $w n1=$n1"; HSFFT_R4_N1=$n1 timeout 200 python tools/sweep.py --kernels --steps 10 $w; done
done
echo "== r2c without the second-pass plan, n1 = 1024 / 512"; HSFFT_NO_SUB2=1 HSFFT_R4_N1=1024 timeout 200 python tools/sweep.py --kernels --steps 10 r2c:524288:512
