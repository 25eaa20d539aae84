#!/bin/bash
mkdir -p gpurun_out
CASES="r2c:32768:8192 r2c:65536:4096 r2c:131072:2048 r2c:262144:1024 r2c:524288:512 r2c:1048576:256 r2ch:1048576:256 conv:524288:64 conv:131072:256"
echo "== default"; timeout 300 python tools/sweep.py --kernels --steps 10 $CASES
echo "== mirrored pass: 8 columns for 1024 points, 16 for 512"; HSFFT_LIB=$PWD/mixed-radix-fast-fourier-transform_b200/libhsfft_b200_c9w.so timeout 300 python tools/sweep.py --kernels --steps 10 r2c:262144:1024 r2c:524288:512 r2c:1048576:256 r2ch:1048576:256
