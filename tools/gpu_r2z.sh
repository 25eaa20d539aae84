#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
CASES="32768:8192 65536:4096 r2c:32768:8192 r2c:65536:4096 r2c:131072:2048 r2c:262144:1024 r2c:524288:512 r2c:1048576:256 r2ch:1048576:256 conv:131072:256"
echo "== default"; timeout 300 python tools/sweep.py --kernels --steps 10 $CASES
echo "== 16 columns for the 256-point second pass"; HSFFT_LIB=$PWD/mixed-radix-fast-fourier-transform_b200/libhsfft_b200_c2w.so timeout 300 python tools/sweep.py --kernels --steps 10 32768:8192 65536:4096 r2c:262144:1024
echo "== same, r2c 2^18 as 1024 x 256"; HSFFT_R4_N1=1024 HSFFT_LIB=$PWD/mixed-radix-fast-fourier-transform_b200/libhsfft_b200_c2w.so timeout 300 python tools/sweep.py --kernels --steps 10 r2c:262144:1024
