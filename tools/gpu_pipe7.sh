#!/bin/bash
mkdir -p gpurun_out
export HSFFT_PIPE=1
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "mixed_radix or bluestein or real or conv" > gpurun_out/pytest_pipe7.log 2>&1; echo "pytest exit $?"; tail -2 gpurun_out/pytest_pipe7.log
CASES="16384:16384 65536:4096 131072:2048 262144:1024 524288:512 1048576:256 r2c:1048576:256 c2r:1048576:256 conv:524288:64 40009:4096"
echo "== PIPE"; timeout 600 python tools/sweep.py $CASES 2>&1 | tee gpurun_out/sweep_pipe7.log
