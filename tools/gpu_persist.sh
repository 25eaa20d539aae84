#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/pytest_persist.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/pytest_persist.log
CASES="3000:65536 1009:16384 600:32768 4096:65536"
echo "== persist"; timeout 300 python tools/sweep.py $CASES 2>&1 | tee gpurun_out/sweep_persist1.log
echo "== HSFFT_PERSIST=0"; HSFFT_PERSIST=0 timeout 300 python tools/sweep.py $CASES 2>&1 | tee gpurun_out/sweep_persist0.log
