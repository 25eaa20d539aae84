#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "mixed_radix or bluestein or real or conv" > gpurun_out/pytest_pipe3.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/pytest_pipe3.log
CASES="16384:16384 32768:8192 65536:4096 131072:2048 262144:1024 524288:512 1048576:256 r2c:1048576:256 c2r:1048576:256 10007:16384"
for s in 0 1; do echo "== SHAPE=$s"; HSFFT_PIPE_SHAPE=$s timeout 600 python tools/sweep.py $CASES 2>&1 | tee gpurun_out/sweep_pipe3_shape$s.log; done
