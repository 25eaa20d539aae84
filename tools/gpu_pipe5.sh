#!/bin/bash
mkdir -p gpurun_out
export HSFFT_PIPE=1 HSFFT_VERBOSE=1
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "mixed_radix or bluestein or real or conv" > gpurun_out/pytest_pipe5.log 2>&1; echo "pytest exit $?"; tail -4 gpurun_out/pytest_pipe5.log
CASES="16384:16384 65536:4096 131072:2048 262144:1024 524288:512 1048576:256 r2c:1048576:256 c2r:1048576:256 conv:524288:64 10007:16384 40009:4096"
for s in 0 1; do echo "== PIPE SHAPE=$s"; HSFFT_PIPE_SHAPE=$s timeout 600 python tools/sweep.py $CASES 2>&1 | tee gpurun_out/sweep_pipe5_shape$s.log; done
echo "== PIPE SHAPE=0 POW2 blu"; HSFFT_BLU_POW2=1 timeout 300 python tools/sweep.py 10007:16384 2>&1 | tee gpurun_out/sweep_pipe5_blupow2.log
