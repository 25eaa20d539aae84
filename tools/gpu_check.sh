#!/bin/bash
# full GPU parity suite + the round's sweep of every BASELINE config
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/pytest_gpu.log
CASES="4096:65536 3000:65536 1024:262144 2048:131072 512:524288 256:1048576 8192:32768 1009:16384 10007:16384 16384:16384 65536:4096 262144:1024 524288:512 1048576:256 r2c:1048576:256 r2ch:1048576:256 c2r:1048576:256 conv:524288:64"
timeout 900 python tools/sweep.py $CASES 2>&1 | tee gpurun_out/sweep_all.log
