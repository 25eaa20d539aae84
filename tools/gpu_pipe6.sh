#!/bin/bash
mkdir -p gpurun_out
export HSFFT_PIPE=1
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "mixed_radix or bluestein or real or conv" > gpurun_out/pytest_pipe6.log 2>&1; echo "pytest exit $?"; tail -4 gpurun_out/pytest_pipe6.log
CASES="65536:4096 262144:1024 524288:512 1048576:256 r2c:1048576:256 c2r:1048576:256 conv:524288:64 40009:4096"
echo "== PIPE"; timeout 600 python tools/sweep.py $CASES 2>&1 | tee gpurun_out/sweep_pipe6.log
python tools/sweep.py 1048576:64 --steps 2 > gpurun_out/plain_p20c.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:hs_pipe -s 3 -c 1 -f -o gpurun_out/prof_p20c python tools/sweep.py 1048576:64 --steps 2 > gpurun_out/ncu_p20c.log 2>&1; tail -1 gpurun_out/ncu_p20c.log
