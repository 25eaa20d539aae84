#!/bin/bash
# persistent-pipeline bring-up: parity subset, then sweeps with the pipeline on / off and a few ring shapes
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/gpu.txt
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/pytest_pipe.log 2>&1; echo "pytest exit $?"; tail -5 gpurun_out/pytest_pipe.log
CASES="16384:16384 32768:8192 65536:4096 131072:2048 262144:1024 524288:512 1048576:256 r2c:1048576:256 c2r:1048576:256 conv:524288:64 10007:16384 40009:4096"
echo "== pipeline (default)"; timeout 600 python tools/sweep.py $CASES 2>&1 | tee gpurun_out/sweep_pipe.log
echo "== HSFFT_PIPE=0"; HSFFT_PIPE=0 timeout 600 python tools/sweep.py $CASES 2>&1 | tee gpurun_out/sweep_nopipe.log
SHORT="65536:4096 1048576:256 r2c:1048576:256 10007:16384"
for kb in 4096 8192 32768; do echo "== SLOT_KB=$kb"; HSFFT_PIPE_SLOT_KB=$kb timeout 300 python tools/sweep.py $SHORT 2>&1 | tee gpurun_out/sweep_pipe_kb$kb.log; done
echo "== NSLOT=4"; HSFFT_PIPE_NSLOT=4 timeout 300 python tools/sweep.py $SHORT 2>&1 | tee gpurun_out/sweep_pipe_ns4.log
echo "== NSLOT=2"; HSFFT_PIPE_NSLOT=2 timeout 300 python tools/sweep.py $SHORT 2>&1 | tee gpurun_out/sweep_pipe_ns2.log
echo "== LAG=2"; HSFFT_PIPE_LAG=2 timeout 300 python tools/sweep.py $SHORT 2>&1 | tee gpurun_out/sweep_pipe_lag2.log
echo "== CTAS=1"; HSFFT_PIPE_CTAS=1 timeout 300 python tools/sweep.py $SHORT 2>&1 | tee gpurun_out/sweep_pipe_ctas1.log
