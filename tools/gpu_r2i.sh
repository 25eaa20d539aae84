#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
echo "== pytest jit two-pass + tiled regressions"; HSFFT_VERBOSE=1 timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "two_pass or tiled or mixed_radix or real_r2c or four_step" > $O/r2i_pytest.log 2>&1; echo "pytest exit $?"; tail -4 $O/r2i_pytest.log; grep "plan-time passes" $O/r2i_pytest.log | sort -u | head -20
S="timeout 300 python tools/sweep.py --steps 10"
$S 10000:16384 12000:16384 49152:4096 50000:4096 100000:2048 196608:1024 1000000:256 6561:16384 2>&1 | tee $O/r2i_jit_twopass.txt
HSFFT_JIT=0 $S 10000:16384 49152:4096 100000:2048 2>&1 | tee $O/r2i_runtime_engine.txt
