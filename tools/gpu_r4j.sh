#!/bin/bash
# whole-warp refinement of the fat-thread rule for three-CTA plans: parity tests of the plan-time kernels, then fat / thin sweeps
mkdir -p gpurun_out
exec > gpurun_out/r4j.log 2>&1
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "plan_time_compiled_kernels or real_r2c_c2r" 2>&1 | tail -3
NS="1800 1920 2000 2250 2304 2400 2592 2880"
CASES=""; for n in $NS; do CASES="$CASES $n:$((67108864 / n))"; done
echo "== fat (default)"; timeout 600 python tools/sweep.py --steps 10 $CASES r2c:4800:16384 c2r:4800:16384 r2c:4000:16384 c2r:4000:16384 2>&1 | cut -c1-170
echo "== thin (HSFFT_JIT_FAT=0)"; HSFFT_JIT_FAT=0 timeout 600 python tools/sweep.py --steps 10 $CASES r2c:4000:16384 c2r:4000:16384 2>&1 | cut -c1-170
