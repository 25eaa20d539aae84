#!/bin/bash
# quick perf iteration: parity tests (fast subset) + sweep with/without carve-out
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "mixed_radix or bluestein or real" > gpurun_out/pytest_quick.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/pytest_quick.log
CASES="4096:65536 3000:65536 1024:262144 2048:131072 8192:32768 512:524288 256:1048576 1009:16384"
echo "== default"; python tools/sweep.py $CASES 2>&1 | tee gpurun_out/sweep_default.log
echo "== HSFFT_CARVEOUT=0"; HSFFT_CARVEOUT=0 python tools/sweep.py $CASES 2>&1 | tee gpurun_out/sweep_nocarve.log
echo "== variant B 4096"; HSFFT_PREFER_CT=1 python tools/sweep.py 4096:65536 2>&1 | tee gpurun_out/sweep_b.log
