#!/bin/bash
# round 2, call H: in-place tests, plan variants for N=8192 / N=3000
mkdir -p gpurun_out
O=gpurun_out
echo "== pytest in-place + split"; timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "in_place or two_streams" > $O/r2h_pytest.log 2>&1; echo "pytest exit $?"; tail -3 $O/r2h_pytest.log
S="timeout 300 python tools/sweep.py --steps 10"
echo "== 8192 default / C / D"; $S 8192:32768; HSFFT_PREFER_CT=$1 $S 8192:32768; HSFFT_PREFER_CT=$2 $S 8192:32768
echo "== 3000 default / MINB 3"; $S 3000:65536; HSFFT_PREFER_CT=$3 $S 3000:65536
echo "== r2c / c2r 16384 through the 8192 variants"; $S r2c:16384:16384 c2r:16384:16384
