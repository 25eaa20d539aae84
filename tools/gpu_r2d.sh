#!/bin/bash
# round 2, call D: the whole GPU test suite, smoke, conv after the four-step change
mkdir -p gpurun_out
O=gpurun_out
echo "== pytest gpu (all)"; timeout 2400 python -m pytest tests -m gpu -x -q > $O/r2d_pytest.log 2>&1; echo "pytest exit $?"; tail -5 $O/r2d_pytest.log
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -12 | tee $O/r2d_smoke.txt
S="timeout 300 python tools/sweep.py --steps 10"
$S conv:524288:64 conv:65536:512 r2c:131072:2048 r2c:32768:8192 convs:268435456:4097 2>&1 | tee $O/r2d_conv.txt
