#!/bin/bash
# plan-time two-pass kernels: columns per tile of the two passes
mkdir -p gpurun_out
CASES="10000:16384 30000:4096 49152:4096 100000:2048 196608:1024 300000:512 11264:16384 24576:8192"
echo "== default"; HSFFT_VERBOSE=1 timeout 600 python tools/sweep.py --kernels --steps 10 $CASES 2>&1 | grep -v "one-shot\|LOGPAD"
for c2 in 8 16; do echo "== C2=$c2"; HSFFT_JIT_C2=$c2 timeout 600 python tools/sweep.py --kernels --steps 10 $CASES 2>&1; done
for c1 in 8 2; do echo "== C1=$c1"; HSFFT_JIT_C1=$c1 timeout 600 python tools/sweep.py --kernels --steps 10 $CASES 2>&1; done
