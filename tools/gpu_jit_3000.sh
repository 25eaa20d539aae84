#!/bin/bash
run() { echo "== $*"; env "$@" HSFFT_NO_CT=1 HSFFT_VERBOSE=1 python tools/sweep.py $CASES 2>&1 | grep -v "^hsfft: pers" | paste - - | sed -E 's/hsfft: plan-time kernel //; s/ of 6545.*//' | cut -c1-200; }
CASES="3000:65536"
run A=1
run HSFFT_JIT_FUSE=1
run HSFFT_JIT_FUSE=1 HSFFT_JIT_MINB=3
run HSFFT_JIT_FUSE=25
run HSFFT_JIT_FUSE=25 HSFFT_JIT_T=120
CASES="4096:65536"
run A=1
CASES="8192:32768"
run A=1
run HSFFT_JIT_T=1024 HSFFT_JIT_MINB=1
CASES="1024:262144"
run A=1
CASES="2048:131072"
run A=1
CASES="512:524288"
run A=1
