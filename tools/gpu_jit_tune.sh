#!/bin/bash
mkdir -p gpurun_out
run() { echo "== $*"; env "$@" HSFFT_VERBOSE=1 python tools/sweep.py $CASES 2>&1 | grep -v "^hsfft: pers" | paste - - | sed -E 's/hsfft: plan-time kernel //; s/ of 6545.*//' | cut -c1-200; }
CASES="2000:65536" 
run A=1
run HSFFT_JIT_MINB=3
run HSFFT_JIT_FUSE=1
run HSFFT_JIT_FUSE=1 HSFFT_JIT_MINB=4
run HSFFT_JIT_FUSE=25
CASES="6000:16384"
run A=1
run HSFFT_JIT_T=750 HSFFT_JIT_MINB=1
run HSFFT_JIT_MINB=2
run HSFFT_JIT_FUSE=1
CASES="5000:16384"
run A=1
run HSFFT_JIT_MINB=2
run HSFFT_JIT_T=312
CASES="360:262144"
run A=1
run HSFFT_JIT_MINB=3
run HSFFT_JIT_THREADS=128
run HSFFT_JIT_THREADS=512
run HSFFT_JIT_FUSE=1 HSFFT_JIT_MINB=4
CASES="1000:65536"
run A=1
run HSFFT_JIT_MINB=4
run HSFFT_JIT_THREADS=512
run HSFFT_JIT_THREADS=128 HSFFT_JIT_MINB=6
run HSFFT_JIT_FUSE=25
