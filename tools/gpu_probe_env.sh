#!/bin/bash
# env-only probe: batch chunking of the long Bluestein (workspace footprint)
mkdir -p gpurun_out; rm -f gpurun_out/probe2.log
run() { echo "== $*" | tee -a gpurun_out/probe2.log; env "$@" 2>&1 | cut -c1-120 | tee -a gpurun_out/probe2.log; }
for mb in 350 700 1400 2800 5600 1000000; do run HSFFT_BLU_CHUNK_MB=$mb python tools/sweep.py 10007:16384 40009:4096 --steps 10; done
run A=1 python tools/sweep.py 10007:2048 10007:4096 10007:8192 --steps 10
