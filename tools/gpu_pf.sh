#!/bin/bash
mkdir -p gpurun_out
CASES="3000:65536 1009:16384 4096:65536 1024:262144 2048:131072 512:524288 256:1048576 8192:32768"
for pf in 1 0 2; do echo "== HSFFT_PREFETCH=$pf"; HSFFT_PREFETCH=$pf timeout 300 python tools/sweep.py $CASES 2>&1 | tee gpurun_out/sweep_pf$pf.log; done
