#!/bin/bash
# plan-time kernels with fewer, fatter threads: HSFFT_JIT_FUSE / _T / _MINB sweeps per length
mkdir -p gpurun_out
exec > gpurun_out/r4b.log 2>&1
run() { # N batch FUSE T MINB
  local e=""; [ "$3" != "-" ] && e="HSFFT_JIT_FUSE=$3 HSFFT_JIT_T=$4 HSFFT_JIT_MINB=$5"
  echo "== N=$1 $e"
  env $e HSFFT_VERBOSE=1 timeout 120 python tools/sweep.py --steps 10 $1:$2 2>&1 | grep -v "LOGPAD cand\|^hsfft: pers" | cut -c1-230
}
run 5000 16384 - ; run 5000 16384 25 200 1; run 5000 16384 25 250 1; run 5000 16384 25 209 1
run 6000 16384 - ; run 6000 16384 25 250 1; run 6000 16384 25 240 1; run 6000 16384 16 250 1
run 4000 16384 - ; run 4000 16384 25 200 1; run 4000 16384 16 167 1
run 6250 8192 - ;  run 6250 8192 25 250 1;  run 6250 8192 25 313 1
run 8000 8192 - ;  run 8000 8192 25 334 1
run 2400 32768 - ; run 2400 32768 25 100 3; run 2400 32768 25 100 2
run 3125 16384 - ; run 3125 16384 25 125 3; run 3125 16384 25 125 2
run 4500 16384 - ; run 4500 16384 25 250 1; run 4500 16384 25 225 1
run 6561 8192 - ;  run 6561 8192 16 243 1;  run 6561 8192 16 365 1
run 7000 8192 - ;  run 7000 8192 25 350 1;  run 7000 8192 16 350 1
run 5120 16384 - ; run 5120 16384 16 171 1; run 5120 16384 16 320 1
run 6400 8192 - ;  run 6400 8192 25 320 1;  run 6400 8192 16 267 1
run 2000 65536 - ; run 2000 65536 25 80 3;  run 2000 65536 25 100 3
run 1000 65536 - ; run 1000 65536 25 40 6
