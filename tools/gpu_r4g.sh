#!/bin/bash
# plan-time kernels: team sizes rounded to whole warps / multiples of 8 against the exact-pass choice of the fat-thread rule
mkdir -p gpurun_out
exec > gpurun_out/r4g.log 2>&1
run() { # N batch T MINB   (T = - : the rule)
  local e=""; [ "$3" != "-" ] && e="HSFFT_JIT_FUSE=25 HSFFT_JIT_T=$3 HSFFT_JIT_MINB=$4"
  echo "== N=$1 $e"
  env $e HSFFT_VERBOSE=1 timeout 120 python tools/sweep.py --steps 10 $1:$2 2>&1 | grep -v "LOGPAD cand\|^hsfft: pers\|one-shot" | cut -c1-170
}
run 3125 16384 - ; run 3125 16384 128 3
run 2500 16384 - ; run 2500 16384 128 3; run 2500 16384 104 3
run 2400 32768 - ; run 2400 32768 104 3; run 2400 32768 128 3; run 2400 32768 96 3
run 2000 32768 - ; run 2000 32768 96 3;  run 2000 32768 128 3
run 5000 16384 - ; run 5000 16384 224 1; run 5000 16384 208 1; run 5000 16384 256 1
run 6250 8192 - ;  run 6250 8192 256 1
run 4500 16384 - ; run 4500 16384 256 1; run 4500 16384 232 1
run 6000 16384 - ; run 6000 16384 256 1
run 4000 16384 - ; run 4000 16384 208 1; run 4000 16384 224 1
