#!/bin/bash
# ncu --set full captures of the round's main kernels (one call: all ncu runs count as one).  The reports are summarised
# on the box (tools/ncu_summary.py, tools/ncu_source.py) and deleted: only the text comes back (64 MiB limit).
mkdir -p gpurun_out
cap() { # name sweep-case kernel-regex skip count [env...]
  name=$1; case=$2; rx=$3; skip=$4; cnt=$5; shift 5
  env "$@" python tools/sweep.py $case --steps 2 > gpurun_out/plain_$name.log 2>&1 && \
  env "$@" ncu --set full --clock-control none --import-source on -k regex:$rx -s $skip -c $cnt -f -o /tmp/prof_$name python tools/sweep.py $case --steps 2 > gpurun_out/ncu_$name.log 2>&1
  echo "ncu $name exit $?"
  (python tools/ncu_summary.py /tmp/prof_$name.ncu-rep; for k in $(seq 0 $((cnt-1))); do echo; python tools/ncu_source.py /tmp/prof_$name.ncu-rep 16 $k; done) > gpurun_out/summary_$name.txt 2>/dev/null
  rm -f /tmp/prof_$name.ncu-rep
}
cap r1b_c2c4096 4096:65536 hs_c2c_smem 3 1
cap r1b_c2c3000 3000:65536 hs_c2c_smem 3 1
cap r1b_blu1009 1009:16384 hs_bluestein 4 1
cap r1b_blu10007 10007:4096 "hs_tiled|hs_blu_mid" 10 3
cap r1b_r2c1m r2c:1048576:64 "hs_tiled|hs_r2c_post" 9 3
cap r1b_c2r1m c2r:1048576:64 "hs_tiled" 6 2
cap r1b_jit1000 1000:65536 hs_jit_c2c 3 1
ls -la gpurun_out/summary_* | head -20
