#!/bin/bash
mkdir -p gpurun_out
run() { # name env kernel-regex
  name=$1; envs=$2; rx=$3
  env $envs python tools/sweep.py 1048576:64 --steps 2 > gpurun_out/plain_$name.log 2>&1 && \
  env $envs ncu --set full --clock-control none --import-source on -k regex:$rx -s 6 -c 2 -f -o gpurun_out/prof_$name python tools/sweep.py 1048576:64 --steps 2 > gpurun_out/ncu_$name.log 2>&1
  echo "ncu $name exit $?"; tail -1 gpurun_out/ncu_$name.log
}
run p20b "HSFFT_PIPE_SHAPE=0" hs_pipe
run t20c4 "HSFFT_PIPE=0 HSFFT_TMA=0" hs_tiled_kernel
run t20tma "HSFFT_PIPE=0" hs_tiled_tma
