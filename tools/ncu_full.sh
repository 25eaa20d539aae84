#!/bin/bash
# one full ncu capture of the top kernel for a workload: tools/ncu_full.sh <name> <N:batch> [kernel regex]
name=$1; case=$2; rx=${3:-hs_}
mkdir -p gpurun_out
python tools/sweep.py $case --steps 2 > gpurun_out/plain_$name.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:$rx -s 3 -c 1 -f -o gpurun_out/prof_$name python tools/sweep.py $case --steps 2 > gpurun_out/ncu_$name.log 2>&1
echo "ncu $name exit $?"; tail -3 gpurun_out/ncu_$name.log
