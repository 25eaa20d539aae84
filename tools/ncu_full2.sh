#!/bin/bash
# tools/ncu_full2.sh <name> <sweep-case> <kernel regex> <skip> <count>
name=$1; case=$2; rx=$3; skip=${4:-6}; cnt=${5:-2}
mkdir -p gpurun_out
python tools/sweep.py $case --steps 2 > gpurun_out/plain_$name.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:$rx -s $skip -c $cnt -f -o gpurun_out/prof_$name python tools/sweep.py $case --steps 2 > gpurun_out/ncu_$name.log 2>&1
echo "ncu $name exit $?"; tail -2 gpurun_out/ncu_$name.log
