#!/bin/bash
# one ncu --set full capture, summarised on the box: gpu_ncu_one.sh name sweep-case kernel-regex skip count [ENV=VAL...]
mkdir -p gpurun_out
name=$1; case=$2; rx=$3; skip=$4; cnt=$5; shift 5
env "$@" python tools/sweep.py $case --steps 2 > gpurun_out/plain_$name.log 2>&1 && \
env "$@" ncu --set full --clock-control none --import-source on -k regex:$rx -s $skip -c $cnt -f -o /tmp/prof_$name python tools/sweep.py $case --steps 2 > gpurun_out/ncu_$name.log 2>&1
echo "ncu $name exit $?"
(python tools/ncu_summary.py /tmp/prof_$name.ncu-rep; for k in $(seq 0 $((cnt-1))); do echo; python tools/ncu_source.py /tmp/prof_$name.ncu-rep 24 $k; done) > gpurun_out/summary_$name.txt 2>/dev/null
rm -f /tmp/prof_$name.ncu-rep
head -60 gpurun_out/summary_$name.txt
