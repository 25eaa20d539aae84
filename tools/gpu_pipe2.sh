#!/bin/bash
mkdir -p gpurun_out
CASES="16384:16384 32768:8192 65536:4096 131072:2048 262144:1024 524288:512 1048576:256 r2c:1048576:256 c2r:1048576:256 10007:16384"
echo "== SHAPE=1"; HSFFT_PIPE_SHAPE=1 timeout 600 python tools/sweep.py $CASES 2>&1 | tee gpurun_out/sweep_pipe_shape1.log
for s in 0 1; do
HSFFT_PIPE_SHAPE=$s python tools/sweep.py 1048576:64 --steps 2 > gpurun_out/plain_pipe20_s$s.log 2>&1 && \
HSFFT_PIPE_SHAPE=$s ncu --set full --clock-control none --import-source on -k regex:hs_pipe -s 3 -c 1 -f -o gpurun_out/prof_pipe20_s$s python tools/sweep.py 1048576:64 --steps 2 > gpurun_out/ncu_pipe20_s$s.log 2>&1
echo "ncu s$s exit $?"; tail -2 gpurun_out/ncu_pipe20_s$s.log
done
