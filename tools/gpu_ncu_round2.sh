#!/bin/bash
# ncu --set full of the tiled passes in their final configuration (512-point tiles: plain kernel, 3 CTAs/SM; 1024-point: TMA first pass, plain second)
bash tools/gpu_ncu_one.sh r1c_t18 262144:256 "hs_tiled" 6 2 > /dev/null 2>&1
bash tools/gpu_ncu_one.sh r1c_c2r1m c2r:1048576:64 "hs_tiled" 6 2 > /dev/null 2>&1
bash tools/gpu_ncu_one.sh r1c_t20 1048576:64 "hs_tiled" 6 2 > /dev/null 2>&1
for n in r1c_t18 r1c_c2r1m r1c_t20; do echo "#### $n"; grep -n "kernel:\|gpu__time_duration\|dram_throughput\|fp64.avg\|l1tex__throughput\|warps_active\|registers_per_thread\|dram__bytes" gpurun_out/summary_$n.txt | cut -c1-190; done
