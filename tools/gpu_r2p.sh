#!/bin/bash
mkdir -p gpurun_out
export HSFFT_SPLIT_TMEM=1
bash tools/gpu_ncu_one.sh r2p_split10007_tmem 10007:4096 split_tmem 1 1 > /dev/null 2>&1
grep -n "gpu__time_duration\|dram__bytes\|fp64_cycles\|lts__t_sector_hit\|lts__throughput\|l1tex__throughput\|warps_active\|issue_active" gpurun_out/summary_r2p_split10007_tmem.txt | head -12
sed -n '/^by reason/,/^top sites/p' gpurun_out/summary_r2p_split10007_tmem.txt | cut -c1-250
sed -n '/^timeline/,$p' gpurun_out/summary_r2p_split10007_tmem.txt
