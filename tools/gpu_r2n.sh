#!/bin/bash
mkdir -p gpurun_out
bash tools/gpu_ncu_one.sh r2n_split10007 10007:4096 conv_bluestein_split 1 1 > /dev/null 2>&1
sed -n '/^timeline/,$p' gpurun_out/summary_r2n_split10007.txt
