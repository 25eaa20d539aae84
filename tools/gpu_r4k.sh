#!/bin/bash
# ncu --set full of two plan-time kernels chosen by the fat-thread rule (N = 2400: three CTAs per SM; N = 5000: persistent, staged input)
bash tools/gpu_ncu_one.sh jit2400_fat 2400:27962 hs_jit_c2c 3 1 > gpurun_out/ncu_one_2400.log 2>&1
bash tools/gpu_ncu_one.sh jit5000_fat 5000:13421 hs_jit_c2c 3 1 > gpurun_out/ncu_one_5000.log 2>&1
