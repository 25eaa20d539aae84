#!/bin/bash
# final checkpoint of the round: smoke, full GPU parity suite, bench lines, launch list (gpu_final2.sh), ncu capture of the new
# N = 3000 kernel, sweep of every configuration
bash tools/gpu_final2.sh > gpurun_out/final6.log 2>&1
bash tools/gpu_ncu_one.sh c2c3000_t125 3000:65536 hs_c2c_smem 3 1 > gpurun_out/ncu_one_3000.log 2>&1
bash tools/gpu_ncu_one.sh c2c8192_3pass 8192:32768 hs_c2c_smem 3 1 > gpurun_out/ncu_one_8192.log 2>&1
CASES="4096:65536 3000:65536 1024:262144 2048:131072 512:524288 256:1048576 8192:32768 1009:16384 10007:16384 40009:4096 4093:16384 16384:16384 32768:8192 65536:4096 131072:2048 262144:1024 524288:512 1048576:256 r2c:1048576:256 r2ch:1048576:256 c2r:1048576:256 conv:524288:64 r2c:8192:32768 c2r:8192:32768 r2c:2048:131072 c2r:2048:131072 r2c:6000:16384 c2r:6000:16384 1000:65536 1001:65536 2187:32768 2400:32768 3125:16384 5000:16384 6000:16384 6144:8192 10000:8192 49152:2048"
timeout 900 python tools/sweep.py $CASES --steps 20 > gpurun_out/sweep_all_r6.log 2>&1
