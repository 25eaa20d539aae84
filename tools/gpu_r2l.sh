#!/bin/bash
# round 2, call L: ncu --set full summaries of the kernels that changed this round, final sweep of every config
mkdir -p gpurun_out
O=gpurun_out
echo "== pytest rader"; timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "rader" > $O/r2l_pytest.log 2>&1; echo "pytest exit $?"; tail -3 $O/r2l_pytest.log
S="timeout 600 python tools/sweep.py --steps 10"
$S 4096:65536 3000:65536 1024:262144 2048:131072 512:524288 256:1048576 8192:32768 1009:16384 10007:16384 4093:16384 8191:16384 40009:4096 \
   16384:16384 65536:4096 131072:2048 262144:1024 524288:512 1048576:256 \
   r2c:1048576:256 r2ch:1048576:256 c2r:1048576:256 r2c:65536:4096 r2c:8192:32768 c2r:8192:32768 r2c:2048:131072 c2r:2048:131072 \
   conv:524288:64 convs:268435456:4097 fft2:1024x1024:256 \
   1000:65536 1001:65536 6000:16384 6561:16384 10000:16384 49152:4096 100000:2048 769:65536 2521:16384 2>&1 | tee $O/r2l_sweep_all.txt
for a in "r2l_r2c1m r2c:1048576:64 hs_tiled 2 2" "r2l_c2r1m c2r:1048576:64 hs_tiled 2 2" "r2l_rader1009 1009:16384 hs_rader 1 1" "r2l_r2c8192 r2c:8192:32768 hs_r2c_smem 1 1" \
         "r2l_conv512k conv:524288:32 hs_ 6 6" "r2l_jit10000 10000:8192 hs_jit_pass 2 2" "r2l_blu8191 8191:4096 bluestein_split 1 1"; do
  bash tools/gpu_ncu_one.sh $a > /dev/null 2>&1; n=$(echo $a | cut -d' ' -f1)
  echo "#### $n"; grep -n "kernel:\|gpu__time_duration\|gpu__dram_throughput\|fp64_cycles\|l1tex__throughput\|warps_active\|registers_per_thread\|dram__bytes" $O/summary_$n.txt | cut -c1-170
done
