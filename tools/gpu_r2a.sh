#!/bin/bash
# round 2, call A: parity of the split chirp-z kernels, A/B of their variants, baseline sweep of every BASELINE config, one ncu capture
mkdir -p gpurun_out
O=gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > $O/r2a_gpu.txt
echo "== pytest (bluestein / rader)"; timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "bluestein or rader or really" > $O/r2a_pytest.log 2>&1; echo "pytest exit $?"; tail -3 $O/r2a_pytest.log
S="timeout 300 python tools/sweep.py --steps 10"
echo "== default"; $S 10007:16384 8209:16384 10239:16384 4099:16384 8191:16384 40009:4096 2>&1 | tee $O/r2a_split_default.txt
for v in 47 48 49; do echo "== PREFER_CT=$v"; HSFFT_PREFER_CT=$v $S 10007:16384 2>&1 | tee $O/r2a_split_ct$v.txt; done
echo "== autosort family"; HSFFT_SPLIT_AUTOSORT=1 $S 10007:16384 4099:16384 8191:16384 2>&1 | tee $O/r2a_split_autosort.txt
echo "== three-kernel route"; HSFFT_BLU_SPLIT=0 $S 10007:16384 4099:16384 8191:16384 2>&1 | tee $O/r2a_split_off.txt
echo "== small lengths: fused M=8192 vs split"; $S 4093:16384 2053:32768 2>&1 | tee $O/r2a_small_default.txt
HSFFT_BLU_SPLIT=2 $S 4093:16384 2053:32768 2>&1 | tee $O/r2a_small_split.txt
HSFFT_BLU_SPLIT=2 HSFFT_SPLIT_AUTOSORT=1 $S 4093:16384 2053:32768 2>&1 | tee $O/r2a_small_split_autosort.txt
echo "== prefetch off"; HSFFT_PREFETCH=0 $S 10007:16384 2>&1 | tee $O/r2a_split_nopf.txt
echo "== all configs"; timeout 600 python tools/sweep.py --steps 10 4096:65536 3000:65536 1024:262144 8192:32768 1009:16384 r2c:1048576:256 c2r:1048576:256 conv:524288:64 131072:2048 524288:512 1048576:256 2>&1 | tee $O/r2a_all_configs.txt
echo "== ncu"; bash tools/gpu_ncu_one.sh r2a_split10007 10007:4096 conv_bluestein_split 1 1
