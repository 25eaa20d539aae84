#!/bin/bash
# round 2, call B: fused real transforms (one-kernel and four-step), split chirp-z defaults
mkdir -p gpurun_out
O=gpurun_out
echo "== pytest (real / conv / bluestein / jit)"; timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "real or convolve or bluestein or plan_time or config1" > $O/r2b_pytest.log 2>&1; echo "pytest exit $?"; tail -3 $O/r2b_pytest.log
S="timeout 300 python tools/sweep.py --steps 10"
echo "== real, fused"; $S r2c:1048576:256 r2ch:1048576:256 c2r:1048576:256 r2c:524288:512 r2c:262144:1024 r2c:131072:2048 r2c:65536:4096 r2c:32768:8192 r2c:8192:32768 c2r:8192:32768 r2c:2048:131072 c2r:2048:131072 r2c:6000:32768 r2c:2000:65536 2>&1 | tee $O/r2b_real_fused.txt
echo "== real, unfused"; HSFFT_NO_R2C_FUSE=1 HSFFT_NO_C2R_FUSE=1 $S r2c:1048576:256 r2c:524288:512 r2c:131072:2048 r2c:32768:8192 r2c:8192:32768 c2r:8192:32768 r2c:2048:131072 c2r:2048:131072 2>&1 | tee $O/r2b_real_unfused.txt
echo "== real, four-step swapped"; HSFFT_R4_SWAP=1 $S r2c:524288:512 r2c:131072:2048 r2c:32768:8192 2>&1 | tee $O/r2b_real_swap.txt
echo "== split defaults"; $S 10007:16384 8191:16384 4099:16384 4093:16384 2053:32768 2039:32768 1031:32768 521:65536 2>&1 | tee $O/r2b_split_default.txt
HSFFT_BLU_SPLIT=2 $S 2039:32768 1031:32768 521:65536 2>&1 | tee $O/r2b_split_mode2.txt
HSFFT_BLU_SPLIT=0 $S 4093:16384 2053:32768 2>&1 | tee $O/r2b_split_off.txt
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -8 | tee $O/r2b_smoke.txt
