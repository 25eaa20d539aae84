#!/bin/bash
# validation of the fat-thread plan-time rule: parity tests that touch it, then the cases the first sweep flagged
mkdir -p gpurun_out
exec > gpurun_out/r4d.log 2>&1
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "plan_time or real or in_place or half_spectrum or compile_time or mixed_radix" 2>&1 | tail -5
timeout 600 python tools/sweep.py --steps 10 1536:43690 1600:41943 3600:18641 3000:65536 8192:32768 c2r:6000:16384 r2c:6000:16384 c2r:10000:8192 r2c:10000:8192 c2r:12000:8192 c2r:9000:8192 r2c:9000:8192 c2r:8000:8192 r2c:8000:8192 c2r:10240:8192 r2c:10240:8192 c2r:12500:8192 r2c:12500:8192 2>&1 | cut -c1-200
