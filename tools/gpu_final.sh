#!/bin/bash
# round checkpoint: smoke, full GPU parity suite, sweep of every BASELINE config, bench lines, ncu launch list of the bench command
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > gpurun_out/gpu.txt 2>&1
echo "== smoke"; python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -3 gpurun_out/smoke.log
echo "== pytest gpu"; timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/pytest_gpu.log
echo "== pytest gpu, pipelines on"; HSFFT_PIPE=1 timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/pytest_gpu_pipe.log 2>&1; echo "pytest exit $?"; tail -2 gpurun_out/pytest_gpu_pipe.log
CASES="4096:65536 3000:65536 1024:262144 2048:131072 512:524288 256:1048576 8192:32768 1009:16384 10007:16384 40009:4096 16384:16384 32768:8192 65536:4096 131072:2048 262144:1024 524288:512 1048576:256 r2c:1048576:256 r2ch:1048576:256 c2r:1048576:256 conv:524288:64 1000:65536 2187:32768 6000:16384"
echo "== sweep"; timeout 900 python tools/sweep.py $CASES --steps 20 2>&1 | tee gpurun_out/sweep_final.log
echo "== sweep pipelines"; HSFFT_PIPE=1 HSFFT_PIPE_SHAPE=1 timeout 600 python tools/sweep.py 16384:16384 32768:8192 65536:4096 131072:2048 262144:1024 524288:512 1048576:256 --steps 20 2>&1 | tee gpurun_out/sweep_final_pipe.log
echo "== bench"; timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"; cat gpurun_out/bench.json
for w in c2c3000 blu1009 blu10007 r2c1m c2r1m; do timeout 600 python bench.py --workload $w --steps 10 --no-e2e --no-cpu > gpurun_out/bench_$w.json 2>> gpurun_out/bench.err; cat gpurun_out/bench_$w.json | cut -c1-400; done
echo "== bench reference arm"; timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref.json 2>&1; cat gpurun_out/bench_ref.json | cut -c1-600
echo "== ncu launch list"
python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu > gpurun_out/ncu.log 2>&1
echo "ncu exit $?"; tail -4 gpurun_out/launches.csv | cut -c1-300
