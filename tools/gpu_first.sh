#!/bin/bash
# first GPU contact: smoke, parity tests, throughput sweep, bench line, ncu launch list
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > gpurun_out/gpu.txt 2>&1
echo "== smoke" ; python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -12 gpurun_out/smoke.log
echo "== pytest gpu"; timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -15 gpurun_out/pytest_gpu.log
echo "== sweep"; timeout 600 python tools/sweep.py 4096:65536 3000:65536 1024:262144 2048:131072 8192:32768 512:524288 256:1048576 1009:16384 10007:16384 1000:65536 32768:4096 > gpurun_out/sweep.log 2>&1; cat gpurun_out/sweep.log
echo "== sweep variant B"; HSFFT_PREFER_CT=1 timeout 300 python tools/sweep.py 4096:65536 > gpurun_out/sweep_b.log 2>&1; cat gpurun_out/sweep_b.log
echo "== sweep generic"; HSFFT_NO_CT=1 timeout 300 python tools/sweep.py 4096:16384 3000:16384 1009:4096 > gpurun_out/sweep_gen.log 2>&1; cat gpurun_out/sweep_gen.log
echo "== bench"; timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"; cat gpurun_out/bench.json; tail -3 gpurun_out/bench.err
echo "== bench reference arm"; timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref.json 2>&1; cat gpurun_out/bench_ref.json
echo "== ncu launch list"
python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu > gpurun_out/ncu.log 2>&1
echo "ncu exit $?"; tail -5 gpurun_out/launches.csv
