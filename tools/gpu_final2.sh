#!/bin/bash
# round-2 checkpoint: smoke, full GPU parity suite, bench lines (own arm + reference arm), ncu launch list of the bench command
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > gpurun_out/gpu.txt 2>&1
echo "== smoke"; python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -3 gpurun_out/smoke.log
echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/pytest_gpu.log
echo "== fft2"; timeout 300 python tools/sweep.py --kernels --steps 20 fft2:512x512:1024 fft2:1024x1024:256
echo "== bench reference arm"; timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "exit $?"; cut -c1-400 gpurun_out/bench_ref.json
echo "== bench"; timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"; cut -c1-600 gpurun_out/bench.json
echo "== ncu launch list"
python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu --no-configs --no-dist > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu --no-configs --no-dist > gpurun_out/ncu.log 2>&1
echo "ncu exit $?"; tail -4 gpurun_out/launches.csv | cut -c1-300
