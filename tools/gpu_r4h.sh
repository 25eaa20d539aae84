#!/bin/bash
# final 1-GPU checkpoint after the last plan change: smoke, full GPU parity suite, reference arm, bench line, launch list
mkdir -p gpurun_out
exec > gpurun_out/final7.log 2>&1
echo "== smoke"; python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -3 gpurun_out/smoke.log
echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/pytest_gpu.log
echo "== bench reference arm"; timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "exit $?"; cut -c1-300 gpurun_out/bench_ref.json
echo "== bench"; timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"; cut -c1-400 gpurun_out/bench.json
echo "== ncu launch list"
python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu --no-configs --no-dist > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu --no-configs --no-dist > gpurun_out/ncu.log 2>&1
echo "ncu exit $?"; tail -2 gpurun_out/launches.csv | cut -c1-200
echo "== r2c/c2r 6000, 3000"; python tools/sweep.py --steps 20 3000:65536 r2c:6000:16384 c2r:6000:16384
