#!/bin/bash
# round 2, call E (2 GPUs): multi-device tests, bench at 2 GPUs with the four-step record, e2e probe per rank
mkdir -p gpurun_out
O=gpurun_out
nvidia-smi -L
echo "== pytest (multi-device + four-step)"; timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "two_devices or two_streams or four_step or half_spectrum" > $O/r2e_pytest.log 2>&1; echo "pytest exit $?"; tail -4 $O/r2e_pytest.log
echo "== smoke"; timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4 | tee $O/r2e_smoke.txt
echo "== bench 2 GPUs"; timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus 2 --steps 10 --warmup 3 > $O/r2e_bench2.json 2> $O/r2e_bench2.err; echo "exit $?"; tail -3 $O/r2e_bench2.err
python - <<'PY'
import json
l=[x for x in open('gpurun_out/r2e_bench2.json') if x.startswith('{')]
d=json.loads(l[-1])
print("headline", d["value"], d["ms_per_step"], d["roofline"]["frac"]); print("e2e", d["e2e"])
print(json.dumps(d.get("dist"), indent=1))
PY
echo "== in-library sharding e2e (one process, 2 devices)"
timeout 600 python - <<'PY'
import ctypes as C, importlib, time, sys, os
sys.path.insert(0, os.getcwd())
import torch
hs = importlib.import_module("mixed-radix-fast-fourier-transform_b200")
L = hs.load()
n, batch = 4096, 65536
hin = torch.empty((batch, n, 2), dtype=torch.float64).pin_memory(); hout = torch.empty_like(hin).pin_memory(); hin.uniform_(-1, 1)
p = hs.Plan(n, 1)
for nd in (1, 2):
    L.hsfft_set_devices(nd)
    for _ in range(2): L.fft_exec_batch(p.p, C.c_void_p(hin.data_ptr()), C.c_void_p(hout.data_ptr()), batch)
    t0 = time.perf_counter()
    for _ in range(3): L.fft_exec_batch(p.p, C.c_void_p(hin.data_ptr()), C.c_void_p(hout.data_ptr()), batch)
    dt = (time.perf_counter() - t0) / 3
    print("fft_exec_batch over %d device(s) from one process: %.1f ms, %.1f GB/s each way, %.1f GFLOP/s" % (nd, dt * 1e3, hin.numel() * 8 / dt / 1e9, batch * 5 * n * 12 / dt / 1e9))
PY
