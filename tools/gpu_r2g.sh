#!/bin/bash
# round 2, call G (8 GPUs): four-step 2^30 exchange variants, bench at 8 GPUs (batch sharding + four-step record), e2e probe
mkdir -p gpurun_out
O=gpurun_out
T="timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29581 tools/dist_bench.py --log2n 30 --steps 3 --p2p"
for v in "HSFFT_DIST_BULK=0 HSFFT_DIST_C=8" "HSFFT_DIST_BULK=0 HSFFT_DIST_C=16" "HSFFT_DIST_BULK=0 HSFFT_DIST_C=32" "HSFFT_DIST_BULK=1 HSFFT_DIST_C=16" "HSFFT_DIST_BULK=1 HSFFT_DIST_C=32"; do
  echo "== $v"; env $v $T 2>/dev/null | grep '^{' | python -c "import json,sys; d=json.loads(sys.stdin.read()); print({k: d[k] for k in ('ms_total','ms_cols','ms_exchange','ms_rows','exchange_gbs_per_gpu_per_direction','rel_rms_err_sampled','ok')})"
done 2>&1 | tee $O/r2g_dist_variants_8gpu.txt
echo "== bench 8 GPUs"; timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29582 bench.py --gpus 8 --steps 10 --warmup 3 > $O/r2g_bench8.json 2> $O/r2g_bench8.err; echo "exit $?"; tail -2 $O/r2g_bench8.err
python - <<'PY'
import json
l=[x for x in open('gpurun_out/r2g_bench8.json') if x.startswith('{')]
d=json.loads(l[-1])
print("headline", d["value"], d["ms_per_step"], d["roofline"]["frac"]); print("e2e", d["e2e"])
for c in d["configs"]: print(c.get("workload"), c.get("ms_per_step"), c.get("frac"), c.get("error"))
print(json.dumps(d.get("dist"), indent=1))
PY
echo "== e2e scaling probe"; timeout 600 python tools/e2e_scaling_probe.py --gpus 1,4,8 --quick 2>&1 | grep "GPU(s)" | tee $O/r2g_e2e_probe_8gpu.txt
