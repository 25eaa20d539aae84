#!/bin/bash
# round 2, call F (2 GPUs): bulk-copy exchange variants, e2e scaling probe
mkdir -p gpurun_out
O=gpurun_out
T="timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29571 tools/dist_bench.py --log2n 28 --steps 3 --p2p"
for v in "HSFFT_DIST_BULK=0 HSFFT_DIST_C=8" "HSFFT_DIST_BULK=0 HSFFT_DIST_C=16" "HSFFT_DIST_BULK=0 HSFFT_DIST_C=32" "HSFFT_DIST_BULK=1 HSFFT_DIST_C=16" "HSFFT_DIST_BULK=1 HSFFT_DIST_C=32"; do
  echo "== $v"; env $v $T 2>/dev/null | grep '^{' | python -c "import json,sys; d=json.loads(sys.stdin.read()); print({k: d[k] for k in ('ms_total','ms_cols','ms_rows','exchange_gbs_per_gpu_per_direction','rel_rms_err_sampled','ok')})"
done 2>&1 | tee $O/r2f_dist_variants_2gpu.txt
echo "== e2e scaling probe"; timeout 600 python tools/e2e_scaling_probe.py --gpus 1,2 2>&1 | tee $O/r2f_e2e_probe_2gpu.txt
