#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
echo "== bench default"; timeout 900 python bench.py > $O/r2c_bench.json 2> $O/r2c_bench.err; echo "exit $?"; tail -3 $O/r2c_bench.err
python - <<'PY'
import json
l=[x for x in open('gpurun_out/r2c_bench.json') if x.startswith('{')]
d=json.loads(l[-1])
print("headline", d["value"], d["ms_per_step"], d["roofline"]["frac"], d["roofline"]["dominant_kernel"], d["roofline"]["kernel_ms"])
print("cpu", d["cpu_baseline"]); print("e2e", d["e2e"])
for c in d["configs"]:
    if "error" in c: print(c); continue
    print("%-45s %8.3f ms frac %.3f launches %.1f dom %s %.3f ms" % (c["workload"], c["ms_per_step"], c["frac"], c["launches_per_step"], c["dominant_kernel"], c["kernel_ms"]))
    for k in c["kernels"]: print("      ", k)
PY
echo "== reference arm"; timeout 600 python bench.py --impl reference --steps 5 --warmup 1 | cut -c1-400
