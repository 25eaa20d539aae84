#!/usr/bin/env python
"""tools/e2e_scaling_probe.py -- where does the host-pointer path stop scaling over GPUs?

    python tools/e2e_scaling_probe.py [--gpus 1,2,4,8] [--gib 1]

For every GPU count G it starts G processes at the same moment, one GPU each, that move the same bytes concurrently host->device and
device->host between pinned buffers with plain cudaMemcpyAsync (torch copy_, no FFT kernel at all), and prints the per-GPU and the
summed rate per direction.  If the sum stays flat from 1 to G GPUs the ceiling is the host (memory bandwidth / PCIe root / IOMMU of
this box), not the library; fft_exec_batch reaches this ceiling at one GPU already (profiles/r1b_e2e_pcie_ceiling.txt).
A second pass pins each process to the CPUs `nvidia-smi topo -m` lists for its GPU before it allocates (first-touch NUMA placement).
"""
import argparse
import os
import subprocess
import sys
import time

CHILD = r'''
import os, sys, time, torch
gib = float(sys.argv[1]); start = float(sys.argv[2]); aff = sys.argv[3]
if aff != "-":
    cpus = set()
    for part in aff.split(","):
        a, _, b = part.partition("-")
        cpus |= set(range(int(a), int(b or a) + 1))
    try:
        os.sched_setaffinity(0, cpus & os.sched_getaffinity(0) or os.sched_getaffinity(0))
    except Exception as e:
        print("affinity failed", e)
n = int(gib * (1 << 30) / 8)
hin = torch.empty(n, dtype=torch.float64).pin_memory(); hout = torch.empty(n, dtype=torch.float64).pin_memory(); hin.fill_(1.0); hout.fill_(0.0)
d1 = torch.empty(n, dtype=torch.float64, device="cuda"); d2 = torch.zeros(n, dtype=torch.float64, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def rnd():
    with torch.cuda.stream(s1): d1.copy_(hin, non_blocking=True)
    with torch.cuda.stream(s2): hout.copy_(d2, non_blocking=True)
rnd(); torch.cuda.synchronize()
while time.time() < start: pass
t0 = time.perf_counter()
for _ in range(8): rnd()
torch.cuda.synchronize()
dt = (time.perf_counter() - t0) / 8
print("RATE %.2f" % (n * 8 / dt / 1e9))
'''


def topo_affinity(ngpu):
    try:
        out = subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True, timeout=30).stdout
    except Exception:
        return ["-"] * ngpu, ""
    aff = ["-"] * ngpu
    hdr = None
    for ln in out.splitlines():
        f = ln.split()
        if not f:
            continue
        if hdr is None and "CPU" in ln and "Affinity" in ln:
            hdr = ln
        if f[0].startswith("GPU") and f[0][3:].isdigit() and int(f[0][3:]) < ngpu:
            cand = [x for x in f[1:] if x and x[0].isdigit() and ("-" in x or "," in x)]
            if cand:
                aff[int(f[0][3:])] = cand[0]
    return aff, out


def run(g, gib, aff):
    start = time.time() + 18.0  # all children begin their timed rounds at the same wall-clock instant
    procs = []
    for i in range(g):
        env = dict(os.environ, CUDA_VISIBLE_DEVICES=str(i))
        procs.append(subprocess.Popen([sys.executable, "-c", CHILD, str(gib), str(start), aff[i]], env=env, stdout=subprocess.PIPE, text=True))
    rates = []
    for p in procs:
        out = p.communicate()[0]
        for ln in out.splitlines():
            if ln.startswith("RATE"):
                rates.append(float(ln.split()[1]))
    return rates


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", default="1,2,4,8")
    ap.add_argument("--gib", type=float, default=1.0)
    ap.add_argument("--quick", action="store_true", help="skip the CPU-affinity pass")
    args = ap.parse_args()
    import torch
    have = torch.cuda.device_count()
    aff, topo = topo_affinity(have)
    print(topo)
    for g in [int(x) for x in args.gpus.split(",") if int(x) <= have]:
        for tag, a in ((("no affinity", ["-"] * have),) if args.quick else (("no affinity", ["-"] * have), ("GPU-local CPUs", aff))):
            r = run(g, args.gib, a)
            if r:
                print("%d GPU(s), %s: per GPU %s GB/s each way, sum %.1f GB/s each way" % (g, tag, " ".join("%.1f" % x for x in r), sum(r)), flush=True)


if __name__ == "__main__":
    main()
