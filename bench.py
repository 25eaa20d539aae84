#!/usr/bin/env python
"""bench.py -- headline benchmark of hsfft-b200 (BASELINE.json metric: batched FFT GFLOP/s
(5 N log2 N) and HBM GB/s vs roofline).

    python bench.py --gpus N --steps K --warmup W [--workload c2c4096] [--impl reference]

One step = one pass of the hot path over one batch: `batch` forward complex transforms of length
N, double precision, through the C-ABI entry point fft_exec_batch_dev with the batch resident in
HBM.  Default workload = BASELINE.json configs[1]: N=4096, batch 65536 (4 GiB in + 4 GiB out per
GPU, far larger than the 126 MB L2, so no L2 flush is needed between steps).

N > 1 (torchrun): every rank owns its own batch of the same size on its own GPU (batch sharding,
no data-path collective) -> weak scaling; value = all ranks' flops / max-over-ranks time.

The JSON line carries `roofline` (dominant kernel's algorithmic bytes / its CUDA-event time vs
MEASURED_PEAKS.json), `cpu_baseline` (the reference's own C code from oracle/_ref -- or the
oracle port when that is absent -- timed on the host cores of this box on a bounded sample) and
`e2e` (same metric through fft_exec_batch with pinned HOST buffers, copies inside the timing).

--impl reference times only the reference's CPU implementation.
"""
import argparse
import ctypes as C
import importlib
import json
import math
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

WORKLOADS = {
    # name: (kind, N, batch)
    "c2c4096": ("c2c", 4096, 65536),
    "c2c3000": ("c2c", 3000, 65536),
    "c2c1024": ("c2c", 1024, 262144),
    "blu1009": ("c2c", 1009, 16384),
    "blu10007": ("c2c", 10007, 16384),
    "r2c1m": ("r2c", 1 << 20, 256),
    "c2r1m": ("c2r", 1 << 20, 256),
}


def flops_per_transform(kind, n):
    return (5.0 if kind == "c2c" else 2.5) * n * math.log2(n)


def bytes_per_transform(kind, n):
    # algorithmic bytes
    if kind == "c2c":
        return 32.0 * n
    if kind == "r2c":
        return 24.0 * n          # N reals in, full N-bin mirrored spectrum out (the reference's contract)
    return 16.0 * (n // 2 + 1) + 8.0 * n


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, power = [], [], set(), []
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); power.append(float(f[3]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons)}


def cpu_reference_lib():
    """The reference's own C implementation built into oracle/_ref (kind 'reference'), else the oracle port."""
    import oracle_lib
    try:
        return oracle_lib.RefLib("repaired"), "reference"
    except Exception:
        return oracle_lib.Oracle(), "port"


def host_cores():
    """Cores this process may use (torchrun exports OMP_NUM_THREADS=1, so ask the affinity mask, not OpenMP)."""
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def time_cpu(kind, n, transforms, repeats, nthreads=0, warm=1, mean=False):
    """Seconds per call of the CPU path on `transforms` transforms (best of `repeats`, or their mean), threads used."""
    import numpy as np
    import oracle_lib
    lib, tag = cpu_reference_lib()
    if nthreads <= 0:
        nthreads = host_cores()
    rng = np.random.default_rng(88172645463325252 % (2 ** 63))
    if kind == "c2c":
        x = rng.uniform(-1, 1, (transforms, 2 * n)).view(np.complex128)
        out = np.empty_like(x)
        fn = lib.L.refdrv_c2c_batch if tag == "reference" else lib.L.orc_c2c_batch
        call = lambda: fn(n, 1, oracle_lib._ptr(x), oracle_lib._ptr(out), transforms, nthreads)
    elif kind == "r2c":
        x = rng.uniform(-1, 1, (transforms, n))
        out = np.empty((transforms, n), dtype=np.complex128)
        fn = lib.L.refdrv_r2c_batch if tag == "reference" else lib.L.orc_r2c_batch
        call = lambda: fn(n, oracle_lib._ptr(x), oracle_lib._ptr(out), transforms, nthreads)
    else:
        x = rng.uniform(-1, 1, (transforms, 2 * (n // 2 + 1))).view(np.complex128)
        out = np.empty((transforms, n))
        fn = lib.L.refdrv_c2r_batch if tag == "reference" else lib.L.orc_c2r_batch
        call = lambda: fn(n, oracle_lib._ptr(x), oracle_lib._ptr(out), transforms, nthreads)
    for _ in range(warm):
        call()  # warm (plan creation is inside every call: one plan per thread, amortised over the sample)
    best, total = float("inf"), 0.0
    for _ in range(repeats):
        t0 = time.perf_counter()
        call()
        dt = time.perf_counter() - t0
        best, total = min(best, dt), total + dt
    threads = nthreads
    return (total / repeats if mean else best), threads, tag


def cpu_sample_size(kind, n):
    """Bounded sample: ~2 core-seconds of reference work per call (SURVEY.md section 6 rates)."""
    per = {"c2c": 1.5e-8, "r2c": 4e-8, "c2r": 3.2e-8}[kind] * n * max(1.0, math.log2(n) / 12.0)
    if n in (1009,):
        per = 150e-6
    if n in (10007,):
        per = 4e-3
    t = int(max(4, min(8192, 2.0 / per)))
    return t


def run_reference_arm(args, kind, n, batch, rank):
    if rank != 0:
        return
    transforms = cpu_sample_size(kind, n)
    # W untimed + exactly K timed passes over the bounded sample; ms_per_step is their mean
    sec, threads, tag = time_cpu(kind, n, transforms, max(1, args.steps), warm=max(0, args.warmup), mean=True)
    gflops = transforms * flops_per_transform(kind, n) / sec / 1e9
    line = {
        "impl": "reference", "metric": "batched FFT GFLOP/s (5N log2 N)", "value": gflops, "unit": "GFLOP/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "%s N=%d batch=%d fwd (CPU sample of %d transforms per step)" % (kind, n, batch, transforms)},
        "cpu_baseline": {"value": gflops, "unit": "GFLOP/s", "cores": threads, "kind": tag,
                         "sample": "%d transforms of N=%d per step, OpenMP over the batch, one plan per thread" % (transforms, n)},
        "e2e": {"value": gflops, "unit": "GFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2c4096", choices=sorted(WORKLOADS))
    ap.add_argument("--batch", type=int, default=0, help="override the workload's batch (parity/debug only)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    kind, n, batch = WORKLOADS[args.workload]
    if args.batch:
        batch = args.batch
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference_arm(args, kind, n, batch, rank)
        return

    import numpy as np
    import torch
    import torch.distributed as dist
    hs = importlib.import_module("mixed-radix-fast-fourier-transform_b200")
    L = hs.load()
    if not torch.cuda.is_available() or L.hsfft_device_count() <= 0:
        raise SystemExit("bench.py: no CUDA device -- hsfft-b200 has no CPU path to measure")
    torch.cuda.set_device(local_rank)
    L.hsfft_set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = torch.device("cuda", local_rank)

    # ---------------- synthetic input, resident in HBM
    g = torch.Generator(device=dev)
    g.manual_seed(88172645463325252 % (2 ** 63) + rank)
    if kind == "c2c":
        x = (torch.rand((batch, n, 2), generator=g, device=dev, dtype=torch.float64) * 2 - 1)
        y = torch.empty((batch, n, 2), device=dev, dtype=torch.float64)
        plan = hs.Plan(n, 1)
        step = lambda: plan.exec_dev(x, y, batch)
        describe = plan.describe()
        in_bytes, out_bytes = x.numel() * 8, y.numel() * 8
    elif kind == "r2c":
        x = torch.rand((batch, n), generator=g, device=dev, dtype=torch.float64) * 2 - 1
        y = torch.empty((batch, n, 2), device=dev, dtype=torch.float64)
        plan = hs.RealPlan(n, 1)
        step = lambda: plan.r2c_dev(x, y, n, 1, batch)
        describe = "r2c N=%d" % n
        in_bytes, out_bytes = x.numel() * 8, y.numel() * 8
    else:
        h1 = n // 2 + 1
        x = torch.rand((batch, h1, 2), generator=g, device=dev, dtype=torch.float64) * 2 - 1
        y = torch.empty((batch, n), device=dev, dtype=torch.float64)
        plan = hs.RealPlan(n, -1)
        step = lambda: plan.c2r_dev(x, h1, y, batch)
        describe = "c2r N=%d" % n
        in_bytes, out_bytes = x.numel() * 8, y.numel() * 8

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    barrier()

    # ---------------- timed region: exactly K steps, CUDA events on the launching (default) stream
    sampler = ClockSampler(local_rank)
    sampler.start()
    time.sleep(0.3)
    launches0 = L.hsfft_launch_count()
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    barrier()
    evs[0].record()
    for i in range(args.steps):
        step()
        evs[i + 1].record()
    barrier()
    launches = L.hsfft_launch_count() - launches0
    total_ms = evs[0].elapsed_time(evs[-1])
    per_step = sorted(evs[i].elapsed_time(evs[i + 1]) for i in range(args.steps))
    clocks = sampler.stop()
    t = torch.tensor([total_ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms = float(t.item())
    ms_per_step = total_ms / args.steps
    value = world * batch * flops_per_transform(kind, n) / (ms_per_step * 1e-3) / 1e9

    # a prime N that runs as Rader's cyclic convolution: also time the reference's own route (padded chirp-z) so the line
    # carries both numbers; the plan reads HSFFT_NO_RADER when it is first executed
    alt_route = None
    if kind == "c2c" and "rader" in describe:
        os.environ["HSFFT_NO_RADER"] = "1"
        try:
            plan_b = hs.Plan(n, 1)
            for _ in range(max(args.warmup, 1)):
                plan_b.exec_dev(x, y, batch)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(args.steps):
                plan_b.exec_dev(x, y, batch)
            e1.record()
            torch.cuda.synchronize()
            ms_b = e0.elapsed_time(e1) / args.steps
            alt_route = {"plan": plan_b.describe(), "ms_per_step": ms_b,
                         "hbm_frac": batch * bytes_per_transform(kind, n) / (ms_b * 1e-3) / 1e9 / measured_peak()[0]}
        finally:
            del os.environ["HSFFT_NO_RADER"]

    # ---------------- roofline of the dominant kernel (the whole step when it is one launch)
    peak, peak_src = measured_peak()
    kernel_ms = ms_per_step if launches == args.steps else (total_ms / max(launches, 1)) * (launches / args.steps)
    alg_bytes = batch * bytes_per_transform(kind, n)
    achieved = alg_bytes / (ms_per_step * 1e-3) / 1e9
    traffic = None
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            traffic = json.load(f).get(args.workload)
    except Exception:
        pass
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "peak_source": peak_src, "kernel_ms": kernel_ms,
                "launches_per_step": launches / args.steps, "algorithmic_bytes_per_step": alg_bytes,
                "median_step_ms": per_step[len(per_step) // 2]}

    # ---------------- end to end through the host-pointer C-ABI call (pinned host buffers)
    e2e = None
    if not args.no_e2e:
        hin = torch.empty(x.shape, dtype=torch.float64).pin_memory()
        hout = torch.empty(y.shape, dtype=torch.float64).pin_memory()
        hin.copy_(x.cpu())
        if kind == "c2c":
            estep = lambda: L.fft_exec_batch(plan.p, C.c_void_p(hin.data_ptr()), C.c_void_p(hout.data_ptr()), batch)
        elif kind == "r2c":
            estep = lambda: L.fft_r2c_exec_batch(plan.p, C.c_void_p(hin.data_ptr()), C.c_void_p(hout.data_ptr()), batch)
        else:
            estep = None
        if estep is not None:
            esteps = max(1, min(args.steps, 5))
            estep()
            barrier()
            t0 = time.perf_counter()
            for _ in range(esteps):
                estep()  # synchronous: returns after the last D2H has landed
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            tt = torch.tensor([dt], device=dev, dtype=torch.float64)
            if world > 1:
                dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            dt = float(tt.item()) / esteps
            e2e = {"value": world * batch * flops_per_transform(kind, n) / dt / 1e9, "unit": "GFLOP/s",
                   "h2d_bytes_per_step": in_bytes, "d2h_bytes_per_step": out_bytes, "ms_per_step": dt * 1e3,
                   "steps": esteps, "api": "fft_exec_batch(host pinned in, host pinned out)"}
            # spot-check the host path against the device path
            ok = torch.allclose(hout.view(-1)[:4096], y.view(-1)[:4096].cpu(), rtol=0, atol=1e-9)
            e2e["matches_device_path"] = bool(ok)
        del hin, hout

    # ---------------- CPU baseline (rank 0, N = 1 only): the reference's own C code on this box
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        try:
            transforms = cpu_sample_size(kind, n)
            sec, threads, tag = time_cpu(kind, n, transforms, 3)
            cpu = {"value": transforms * flops_per_transform(kind, n) / sec / 1e9, "unit": "GFLOP/s", "cores": threads,
                   "kind": tag, "sample": "%d transforms of N=%d, best of 3, OpenMP over the batch" % (transforms, n)}
            sec1, _, _ = time_cpu(kind, n, max(4, transforms // 8), 2, nthreads=1)
            cpu["single_thread_value"] = max(4, transforms // 8) * flops_per_transform(kind, n) / sec1 / 1e9
        except Exception as e:  # the checker being absent must not hide the GPU number
            cpu = {"value": None, "unit": "GFLOP/s", "cores": 0, "kind": "unavailable", "sample": repr(e)}

    if rank == 0:
        line = {
            "metric": "batched FFT GFLOP/s (5N log2 N)", "value": value, "unit": "GFLOP/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "%s N=%d batch=%d per GPU, forward, out of place (%s)" % (kind, n, batch, args.workload),
                       "plan": describe, "l2": "inputs larger than L2 (no flush needed)", "parallelism": "batch-shard x%d" % world},
            "hbm_gbs": achieved * world, "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e,
            "gpu_launches": int(launches), "clocks": clocks,
        }
        if alt_route:
            line["config"]["chirp_z_route"] = alt_route
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
