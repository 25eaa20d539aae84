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
    "conv512k": ("conv", 1 << 19, 64),     # 2^19-sample rows convolved pairwise ("full", linear): padded length 2^20
}
# every BASELINE.json config that fits one GPU: the default run reports them all in the one JSON line (`configs`)
ALL_CONFIGS = ["c2c4096", "c2c3000", "blu1009", "blu10007", "r2c1m", "c2r1m", "conv512k"]


def flops_per_transform(kind, n):
    if kind == "conv":          # three real transforms of the padded length P = 2n
        return 3 * 2.5 * (2 * n) * math.log2(2 * n)
    return (5.0 if kind == "c2c" else 2.5) * n * math.log2(n)


def bytes_per_transform(kind, n):
    # algorithmic bytes
    if kind == "c2c":
        return 32.0 * n
    if kind == "conv":
        return 8.0 * (2 * n) + 8.0 * (2 * n - 1)
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


class Workload:
    """One BASELINE config resident in HBM: synthetic input, output, plan and a step() that launches one batch."""

    def __init__(self, hs, L, torch, dev, name, rank, batch=0):
        self.name = name
        self.kind, self.n, self.batch = WORKLOADS[name]
        if batch:
            self.batch = batch
        kind, n, batch = self.kind, self.n, self.batch
        g = torch.Generator(device=dev)
        g.manual_seed(88172645463325252 % (2 ** 63) + rank)
        rnd = lambda shape: torch.rand(shape, generator=g, device=dev, dtype=torch.float64) * 2 - 1
        self.plan = self.plan2 = None
        if kind == "c2c":
            self.x, self.y = rnd((batch, n, 2)), torch.empty((batch, n, 2), device=dev, dtype=torch.float64)
            self.plan = hs.Plan(n, 1)
            self.step = lambda: self.plan.exec_dev(self.x, self.y, batch)
            self.describe = self.plan.describe()
        elif kind == "r2c":
            self.x, self.y = rnd((batch, n)), torch.empty((batch, n, 2), device=dev, dtype=torch.float64)
            self.plan = hs.RealPlan(n, 1)
            self.step = lambda: self.plan.r2c_dev(self.x, self.y, n, 1, batch)
            self.describe = "r2c N=%d, full N-bin spectrum" % n
        elif kind == "c2r":
            h1 = n // 2 + 1
            self.x, self.y = rnd((batch, h1, 2)), torch.empty((batch, n), device=dev, dtype=torch.float64)
            self.plan = hs.RealPlan(n, -1)
            self.step = lambda: self.plan.c2r_dev(self.x, h1, self.y, batch)
            self.describe = "c2r N=%d" % n
        else:  # conv: rows of n samples convolved pairwise, "full" linear
            self.x, self.x2 = rnd((batch, n)), rnd((batch, n))
            self.y = torch.empty((batch, 2 * n), device=dev, dtype=torch.float64)
            self.step = lambda: L.fft_convolve_batch_dev(b"full", b"linear", C.c_void_p(self.x.data_ptr()), n, C.c_void_p(self.x2.data_ptr()), n,
                                                         C.c_void_p(self.y.data_ptr()), 2 * n, batch, None)
            self.describe = "fft_convolve %d (*) %d, padded length %d" % (n, n, 2 * n)
        self.in_bytes = self.x.numel() * 8 + (self.x2.numel() * 8 if kind == "conv" else 0)
        self.out_bytes = self.y.numel() * 8
        self.flops = self.batch * flops_per_transform(kind, n)
        self.alg_bytes = self.batch * bytes_per_transform(kind, n)

    def close(self):
        if self.plan is not None:
            self.plan.close()
        self.plan = None
        self.x = self.y = self.x2 = None


def kernel_breakdown(hs, torch, wl, steps):
    """Per-kernel device time of `steps` steps: CUDA events recorded by the library on the launching stream right after
    every kernel (hsfft_timing_begin / _report).  Returns (list of kernels, dominant kernel)."""
    torch.cuda.synchronize()
    with hs.launch_timing(0) as t:
        for _ in range(steps):
            wl.step()
    total = sum(ms for _, _, ms in t.kernels) or 1.0
    ks = [{"kernel": tag, "launches_per_step": cnt / steps, "ms_per_launch": ms / cnt, "share_of_step": ms / total} for tag, cnt, ms in t.kernels]
    dom = max(ks, key=lambda k: k["share_of_step"]) if ks else None
    return ks, dom


def traffic_record(name):
    """dram__bytes_read.sum + dram__bytes_write.sum per step from the committed ncu --set full capture (profiles/traffic.json
    names the capture each figure comes from); ncu cannot run inside a timed bench, so this is a recorded constant."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            t = json.load(f)
        v = t.get(name)
        if isinstance(v, dict):
            return v.get("bytes"), v.get("capture")
        return v, t.get("_note")
    except Exception:
        return None, None


def measure(hs, L, torch, dist, dev, wl, steps, warmup, world, sample_clocks=None):
    """W warm-up steps, then exactly K steps bracketed by barrier + synchronize; CUDA events on the launching stream; max over ranks."""
    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(warmup):
        wl.step()
    barrier()
    if sample_clocks:
        sample_clocks.start()
        time.sleep(0.3)
    launches0 = L.hsfft_launch_count()
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
    barrier()
    evs[0].record()
    for i in range(steps):
        wl.step()
        evs[i + 1].record()
    barrier()
    launches = L.hsfft_launch_count() - launches0
    total_ms = evs[0].elapsed_time(evs[-1])
    per_step = sorted(evs[i].elapsed_time(evs[i + 1]) for i in range(steps))
    clocks = sample_clocks.stop() if sample_clocks else None
    t = torch.tensor([total_ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item()) / steps
    return {"ms_per_step": ms, "median_step_ms": per_step[len(per_step) // 2], "launches": launches, "clocks": clocks}


def roofline_record(hs, torch, wl, m, steps, peak, peak_src):
    kernels, dom = kernel_breakdown(hs, torch, wl, max(2, min(steps, 5)))
    achieved = wl.alg_bytes / (m["ms_per_step"] * 1e-3) / 1e9
    traffic, capture = traffic_record(wl.name)
    return {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
            "traffic_capture": capture, "peak_source": peak_src,
            "kernel_ms": dom["ms_per_launch"] if dom else m["ms_per_step"], "dominant_kernel": dom["kernel"] if dom else None,
            "kernels": kernels, "launches_per_step": m["launches"] / steps, "algorithmic_bytes_per_step": wl.alg_bytes,
            "median_step_ms": m["median_step_ms"]}


def four_step_record(hs, torch, dist, dev, rank, local_rank, world, steps):
    """BASELINE config 5 on all ranks: one c2c of N = 2^27 * world points (2^30 on 8 GPUs), four-step, the exchange fused into
    phase A (NVLink peer stores) and, for reference, as an NCCL all-to-all.  Device-timed phases (max over ranks), sampled-bin
    parity against direct DFT sums, Parseval."""
    dmod = importlib.import_module("mixed-radix-fast-fourier-transform_b200.dist")
    log2n = 27 + int(math.log2(world))
    out = {"workload": "single c2c N=2^%d fp64, four-step on %d GPUs" % (log2n, world)}
    for mode in ("p2p", "nccl"):
        try:
            out[mode] = dmod.bench_four_step(hs, log2n, max(1, min(steps, 3)), mode == "p2p", rank, world, dev)
        except Exception as e:  # keep the batch-sharding line even if the exchange cannot be set up on this box
            out[mode] = {"error": repr(e)[:300]}
    return out


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
    ap.add_argument("--no-configs", action="store_true", help="skip the other BASELINE configs after the headline workload")
    ap.add_argument("--no-dist", action="store_true", help="skip the four-step record at --gpus > 1")
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

    # ---------------- CPU baseline first (rank 0, N = 1 only), before any GPU work heats the box: the reference's own C
    # code on this box's host cores, mean of 10 passes over a bounded sample
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu and kind != "conv":
        try:
            transforms = cpu_sample_size(kind, n)
            sec, threads, tag = time_cpu(kind, n, transforms, 10, warm=2, mean=True)
            cpu = {"value": transforms * flops_per_transform(kind, n) / sec / 1e9, "unit": "GFLOP/s", "cores": threads,
                   "kind": tag, "sample": "%d transforms of N=%d per pass, mean of 10 passes after 2 warm-up passes, OpenMP over the batch" % (transforms, n)}
            sec1, _, _ = time_cpu(kind, n, max(4, transforms // 8), 3, nthreads=1)
            cpu["single_thread_value"] = max(4, transforms // 8) * flops_per_transform(kind, n) / sec1 / 1e9
        except Exception as e:  # the checker being absent must not hide the GPU number
            cpu = {"value": None, "unit": "GFLOP/s", "cores": 0, "kind": "unavailable", "sample": repr(e)}

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
    peak, peak_src = measured_peak()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------- headline workload: synthetic input resident in HBM, exactly K timed steps
    wl = Workload(hs, L, torch, dev, args.workload, rank, args.batch)
    m = measure(hs, L, torch, dist, dev, wl, args.steps, args.warmup, world, ClockSampler(local_rank))
    ms_per_step, launches, clocks = m["ms_per_step"], m["launches"], m["clocks"]
    value = world * wl.flops / (ms_per_step * 1e-3) / 1e9
    roofline = roofline_record(hs, torch, wl, m, args.steps, peak, peak_src)
    achieved = roofline["achieved"]
    x, y, describe, in_bytes, out_bytes = wl.x, wl.y, wl.describe, wl.in_bytes, wl.out_bytes

    # a prime N that runs as Rader's cyclic convolution: also time the reference's own route (padded chirp-z) so the line
    # carries both numbers; the plan reads HSFFT_NO_RADER when it is first executed
    alt_route = None
    if kind == "c2c" and "rader" in describe:
        alt_route = chirp_z_route(hs, torch, wl, args, peak)

    # ---------------- end to end through the host-pointer C-ABI call (pinned host buffers)
    e2e = None
    if not args.no_e2e and kind in ("c2c", "r2c"):
        hin = torch.empty(x.shape, dtype=torch.float64).pin_memory()
        hout = torch.empty(y.shape, dtype=torch.float64).pin_memory()
        hin.copy_(x.cpu())
        if kind == "c2c":
            estep = lambda: L.fft_exec_batch(wl.plan.p, C.c_void_p(hin.data_ptr()), C.c_void_p(hout.data_ptr()), wl.batch)
        else:
            estep = lambda: L.fft_r2c_exec_batch(wl.plan.p, C.c_void_p(hin.data_ptr()), C.c_void_p(hout.data_ptr()), wl.batch)
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
        e2e = {"value": world * wl.flops / dt / 1e9, "unit": "GFLOP/s",
               "h2d_bytes_per_step": in_bytes, "d2h_bytes_per_step": out_bytes, "ms_per_step": dt * 1e3,
               "steps": esteps, "api": "fft_exec_batch(host pinned in, host pinned out)"}
        # spot-check the host path against the device path
        ok = torch.allclose(hout.view(-1)[:4096], y.view(-1)[:4096].cpu(), rtol=0, atol=1e-9)
        e2e["matches_device_path"] = bool(ok)
        del hin, hout
    del x, y
    wl.close()
    torch.cuda.empty_cache()

    # ---------------- the other BASELINE configs, same method (device resident, CUDA events, max over ranks), fewer steps
    configs = []
    if not args.no_configs:
        for name in ALL_CONFIGS:
            if name == args.workload:
                continue
            try:
                w2 = Workload(hs, L, torch, dev, name, rank)
                k2 = max(3, min(args.steps, 10))
                m2 = measure(hs, L, torch, dist, dev, w2, k2, max(3, min(args.warmup, 3)), world)
                r2 = roofline_record(hs, torch, w2, m2, k2, peak, peak_src)
                configs.append({"workload": "%s N=%d batch=%d per GPU (%s)" % (w2.kind, w2.n, w2.batch, name), "plan": w2.describe,
                                "steps": k2, "ms_per_step": m2["ms_per_step"], "value": world * w2.flops / (m2["ms_per_step"] * 1e-3) / 1e9,
                                "unit": "GFLOP/s", "hbm_gbs": r2["achieved"] * world, "frac": r2["frac"], "traffic": r2["traffic"],
                                "algorithmic_bytes_per_step": w2.alg_bytes, "launches_per_step": r2["launches_per_step"],
                                "dominant_kernel": r2["dominant_kernel"], "kernel_ms": r2["kernel_ms"], "kernels": r2["kernels"]})
                w2.close()
                del w2
                torch.cuda.empty_cache()
            except Exception as e:
                configs.append({"workload": name, "error": repr(e)[:300]})

    # ---------------- the one path with a real exchange step: a single huge transform over all ranks (N > 1 only)
    dist_rec = None
    if world > 1 and not args.no_dist:
        dist_rec = four_step_record(hs, torch, dist, dev, rank, local_rank, world, args.steps)

    if rank == 0:
        line = {
            "metric": "batched FFT GFLOP/s (5N log2 N)", "value": value, "unit": "GFLOP/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "%s N=%d batch=%d per GPU, forward, out of place (%s)" % (kind, n, wl.batch, args.workload),
                       "plan": describe, "l2": "inputs larger than L2 (no flush needed)", "parallelism": "batch-shard x%d" % world},
            "hbm_gbs": achieved * world, "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e,
            "gpu_launches": int(launches), "clocks": clocks, "configs": configs,
        }
        if dist_rec:
            line["dist"] = dist_rec
        if alt_route:
            line["config"]["chirp_z_route"] = alt_route
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def chirp_z_route(hs, torch, wl, args, peak):
    os.environ["HSFFT_NO_RADER"] = "1"
    try:
        plan_b = hs.Plan(wl.n, 1)
        for _ in range(max(args.warmup, 1)):
            plan_b.exec_dev(wl.x, wl.y, wl.batch)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            plan_b.exec_dev(wl.x, wl.y, wl.batch)
        e1.record()
        torch.cuda.synchronize()
        ms_b = e0.elapsed_time(e1) / args.steps
        return {"plan": plan_b.describe(), "ms_per_step": ms_b, "hbm_frac": wl.alg_bytes / (ms_b * 1e-3) / 1e9 / peak}
    finally:
        del os.environ["HSFFT_NO_RADER"]


if __name__ == "__main__":
    main()
