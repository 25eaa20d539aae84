#!/usr/bin/env python
"""tools/e2e_probe.py -- host-pointer path (fft_exec_batch, pinned buffers) against the PCIe ceiling: the same bytes moved by
plain concurrent H2D + D2H copies with no kernels in between."""
import ctypes as C, importlib, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
hs = importlib.import_module("mixed-radix-fast-fourier-transform_b200")
L = hs.load()
n, batch = 4096, 65536
hin = torch.empty((batch, n, 2), dtype=torch.float64).pin_memory()
hout = torch.empty((batch, n, 2), dtype=torch.float64).pin_memory()
hin.uniform_(-1, 1)
p = hs.Plan(n, 1)
for _ in range(2):
    L.fft_exec_batch(p.p, C.c_void_p(hin.data_ptr()), C.c_void_p(hout.data_ptr()), batch)
t0 = time.perf_counter()
for _ in range(3):
    L.fft_exec_batch(p.p, C.c_void_p(hin.data_ptr()), C.c_void_p(hout.data_ptr()), batch)
dt = (time.perf_counter() - t0) / 3
print("fft_exec_batch: %.1f ms, %.1f GB/s each way (HSFFT_CHUNK_MB=%s)" % (dt * 1e3, hin.numel() * 8 / dt / 1e9, os.environ.get("HSFFT_CHUNK_MB", "64")))
if not os.environ.get("HSFFT_CHUNK_MB"):
    d1 = torch.empty_like(hin, device="cuda"); d2 = torch.empty_like(hin, device="cuda")
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(3):
        with torch.cuda.stream(s1):
            d1.copy_(hin, non_blocking=True)
        with torch.cuda.stream(s2):
            hout.copy_(d2, non_blocking=True)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 3
    print("plain concurrent H2D + D2H of the same bytes: %.1f ms, %.1f GB/s each way" % (dt * 1e3, hin.numel() * 8 / dt / 1e9))
