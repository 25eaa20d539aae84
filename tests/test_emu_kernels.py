"""CPU tests that run the product's KERNEL SOURCE (csrc/*.cuh) under the thread emulation of
tests/emu (one OS thread per CUDA thread, std::barrier for __syncthreads) and compare with the
oracle.  This checks index arithmetic, barrier placement, the planner and the numerics of every
kernel family without a GPU; the same cases run on the B200 in test_gpu_parity.py."""
import os

import numpy as np
import pytest

from oracle_lib import rand_cplx, rel_rms, tol, xorshift_uniform


def run_c2c(hs, lib, oracle, n, sgn, howmany=2, seed=0):
    x = rand_cplx(howmany, n, seed=77 + n + seed)
    with hs.Plan(n, sgn, lib=lib) as p:
        desc = p.describe()
        y = p.exec(x)
    r = oracle.c2c(x, sgn)
    return max(rel_rms(y[i], r[i]) for i in range(howmany)), desc


@pytest.mark.parametrize("n", [4096, 3000, 1024, 2048, 512, 256, 8192])
def test_compile_time_plans(hs, emu_lib, oracle, n):
    for sgn in (1, -1):
        err, desc = run_c2c(hs, emu_lib, oracle, n, sgn, howmany=5 if n <= 1024 else 2)
        assert "compile-time" in desc and err <= tol(n), (n, sgn, err, desc)


def test_n3000_is_within_rounding_of_the_reference_not_just_tolerance(hs, emu_lib, oracle):
    """Same butterflies, same truncated constants, same stage twiddles: ~1e-16, far below 1.16e-11."""
    err, _ = run_c2c(hs, emu_lib, oracle, 3000, 1)
    assert err <= 1e-15


@pytest.mark.parametrize("n", [67, 131, 257, 509, 1031, 2039, 4093])
def test_fused_bluestein_every_pad_length_one_sm_holds(hs, emu_lib, oracle, n, monkeypatch):
    """M = 256, 512, 1024, 4096 (palindromic radix lists: the inverse plan is the forward plan) and 8192."""
    monkeypatch.setenv("HSFFT_NO_RADER", "1")  # N = 257 would otherwise run as a 256-point cyclic convolution
    monkeypatch.setenv("HSFFT_BLU_SPLIT", "0")  # N = 4093 would otherwise run as two 4096-point convolutions
    for sgn in (1, -1):
        err, desc = run_c2c(hs, emu_lib, oracle, n, sgn, howmany=5)
        assert "bluestein" in desc and "fused compile-time" in desc and err <= tol(n), (n, sgn, err, desc)


@pytest.mark.parametrize("n", [1009, 257])
@pytest.mark.parametrize("stage", ["0", "2"])
def test_rader_prime_lengths_with_a_smooth_predecessor(hs, emu_lib, oracle, n, stage, monkeypatch):
    """N prime, N-1 = 1008 / 256 held in shared memory: gather by g^q, FFT, * Bk, IFFT, scatter by g^-p (hs_fused.cuh).
    Odd batch: the last CTA has idle teams.  Same DFT as the reference's Bluestein route, to rounding."""
    monkeypatch.setenv("HSFFT_RADER_STAGE", stage)
    for sgn in (1, -1):
        err, desc = run_c2c(hs, emu_lib, oracle, n, sgn, howmany=7)
        assert "rader" in desc and err <= tol(n) and err < 5e-15, (n, sgn, err, desc)


def test_rader_is_only_taken_for_primes_with_a_kernel(hs, emu_lib, oracle):
    for n in (1008 + 1 + 1008, 1021, 769):  # 2017 = prime but 2016 has no kernel; 1021, 769: primes, N-1 without a kernel
        with hs.Plan(n, 1, lib=emu_lib) as p:
            assert "rader" not in p.describe(), (n, p.describe())


@pytest.mark.parametrize("n", [1009, 513, 600, 1000])
def test_fused_bluestein_and_neighbours(hs, emu_lib, oracle, n, monkeypatch):
    monkeypatch.setenv("HSFFT_NO_RADER", "1")
    for sgn in (1, -1):
        err, desc = run_c2c(hs, emu_lib, oracle, n, sgn)
        assert err <= tol(n), (n, sgn, err, desc)


@pytest.mark.parametrize("n", [2, 3, 4, 5, 7, 8, 15, 20, 64, 11, 13, 26, 143, 53, 1001, 19, 59, 257])
def test_runtime_planned_engine_small(hs, emu_lib, oracle, n):
    for sgn in (1, -1):
        err, desc = run_c2c(hs, emu_lib, oracle, n, sgn, howmany=3)
        assert err <= tol(n), (n, sgn, err, desc)


@pytest.mark.parametrize("n", [6561, 10000, 24576])
def test_runtime_planned_engine_multipass(hs, emu_lib, oracle, n):
    err, desc = run_c2c(hs, emu_lib, oracle, n, 1, howmany=2)
    assert ("pass" in desc) and err <= tol(n), (n, err, desc)


@pytest.mark.parametrize("n", [16384, 65536])
def test_compile_time_two_pass(hs, emu_lib, oracle, n):
    err, desc = run_c2c(hs, emu_lib, oracle, n, -1, howmany=1)
    assert "two-pass" in desc and err <= tol(n), (err, desc)


@pytest.mark.parametrize("n", [20011, 40009])
def test_long_bluestein_three_fused_kernels(hs, emu_lib, oracle, n):
    """Lengths above the split chirp-z kernels (N > 10240): chirp + pass 1, pass 2 * Hk + inverse pass 1, inverse pass 2 + chirp."""
    for sgn in (1, -1):
        err, desc = run_c2c(hs, emu_lib, oracle, n, sgn, howmany=2)
        assert "3 fused tiled kernels" in desc and err <= tol(n), (n, sgn, err, desc)


def test_long_bluestein_route_still_serves_mid_lengths_when_split_is_off(hs, emu_lib, oracle, monkeypatch):
    monkeypatch.setenv("HSFFT_BLU_SPLIT", "0")
    for n in (4099, 10007):
        err, desc = run_c2c(hs, emu_lib, oracle, n, 1, howmany=2)
        assert "3 fused tiled kernels" in desc and err <= tol(n), (n, err, desc)


@pytest.mark.parametrize("n", [2053, 4093, 4099, 8191, 8209, 10007, 10239])
def test_split_chirp_z_one_kernel_two_half_length_convolutions(hs, emu_lib, oracle, n):
    """N <= M' = 4096 / 8192 / 10240: the padded convolution of length 2M' as two convolutions of length M' inside one CTA
    (hs_fused.cuh: the auto-sort pair, M' <= 8192; hs_conv_fft.cuh: in-place transforms with a digit-reversed spectrum,
    M' = 10240), odd batch."""
    want = "in-place" if n > 6144 else "auto-sort"
    for sgn in (1, -1):
        err, desc = run_c2c(hs, emu_lib, oracle, n, sgn, howmany=3)
        assert "split chirp-z" in desc and want in desc and err <= tol(n) and err < 5e-15, (n, sgn, err, desc)


def test_split_chirp_z_with_and_without_tensor_memory_parking(hs, emu_lib, oracle):
    """M' = 10240: the persistent kernel that parks a1 / y0 in tensor memory and prefetches the next transform with cp.async
    (five transforms over the emulation's two CTAs: a CTA walks 3 / 2 of them) against the kernel that parks in the L2.
    HSFFT_SPLIT_TMEM is read once per process, so the L2 variant runs in a child process."""
    import subprocess
    import sys
    err, desc = run_c2c(hs, emu_lib, oracle, 10007, 1, howmany=5)
    assert err <= tol(10007) and err < 5e-15, (err, desc)
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = ("import importlib,sys;sys.path.insert(0,%r);sys.path.insert(0,%r+'/tests');import oracle_lib as ol;"
            "hs=importlib.import_module('mixed-radix-fast-fourier-transform_b200');L=hs.load(%r+'/tests/emu/_build/libhsfft_emu.so');"
            "x=ol.rand_cplx(3,10007,seed=1);p=hs.Plan(10007,-1,lib=L);y=p.exec(x);r=ol.Oracle().c2c(x,-1);"
            "print('ERR',max(ol.rel_rms(y[i],r[i]) for i in range(3)))") % (root, root, root)
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=dict(os.environ, HSFFT_SPLIT_TMEM="0"), timeout=300)
    assert "ERR" in r.stdout and float(r.stdout.split("ERR")[1]) < 5e-15, r.stdout + r.stderr[-800:]


def test_smooth_bluestein_pad_uses_exact_radix5_and_stays_bluestein_only(hs, emu_lib, oracle, monkeypatch):
    """N in 8193..10240 without the split kernel: the device pads to 20480 = 128 x 160 (full-precision radix-5 stages)
    instead of 32768.  A user-facing N = 20480 transform must NOT take that plan: the reference's own radix-5 literals apply there."""
    monkeypatch.setenv("HSFFT_BLU_SPLIT", "0")
    for n in (8209, 10239):
        err, desc = run_c2c(hs, emu_lib, oracle, n, 1, howmany=2)
        assert "M=20480=160x128" in desc and err <= 1e-14, (n, err, desc)
    err, desc = run_c2c(hs, emu_lib, oracle, 20480, 1, howmany=1)
    assert "run-time" in desc and err <= 1e-14, (err, desc)   # within rounding of the reference's truncated constants
    monkeypatch.setenv("HSFFT_BLU_POW2", "1")
    err, desc = run_c2c(hs, emu_lib, oracle, 8209, -1, howmany=1)
    assert "M=32768" in desc and err <= tol(8209)


def test_three_pass_plan(hs, emu_lib, oracle):
    n = 3 << 17   # 393216 = [8]*5 + [4, 3]: no compile-time plan, three run-time passes
    err, desc = run_c2c(hs, emu_lib, oracle, n, -1, howmany=1)
    assert "3 run-time" in desc and err <= tol(n), (err, desc)


@pytest.mark.parametrize("n", [4, 8, 16, 32, 64, 2018, 4096])
def test_real_transforms(hs, emu_lib, oracle, n):
    x = xorshift_uniform(3 * n, seed=5 + n).reshape(3, n)
    with hs.RealPlan(n, 1, lib=emu_lib) as f, hs.RealPlan(n, -1, lib=emu_lib) as g:
        X = f.r2c(x)
        Xr = oracle.r2c(x)
        assert rel_rms(X, Xr) <= tol(n)
        assert np.array_equal(X[:, n - 1: n // 2: -1], np.conj(X[:, 1: n // 2]))   # mirror bins written
        y = g.c2r(X)
        assert rel_rms(y, oracle.c2r(Xr, n)) <= tol(n)
        assert rel_rms(y / n, x) <= tol(n)
        assert rel_rms(g.c2r(f.r2c(x[0])) / n, x[0]) <= tol(n)                   # single-transform entry points


@pytest.mark.parametrize("n", [512, 1024, 2048, 4096, 6000, 8192, 16384])
def test_real_transforms_ride_on_the_one_kernel_complex_plan(hs, emu_lib, oracle, n):
    """H = N/2 with a one-kernel plan: r2c splits the spectrum in shared memory after the last pass, c2r merges on the first
    pass's loads -- ONE launch each, no half-length spectrum / merged signal in HBM (reference src/real.c:109-132, :171-179).
    Batch 5: the last CTA of the plans that hold several transforms per CTA has idle teams."""
    x = xorshift_uniform(5 * n, seed=9 + n).reshape(5, n)
    with hs.RealPlan(n, 1, lib=emu_lib) as f, hs.RealPlan(n, -1, lib=emu_lib) as g:
        f.r2c(x[:2]); g.c2r(oracle.r2c(x[:2]))          # plans built
        l0 = emu_lib.hsfft_launch_count()
        X = f.r2c(x)
        l1 = emu_lib.hsfft_launch_count()
        Xr = oracle.r2c(x)
        y = g.c2r(Xr)
        l2 = emu_lib.hsfft_launch_count()
        assert (l1 - l0, l2 - l1) == (1, 1), (l1 - l0, l2 - l1)
        assert max(rel_rms(X[i], Xr[i]) for i in range(5)) <= min(tol(n), 5e-15)
        assert np.array_equal(X[:, n - 1: n // 2: -1], np.conj(X[:, 1: n // 2]))
        assert max(rel_rms(y[i], oracle.c2r(Xr, n)[i]) for i in range(5)) <= min(tol(n), 5e-15)


@pytest.mark.parametrize("n,swap,staged", [(1 << 15, 0, 0), (1 << 16, 0, 0), (1 << 17, 0, 0), (1 << 17, 1, 0), (1 << 17, 0, 1), (1 << 18, 0, 0),
                                           (1 << 19, 0, 0), (1 << 19, 0, 1), (1 << 20, 0, 1)])
def test_real_four_step_split_on_pass_1_mirror_on_pass_2(hs, emu_lib, oracle, n, swap, staged, monkeypatch):
    """Long forward real transforms (hs_tiled.cuh IO 8 / 9): two launches, no split kernel; full spectrum with an exact
    Hermitian mirror, and the half spectrum (bins 0..N/2) with a row stride larger than N/2+1 straight through the
    device-pointer entry point (host arrays stand in for device memory under the emulation).  Every split the planner
    picks (128 x 256 ... 1024 x 256, 1024 x 512 with the second-pass 512-point plan, 1024 x 1024); `staged`: the first
    pass on the persistent TMA-staged kernel (what the GPU runs for n1 = 512 / 1024), with the split as its epilogue."""
    import ctypes as C
    if swap:
        monkeypatch.setenv("HSFFT_R4_SWAP", "1")
    monkeypatch.setenv("HSFFT_R4_TMA", "1" if staged else "0")
    b = 2 if n <= (1 << 17) else 1
    x = xorshift_uniform(b * n, seed=3 + n).reshape(b, n)
    Xr = oracle.r2c(x)
    with hs.RealPlan(n, 1, lib=emu_lib) as f:
        l0 = emu_lib.hsfft_launch_count()
        X = f.r2c(x)
        assert emu_lib.hsfft_launch_count() - l0 == 2
        assert max(rel_rms(X[i], Xr[i]) for i in range(b)) <= min(tol(n), 5e-15)
        assert np.array_equal(X[:, n - 1: n // 2: -1], np.conj(X[:, 1: n // 2]))
        stride = n // 2 + 5
        half = np.full((b, stride), np.nan + 0j)
        emu_lib.fft_r2c_exec_batch_dev(f.p, x.ctypes.data_as(C.c_void_p), half.ctypes.data_as(C.c_void_p), stride, 0, b, None)
        assert np.array_equal(half[:, : n // 2 + 1], X[:, : n // 2 + 1]) and np.isnan(half[:, n // 2 + 1:]).all()


@pytest.mark.parametrize("n,stage", [(1024, "0"), (1024, "2"), (4096, "0"), (4096, "2"), (6000, "2")])
def test_one_kernel_c2r_merge_direct_and_from_a_staged_row(hs, emu_lib, oracle, n, stage, monkeypatch):
    """hs_c2r_smem_kernel<..., STG>: the merge either loads bins k and H-k straight from global memory (HSFFT_C2R_STAGE=0)
    or reads a copy of bins 0..H-1 staged in the team's buffer by asynchronous copies, bin H from global memory (=2: for
    every plan; the default picks per plan).  Arbitrary complex bins: the reference's merge defines the result."""
    monkeypatch.setenv("HSFFT_C2R_STAGE", stage)
    X = rand_cplx(5, n, seed=n + int(stage))
    with hs.RealPlan(n, -1, lib=emu_lib) as g:
        y = g.c2r(X)
    yr = oracle.c2r(X, n)
    assert max(rel_rms(y[i], yr[i]) for i in range(5)) <= min(tol(n), 5e-15)


@pytest.mark.parametrize("la,lb", [(5, 3), (100, 37), (512, 512), (1000, 24)])
def test_convolution(hs, emu_lib, oracle, la, lb):
    a, b = xorshift_uniform(la, seed=la), xorshift_uniform(lb, seed=lb + 1)
    for t in ("full", "same", "valid"):
        n, y = hs.convolve(a, b, t, "linear", lib=emu_lib)
        nr, yr = oracle.convolve(a, b, t, "linear")
        assert n == nr and rel_rms(y, yr) <= 1e-14
    n, y = hs.convolve(a, b, "full", "circular", lib=emu_lib)
    nr, yr = oracle.convolve(a, b, "full", "circular")
    assert n == nr and rel_rms(y, yr) <= 1e-14
    assert hs.convolve(a, b, "bogus", "linear", lib=emu_lib)[0] == -1
    assert hs.convolve(a, b, "full", "nope", lib=emu_lib)[0] == -1


def test_long_convolution_fused_padding_product_and_slice(hs, emu_lib, oracle):
    """P >= 2^15: the zero padding rides on pass 1's loads, the product and 1/P on the merge loads of the inverse
    transform, the slice on its stores.  Odd lengths: rows of a batch are then only 8-byte aligned."""
    la, lb = 20001, 12999                       # conv_length 32999 -> P = 65536
    a = xorshift_uniform(2 * la, seed=3).reshape(2, la)
    b = xorshift_uniform(2 * lb, seed=4).reshape(2, lb)
    c0 = hs.launch_count(lib=emu_lib) if hasattr(hs, "launch_count") else 0
    n, y = hs.convolve_batch(a, b, lib=emu_lib)
    assert n == la + lb - 1
    for i in range(2):
        nr, yr = oracle.convolve(a[i], b[i], "full", "linear")
        assert nr == n and rel_rms(y[i], yr) <= 1e-13
    for t in ("same", "valid"):
        n, y1 = hs.convolve(a[0], b[0], t, "linear", lib=emu_lib)
        nr, yr = oracle.convolve(a[0], b[0], t, "linear")
        assert n == nr and rel_rms(y1, yr) <= 1e-13, t


@pytest.mark.parametrize("Lx,Lh,P", [(100000, 1001, 32768), (70001, 333, 0), (40000, 16000, 32768)])
def test_overlap_save_streaming_convolution(hs, emu_lib, Lx, Lh, P):
    """fft_convolve_stream: windows read in place (zero outside the signal), one shared filter spectrum, each block's
    valid samples stored straight into y; the last block is clipped."""
    x, h = xorshift_uniform(Lx, seed=Lx), xorshift_uniform(Lh, seed=Lh)
    n, y = hs.convolve_stream(x, h, P, lib=emu_lib)
    assert n == Lx + Lh - 1 and rel_rms(y, np.convolve(x, h)) <= 1e-13
    assert hs.convolve_stream(x, xorshift_uniform(20000, 1), 32768, lib=emu_lib)[0] == -1   # filter longer than half a window


def test_convolution_batch(hs, emu_lib):
    a = xorshift_uniform(4 * 50, 3).reshape(4, 50)
    b = xorshift_uniform(4 * 20, 4).reshape(4, 20)
    n, y = hs.convolve_batch(a, b, lib=emu_lib)
    assert n == 69
    for i in range(4):
        assert rel_rms(y[i], np.convolve(a[i], b[i])) <= 1e-14


def test_batch_tail_and_chunking(hs, emu_lib, oracle):
    """Batches that do not fill the last CTA (X transforms per CTA) and single transforms."""
    for n, howmany in ((1024, 7), (256, 19), (64, 33)):
        x = rand_cplx(howmany, n, seed=n)
        with hs.Plan(n, 1, lib=emu_lib) as p:
            assert rel_rms(p.exec(x), oracle.c2c(x, 1)) <= tol(n)


def test_plan_address_reuse_rebuilds_device_state(hs, emu_lib, oracle):
    """Callers may free() a plan directly (reference src/real.c:264); a new plan at the same address
    with other (N, sgn) must not inherit the old device tables."""
    import ctypes as C
    libc = C.CDLL(None)
    libc.free.argtypes = [C.c_void_p]
    seen = 0
    for n in (64, 128, 64, 256, 128):
        p = hs.Plan(n, 1, lib=emu_lib)
        x = rand_cplx(1, n, seed=n + seen)[0]
        assert rel_rms(p.exec(x), oracle.c2c(x, 1)) <= tol(n)
        libc.free(C.cast(p.p, C.c_void_p))      # bare free: the registry entry goes stale
        p.p = None
        seen += 1


def oracle_fft2(oracle, x, sgn):
    """2-D reference composed from the reference's own 1-D transform: rows, then columns."""
    b, n0, n1 = x.shape
    r1 = oracle.c2c(np.ascontiguousarray(x.reshape(b * n0, n1)), sgn).reshape(b, n0, n1)
    r2 = oracle.c2c(np.ascontiguousarray(r1.transpose(0, 2, 1).reshape(b * n1, n0)), sgn).reshape(b, n1, n0)
    return np.ascontiguousarray(r2.transpose(0, 2, 1))


@pytest.mark.parametrize("n0,n1", [(128, 64), (256, 48), (100, 36), (37, 50), (512, 16)])
def test_two_dimensional_transforms(hs, emu_lib, oracle, n0, n1):
    """fft2_*: rows as batched 1-D transforms; columns by the strided tiled pass (n0 = 128..1024, n1 % 16 == 0) or through
    a transpose.  Checked against the reference's 1-D transform applied along both axes."""
    for sgn in (1, -1):
        x = rand_cplx(2, n0 * n1, seed=n0 + n1 + sgn).reshape(2, n0, n1)
        with hs.Plan2D(n0, n1, sgn, lib=emu_lib) as p:
            y = p.exec(x)
            desc = p.describe()
        assert ("no transpose" in desc) == (n0 in (128, 256, 512, 1024) and n1 % {1024: 4, 512: 8}.get(n0, 16) == 0), desc
        assert rel_rms(y, oracle_fft2(oracle, x, sgn)) <= tol(n0 * n1), (n0, n1, sgn, desc)


def test_batch_sharding_over_devices_and_plans_shared_between_threads(hs, emu_lib, oracle, monkeypatch):
    """hsfft_set_devices(n): the host-pointer batch calls cut the batch into n contiguous ranges, one host thread and one
    (plan, device) state each (the emulation pretends to have 3 devices).  Also several host threads executing ONE plan at
    once, which the reference allows for mixed-radix plans: calls serialise on the plan's lock instead of sharing staging buffers."""
    import threading
    monkeypatch.setenv("HSFFT_EMU_DEVICES", "3")
    emu_lib.hsfft_set_devices(3)
    try:
        assert emu_lib.hsfft_get_devices() == 3
        n = 1000
        x = rand_cplx(11, n, seed=4)
        with hs.Plan(n, 1, lib=emu_lib) as p:
            y = p.exec(x)                          # 11 transforms over 3 devices: 4 + 4 + 3
            assert max(rel_rms(y[i], oracle.c2c(x, 1)[i]) for i in range(11)) <= tol(n)
        nr = 2048
        xr = xorshift_uniform(7 * nr, seed=8).reshape(7, nr)
        with hs.RealPlan(nr, 1, lib=emu_lib) as f, hs.RealPlan(nr, -1, lib=emu_lib) as g:
            X = f.r2c(xr)
            assert rel_rms(X, oracle.r2c(xr)) <= tol(nr)
            assert rel_rms(g.c2r(X) / nr, xr) <= tol(nr)
        a, b = xorshift_uniform(6 * 300, seed=1).reshape(6, 300), xorshift_uniform(6 * 41, seed=2).reshape(6, 41)
        cnt, yc = hs.convolve_batch(a, b, lib=emu_lib)
        assert cnt == 340 and all(rel_rms(yc[i], np.convolve(a[i], b[i])) <= 1e-13 for i in range(6))
    finally:
        emu_lib.hsfft_set_devices(1)
    n = 4096
    x = rand_cplx(4, n, seed=9)
    ref = oracle.c2c(x, 1)
    out, errs = [None] * 4, []
    with hs.Plan(n, 1, lib=emu_lib) as p:
        def work(i):
            try:
                for _ in range(3):
                    out[i] = p.exec(x[i])
            except Exception as e:   # pragma: no cover
                errs.append(e)
        th = [threading.Thread(target=work, args=(i,)) for i in range(4)]
        [t.start() for t in th]
        [t.join() for t in th]
    assert not errs and all(rel_rms(out[i], ref[i]) <= tol(n) for i in range(4))


@pytest.mark.parametrize("ndev,N1,N2,sgn", [(1, 256, 512, 1), (2, 256, 512, -1), (4, 16384, 128, 1)])
def test_single_process_four_step_entry(hs, emu_lib, oracle, ndev, N1, N2, sgn, monkeypatch):
    """fft_exec_dist: host pointers, natural order, one process driving `ndev` devices whose receive buffers are mutually
    addressable (peer access on the GPU box, one address space here): scatter, phase A with the exchange fused into its
    stores, barrier, phase B, transpose, gather."""
    import ctypes as C
    monkeypatch.setenv("HSFFT_EMU_DEVICES", str(ndev))
    x = rand_cplx(1, N1 * N2, seed=77 + ndev)[0]
    y = np.empty_like(x)
    assert emu_lib.fft_exec_dist(N1, N2, sgn, x.ctypes.data_as(C.c_void_p), y.ctypes.data_as(C.c_void_p), ndev) == 0
    assert rel_rms(y, oracle.c2c(x, sgn)) <= tol(N1 * N2)


@pytest.mark.parametrize("n", [64, 1000, 4096, 1009, 10007, 4099, 16384, 6561, 20011, 3 << 14])
def test_in_place_execution(hs, emu_lib, oracle, n):
    """fft_exec_batch_inplace_dev (additive): one buffer in, the same buffer out, for every kernel family -- the plans that
    cannot run on a single buffer (run-time multi-pass, split chirp-z) go through a copy of the input."""
    import ctypes as C
    x = rand_cplx(3, n, seed=n)
    buf = x.copy()
    with hs.Plan(n, 1, lib=emu_lib) as p:
        emu_lib.fft_exec_batch_inplace_dev(p.p, buf.ctypes.data_as(C.c_void_p), 3, None)
        desc = p.describe()
    r = oracle.c2c(x, 1)
    assert max(rel_rms(buf[i], r[i]) for i in range(3)) <= tol(n), (n, desc)


@pytest.mark.parametrize("r", [11, 13, 17])
def test_odd_radix_butterflies_with_compile_time_tables(emu_lib, r):
    """hs_bfly<11 / 13 / 17>: the reference's generic butterfly unrolled with literal cos / sin tables (plan-time kernels
    for lengths with these factors): a plain r-point DFT to rounding, both directions."""
    import ctypes as C
    x = rand_cplx(1, r, seed=r)[0]
    k = np.arange(r)
    for sgn in (1, -1):
        v = x.copy()
        assert emu_lib.hs_emu_bfly_odd(r, sgn, v.ctypes.data_as(C.c_void_p)) == 0
        ref = np.exp(-2j * np.pi * sgn * np.outer(k, k) / r) @ x
        assert rel_rms(v, ref) <= 5e-16 * r

