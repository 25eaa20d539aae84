"""GPU parity tests: the CUDA path, called through the C ABI (libhsfft_b200.so), against the oracle
on the same seeded inputs.  Tolerance = north_star's: relative RMS <= 1e-12 * log2(N), forward and
inverse, per transform.  Run on the B200 box with `pytest -m gpu`.
"""
import math
import os

import numpy as np
import pytest

import oracle_lib
from oracle_lib import rand_cplx, rel_rms, tol, xorshift_uniform

pytestmark = pytest.mark.gpu


def per_transform_worst(y, r):
    return max(rel_rms(y[i], r[i]) for i in range(y.shape[0]))


# BASELINE.json config sizes + the reference's own self-test sizes (src/real.c:278,360) + edge sizes
C2C_SIZES = [2, 3, 4, 5, 7, 8, 15, 16, 20, 64, 100, 128, 256, 512, 1000, 1024, 2048, 3000, 4096, 8192,
             11, 13, 26, 143, 1001, 53 * 8, 17 * 29, 6561, 10000, 16384, 32768, 1 << 17, 1 << 19, 1 << 20, 3 << 14]
BLUESTEIN_SIZES = [19, 59, 67, 131, 257, 509, 513, 1009, 1031, 2018, 2039, 2053, 4093, 4099, 8209, 10007, 10239, 40009, 100003, 300007]


@pytest.mark.parametrize("n", C2C_SIZES)
@pytest.mark.parametrize("sgn", [1, -1])
def test_c2c_mixed_radix(hs, product_lib, oracle, n, sgn):
    howmany = 5 if n <= 8192 else 2
    x = rand_cplx(howmany, n, seed=1234 + n)
    with hs.Plan(n, sgn) as p:
        assert p.lt == 0
        y = p.exec(x)
        y1 = p.exec(x[0])  # the reference-style single fft_exec call
    r = oracle.c2c(x, sgn)
    assert per_transform_worst(y, r) <= tol(n), p.describe()
    assert rel_rms(y1, r[0]) <= tol(n)


@pytest.mark.parametrize("n", BLUESTEIN_SIZES)
@pytest.mark.parametrize("sgn", [1, -1])
def test_c2c_bluestein(hs, product_lib, oracle, n, sgn):
    howmany = 3
    x = rand_cplx(howmany, n, seed=99 + n)
    with hs.Plan(n, sgn) as p:
        assert p.lt == 1
        y = p.exec(x)
    r = oracle.c2c(x, sgn)
    assert per_transform_worst(y, r) <= tol(n)


@pytest.mark.parametrize("n", [1009, 257])
@pytest.mark.parametrize("sgn", [1, -1])
def test_c2c_prime_rader_and_bluestein_routes_agree(hs, product_lib, oracle, n, sgn):
    """N = 1009 / 257 run as a cyclic convolution of length N-1 (Rader) by default and as the reference's padded chirp-z
    with HSFFT_NO_RADER=1 (read at plan build); both must match the reference, and each other far inside the tolerance."""
    x = rand_cplx(37, n, seed=7 + n)
    with hs.Plan(n, sgn) as p:
        d1, y1 = p.describe(), p.exec(x)
    os.environ["HSFFT_NO_RADER"] = "1"
    try:
        with hs.Plan(n, sgn) as p:
            d2, y2 = p.describe(), p.exec(x)
    finally:
        del os.environ["HSFFT_NO_RADER"]
    assert "rader" in d1 and "bluestein" in d2, (d1, d2)
    r = oracle.c2c(x, sgn)
    assert per_transform_worst(y1, r) <= tol(n) and per_transform_worst(y2, r) <= tol(n)
    assert per_transform_worst(y1, y2) <= 1e-14


@pytest.mark.parametrize("n", [1024, 3000, 4096, 1009])
def test_generic_engine_matches_compile_time_kernels(hs, product_lib, oracle, n):
    """Same sizes through the run-time planned engine (HSFFT_NO_CT=1 is read at plan build)."""
    x = rand_cplx(3, n, seed=5 + n)
    os.environ["HSFFT_NO_CT"] = "1"
    try:
        with hs.Plan(n, 1) as p:
            desc = p.describe()
            y = p.exec(x)
    finally:
        del os.environ["HSFFT_NO_CT"]
    assert "compile-time" not in desc
    assert per_transform_worst(y, oracle.c2c(x, 1)) <= tol(n)


JIT_SIZES = [24, 96, 360, 640, 1000, 1001, 1144, 1536, 1800, 2000, 2002, 2187, 2400, 2401, 2560, 3072, 3125, 3600, 4500, 5120, 5632, 5000, 6000, 6250, 6561, 6912, 7000, 8000]


@pytest.mark.parametrize("n", JIT_SIZES)
def test_plan_time_compiled_kernels(hs, product_lib, oracle, n):
    """Lengths without a static plan get a kernel compiled with NVRTC at plan creation (csrc/hs_jit.cu): same stages,
    butterflies and literals as the static plans, so the result is within rounding of the reference."""
    for sgn in (1, -1):
        x = rand_cplx(7, n, seed=n + sgn)
        with hs.Plan(n, sgn) as p:
            desc = p.describe()
            y = p.exec(x)
        assert "NVRTC" in desc or n in (6561,), desc   # 3^8 needs 729 threads of 9 points: left to the run-time engine
        assert per_transform_worst(y, oracle.c2c(x, sgn)) <= min(tol(n), 2e-15 * math.log2(n) * 4), (n, sgn, desc)


@pytest.mark.parametrize("n", [4000, 4500, 6250, 6561])   # (N = 2400 runs three CTAs per SM since the fat-thread rule: no persistent twin)
def test_plan_time_kernels_persistent_staged_form(hs, product_lib, oracle, n, monkeypatch):
    """Batches of a few waves and more run the plan-time kernel as persistent CTAs whose next input is staged by the copy
    engine (csrc/hs_smem_fft.cuh HsStageRefill; HSFFT_JIT_STAGED=3 takes that form without the plan-time measurement):
    first / last transform, the ones around the end of the first wave and the tail against the oracle."""
    monkeypatch.setenv("HSFFT_JIT_STAGED", "3")
    b = 2500
    idx = [0, 1, 147, 148, 149, 295, 296, 297, 1183, 1184, 1185, b // 2, b - 2, b - 1]
    import torch
    dev = torch.device("cuda", 0)
    for sgn in (1, -1):
        x = rand_cplx(b, n, seed=n - sgn)
        xd = torch.from_numpy(x).to(dev)
        yd = torch.empty_like(xd)
        with hs.Plan(n, sgn) as p:
            with hs.launch_timing() as t:
                p.exec_dev(xd, yd, b)
                torch.cuda.synchronize()
        assert any("staged" in k[0] for k in t.kernels), t.kernels
        y = yd.cpu().numpy()
        assert per_transform_worst(y[idx], oracle.c2c(x[idx], sgn)) <= tol(n), (n, sgn)


def test_config1_n1024_forward_inverse_roundtrip(hs, product_lib, oracle):
    """BASELINE.json config 1: N=1024, forward then inverse, /N round trip."""
    n = 1024
    x = rand_cplx(1, n)[0]
    with hs.Plan(n, 1) as f, hs.Plan(n, -1) as b:
        X = f.exec(x)
        back = b.exec(X) / n
    assert rel_rms(X, oracle.c2c(x, 1)) <= tol(n)
    assert rel_rms(back, x) <= tol(n)


def test_linearity_and_parseval_full_batch(hs, product_lib):
    """Size-independent properties at the full BASELINE batch (device resident, N=4096 x 65536)."""
    import torch
    n, batch = 4096, 65536
    dev = torch.device("cuda", 0)
    g = torch.Generator(device=dev); g.manual_seed(1)
    x = torch.rand((batch, n), generator=g, device=dev, dtype=torch.float64).to(torch.complex128)
    x = x + 1j * torch.rand((batch, n), generator=g, device=dev, dtype=torch.float64)
    y = torch.empty_like(x)
    with hs.Plan(n, 1) as p:
        p.exec_dev(x, y, batch)
        torch.cuda.synchronize()
        # Parseval per transform
        ex = (x.abs() ** 2).sum(dim=1)
        ey = (y.abs() ** 2).sum(dim=1) / n
        assert float(((ex - ey).abs() / ex).max()) < 1e-12
        # first / last / random transforms against torch's own fp64 FFT (independent cross-check)
        idx = torch.tensor([0, 1, 777, 40000, batch - 1], device=dev)
        ref = torch.fft.fft(x[idx], dim=1)
        err = (y[idx] - ref).abs().pow(2).sum(dim=1).sqrt() / ref.abs().pow(2).sum(dim=1).sqrt()
        assert float(err.max()) < 1e-12 * math.log2(n)
        # inverse round trip of the whole batch
        z = torch.empty_like(x)
        with hs.Plan(n, -1) as q:
            q.exec_dev(y, z, batch)
        torch.cuda.synchronize()
        rt = ((z / n - x).abs().pow(2).sum(dim=1).sqrt() / x.abs().pow(2).sum(dim=1).sqrt()).max()
        assert float(rt) < 1e-12 * math.log2(n)


@pytest.mark.parametrize("n", [4, 8, 16, 32, 64, 2018, 4096, 4800, 6000, 8192, 10000, 16384, 1 << 15, 1 << 20])   # 4800 / 10000: plan-time kernels (at 10000 the c2r kernel runs on its own plan)
def test_real_r2c_c2r(hs, product_lib, oracle, n):
    x = xorshift_uniform(3 * n, seed=n).reshape(3, n)
    with hs.RealPlan(n, 1) as f, hs.RealPlan(n, -1) as g:
        X = f.r2c(x)
        Xr = oracle.r2c(x)
        assert per_transform_worst(X, Xr) <= tol(n)
        # Hermitian mirror written in full (reference src/real.c:128-132)
        assert np.array_equal(X[:, n - 1: n // 2: -1], np.conj(X[:, 1: n // 2]))
        y = g.c2r(X)
        yr = oracle.c2r(Xr, n)
        assert per_transform_worst(y, yr) <= tol(n)
        # round trip: as good as the reference's own (its 11-digit radix-3/5 constants put N = 4800, 6000, 10000 above tol(n))
        assert rel_rms(y / n, x) <= max(tol(n), 1.05 * rel_rms(yr / n, x))


@pytest.mark.parametrize("la,lb", [(5, 3), (100, 37), (512, 512), (1000, 24), (4096, 4096), (20001, 12999), (300000, 200001), (524288, 524288)])
@pytest.mark.parametrize("type_", ["full", "same", "valid"])
def test_convolve_linear(hs, product_lib, oracle, la, lb, type_):
    a = xorshift_uniform(la, seed=la)
    b = xorshift_uniform(lb, seed=lb + 1)
    n, y = hs.convolve(a, b, type_, "linear")
    nr, yr = oracle.convolve(a, b, type_, "linear")
    assert n == nr
    assert rel_rms(y, yr) <= 1e-12 * math.log2(max(la + lb, 4))


def test_convolve_batch_fused_path_matches_unfused(hs, product_lib):
    """Batched long convolution (BASELINE config 4): the fused path (padding / product / slice on the transforms' own
    loads and stores) against the unfused kernel sequence, odd lengths so that rows are only 8-byte aligned."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = ("import importlib,sys,numpy as np;sys.path.insert(0,%r);sys.path.insert(0,%r+'/tests');"
            "hs=importlib.import_module('mixed-radix-fast-fourier-transform_b200');from oracle_lib import xorshift_uniform as xu;"
            "a=xu(5*70001,1).reshape(5,70001);b=xu(5*33333,2).reshape(5,33333);n,y=hs.convolve_batch(a,b);"
            "np.save(sys.argv[1],y);print(n)") % (root, root)
    outs = []
    for fuse in ("1", "0"):
        env = dict(os.environ)
        if fuse == "0":
            env["HSFFT_NO_CONV_FUSE"] = "1"
        f = "/tmp/hsfft_conv_%s.npy" % fuse
        r = subprocess.run([sys.executable, "-c", code, f], capture_output=True, text=True, env=env, timeout=300)
        assert r.returncode == 0 and r.stdout.strip() == str(70001 + 33333 - 1), r.stdout + r.stderr[-800:]
        outs.append(np.load(f))
    assert rel_rms(outs[0], outs[1]) <= 1e-14
    a = xorshift_uniform(5 * 70001, 1).reshape(5, 70001)
    b = xorshift_uniform(5 * 33333, 2).reshape(5, 33333)
    for i in (0, 4):
        assert rel_rms(outs[0][i], np.convolve(a[i], b[i])) <= 1e-12


@pytest.mark.parametrize("Lx,Lh,P", [(3000001, 4097, 0), (1 << 22, 12345, 65536), (500000, 40000, 1 << 17)])
def test_overlap_save_streaming_convolution(hs, product_lib, Lx, Lh, P):
    """Additive API: long signal, short filter, checked against numpy's FFT convolution."""
    x, h = xorshift_uniform(Lx, seed=Lx), xorshift_uniform(Lh, seed=Lh)
    n, y = hs.convolve_stream(x, h, P)
    nfft = 1 << int(np.ceil(np.log2(Lx + Lh - 1)))
    ref = np.fft.irfft(np.fft.rfft(x, nfft) * np.fft.rfft(h, nfft), nfft)[: Lx + Lh - 1]
    assert n == Lx + Lh - 1 and rel_rms(y, ref) <= 1e-12


@pytest.mark.parametrize("n0,n1", [(128, 64), (256, 3000), (1024, 1024), (100, 36), (1009, 250), (512, 1000)])
def test_two_dimensional_transforms(hs, product_lib, oracle, n0, n1):
    """Additive API: 2-D transforms against the reference's 1-D transform along both axes."""
    b = 2
    for sgn in (1, -1):
        x = rand_cplx(b, n0 * n1, seed=n0 + n1 + sgn).reshape(b, n0, n1)
        with hs.Plan2D(n0, n1, sgn) as p:
            y = p.exec(x)
        r1 = oracle.c2c(np.ascontiguousarray(x.reshape(b * n0, n1)), sgn).reshape(b, n0, n1)
        r2 = oracle.c2c(np.ascontiguousarray(r1.transpose(0, 2, 1).reshape(b * n1, n0)), sgn).reshape(b, n1, n0)
        assert rel_rms(y, r2.transpose(0, 2, 1)) <= tol(n0 * n1), (n0, n1, sgn)


def test_convolve_circular_and_errors(hs, product_lib, oracle):
    a = xorshift_uniform(100, seed=1)
    b = xorshift_uniform(37, seed=2)
    n, y = hs.convolve(a, b, "full", "circular")
    nr, yr = oracle.convolve(a, b, "full", "circular")
    assert n == nr == 128 and rel_rms(y, yr) <= 1e-11
    assert hs.convolve(a, b, "bogus", "linear")[0] == -1
    assert hs.convolve(a, b, "full", "nope")[0] == -1


def test_kernels_really_launched(hs, product_lib):
    before = hs.launch_count()
    with hs.Plan(4096, 1) as p:
        p.exec(rand_cplx(2, 4096))
    assert hs.launch_count() > before


def test_four_step_distributed_phases_single_gpu(hs, product_lib, oracle):
    """hsfft_dist_* local phases on one GPU (world = 1: the exchange is a copy); N = N1*N2 = 2^17 and 2^28-style split."""
    import importlib
    import torch
    dmod = importlib.import_module("mixed-radix-fast-fourier-transform_b200.dist")
    dev = torch.device("cuda", 0)
    for N1, N2, sgn in ((256, 512, 1), (16384, 128, -1), (128, 16384, 1)):
        N = N1 * N2
        x = rand_cplx(1, N, seed=N1 + N2)[0]
        plan = dmod.DistPlan(hs, N1, N2, sgn, rank=0, world=1, device=dev)
        a = torch.view_as_real(torch.from_numpy(dmod.scatter_input(x, N1, N2, 0, 1))).reshape(-1, 2).contiguous().to(dev)
        out = torch.empty((N1 * N2, 2), dtype=torch.float64, device=dev)
        plan.forward(a, out)
        torch.cuda.synchronize()
        X = dmod.gather_output([torch.view_as_complex(out).reshape(N1, N2).cpu().numpy()], N1, N2)
        assert rel_rms(X, oracle.c2c(x, sgn)) <= tol(N), (N1, N2)
        plan.close()


@pytest.mark.parametrize("N1,N2,world,sgn", [(256, 512, 2, 1), (16384, 256, 4, -1), (32768, 1024, 8, 1)])
def test_four_step_exchange_fused_into_phase_a_single_gpu(hs, product_lib, oracle, N1, N2, world, sgn):
    """The peer-store exchange (phase A's last pass twiddles and stores each row into the owning rank's receive buffer):
    every rank's plan driven from this process on one GPU, so "peer" buffers are ordinary device pointers."""
    import importlib
    import torch
    dmod = importlib.import_module("mixed-radix-fast-fourier-transform_b200.dist")
    x = rand_cplx(1, N1 * N2, seed=N1 + N2 + world)[0]
    X = dmod.local_group_forward(hs, x, N1, N2, sgn, world, device=torch.device("cuda", 0))
    assert rel_rms(X, oracle.c2c(x, sgn)) <= tol(N1 * N2)


def test_four_step_p2p_world1_matches_all_to_all_path(hs, product_lib, oracle):
    import importlib
    import torch
    dmod = importlib.import_module("mixed-radix-fast-fourier-transform_b200.dist")
    dev = torch.device("cuda", 0)
    N1, N2 = 16384, 512
    x = rand_cplx(1, N1 * N2, seed=5)[0]
    plan = dmod.DistPlan(hs, N1, N2, 1, rank=0, world=1, device=dev).enable_p2p()
    a = torch.view_as_real(torch.from_numpy(dmod.scatter_input(x, N1, N2, 0, 1))).reshape(-1, 2).contiguous().to(dev)
    out = torch.empty((N1 * N2, 2), dtype=torch.float64, device=dev)
    for _ in range(3):   # alternates the two receive buffers
        out.zero_()
        plan.forward_p2p(a, out)
        torch.cuda.synchronize()
        X = dmod.gather_output([torch.view_as_complex(out).reshape(N1, N2).cpu().numpy()], N1, N2)
        assert rel_rms(X, oracle.c2c(x, 1)) <= tol(N1 * N2)
    plan.close()


def test_four_step_p2p_two_processes_two_gpus():
    """The real thing: two processes, two GPUs, receive buffers mapped through CUDA IPC, phase A storing into the peer
    over NVLink.  Skipped on a one-GPU box (tools/dist_bench.py --p2p is the same path, verified at 2^28 / 2^30)."""
    import subprocess
    import sys
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", "29541", os.path.join(root, "tools", "dist_bench.py"), "--log2n", "28", "--steps", "1", "--p2p"],
                       capture_output=True, text=True, timeout=600)
    assert '"ok": true' in r.stdout, r.stdout[-1500:] + r.stderr[-1500:]


@pytest.mark.parametrize("tma", ["0", "1"])
def test_tiled_passes_with_and_without_tma(tma):
    """The plain and the TMA-staged tiled kernels must agree with the oracle on every two-pass size class."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = r'''
import importlib, sys
sys.path.insert(0, %r); sys.path.insert(0, %r + "/tests")
import oracle_lib as ol
hs = importlib.import_module("mixed-radix-fast-fourier-transform_b200")
orc = ol.Oracle()
for n in (1 << 14, 1 << 15, 1 << 17, 1 << 19, 1 << 20):
    x = ol.rand_cplx(2, n, seed=n)
    with hs.Plan(n, -1) as p:
        y = p.exec(x)
    e = ol.rel_rms(y, orc.c2c(x, -1))
    assert e <= ol.tol(n), (n, e)
print("TILED_OK")
''' % (root, root)
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=dict(os.environ, HSFFT_TMA=tma), timeout=600)
    assert "TILED_OK" in r.stdout, r.stdout[-1500:] + r.stderr[-1500:]


# ------------------------------------------------------------------------------------------------------------------
# Parity at the full BASELINE batches (SURVEY.md section 8d "verification at scale"): the whole batch runs device
# resident through the C ABI; a >= 1 % random sample of the transforms, including the first and the last one (the
# tail CTA, the last chunk of a chunked plan), is compared with the oracle on the same inputs.

def _sample_indices(batch, frac=0.01, minimum=3, seed=0):
    rng = np.random.default_rng(seed)
    k = max(minimum, int(math.ceil(batch * frac)))
    idx = set(rng.choice(batch, size=min(k, batch), replace=False).tolist()) | {0, batch - 1}
    return np.array(sorted(idx), dtype=np.int64)


def _device_rand(shape, seed):
    import torch
    g = torch.Generator(device="cuda")
    g.manual_seed(seed)
    return torch.rand(shape, generator=g, device="cuda", dtype=torch.float64) * 2 - 1


@pytest.mark.parametrize("n,batch", [(3000, 65536), (1009, 16384), (10007, 16384), (4096, 65536)])
@pytest.mark.parametrize("sgn", [1, -1])
def test_full_batch_c2c_sample_against_oracle(hs, product_lib, oracle, n, batch, sgn):
    import torch
    x = _device_rand((batch, n, 2), seed=n + sgn)
    y = torch.empty_like(x)
    with hs.Plan(n, sgn) as p:
        p.exec_dev(x, y, batch)
        torch.cuda.synchronize()
        desc = p.describe()
    idx = _sample_indices(batch, seed=n)
    ti = torch.from_numpy(idx).cuda()
    xs = torch.view_as_complex(x[ti].contiguous()).cpu().numpy()
    ys = torch.view_as_complex(y[ti].contiguous()).cpu().numpy()
    r = oracle.c2c(xs, sgn)
    assert len(idx) >= batch // 100 and idx[0] == 0 and idx[-1] == batch - 1
    assert per_transform_worst(ys, r) <= tol(n), (n, sgn, desc)


@pytest.mark.parametrize("n,batch", [(1 << 20, 256), (8192, 32768), (2048, 131072)])
def test_full_batch_real_sample_against_oracle(hs, product_lib, oracle, n, batch):
    import torch
    x = _device_rand((batch, n), seed=n)
    X = torch.empty((batch, n, 2), device="cuda", dtype=torch.float64)
    back = torch.empty((batch, n), device="cuda", dtype=torch.float64)
    h1 = n // 2 + 1
    with hs.RealPlan(n, 1) as f, hs.RealPlan(n, -1) as g:
        f.r2c_dev(x, X, n, 1, batch)
        Xh = X[:, :h1, :].contiguous()            # rows of N/2+1 bins: the c2r input layout with in_stride = N/2+1
        g.c2r_dev(Xh, h1, back, batch)
        torch.cuda.synchronize()
    idx = _sample_indices(batch, seed=n)
    ti = torch.from_numpy(idx).cuda()
    xs = x[ti].cpu().numpy()
    Xs = torch.view_as_complex(X[ti].contiguous()).cpu().numpy()
    Xr = oracle.r2c(xs)
    assert per_transform_worst(Xs, Xr) <= tol(n)
    assert np.array_equal(Xs[:, n - 1: n // 2: -1], np.conj(Xs[:, 1: n // 2]))
    bs = back[ti].cpu().numpy()
    assert per_transform_worst(bs, oracle.c2r(Xs, n)) <= tol(n)
    # round trip of the WHOLE batch on the device (size-independent property)
    rt = ((back / n - x).pow(2).sum(dim=1).sqrt() / x.pow(2).sum(dim=1).sqrt()).max()
    assert float(rt) <= tol(n)


def test_full_batch_convolution_sample_against_oracle(hs, product_lib, oracle):
    """BASELINE config 4: 2^19 (*) 2^19 samples, 64 rows, device resident through fft_convolve_batch_dev."""
    import ctypes as C
    import torch
    L = hs.load()
    la = lb = 1 << 19
    batch = 64
    a, b = _device_rand((batch, la), seed=1), _device_rand((batch, lb), seed=2)
    y = torch.zeros((batch, la + lb), device="cuda", dtype=torch.float64)
    n = L.fft_convolve_batch_dev(b"full", b"linear", C.c_void_p(a.data_ptr()), la, C.c_void_p(b.data_ptr()), lb, C.c_void_p(y.data_ptr()),
                                 la + lb, batch, None)
    torch.cuda.synchronize()
    assert n == la + lb - 1
    for i in (0, 31, batch - 1):
        nr, yr = oracle.convolve(a[i].cpu().numpy(), b[i].cpu().numpy(), "full", "linear")
        assert nr == n and rel_rms(y[i, :n].cpu().numpy(), yr) <= 1e-12 * 20


# ------------------------------------------------------------------------------------------------------------------
# The CUDA path against the REFERENCE'S OWN CODE, with no port in between: oracle/_ref (the reference's three .c files
# compiled where they lie; travels to the GPU box as a binary) and the committed outputs of that same build.

@pytest.mark.parametrize("n", [1024, 3000, 4096, 1009, 10007, 2018, 143])
def test_cuda_path_against_the_compiled_reference(hs, product_lib, ref_repaired, n):
    x = rand_cplx(3, n, seed=31 + n)
    for sgn in (1, -1):
        with hs.Plan(n, sgn) as p:
            y = p.exec(x)
        r = ref_repaired.c2c(x, sgn)
        assert per_transform_worst(y, r) <= tol(n), (n, sgn)
    if n % 2 == 0:
        xr = xorshift_uniform(2 * n, seed=n).reshape(2, n)
        with hs.RealPlan(n, 1) as f, hs.RealPlan(n, -1) as g:
            X = f.r2c(xr)
            Xr = ref_repaired.r2c(xr)
            assert per_transform_worst(X, Xr) <= tol(n)
            assert per_transform_worst(g.c2r(X), ref_repaired.c2r(Xr, n)) <= tol(n)


def test_cuda_path_against_the_committed_reference_outputs(hs, product_lib):
    """tests/golden/ref_vectors.npz was written by tests/golden/make_golden.py from oracle/_ref in the build container."""
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_vectors.npz"))
    checked = 0
    for n in [2, 3, 4, 5, 7, 8, 15, 16, 20, 64, 11, 13, 26, 143, 19, 59, 257, 1009, 1024, 3000, 4096, 10007]:
        x = rand_cplx(1, n, seed=424242 + n)[0]
        for sgn, tag in ((1, "f"), (-1, "i")):
            key = "c2c_%d_%s" % (n, tag)
            with hs.Plan(n, sgn) as p:
                y = p.exec(x)
            if n <= 64:
                assert rel_rms(y, z[key]) <= tol(n), key
            else:
                idx = z[key + "_idx"]
                scale = np.sqrt(np.mean(np.abs(y) ** 2))
                assert np.max(np.abs(y[idx] - z[key])) <= 40 * tol(n) * scale, key
                assert abs(y.sum() - z[key + "_sum"][0]) <= tol(n) * scale * n, key
            checked += 1
    for n in [4, 8, 16, 32, 64, 2018, 4096]:
        x = xorshift_uniform(n, seed=777 + n)
        with hs.RealPlan(n, 1) as f, hs.RealPlan(n, -1) as g:
            X = f.r2c(x)
            y = g.c2r(X)
        if n <= 64:
            assert rel_rms(X, z["r2c_%d" % n]) <= tol(n) and rel_rms(y, z["c2r_%d" % n]) <= tol(n)
        else:
            idx = z["r2c_%d_idx" % n]
            sX, sy = np.sqrt(np.mean(np.abs(X) ** 2)), np.sqrt(np.mean(np.abs(y) ** 2))
            assert np.max(np.abs(X[idx] - z["r2c_%d" % n])) <= 40 * tol(n) * sX
            assert np.max(np.abs(y[idx] - z["c2r_%d" % n])) <= 40 * tol(n) * sy
        checked += 1
    for la, lb in [(5, 3), (100, 37), (512, 512)]:
        a, b = xorshift_uniform(la, seed=la), xorshift_uniform(lb, seed=lb + 1)
        for t in ("full", "same", "valid"):
            _, y = hs.convolve(a, b, t, "linear")
            assert rel_rms(y, z["conv_%d_%d_%s" % (la, lb, t)]) <= 1e-12 * math.log2(la + lb + 2)
            checked += 1
    assert checked >= 60


# ------------------------------------------------------------------------------------------------------------------
# Additive layouts, called directly: half spectrum with a row stride, strided c2r input.

@pytest.mark.parametrize("n", [64, 2018, 4096, 6000, 10000, 1 << 15, 1 << 17, 1 << 20])
def test_half_spectrum_and_row_strides_direct(hs, product_lib, oracle, n):
    import torch
    b = 3
    h1 = n // 2 + 1
    x = _device_rand((b, n), seed=5 + n)
    Xr = oracle.r2c(x.cpu().numpy())
    for stride in (h1, h1 + 7):
        out = torch.full((b, stride, 2), float("nan"), device="cuda", dtype=torch.float64)
        with hs.RealPlan(n, 1) as f:
            f.r2c_dev(x, out, stride, 0, b)
            torch.cuda.synchronize()
        got = torch.view_as_complex(out).cpu().numpy()
        assert per_transform_worst(got[:, :h1], Xr[:, :h1]) <= tol(n), (n, stride)
        assert np.isnan(got[:, h1:]).all()                      # nothing written beyond bin N/2 of a row
        back = torch.empty((b, n), device="cuda", dtype=torch.float64)
        with hs.RealPlan(n, -1) as g:
            g.c2r_dev(out, stride, back, b)                       # in_stride > N/2+1: the NaN padding must never be read
            torch.cuda.synchronize()
        # against the reference's own c2r of the reference's spectrum (a round trip is NOT the identity to 1e-12 log2 N where the
        # reference's truncated radix-3/5 constants act twice, e.g. N/2 = 3000)
        assert per_transform_worst(back.cpu().numpy(), oracle.c2r(Xr, n)) <= tol(n), (n, stride)


# ------------------------------------------------------------------------------------------------------------------
# Re-entrancy and several devices from one process (the reference's mixed-radix fft_exec only reads its plan, so callers
# may share a plan between threads; kernel attributes and plan state are per device here).

def test_one_plan_two_streams_and_four_threads(hs, product_lib, oracle):
    import threading
    import torch
    n, b = 1 << 16, 4                                 # two-pass plan: needs a workspace, one per caller stream
    xs = [_device_rand((b, n, 2), seed=s) for s in (1, 2)]
    ys = [torch.empty_like(x) for x in xs]
    st = [torch.cuda.Stream(), torch.cuda.Stream()]
    with hs.Plan(n, 1) as p:
        for rep in range(5):
            for i in (0, 1):
                p.exec_dev(xs[i], ys[i], b, stream=st[i].cuda_stream)
        torch.cuda.synchronize()
        for i in (0, 1):
            got = torch.view_as_complex(ys[i]).cpu().numpy()
            assert per_transform_worst(got, oracle.c2c(torch.view_as_complex(xs[i]).cpu().numpy(), 1)) <= tol(n)
    n = 10007
    x = rand_cplx(8, n, seed=3)
    ref = oracle.c2c(x, 1)
    out, errs = [None] * 8, []
    with hs.Plan(n, 1) as p:
        def work(i):
            try:
                for _ in range(4):
                    out[i] = p.exec(x[i])
            except Exception as e:   # pragma: no cover
                errs.append(e)
        th = [threading.Thread(target=work, args=(i,)) for i in range(8)]
        [t.start() for t in th]
        [t.join() for t in th]
    assert not errs and all(rel_rms(out[i], ref[i]) <= tol(n) for i in range(8))


def test_two_devices_driven_by_one_process(hs, product_lib, oracle):
    """One process, two GPUs: the same plan executes on device 0 and then on device 1 (kernel attributes and plan state are
    kept per device), hsfft_set_devices(2) shards host-pointer batches by batch index, and fft_exec_dist runs the four-step
    with the exchange as NVLink peer stores between the two devices."""
    import ctypes as C
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    L = hs.load()
    x = rand_cplx(6, 8192, seed=12)
    r = oracle.c2c(x, 1)
    with hs.Plan(8192, 1) as p:                        # 139 KB of dynamic shared memory: the attribute must be set on both devices
        for dev in (0, 1, 0):
            L.hsfft_set_device(dev)
            assert per_transform_worst(p.exec(x), r) <= tol(8192)
    L.hsfft_set_device(0)
    L.hsfft_set_devices(2)
    try:
        for n, b in ((4096, 37), (10007, 9), (1 << 17, 5)):
            x = rand_cplx(b, n, seed=n)
            with hs.Plan(n, -1) as p:
                assert per_transform_worst(p.exec(x), oracle.c2c(x, -1)) <= tol(n)
        xr = xorshift_uniform(9 * 8192, seed=4).reshape(9, 8192)
        with hs.RealPlan(8192, 1) as f:
            assert per_transform_worst(f.r2c(xr), oracle.r2c(xr)) <= tol(8192)
    finally:
        L.hsfft_set_devices(1)
    for N1, N2, sgn in ((256, 512, 1), (16384, 1024, -1)):
        x = rand_cplx(1, N1 * N2, seed=N1)[0]
        y = np.empty_like(x)
        assert L.fft_exec_dist(N1, N2, sgn, x.ctypes.data_as(C.c_void_p), y.ctypes.data_as(C.c_void_p), 2) == 0
        assert rel_rms(y, oracle.c2c(x, sgn)) <= tol(N1 * N2)


@pytest.mark.parametrize("n", [64, 1000, 3000, 4096, 8192, 1009, 2039, 4099, 10007, 40009, 16384, 1 << 20, 6561, 3 << 14, 1001])
def test_in_place_execution(hs, product_lib, oracle, n):
    """fft_exec_batch_inplace_dev: one device buffer in and out, every kernel family."""
    import torch
    b = 4 if n <= 16384 else 2
    x = rand_cplx(b, n, seed=n + 1)
    buf = torch.view_as_real(torch.from_numpy(x.copy())).cuda()
    for sgn in (1, -1):
        buf.copy_(torch.view_as_real(torch.from_numpy(x)))
        with hs.Plan(n, sgn) as p:
            p.exec_inplace_dev(buf, b)
            torch.cuda.synchronize()
            desc = p.describe()
        got = torch.view_as_complex(buf).cpu().numpy()
        assert per_transform_worst(got, oracle.c2c(x, sgn)) <= tol(n), (n, sgn, desc)


JIT_TWOPASS_SIZES = [11 * 1024, 13 * 4096, 17 * 512, 10000, 12000, 3 << 14, 50000, 100000, 3 << 16, 250000, 1000000]


@pytest.mark.parametrize("n", JIT_TWOPASS_SIZES)
def test_plan_time_compiled_two_pass_kernels(hs, product_lib, oracle, n):
    """Smooth lengths above one SM's shared memory without a static plan: the factor list is cut in two and both tiled
    passes are compiled with NVRTC at plan creation (csrc/hs_jit.cu) -- same stages, butterflies and literals in the
    reference's order, partial last tiles where a sub-length is not a multiple of the tile width."""
    for sgn in (1, -1):
        x = rand_cplx(3, n, seed=n + sgn)
        with hs.Plan(n, sgn) as p:
            desc = p.describe()
            y = p.exec(x)
        assert "two tiled passes compiled" in desc, desc
        assert per_transform_worst(y, oracle.c2c(x, sgn)) <= tol(n), (n, sgn, desc)


RADER_JIT_PRIMES = [127, 251, 337, 401, 449, 577, 641, 769, 1153, 1871, 2017, 2521, 3137, 3361, 4001, 4051]


@pytest.mark.parametrize("n", RADER_JIT_PRIMES)
def test_rader_kernels_compiled_for_primes_with_a_smooth_predecessor(hs, product_lib, oracle, n):
    """Every prime N <= 4097 whose N-1 is a product of compile-time radices runs as ONE cyclic convolution of length N-1
    (hs_fused.cuh's Rader kernel, compiled for this N at plan creation), not as the padded chirp-z of twice the length."""
    x = rand_cplx(9, n, seed=n)
    os.environ["HSFFT_RADER_JIT"] = "2"       # every prime that can have a kernel (the default keeps those that measured faster)
    try:
        for sgn in (1, -1):
            with hs.Plan(n, sgn) as p:
                desc = p.describe()
                y = p.exec(x)
            assert "rader" in desc and "plan-time" in desc, desc
            assert per_transform_worst(y, oracle.c2c(x, sgn)) <= tol(n), (n, sgn, desc)
    finally:
        del os.environ["HSFFT_RADER_JIT"]
    with hs.Plan(n, 1) as p:
        desc = p.describe()
        y = p.exec(x)
    assert ("rader" in desc) == (n in (401, 449, 577, 641, 769, 1153, 1871, 2521, 3137, 3361, 4001)), desc
    assert per_transform_worst(y, oracle.c2c(x, 1)) <= tol(n)

