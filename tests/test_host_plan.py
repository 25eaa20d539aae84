"""CPU tests of the product's HOST side (plain C in libhsfft_b200.so): the library loads, exports
every symbol include/*.h declares, and builds plans whose public fields equal the reference's.
No compute call is made here (that needs a GPU; see test_gpu_parity.py)."""
import ctypes as C
import os
import re
import subprocess
import sys

import numpy as np
import pytest

import oracle_lib as ol

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "ref_vectors.npz"))


def test_library_exports_every_declared_symbol(hs, product_lib):
    for name in hs.EXPORTS:
        assert hasattr(product_lib, name), name
    # and every function prototype in the public headers is in that list
    declared = set()
    for h in ("highspeedFFT.h", "real.h", "hsfft_b200.h"):
        src = open(os.path.join(ROOT, "include", h)).read()
        src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
        for m in re.finditer(r"^[ \t]*(?:const\s+)?[A-Za-z_][\w \*]*?[\s\*]([A-Za-z_]\w*)\s*\([^;{]*\)\s*;", src, flags=re.M):
            declared.add(m.group(1))
    assert declared and declared <= set(hs.EXPORTS), declared - set(hs.EXPORTS)


def test_struct_layout_matches_reference_abi(hs):
    """N@0 sgn@4 factors@8 lf@264 lt@268 twiddle@272, sizeof 288."""
    S = hs.FftSet
    assert (S.N.offset, S.sgn.offset, S.factors.offset, S.lf.offset, S.lt.offset, S.twiddle.offset) == (0, 4, 8, 264, 268, 272)
    assert C.sizeof(S) == 288
    code = r'''
#include <stdio.h>
#include <stddef.h>
#include "real.h"
int main(void){printf("%zu %zu %zu %zu %zu %zu %zu %zu\n", offsetof(struct fft_set,N), offsetof(struct fft_set,sgn),
 offsetof(struct fft_set,factors), offsetof(struct fft_set,lf), offsetof(struct fft_set,lt), offsetof(struct fft_set,twiddle),
 sizeof(struct fft_set), offsetof(struct fft_real_set,twiddle2)); return 0;}'''
    exe = "/tmp/hs_layout_test"
    subprocess.run(["/usr/bin/gcc", "-x", "c", "-", "-I", os.path.join(ROOT, "include"), "-o", exe], input=code.encode(), check=True)
    out = subprocess.run([exe], capture_output=True, text=True, check=True).stdout.split()
    assert [int(v) for v in out] == [0, 4, 8, 264, 268, 272, 288, 8]


def test_routing_and_factors_match_oracle_sweep(hs, product_lib, oracle):
    for n in list(range(1, 3000)) + [4096, 10007, 32768, 1 << 19, 1 << 20, 3 ** 9, 19 * 64, 53 * 53 * 4]:
        assert hs.dividebyN(n) == oracle.dividebyN(n), n
        assert hs.factors(n) == oracle.factors(n), n


def test_routing_matches_reference_build(hs, product_lib, ref_repaired):
    for n in list(range(1, 1100)) + [2048, 3000, 4096, 10007]:
        assert hs.dividebyN(n) == ref_repaired.dividebyN(n), n
        assert hs.factors(n) == ref_repaired.factors(n), n


@pytest.mark.parametrize("n", [15, 17, 64, 1024, 3000, 4096, 1009])
@pytest.mark.parametrize("sgn", [1, -1])
def test_plan_fields_equal_reference(hs, product_lib, oracle, n, sgn):
    with hs.Plan(n, sgn) as p:
        assert p.p.contents.N == n and p.p.contents.sgn == sgn
        assert p.factors == list(GOLD["plan_%d_factors" % n])
        assert p.lt == int(GOLD["plan_%d_lt" % n][0])
        if p.lt == 0:
            tw = p.twiddles(n - 1)
            ref = GOLD["plan_%d_twiddle" % n]
            ref = ref if sgn == 1 else np.conj(ref)
            assert np.array_equal(tw, ref)                       # same cos/sin expression -> same bits
            assert np.array_equal(tw, oracle.stage_twiddles(n, p.factors) if sgn == 1 else np.conj(oracle.stage_twiddles(n, p.factors)))


def test_bluestein_plan_records_padded_length(hs, product_lib):
    with hs.Plan(1009, 1) as p:
        assert p.lt == 1 and p.factors == [8, 8, 8, 4]            # M = 2048
    with hs.Plan(10007, -1) as p:
        assert p.lt == 1 and p.factors == [8] * 5                  # M = 32768
    with hs.Plan(19, 1) as p:                                      # 19 is not in the router's list
        assert p.lt == 1


def test_real_plan_fields(hs, product_lib):
    n = 64
    with hs.RealPlan(n, 1) as r:
        assert r.p.contents.cobj.contents.N == n // 2
        tw2 = r.twiddle2()
        k = np.arange(n // 2)
        assert np.allclose(tw2, np.exp(2j * np.pi * k / n), atol=1e-15)


def test_convolution_length_helpers(hs, product_lib):
    L = product_lib
    assert [L.next_power_of_two(v) for v in (0, 1, 2, 3, 1023, 1024, 1025)] == [1, 1, 2, 4, 1024, 1024, 2048]
    assert L.find_optimal_fft_length(1047, b"linear", 1000, 48) == 2048
    assert L.find_optimal_fft_length(1047, b"circular", 1000, 48) == 1024


def _run_snippet(code):
    return subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, cwd=ROOT)


PRE = "import importlib,sys; sys.path.insert(0,'.'); hs=importlib.import_module('mixed-radix-fast-fourier-transform_b200'); L=hs.load();"


def test_error_convention_print_and_exit():
    """Reference convention: message on stderr + exit(EXIT_FAILURE)."""
    r = _run_snippet(PRE + "L.fft_real_init(7, 1)")
    assert r.returncode == 1 and "must be positive and even" in r.stderr
    r = _run_snippet(PRE + "L.fft_init(0, 1)")
    assert r.returncode == 1 and "must be positive" in r.stderr
    r = _run_snippet(PRE + "L.fft_init(1, 1)")                    # reference: longvectorN rejects num_factors == 0
    assert r.returncode == 1 and "longvectorN" in r.stderr
    r = _run_snippet(PRE + "import ctypes; L.fft_exec(None, None, None)")
    assert r.returncode == 1 and "Invalid FFT object or data pointers" in r.stderr
    r = _run_snippet(PRE + "L.find_optimal_fft_length(8, b'bogus', 4, 4)")
    assert r.returncode == 1 and "Invalid convolution type" in r.stderr


def test_no_cpu_fallback_exec_fails_loudly_without_gpu(hs, product_lib):
    """The product path must not silently compute on the CPU."""
    if product_lib.hsfft_device_count() > 0:
        pytest.skip("a CUDA device is present")
    r = _run_snippet(PRE + "import numpy as np; p=hs.Plan(64,1); p.exec(np.zeros(64,dtype=np.complex128)); print('computed')")
    assert r.returncode == 1 and "no CPU fallback" in r.stderr and "computed" not in r.stdout


def test_product_does_not_link_or_import_the_oracle():
    pkg = os.path.join(ROOT, "mixed-radix-fast-fourier-transform_b200")
    for dirpath, _, files in os.walk(pkg):
        if os.path.basename(dirpath) == "build":
            continue
        for f in files:
            if f.endswith((".c", ".h", ".cu", ".cuh", ".py")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "hsfft_oracle" not in txt and "liboracle" not in txt and "oracle_lib" not in txt, f
    ldd = subprocess.run(["ldd", os.path.join(pkg, "libhsfft_b200.so")], capture_output=True, text=True).stdout
    assert "oracle" not in ldd and "cufft" not in ldd.lower()


def test_plan_time_kernel_source_is_generated_and_compiles_with_nvrtc(hs, product_lib):
    """csrc/hs_jit.cu: the translation unit generated for a length without a static plan, and -- NVRTC needs no GPU --
    its compilation for sm_100a from the library's own headers."""
    import ctypes as C
    L = product_lib
    L.hs_cu_jit_source.argtypes = [C.c_int, C.c_int, C.POINTER(C.c_int), C.c_int, C.c_char_p, C.c_int]
    L.hs_cu_jit_source.restype = C.c_int
    buf = C.create_string_buffer(16384)
    rx = (C.c_int * 4)(5, 5, 5, 8)                      # N = 1000: the reference's factor list [8,5,5,5] reversed
    assert L.hs_cu_jit_source(1000, 1, rx, 4, buf, 16384) > 0
    src = buf.value.decode()
    assert "HsPlan<1000, 125, 1, 3, -2, 1111, 5, 5, 5, 8>" in src and "hs_fft_team<PJ, 1>" in src
    assert L.hs_cu_jit_source(1472, 1, (C.c_int * 3)(8, 8, 23), 3, buf, 16384) == 0       # radix 23: not a template radix
    buf2 = C.create_string_buffer(16384)
    assert L.hs_cu_jit_source(1001, 1, (C.c_int * 3)(7, 11, 13), 3, buf2, 16384) > 0      # radix 11, 13, 17: unrolled generic butterflies
    src2 = buf2.value.decode()
    assert "7, 11, 13>" in src2 and "hs_jit_r2c" in src2 and "hs_jit_c2r" in src2
    try:
        rt = C.CDLL("libnvrtc.so.12")
    except OSError:
        try:
            rt = C.CDLL("/usr/local/cuda/lib64/libnvrtc.so.12")
        except OSError:
            pytest.skip("no NVRTC in this environment")
    csrc = os.path.join(ROOT, "mixed-radix-fast-fourier-transform_b200", "csrc")
    names = ["hs_common.cuh", "hs_abi.h", "hs_butterflies.cuh", "hs_smem_fft.cuh"]
    texts = [open(os.path.join(csrc, n), "rb").read() for n in names]
    prog = C.c_void_p()
    assert rt.nvrtcCreateProgram(C.byref(prog), src.encode(), b"p.cu", 4, (C.c_char_p * 4)(*texts), (C.c_char_p * 4)(*[n.encode() for n in names])) == 0
    opts = (C.c_char_p * 4)(b"--gpu-architecture=sm_100a", b"-std=c++17", b"-default-device", b"-DHS_JIT=1")
    rc = rt.nvrtcCompileProgram(prog, 4, opts)
    size = C.c_size_t()
    rt.nvrtcGetCUBINSize(prog, C.byref(size))
    assert rc == 0 and size.value > 1000
    prog2 = C.c_void_p()
    assert rt.nvrtcCreateProgram(C.byref(prog2), src2.encode(), b"p2.cu", 4, (C.c_char_p * 4)(*texts), (C.c_char_p * 4)(*[n.encode() for n in names])) == 0
    assert rt.nvrtcCompileProgram(prog2, 4, opts) == 0


def test_plan_time_planner_fat_thread_rule(hs, product_lib, monkeypatch):
    """csrc/hs_jit.cu choose(): lengths from 1800 up get fewer, fatter threads (pairs fused up to 25 points, 16-32 points per
    thread, three CTAs per SM where they fit, whole warps where a warp would idle); the c2r kernel keeps the thin plan where
    the fat one runs one CTA per SM; shorter lengths and HSFFT_JIT_FAT=0 keep one butterfly of the widest pass per thread."""
    import ctypes as C
    import re
    L = product_lib
    L.hs_cu_jit_source.argtypes = [C.c_int, C.c_int, C.POINTER(C.c_int), C.c_int, C.c_char_p, C.c_int]
    L.hs_cu_jit_source.restype = C.c_int

    def plans(n, rx):
        buf = C.create_string_buffer(32768)
        assert L.hs_cu_jit_source(n, 1, (C.c_int * len(rx))(*rx), len(rx), buf, 32768) > 0
        src = buf.value.decode()
        pj = re.search(r"using PJ = HsPlan<([^>]*)>", src).group(1)
        pr = re.search(r"using PR = HsPlan<([^>]*)>", src).group(1)
        minb = int(re.search(r"__launch_bounds__\(PJ::threads, (\d+)\)\s*hs_jit_c2c\(", src).group(1))
        return pj, pr, minb

    # (length, stage radices in execution order) -> threads, pass grouping, CTAs per SM
    pj, pr, minb = plans(5000, [5, 5, 5, 5, 8])
    assert pj.startswith("5000, 200, 1,") and ", 221, 5, 5, 5, 5, 8" in pj and minb == 1
    assert pr.startswith("5000, 625, 1,") and ", 11111, " in pr              # c2r: the thin plan (no persistent twin)
    pj, pr, minb = plans(3125, [5, 5, 5, 5, 5])
    assert pj.startswith("3125, 125, 1,") and ", 221, " in pj and minb == 3 and pr == pj   # every pass exact: stays at 125
    pj, pr, minb = plans(2400, [3, 4, 5, 5, 8])
    assert pj.startswith("2400, 128, 1,") and ", 221, " in pj and minb == 3 and pr == pj   # 100 -> whole warps
    pj, pr, minb = plans(3600, [2, 3, 3, 5, 5, 8])
    assert pj.startswith("3600, 240, 1,") and minb == 2                       # <= 20 points per thread: 128 registers, two CTAs
    pj, pr, minb = plans(1536, [3, 8, 8, 8])
    assert pj.startswith("1536, 192, 1,") and ", 1111, " in pj                # below the threshold: thin threads
    pj, pr, minb = plans(6561, [3] * 8)
    assert pj.startswith("6561, 729, 1,")                                     # no team size with an exact pass in range: rule above
    monkeypatch.setenv("HSFFT_JIT_FAT", "0")
    pj, pr, minb = plans(5000, [5, 5, 5, 5, 8])
    assert pj.startswith("5000, 625, 1,") and pr == pj


def test_compile_time_kernel_table_is_self_consistent(hs, product_lib):
    """csrc/hs_ct_table.cuh: every two-pass descriptor factors its length, and both of its stage lists belong to a sub-plan
    the host side can launch by id (family HS_CT_SUB for first passes, HS_CT_SUB or HS_CT_SUB2 for second passes) -- the
    fused real / convolution paths pair a descriptor's twiddle tables with those ids (hsfft_plan.c find_sub_like).
    Host tables only: no GPU needed."""
    import ctypes as C

    class Desc(C.Structure):
        _fields_ = [("id", C.c_int), ("family", C.c_int), ("n", C.c_int), ("nst", C.c_int), ("radix", C.c_int * 8), ("split_min", C.c_int),
                    ("threads", C.c_int), ("smem_bytes", C.c_int), ("n1", C.c_int), ("n2", C.c_int), ("nst2", C.c_int), ("radix2", C.c_int * 8),
                    ("split_min2", C.c_int), ("exact", C.c_int)]

    L = product_lib
    L.hs_cu_ct_count.restype = C.c_int
    L.hs_cu_ct_get.restype = C.POINTER(Desc)
    L.hs_cu_ct_get.argtypes = [C.c_int]
    descs = [L.hs_cu_ct_get(i).contents for i in range(L.hs_cu_ct_count())]
    assert len({d.id for d in descs}) == len(descs)
    SUB, SUB2, TWOPASS = 3, 7, 2
    subs = [(d.family, d.n, tuple(d.radix[: d.nst])) for d in descs if d.family in (SUB, SUB2)]
    two = [d for d in descs if d.family == TWOPASS]
    assert len(two) >= 7 and any(f == SUB2 for f, _, _ in subs)
    for d in descs:
        prod = 1
        for r in d.radix[: d.nst]:
            prod *= r
        assert prod == (d.n1 if d.family == TWOPASS else d.n), (d.id, prod)
    for d in two:
        assert d.n1 * d.n2 == d.n
        prod2 = 1
        for r in d.radix2[: d.nst2]:
            prod2 *= r
        assert prod2 == d.n2
        assert (SUB, d.n1, tuple(d.radix[: d.nst])) in subs, ("first pass without a sub-plan", d.id)
        assert any(f in (SUB, SUB2) and n == d.n2 and rx == tuple(d.radix2[: d.nst2]) for f, n, rx in subs), ("second pass without a sub-plan", d.id)
