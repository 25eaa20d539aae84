"""CPU tests that PIN THE ORACLE (oracle/hsfft_oracle.c):
  * against the reference's own C sources built in place (oracle/_ref) -- when available,
  * against committed outputs of that build (tests/golden/ref_vectors.npz),
  * against the known-answer vectors of the reference's gtest file,
  * against an independent long-double DFT.
"""
import os

import numpy as np
import pytest

import oracle_lib as ol
from oracle_lib import rand_cplx, rel_rms, tol, xorshift_uniform

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "ref_vectors.npz"))


def test_oracle_bit_identical_to_repaired_reference(oracle, ref_repaired):
    """Every N in 2..600 (incl. multiples of 13 and 2^k+1): same operation order -> same bits."""
    for n in range(2, 601):
        x = rand_cplx(1, n, seed=1234567 + n)
        for sgn in (1, -1):
            assert rel_rms(oracle.c2c(x, sgn), ref_repaired.c2c(x, sgn)) <= 1e-15, (n, sgn)


def test_oracle_matches_unmodified_reference_where_defined(oracle, ref_asis):
    cnt = 0
    for n in list(range(2, 301)) + [1000, 1009, 3000, 4096, 10007]:
        if n % 13 == 0 or not ol.asis_is_authoritative(n, oracle):
            continue
        x = rand_cplx(1, n, seed=5 + n)
        assert rel_rms(oracle.c2c(x, 1), ref_asis.c2c(x, 1)) <= 1e-15, n
        cnt += 1
    assert cnt > 150


def test_unmodified_reference_bug2_is_real(oracle, ref_asis):
    """N=1024=[8,8,8,2]: the as-is two-point leaf reads its output buffer first."""
    x = rand_cplx(1, 1024)
    a = ref_asis.c2c(x, 1, prefill=0.0)
    b = ref_asis.c2c(x, 1, prefill=3.0)
    assert rel_rms(a, b) > 0.1                      # depends on the caller's previous out[] contents
    assert rel_rms(a, oracle.c2c(x, 1)) > 0.1       # and is not the DFT


@pytest.mark.parametrize("n", [2, 3, 4, 5, 7, 8, 15, 16, 20, 64, 11, 13, 26, 143, 19, 59, 257, 1009, 1024, 3000, 4096, 10007])
def test_oracle_vs_golden_reference_outputs(oracle, n):
    x = rand_cplx(1, n, seed=424242 + n)[0]
    for sgn, tag in ((1, "f"), (-1, "i")):
        y = oracle.c2c(x, sgn)
        key = "c2c_%d_%s" % (n, tag)
        if n <= 64:
            assert rel_rms(y, GOLD[key]) <= 1e-15
        else:
            assert rel_rms(y[GOLD[key + "_idx"]], GOLD[key]) <= 1e-15
            assert abs(y.sum() - GOLD[key + "_sum"][0]) <= 1e-9 * max(1.0, abs(GOLD[key + "_sum"][0]))


@pytest.mark.parametrize("n", [4, 8, 16, 32, 64, 2018, 4096])
def test_oracle_real_vs_golden(oracle, n):
    x = xorshift_uniform(n, seed=777 + n)
    X = oracle.r2c(x)
    y = oracle.c2r(X, n)
    if n <= 64:
        assert rel_rms(X, GOLD["r2c_%d" % n]) <= 1e-15 and rel_rms(y, GOLD["c2r_%d" % n]) <= 1e-15
    else:
        idx = GOLD["r2c_%d_idx" % n]
        assert rel_rms(X[idx], GOLD["r2c_%d" % n]) <= 1e-15 and rel_rms(y[idx], GOLD["c2r_%d" % n]) <= 1e-15


@pytest.mark.parametrize("la,lb", [(5, 3), (100, 37), (512, 512)])
def test_oracle_convolve_vs_golden_and_numpy(oracle, la, lb):
    a, b = xorshift_uniform(la, seed=la), xorshift_uniform(lb, seed=lb + 1)
    for t in ("full", "same", "valid"):
        n, y = oracle.convolve(a, b, t, "linear")
        assert rel_rms(y, GOLD["conv_%d_%d_%s" % (la, lb, t)]) <= 1e-15
    n, y = oracle.convolve(a, b, "full", "linear")
    assert n == la + lb - 1 and rel_rms(y, np.convolve(a, b)) <= 1e-12


@pytest.mark.parametrize("n", list(range(2, 66)) + [100, 121, 125, 128, 243, 257, 343, 513, 1000, 1009, 2048, 3000])
def test_oracle_vs_independent_long_double_dft(oracle, n):
    """Shared-bug guard.  5e-11: the reference's 11-digit radix-3/5/7 constants cost up to ~2e-11."""
    x = rand_cplx(1, n, seed=99 + n)[0]
    for sgn in (1, -1):
        assert rel_rms(oracle.c2c(x, sgn), oracle.naive_dft(x, sgn)) <= 5e-11


def test_truncated_constants_are_part_of_the_contract(oracle):
    """N=3000: the reference is 1.5e-11 away from the exact DFT -- more than the 1.16e-11 parity tolerance."""
    x = rand_cplx(1, 3000, seed=3)[0]
    err = rel_rms(oracle.c2c(x, 1), np.fft.fft(x))
    assert 5e-12 < err < 5e-11


# ---------------------------------------------------------------- reference gtest known answers
def test_kat_impulse_gives_ones(oracle):
    """test/test_mixedRadixFFT.cpp:444-501 (leaf butterflies N=2,3) extended to all leaf radices."""
    for n in (2, 3, 4, 5, 7, 8, 11):
        x = np.zeros(n, dtype=np.complex128)
        x[0] = 1.0
        assert np.allclose(oracle.c2c(x, 1), np.ones(n), atol=1e-15)


def test_kat_r2c_sine_n64(oracle):
    """test/test_mixedRadixFFT.cpp:794-836: sine f=2, N=64 -> bin 2 = -32i, Hermitian mirror in out[N-i]."""
    n = 64
    x = np.sin(2.0 * np.pi * 2.0 * np.arange(n) / n)
    X = oracle.r2c(x)
    assert abs(X[0]) < 1e-10 and abs(X[32]) < 1e-10
    assert abs(X[2] - (-32.0j)) < 1e-10
    assert np.allclose(X[n - np.arange(1, 32)], np.conj(X[1:32]), atol=0)


def test_kat_inverse_n15_samples(oracle):
    """test/test_mixedRadixFFT.cpp:1702-1707: the expected inverse output is sin(2 pi 2 i / 15)."""
    expected = np.array([0.0, 0.743145, 0.994522, 0.587785, -0.207912, -0.866025, -0.951057, -0.406737, 0.406737,
                         0.951057, 0.866025, 0.207912, -0.587785, -0.994522, -0.743145])
    n = 15
    X = oracle.c2c(np.sin(2 * np.pi * 2 * np.arange(n) / n).astype(np.complex128), 1)
    back = oracle.c2c(X, -1) / n
    assert np.max(np.abs(back.real - expected)) < 1e-6 and np.max(np.abs(back.imag)) < 1e-10


@pytest.mark.parametrize("n", [4, 8, 16, 32, 64])
def test_kat_real_roundtrip(oracle, n):
    """test/test_mixedRadixFFT.cpp:1259-1386 and src/real.c:360-420: c2r(r2c(x))/N == x, MSE < 1e-10."""
    x = np.sin(2.0 * np.pi * 2.0 * np.arange(n) / n)
    back = oracle.c2r(oracle.r2c(x), n) / n
    assert np.mean((back - x) ** 2) < 1e-10


@pytest.mark.parametrize("n", [4, 8, 15, 20, 64])
def test_kat_complex_roundtrip_selftest_sizes(oracle, n):
    """src/real.c:278-357 (the commented-out main): forward, inverse, /N, MSE < 1e-10."""
    x = (np.sin(2.0 * np.pi * 2.0 * np.arange(n) / n)).astype(np.complex128)
    back = oracle.c2c(oracle.c2c(x, 1), -1) / n
    assert np.mean(np.abs(back - x) ** 2) < 1e-10


def test_plan_routing_and_factor_order(oracle):
    """SURVEY section 8a4 [measured] factor lists; 19 is not in the router's list."""
    assert oracle.factors(4096) == [8, 8, 8, 8]
    assert oracle.factors(3000) == [8, 5, 5, 5, 3]
    assert oracle.factors(1024) == [8, 8, 8, 2]
    assert oracle.factors(2048) == [8, 8, 8, 4]
    assert oracle.factors(1 << 19) == [8] * 6 + [2]
    assert oracle.dividebyN(19) == 0 and oracle.dividebyN(17) == 1 and oracle.dividebyN(1009) == 0
    assert oracle.bluestein_pad(1009) == 2048 and oracle.bluestein_pad(10007) == 32768 and oracle.bluestein_pad(257) == 1024
