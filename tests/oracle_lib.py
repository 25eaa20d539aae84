"""ctypes bindings for the CHECKERS: oracle/liboracle.so (our C restatement) and
oracle/_ref/libhsfft_ref_{asis,repaired}.so (the reference's own sources built by
oracle/Makefile).  Test infrastructure only -- the product never imports this.
"""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")

c_dbl_p = C.POINTER(C.c_double)


def _ptr(a):
    return a.ctypes.data_as(c_dbl_p)


def build_oracle():
    """(Re)build liboracle.so and, when /root/reference is present, oracle/_ref."""
    subprocess.run(["make", "-s", "-C", ORACLE_DIR], check=True, capture_output=True)


def xorshift_uniform(n, seed=88172645463325252):
    """The SURVEY section 8d input stream: xorshift64, uniform [-1, 1)."""
    out = np.empty(n, dtype=np.float64)
    x = np.uint64(seed)
    # vectorised in blocks is not possible for a serial PRNG; keep a fast python loop for
    # small n and fall back to numpy's PCG (seeded from the same constant) for large n.
    if n > 200_000:
        rng = np.random.default_rng(seed % (2**63))
        return rng.uniform(-1.0, 1.0, n)
    xi = int(x)
    mask = (1 << 64) - 1
    for i in range(n):
        xi ^= (xi << 13) & mask
        xi ^= xi >> 7
        xi ^= (xi << 17) & mask
        out[i] = (xi >> 11) * (2.0 / 9007199254740992.0) - 1.0
    return out


def rand_cplx(howmany, n, seed=88172645463325252):
    return xorshift_uniform(2 * howmany * n, seed).view(np.complex128).reshape(howmany, n)


class Oracle:
    """Our restatement (oracle/hsfft_oracle.c)."""

    def __init__(self):
        path = os.path.join(ORACLE_DIR, "liboracle.so")
        if not os.path.exists(path):
            build_oracle()
        L = self.L = C.CDLL(path)
        L.orc_c2c_batch.argtypes = [C.c_int, C.c_int, c_dbl_p, c_dbl_p, C.c_long, C.c_int]
        L.orc_r2c_batch.argtypes = [C.c_int, c_dbl_p, c_dbl_p, C.c_long, C.c_int]
        L.orc_c2r_batch.argtypes = [C.c_int, c_dbl_p, c_dbl_p, C.c_long, C.c_int]
        L.orc_naive_dft.argtypes = [C.c_int, C.c_int, c_dbl_p, c_dbl_p]
        L.orc_convolve.argtypes = [C.c_char_p, C.c_char_p, c_dbl_p, C.c_int, c_dbl_p, C.c_int, c_dbl_p]
        L.orc_convolve.restype = C.c_int
        L.orc_dividebyN.argtypes = [C.c_int]
        L.orc_factors.argtypes = [C.c_int, C.POINTER(C.c_int)]
        L.orc_bluestein_pad.argtypes = [C.c_int]
        L.orc_stage_twiddles.argtypes = [c_dbl_p, C.c_int, C.POINTER(C.c_int), C.c_int]
        L.orc_num_threads.restype = C.c_int

    def num_threads(self):
        return self.L.orc_num_threads()

    def c2c(self, x, sgn, nthreads=0):
        x = np.ascontiguousarray(x, dtype=np.complex128)
        x2 = x.reshape(-1, x.shape[-1])
        out = np.empty_like(x2)
        self.L.orc_c2c_batch(x2.shape[1], sgn, _ptr(x2), _ptr(out), x2.shape[0], nthreads)
        return out.reshape(x.shape)

    def r2c(self, x, nthreads=0):
        x = np.ascontiguousarray(x, dtype=np.float64)
        x2 = x.reshape(-1, x.shape[-1])
        out = np.empty(x2.shape, dtype=np.complex128)
        self.L.orc_r2c_batch(x2.shape[1], _ptr(x2), _ptr(out), x2.shape[0], nthreads)
        return out.reshape(x.shape)

    def c2r(self, X, n, nthreads=0):
        """X[..., :n/2+1] is read."""
        X = np.ascontiguousarray(np.asarray(X, dtype=np.complex128)[..., : n // 2 + 1])
        X2 = X.reshape(-1, n // 2 + 1)
        out = np.empty((X2.shape[0], n), dtype=np.float64)
        self.L.orc_c2r_batch(n, _ptr(X2), _ptr(out), X2.shape[0], nthreads)
        return out.reshape(X.shape[:-1] + (n,))

    def naive_dft(self, x, sgn):
        x = np.ascontiguousarray(x, dtype=np.complex128)
        out = np.empty_like(x)
        self.L.orc_naive_dft(x.shape[0], sgn, _ptr(x), _ptr(out))
        return out

    def convolve(self, a, b, type_="full", conv_type="linear"):
        a = np.ascontiguousarray(a, dtype=np.float64)
        b = np.ascontiguousarray(b, dtype=np.float64)
        out = np.zeros(4 * (len(a) + len(b)) + 8, dtype=np.float64)
        n = self.L.orc_convolve(type_.encode(), conv_type.encode(), _ptr(a), len(a), _ptr(b), len(b), _ptr(out))
        return n, out[: max(n, 0)].copy()

    def factors(self, n):
        arr = (C.c_int * 64)()
        lf = self.L.orc_factors(n, arr)
        return list(arr[:lf])

    def dividebyN(self, n):
        return self.L.orc_dividebyN(n)

    def bluestein_pad(self, n):
        return self.L.orc_bluestein_pad(n)

    def stage_twiddles(self, n, fac):
        tw = np.zeros(n + 1, dtype=np.complex128)
        arr = (C.c_int * 64)(*fac)
        self.L.orc_stage_twiddles(_ptr(tw), n, arr, len(fac))
        return tw[: n - 1]


class FftSet(C.Structure):
    """struct fft_set (reference src/highspeedFFT.h:36-43)."""

    _fields_ = [("N", C.c_int), ("sgn", C.c_int), ("factors", C.c_int * 64), ("lf", C.c_int), ("lt", C.c_int),
                ("twiddle", C.c_double * 2)]


class RefLib:
    """The reference's own C code (flavour 'asis' or 'repaired'), via oracle/_ref."""

    def __init__(self, flavour):
        path = os.path.join(ORACLE_DIR, "_ref", "libhsfft_ref_%s.so" % flavour)
        if not os.path.exists(path):
            build_oracle()
        if not os.path.exists(path):
            raise FileNotFoundError(path)
        L = self.L = C.CDLL(path)
        self.flavour = flavour
        L.refdrv_c2c_batch.argtypes = [C.c_int, C.c_int, c_dbl_p, c_dbl_p, C.c_long, C.c_int]
        L.refdrv_r2c_batch.argtypes = [C.c_int, c_dbl_p, c_dbl_p, C.c_long, C.c_int]
        L.refdrv_c2r_batch.argtypes = [C.c_int, c_dbl_p, c_dbl_p, C.c_long, C.c_int]
        L.refdrv_convolve_batch.argtypes = [C.c_char_p, C.c_char_p, c_dbl_p, C.c_int, c_dbl_p, C.c_int, c_dbl_p,
                                            C.c_int, C.c_long, C.c_int]
        L.refdrv_num_threads.restype = C.c_int
        L.fft_convolve.argtypes = [C.c_char_p, C.c_char_p, c_dbl_p, C.c_int, c_dbl_p, C.c_int, c_dbl_p]
        L.fft_convolve.restype = C.c_int
        L.fft_init.argtypes = [C.c_int, C.c_int]
        L.fft_init.restype = C.POINTER(FftSet)
        L.free_fft.argtypes = [C.POINTER(FftSet)]
        L.factors.argtypes = [C.c_int, C.POINTER(C.c_int)]
        L.dividebyN.argtypes = [C.c_int]

    @staticmethod
    def available(flavour="repaired"):
        return os.path.exists(os.path.join(ORACLE_DIR, "_ref", "libhsfft_ref_%s.so" % flavour)) or os.path.exists(
            "/root/reference/src/highSpeedFFT.c")

    def num_threads(self):
        return self.L.refdrv_num_threads()

    def c2c(self, x, sgn, nthreads=0, prefill=None):
        x = np.ascontiguousarray(x, dtype=np.complex128)
        x2 = x.reshape(-1, x.shape[-1]).copy()
        out = np.zeros_like(x2) if prefill is None else np.full_like(x2, prefill)
        self.L.refdrv_c2c_batch(x2.shape[1], sgn, _ptr(x2), _ptr(out), x2.shape[0], nthreads)
        return out.reshape(x.shape)

    def r2c(self, x, nthreads=0):
        x = np.ascontiguousarray(x, dtype=np.float64)
        x2 = x.reshape(-1, x.shape[-1]).copy()
        out = np.zeros(x2.shape, dtype=np.complex128)
        self.L.refdrv_r2c_batch(x2.shape[1], _ptr(x2), _ptr(out), x2.shape[0], nthreads)
        return out.reshape(x.shape)

    def c2r(self, X, n, nthreads=0):
        X = np.ascontiguousarray(np.asarray(X, dtype=np.complex128)[..., : n // 2 + 1])
        X2 = X.reshape(-1, n // 2 + 1).copy()
        out = np.zeros((X2.shape[0], n), dtype=np.float64)
        self.L.refdrv_c2r_batch(n, _ptr(X2), _ptr(out), X2.shape[0], nthreads)
        return out.reshape(X.shape[:-1] + (n,))

    def convolve(self, a, b, type_="full", conv_type="linear"):
        a = np.ascontiguousarray(a, dtype=np.float64).copy()
        b = np.ascontiguousarray(b, dtype=np.float64).copy()
        out = np.zeros(4 * (len(a) + len(b)) + 8, dtype=np.float64)
        n = self.L.fft_convolve(type_.encode(), conv_type.encode(), _ptr(a), len(a), _ptr(b), len(b), _ptr(out))
        return n, out[: max(n, 0)].copy()

    def plan_fields(self, n, sgn):
        p = self.L.fft_init(n, sgn)
        s = p.contents
        d = dict(N=s.N, sgn=s.sgn, lf=s.lf, lt=s.lt, factors=list(s.factors[: s.lf]))
        ntw = (n if s.lt == 0 else 0)
        if ntw:
            base = C.addressof(s) + FftSet.twiddle.offset
            d["twiddle"] = np.ctypeslib.as_array((C.c_double * (2 * (ntw - 1))).from_address(base)).copy().view(
                np.complex128)
        self.L.free_fft(p)
        return d

    def factors(self, n):
        arr = (C.c_int * 64)()
        lf = self.L.factors(n, arr)
        return list(arr[:lf])

    def dividebyN(self, n):
        return self.L.dividebyN(n)


def rel_rms(y, r):
    y = np.asarray(y).astype(np.complex128).ravel()
    r = np.asarray(r).astype(np.complex128).ravel()
    den = np.sqrt(np.sum(np.abs(r) ** 2))
    num = np.sqrt(np.sum(np.abs(y - r) ** 2))
    if den == 0:
        return 0.0 if num == 0 else float("inf")
    return float(num / den)


def tol(n):
    """north_star tolerance: relative RMS <= 1e-12 * log2(N)."""
    return 1e-12 * np.log2(n)


def asis_is_authoritative(n, oracle):
    """the unmodified reference is well defined iff the factor list (of N, or of the
    Bluestein pad M) does not end in 2, has no 13, has no NON-final 3/4/7 and (Bluestein) N != 2^k+1."""
    if oracle.dividebyN(n):
        fac = oracle.factors(n)
    else:
        m = oracle.bluestein_pad(n)
        k = 1
        while k < n:
            k *= 2
        alloc = 2 * k if k < 2 * n - 2 else k
        if alloc != m:
            return False
        fac = oracle.factors(m)
    if fac[-1] == 2 or 13 in fac:
        return False
    return not any(f in (2, 3, 4, 7) for f in fac[:-1])
