"""hsfft-b200: Python binding of the C-ABI library libhsfft_b200.so.

The product is the shared library (extern "C", include/*.h) -- this module only mirrors the
reference's interface one to one for tests, benchmarks and Python callers:

    reference C (src/highspeedFFT.h, src/real.h, src/convolve.c)     here
    ------------------------------------------------------------     -------------------------------
    fft_init / fft_exec / free_fft                                   Plan(N, sgn).exec(x) / .close()
    fft_real_init / fft_r2c_exec / fft_c2r_exec / free_real_fft      RealPlan(N, sgn).r2c(x) / .c2r(X)
    fft_convolve                                                     convolve(a, b, type, conv_type)
    dividebyN / factors / longvectorN                                dividebyN / factors / plan fields

There is no CPU path here: the library needs a CUDA device and exits otherwise (the reference's own
error convention), and importing this module fails loudly when the library has not been built.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libhsfft_b200.so")

c_dbl_p = C.POINTER(C.c_double)
c_int_p = C.POINTER(C.c_int)


class FftSet(C.Structure):
    """struct fft_set (include/highspeedFFT.h; reference src/highspeedFFT.h:36-43)."""

    _fields_ = [("N", C.c_int), ("sgn", C.c_int), ("factors", C.c_int * 64), ("lf", C.c_int), ("lt", C.c_int),
                ("twiddle", C.c_double * 2)]


class FftRealSet(C.Structure):
    """struct fft_real_set (include/real.h; reference src/real.h:28-32)."""

    _fields_ = [("cobj", C.POINTER(FftSet)), ("twiddle2", C.c_double * 2)]


# every symbol include/*.h declares; tests check that the library exports all of them
EXPORTS = [
    "fft_init", "fft_exec", "free_fft", "divideby", "dividebyN", "factors", "twiddle", "longvectorN",
    "fft_real_init", "fft_r2c_exec", "fft_c2r_exec", "free_real_fft",
    "generate_signal", "compute_mse", "print_complex", "generate_real_signal", "compute_mse_real", "print_real",
    "next_power_of_two", "find_optimal_fft_length", "fft_convolve",
    "fft_exec_batch", "fft_r2c_exec_batch", "fft_c2r_exec_batch", "fft_convolve_batch",
    "fft_exec_batch_dev", "fft_r2c_exec_batch_dev", "fft_c2r_exec_batch_dev", "fft_convolve_batch_dev",
    "fft_convolve_stream", "fft_convolve_stream_dev",
    "fft2_init", "fft2_exec", "fft2_exec_batch", "fft2_exec_batch_dev", "fft2_describe", "free_fft2",
    "hsfft_device_count", "hsfft_set_device", "hsfft_get_device", "hsfft_malloc_device", "hsfft_free_device",
    "hsfft_malloc_host", "hsfft_free_host", "hsfft_copy_h2d", "hsfft_copy_d2h", "hsfft_sync", "hsfft_launch_count",
    "hsfft_plan_describe", "hsfft_version", "hsfft_timing_begin", "hsfft_timing_report",
    "hsfft_set_devices", "hsfft_get_devices", "fft_exec_dist", "fft_exec_batch_inplace_dev",
    "hsfft_dist_init", "hsfft_dist_cols_dev", "hsfft_dist_rows_dev", "hsfft_dist_slab_elems", "hsfft_dist_free",
    "hsfft_ipc_get_handle", "hsfft_dist_set_peers", "hsfft_dist_set_peer_pointers", "hsfft_dist_cols_exchange_dev",
]


def _bind(L):
    vp, i, l, ll, sz, cp = C.c_void_p, C.c_int, C.c_long, C.c_longlong, C.c_size_t, C.c_char_p
    P, RP = C.POINTER(FftSet), C.POINTER(FftRealSet)
    sig = {
        "fft_init": (P, [i, i]), "fft_exec": (None, [P, vp, vp]), "free_fft": (None, [P]),
        "divideby": (i, [i, i]), "dividebyN": (i, [i]), "factors": (i, [i, c_int_p]),
        "twiddle": (None, [vp, i, i]), "longvectorN": (None, [vp, i, c_int_p, i]),
        "fft_real_init": (RP, [i, i]), "fft_r2c_exec": (None, [RP, vp, vp]), "fft_c2r_exec": (None, [RP, vp, vp]),
        "free_real_fft": (None, [RP]),
        "next_power_of_two": (i, [i]), "find_optimal_fft_length": (i, [i, cp, i, i]),
        "fft_convolve": (i, [cp, cp, vp, i, vp, i, vp]),
        "fft_exec_batch": (None, [P, vp, vp, l]), "fft_r2c_exec_batch": (None, [RP, vp, vp, l]),
        "fft_c2r_exec_batch": (None, [RP, vp, vp, l]),
        "fft_convolve_batch": (i, [cp, cp, vp, i, vp, i, vp, l, l]),
        "fft_exec_batch_dev": (None, [P, vp, vp, l, vp]),
        "fft_r2c_exec_batch_dev": (None, [RP, vp, vp, l, i, l, vp]),
        "fft_c2r_exec_batch_dev": (None, [RP, vp, l, vp, l, vp]),
        "fft_convolve_batch_dev": (i, [cp, cp, vp, i, vp, i, vp, l, l, vp]),
        "fft2_init": (vp, [i, i, i]), "fft2_exec": (None, [vp, vp, vp]), "fft2_exec_batch": (None, [vp, vp, vp, l]),
        "fft2_exec_batch_dev": (None, [vp, vp, vp, l, vp]), "fft2_describe": (cp, [vp]), "free_fft2": (None, [vp]),
        "fft_convolve_stream": (l, [vp, l, vp, i, vp, i]), "fft_convolve_stream_dev": (l, [vp, l, vp, i, vp, i, vp]),
        "hsfft_device_count": (i, []), "hsfft_set_device": (None, [i]), "hsfft_get_device": (i, []),
        "hsfft_malloc_device": (vp, [sz]), "hsfft_free_device": (None, [vp]),
        "hsfft_malloc_host": (vp, [sz]), "hsfft_free_host": (None, [vp]),
        "hsfft_copy_h2d": (None, [vp, vp, sz]), "hsfft_copy_d2h": (None, [vp, vp, sz]), "hsfft_sync": (None, []),
        "hsfft_launch_count": (ll, []), "hsfft_timing_begin": (None, [vp]), "hsfft_set_devices": (None, [i]), "hsfft_get_devices": (i, []),
        "fft_exec_dist": (i, [i, i, i, vp, vp, i]), "fft_exec_batch_inplace_dev": (None, [P, vp, l, vp]), "hsfft_timing_report": (i, [cp, i]), "hsfft_plan_describe": (cp, [P]), "hsfft_version": (cp, []),
        "hsfft_dist_init": (vp, [i, i, i, i, i]), "hsfft_dist_cols_dev": (None, [vp, vp, vp, vp, vp]),
        "hsfft_dist_rows_dev": (None, [vp, vp, vp, vp, vp]), "hsfft_dist_slab_elems": (ll, [vp]), "hsfft_dist_free": (None, [vp]),
        "hsfft_ipc_get_handle": (None, [vp, vp]), "hsfft_dist_set_peers": (None, [vp, vp, vp, i]),
        "hsfft_dist_set_peer_pointers": (None, [vp, vp, i]), "hsfft_dist_cols_exchange_dev": (None, [vp, vp, vp, vp]),
    }
    for name, (res, args) in sig.items():
        f = getattr(L, name)
        f.restype = res
        f.argtypes = args
    return L


_LIBS = {}


def load(path=None):
    """Load (once) and return the bound C library.  Raises if it has not been built."""
    path = os.path.abspath(path or os.environ.get("HSFFT_LIB") or LIB_PATH)  # HSFFT_LIB: A/B runs of build variants
    if path not in _LIBS:
        if not os.path.exists(path):
            raise ImportError(
                "%s is missing: build it with `python mixed-radix-fast-fourier-transform_b200/build.py` "
                "(hsfft-b200 has no CPU fallback)" % path)
        _LIBS[path] = _bind(C.CDLL(path))
    return _LIBS[path]


def _vp(a):
    return a.ctypes.data_as(C.c_void_p)


def _dev_ptr(t):
    """Device pointer of a torch CUDA tensor (or a raw int)."""
    return C.c_void_p(t if isinstance(t, int) else t.data_ptr())


class Plan:
    """fft_init / fft_exec / free_fft (reference src/highspeedFFT.h:34,45,59)."""

    def __init__(self, N, sgn, lib=None):
        self.L = lib or load()
        self.N, self.sgn = int(N), int(sgn)
        self.p = self.L.fft_init(self.N, self.sgn)
        if not self.p:
            raise MemoryError("fft_init returned NULL")

    # public struct fields, as the reference's callers read them
    @property
    def lt(self):
        return self.p.contents.lt

    @property
    def lf(self):
        return self.p.contents.lf

    @property
    def factors(self):
        return list(self.p.contents.factors[: self.p.contents.lf])

    def twiddles(self, count):
        base = C.addressof(self.p.contents) + FftSet.twiddle.offset
        return np.ctypeslib.as_array((C.c_double * (2 * count)).from_address(base)).copy().view(np.complex128)

    def describe(self):
        return self.L.hsfft_plan_describe(self.p).decode()

    def exec(self, x):
        """One transform (fft_exec) or a batch (fft_exec_batch) of host data; x[..., N] complex128."""
        x = np.ascontiguousarray(x, dtype=np.complex128)
        assert x.shape[-1] == self.N
        out = np.empty_like(x)
        howmany = x.size // self.N
        if howmany == 1:
            self.L.fft_exec(self.p, _vp(x), _vp(out))
        else:
            self.L.fft_exec_batch(self.p, _vp(x), _vp(out), howmany)
        return out

    def exec_host_ptr(self, in_ptr, out_ptr, howmany):
        self.L.fft_exec_batch(self.p, C.c_void_p(in_ptr), C.c_void_p(out_ptr), howmany)

    def exec_dev(self, d_in, d_out, howmany, stream=0):
        """Device-resident batch: d_in/d_out are torch CUDA complex128 tensors (or raw pointers)."""
        self.L.fft_exec_batch_dev(self.p, _dev_ptr(d_in), _dev_ptr(d_out), howmany, C.c_void_p(stream))

    def exec_inplace_dev(self, d_data, howmany, stream=0):
        self.L.fft_exec_batch_inplace_dev(self.p, _dev_ptr(d_data), howmany, C.c_void_p(stream))

    def close(self):
        if self.p:
            self.L.free_fft(self.p)
            self.p = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()


class RealPlan:
    """fft_real_init / fft_r2c_exec / fft_c2r_exec / free_real_fft (reference src/real.h:46-84)."""

    def __init__(self, N, sgn, lib=None):
        self.L = lib or load()
        self.N, self.sgn = int(N), int(sgn)
        self.p = self.L.fft_real_init(self.N, self.sgn)

    def twiddle2(self):
        base = C.addressof(self.p.contents) + FftRealSet.twiddle2.offset
        return np.ctypeslib.as_array((C.c_double * self.N).from_address(base)).copy().view(np.complex128)

    def r2c(self, x):
        """x[..., N] float64 -> [..., N] complex128 (full mirrored spectrum, like the reference)."""
        x = np.ascontiguousarray(x, dtype=np.float64)
        assert x.shape[-1] == self.N
        out = np.empty(x.shape, dtype=np.complex128)
        howmany = x.size // self.N
        if howmany == 1:
            self.L.fft_r2c_exec(self.p, _vp(x), _vp(out))
        else:
            self.L.fft_r2c_exec_batch(self.p, _vp(x), _vp(out), howmany)
        return out

    def c2r(self, X):
        """X[..., N] complex128 (bins 0..N/2 are read) -> [..., N] float64, unnormalised."""
        X = np.ascontiguousarray(X, dtype=np.complex128)
        assert X.shape[-1] == self.N
        out = np.empty(X.shape, dtype=np.float64)
        howmany = X.size // self.N
        if howmany == 1:
            self.L.fft_c2r_exec(self.p, _vp(X), _vp(out))
        else:
            self.L.fft_c2r_exec_batch(self.p, _vp(X), _vp(out), howmany)
        return out

    def r2c_dev(self, d_in, d_out, out_stride, full, howmany, stream=0):
        self.L.fft_r2c_exec_batch_dev(self.p, _dev_ptr(d_in), _dev_ptr(d_out), out_stride, int(full), howmany,
                                      C.c_void_p(stream))

    def c2r_dev(self, d_in, in_stride, d_out, howmany, stream=0):
        self.L.fft_c2r_exec_batch_dev(self.p, _dev_ptr(d_in), in_stride, _dev_ptr(d_out), howmany, C.c_void_p(stream))

    def close(self):
        if self.p:
            self.L.free_real_fft(self.p)
            self.p = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()


def convolve(a, b, type_="full", conv_type="linear", lib=None):
    """fft_convolve (reference src/convolve.c:74): returns (length, samples)."""
    L = lib or load()
    a = np.ascontiguousarray(a, dtype=np.float64)
    b = np.ascontiguousarray(b, dtype=np.float64)
    out = np.zeros(4 * (a.size + b.size) + 8, dtype=np.float64)
    n = L.fft_convolve(type_.encode() if type_ is not None else None, conv_type.encode(), _vp(a), a.size, _vp(b), b.size, _vp(out))
    return n, out[: max(n, 0)].copy()


def convolve_batch(a, b, type_="full", conv_type="linear", lib=None):
    """Rows of a[howmany, la] convolved with rows of b[howmany, lb]."""
    L = lib or load()
    a = np.ascontiguousarray(a, dtype=np.float64)
    b = np.ascontiguousarray(b, dtype=np.float64)
    howmany, la = a.shape
    lb = b.shape[1]
    stride = 2 * (la + lb)
    out = np.zeros((howmany, stride), dtype=np.float64)
    n = L.fft_convolve_batch(type_.encode(), conv_type.encode(), _vp(a), la, _vp(b), lb, _vp(out), stride, howmany)
    return n, out[:, : max(n, 0)].copy()


class Plan2D:
    """fft2_init / fft2_exec_batch / free_fft2: row-major n0 x n1 complex transforms (additive API)."""

    def __init__(self, n0, n1, sgn, lib=None):
        self.L = lib or load()
        self.n0, self.n1 = n0, n1
        self.p = self.L.fft2_init(n0, n1, sgn)

    def exec(self, x):
        x = np.ascontiguousarray(x, dtype=np.complex128)
        howmany = x.size // (self.n0 * self.n1)
        y = np.empty_like(x)
        self.L.fft2_exec_batch(self.p, _vp(x), _vp(y), howmany)
        return y

    def exec_dev(self, x, y, howmany):
        import torch
        self.L.fft2_exec_batch_dev(self.p, _dev_ptr(x), _dev_ptr(y), howmany, C.c_void_p(torch.cuda.current_stream().cuda_stream))

    def describe(self):
        return self.L.fft2_describe(self.p).decode()

    def close(self):
        if self.p:
            self.L.free_fft2(self.p)
            self.p = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()


def convolve_stream(x, h, P=0, lib=None):
    """fft_convolve_stream: overlap-save convolution of a long signal with a short filter ("full")."""
    L = lib or load()
    x = np.ascontiguousarray(x, dtype=np.float64)
    h = np.ascontiguousarray(h, dtype=np.float64)
    y = np.zeros(x.size + h.size - 1, dtype=np.float64)
    n = L.fft_convolve_stream(_vp(x), x.size, _vp(h), h.size, _vp(y), P)
    return n, y[: max(n, 0)]


def dividebyN(n, lib=None):
    return (lib or load()).dividebyN(n)


def factors(n, lib=None):
    arr = (C.c_int * 64)()
    lf = (lib or load()).factors(n, arr)
    return list(arr[:lf])


def device_count(lib=None):
    return (lib or load()).hsfft_device_count()


def launch_count(lib=None):
    return (lib or load()).hsfft_launch_count()


class launch_timing:
    """with launch_timing(stream) as t: ... ; t.kernels -> [(tag, launches, total_ms)] in launch order (CUDA events on the stream)."""

    def __init__(self, stream=0, lib=None):
        self.L, self.stream, self.kernels = lib or load(), stream, []

    def __enter__(self):
        self.L.hsfft_timing_begin(C.c_void_p(self.stream))
        return self

    def __exit__(self, *a):
        buf = C.create_string_buffer(1 << 16)
        n = self.L.hsfft_timing_report(buf, len(buf))
        self.kernels = []
        if n > 0:
            for f in buf.value.decode().split(";"):
                tag, cnt, ms = f.rsplit(":", 2)
                self.kernels.append((tag, int(cnt), float(ms)))
