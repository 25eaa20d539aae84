import importlib
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

PKG = "mixed-radix-fast-fourier-transform_b200"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _has_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def hs():
    return importlib.import_module(PKG)


@pytest.fixture(scope="session")
def oracle():
    import oracle_lib
    return oracle_lib.Oracle()


@pytest.fixture(scope="session")
def ref_repaired():
    import oracle_lib
    if not oracle_lib.RefLib.available("repaired"):
        pytest.skip("oracle/_ref not built and /root/reference absent")
    return oracle_lib.RefLib("repaired")


@pytest.fixture(scope="session")
def ref_asis():
    import oracle_lib
    if not oracle_lib.RefLib.available("asis"):
        pytest.skip("oracle/_ref not built and /root/reference absent")
    return oracle_lib.RefLib("asis")


@pytest.fixture(scope="session")
def product_lib(hs):
    """The product C-ABI library (built in tree by __graft_entry__.build / build.py)."""
    if not os.path.exists(hs.LIB_PATH):
        spec = importlib.util.spec_from_file_location("hsfft_build", os.path.join(ROOT, PKG, "build.py"))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        mod.build()
    return hs.load()


@pytest.fixture(scope="session")
def emu_lib(hs):
    """CPU emulation of the kernels (tests/emu): the product's kernel source run on OS threads."""
    d = os.path.join(ROOT, "tests", "emu")
    if os.environ.get("HSFFT_EMU_ASAN"):
        # AddressSanitizer build of the emulation: run as
        #   HSFFT_EMU_ASAN=1 LD_PRELOAD=$(gcc -print-file-name=libasan.so) ASAN_OPTIONS=detect_leaks=0 pytest tests/test_emu_kernels.py
        subprocess.run(["make", "-s", "-C", d, "SAN=1"], check=True, capture_output=True)
        return hs.load(os.path.join(d, "_build_asan", "libhsfft_emu.so"))
    subprocess.run(["make", "-s", "-C", d], check=True, capture_output=True)
    return hs.load(os.path.join(d, "_build", "libhsfft_emu.so"))
