import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def _have_gpu() -> bool:
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _have_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device here")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


@pytest.fixture(scope="session")
def built():
    """Build (if stale) and return the path of the product library."""
    from fastglobetrotter_b200 import build as b
    return b.build()


@pytest.fixture(scope="session")
def lib(built):
    from fastglobetrotter_b200.companion import PackedCompanion
    return PackedCompanion(built)


@pytest.fixture(params=["rle", "dense"])
def each_encoding(lib, request):
    """run a test once with run-length-encoded paintings (the library's preferred format) and once with dense label
    matrices (chunked on the device by the label-scan kernels)"""
    lib.encode = request.param
    yield request.param
    lib.encode = "rle"


@pytest.fixture(params=["windowed", "two_phase", "two_phase_batched", "two_phase_blocks", "two_phase_window"])
def each_null_impl(lib, request):
    """run a NULL test with k_null_strict (null_impl 1), with the two-phase path (null_impl 2: every chunk pair evaluated once
    into records, ordered commit), with the two-phase path forced into many small batches and many narrow commit tasks, and with
    the labels of B taken in blocks (the large-K variant of phase 1)"""
    lib.set_option("null_impl", 1 if request.param == "windowed" else 2)
    if request.param == "two_phase_batched":
        lib.set_option("null_batch", 3000)
        lib.set_option("null_tasks", 100000)
    lib.set_option("null_block", {"two_phase": 0, "two_phase_blocks": 1}.get(request.param, -1))
    if request.param == "two_phase_blocks":
        lib.set_option("null_tasks", 400)          # label pairs with one and with several bin ranges inside a block
    lib.set_option("null_window", {"two_phase_window": 1, "two_phase": 0}.get(request.param, -1))
    yield request.param
    lib.set_option("null_impl", 0)
    lib.set_option("null_block", -1)
    lib.set_option("null_window", -1)
    lib.set_option("null_batch", 0)
    lib.set_option("null_tasks", 0)


@pytest.fixture(scope="session")
def oracle():
    from oracle.oracle import Oracle
    return Oracle()


@pytest.fixture(scope="session")
def ref_libs():
    """(pristine, hooked) compiled reference, built here if /root/reference exists."""
    from oracle import build_ref
    ok = build_ref.build()
    if not ok:
        pytest.skip("compiled reference unavailable")
    from helpers import REF_LIB, REF_HOOKED
    from fastglobetrotter_b200.companion import Companion
    return Companion(REF_LIB), Companion(REF_HOOKED)
