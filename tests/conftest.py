"""Shared fixtures.  `-m "not gpu"` runs on a CPU box (oracle vs reference/golden, host logic, ABI
surface); `-m gpu` needs a B200 and goes through the C ABI of libpixelart_b200.so."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200, sm_100a)")


@pytest.fixture(scope="session")
def oracle():
    from oracle.oracle import Oracle, build
    build(ref=os.path.isdir("/root/reference"))
    return Oracle()


@pytest.fixture(scope="session")
def ref_host():
    from oracle.oracle import RefHost
    if not RefHost.available(fma=True):
        pytest.skip("oracle/_ref/libref_host_fma.so not built (needs /root/reference)")
    return RefHost(fma=True)


@pytest.fixture(scope="session")
def ref_host_plain():
    from oracle.oracle import RefHost
    if not RefHost.available(fma=False):
        pytest.skip("oracle/_ref/libref_host.so not built (needs /root/reference)")
    return RefHost(fma=False)


@pytest.fixture(scope="session")
def lib():
    """The product library, built in-tree if stale."""
    from pixel_art_remaster_gpu_b200 import build as b
    b.build_library()
    import pixel_art_remaster_gpu_b200 as par
    par.load_library()
    return par


@pytest.fixture(scope="session")
def ctx(lib):
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    c = lib.Remaster(device=0, max_width=512, max_height=448, max_frames=16)
    yield c
    c.close()


def valid_vertex_mask(count, closing=False):
    """(N, 45) mask of the polygon slots that carry vertices."""
    return np.arange(45)[None, :] < (count[:, None] + (1 if closing else 0))
