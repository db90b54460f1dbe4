"""Stage D+E of the PRODUCT without a GPU: csrc/polygon.cuh is __host__ __device__, so the same code the
kernels run is compiled for the host (tools/host_polygons.cpp) and compared with the oracle / golden
fixtures vertex by vertex."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT, valid_vertex_mask
from pixel_art_remaster_gpu_b200 import synth


@pytest.fixture(scope="module")
def hostpoly(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("hostpoly") / "libhostpoly.so")
    inc = "/usr/local/cuda/include"
    if not os.path.exists(os.path.join(inc, "cuda.h")):
        pytest.skip("CUDA headers not found")
    subprocess.run(["g++", "-O2", "-shared", "-fPIC", "-I", inc, os.path.join(ROOT, "tools", "host_polygons.cpp"),
                    os.path.join(ROOT, "pixel_art_remaster_gpu_b200", "csrc", "cell_table.cpp"), "-o", out], check=True)
    return C.CDLL(out)


def _run(lib, img, graph, subdivide):
    H, W = img.shape[:2]
    ws = img.strides[0]
    flat = np.zeros(H * ws + ws + 64, np.uint8)
    flat[: H * ws] = np.lib.stride_tricks.as_strided(img, shape=(H * ws,), strides=(1,)) if ws != 3 * W else img.reshape(-1)
    poly = np.zeros((H * W, 45, 2), np.float32)
    cnt = np.zeros(H * W, np.int32)
    lib.host_polygons(flat.ctypes.data_as(C.c_void_p), np.ascontiguousarray(graph).ctypes.data_as(C.c_void_p), W, H, ws, int(subdivide),
                      poly.ctypes.data_as(C.c_void_p), cnt.ctypes.data_as(C.c_void_p))
    return poly, cnt


@pytest.mark.parametrize("make", [lambda: synth.snes_frame(96, 80, 3), lambda: synth.adversarial_sprite(120, 90), lambda: synth.snes_frame(50, 37, 5),
                                  lambda: synth.pad_rows(synth.snes_frame(33, 21, 9), 104), lambda: synth.snes_frame(1, 30, 4)])
@pytest.mark.parametrize("subdivide", [False, True])
def test_product_polygons_equal_oracle(hostpoly, oracle, make, subdivide):
    img = make()
    want = oracle.pipeline(img, subdivide, True, 4, ("graph", "poly", "poly_count"))
    poly, cnt = _run(hostpoly, img, want["graph"], subdivide)
    assert np.array_equal(cnt, want["poly_count"])
    m = valid_vertex_mask(cnt)
    assert np.array_equal(poly[m], want["poly"][m])
