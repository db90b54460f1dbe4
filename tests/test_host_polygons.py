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
                    os.path.join(ROOT, "pixel_art_remaster_gpu_b200", "csrc", "cell_table.cpp"),
                    os.path.join(ROOT, "pixel_art_remaster_gpu_b200", "csrc", "smooth_table.cpp"), "-o", out], check=True)
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


@pytest.mark.parametrize("name", ["snes", "adversarial", "noise", "padded"])
@pytest.mark.parametrize("scale,halo", [(4, 1), (8, 3), (3, 1)])
def test_smoothing_pieces_equal_polygon_coverage(hostpoly, oracle, name, scale, halo):
    """The smoothing tables of the raster kernel (csrc/smooth_table.h: coverage as an XOR of CUT and LINK pieces) against
    the coverage of the polygon itself, on the host: every smoothed cell, every sample of its (S + 2 halo)^2 window."""
    rng = np.random.default_rng(5)
    if name == "snes":
        img = synth.snes_frame(96, 80, 3)
    elif name == "adversarial":
        img = synth.adversarial_sprite(120, 90)
    elif name == "padded":
        img = synth.pad_rows(synth.snes_frame(33, 21, 9), 104)
    else:
        img = np.ascontiguousarray(rng.integers(0, 256, (3, 3), dtype=np.uint8)[rng.integers(0, 3, (64, 72))])
    graph = oracle.pipeline(img, True, True, 4, ("graph",))["graph"]
    H, W = img.shape[:2]
    ws = img.strides[0]
    flat = np.zeros(H * ws + ws + 64, np.uint8)
    flat[: H * ws] = np.lib.stride_tricks.as_strided(img, shape=(H * ws,), strides=(1,)) if ws != 3 * W else img.reshape(-1)
    out = (C.c_long * 4)()
    hostpoly.host_smooth_check(flat.ctypes.data_as(C.c_void_p), np.ascontiguousarray(graph).ctypes.data_as(C.c_void_p), W, H, ws, scale, halo, out)
    smoothed, geometric, wrong, classes = out[0], out[1], out[2], out[3]
    assert smoothed > 0 and wrong == 0, (smoothed, geometric, wrong)
    assert classes == 188
    if name in ("snes", "padded"):
        assert geometric <= 0.01 * smoothed
