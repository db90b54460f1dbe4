"""The reference's own CUDA build (kernel.cu compiled unmodified for sm_100a, oracle/_ref/libref_cuda.so)
run on the B200 next to the oracle and the product.  Settles SURVEY App. B's open items:
  (i)   ref_cuda == oracle on graph_aux / graph / polygon counts / triangle lists outside the pixels
        where the reference itself reads out of bounds (B-3);
  (ii)  -use_fast_math (div.approx / sqrt.approx) does not change stage E on B200 (B-2);
  (iii) the device's RGBtoYUV for all 2^24 colours == the fused-Y oracle (B-1).
"""
import numpy as np
import pytest

from pixel_art_remaster_gpu_b200 import synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ref_cuda():
    import torch
    from oracle.oracle import RefCuda
    if not torch.cuda.is_available():
        pytest.skip("no GPU")
    if not RefCuda.available():
        pytest.skip("oracle/_ref/libref_cuda.so not built (needs /root/reference at build time)")
    return RefCuda()


def test_device_yuv_all_colours(ref_cuda, oracle, lib):
    dev = ref_cuda.yuv_all()
    assert np.array_equal(dev, oracle.yuv_all(True))
    rng = np.random.default_rng(3)
    for c in rng.integers(0, 1 << 24, 5000):
        assert lib.yuv_word(int(c) & 255, (int(c) >> 8) & 255, int(c) >> 16) == int(dev[int(c)])


def _undefined_mask(W, H, stage_e):
    """Pixels whose reference output depends on out-of-bounds reads, dilated by one."""
    bad = np.zeros((H, W), bool)
    bad[0, 0] = bad[H - 1, W - 1] = True              # cells_Kernel reads graph[n-1] / graph[n+1], kernel.cu:205
    if stage_e:
        bad[H - 1, :] = True                          # checkTJunction reads past the image end on the top row
        bad[H - 2, W - 1] = True                      # ... and for pixel (W-1, H-2), subdivision_functions.cu:187
    grown = bad.copy()
    for dy in (-1, 0, 1):
        for dx in (-1, 0, 1):
            grown |= np.roll(np.roll(bad, dy, 0), dx, 1)
    return grown.reshape(-1)


@pytest.mark.parametrize("name,make", [("g1", lambda: synth.snes_frame(256, 224, synth.BASE_SEED)),
                                       ("g2", lambda: synth.snes_frame(320, 240, synth.BASE_SEED + 2)),
                                       ("g5", lambda: synth.adversarial_sprite(192, 160))])
@pytest.mark.parametrize("subdivide", [False, True], ids=["hull", "subdivided"])
def test_reference_cuda_equals_oracle_and_product(ref_cuda, oracle, lib, name, make, subdivide):
    img = make()
    H, W = img.shape[:2]
    want = oracle.pipeline(img, subdivide=subdivide, want=("graph_aux", "graph", "poly_count", "tri", "ntri"))
    assert np.array_equal(ref_cuda.graph_aux(img), want["graph_aux"])            # stages A+B, every pixel
    ref = ref_cuda.launch(img, subdivide, ("graph", "edge_count", "diagram"))
    assert np.array_equal(ref["graph"], want["graph"])                           # stage C, every pixel
    ok = ~_undefined_mask(W, H, subdivide)
    assert np.array_equal(ref["edge_count"][ok], want["poly_count"][ok])         # stages D/E vertex counts
    m = (np.arange(45)[None, :] < 3 * want["ntri"][:, None]) & ok[:, None]
    assert np.array_equal(ref["diagram"][m], want["tri"][m])                     # stage F triangle lists
    # and the product's implementation of the reference's entry point, against the reference itself
    graph, count, diagram = lib.launch_kernel(img, subdivide=subdivide)
    assert np.array_equal(graph, ref["graph"])
    assert np.array_equal(count[ok], ref["edge_count"][ok])
    assert np.array_equal(diagram[m], ref["diagram"][m])
