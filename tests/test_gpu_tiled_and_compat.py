"""Tiled multi-strip mode (par_group) and the reference's own entry point (launch_kernel), on the GPU."""
import numpy as np
import pytest

from conftest import valid_vertex_mask
from pixel_art_remaster_gpu_b200 import synth

pytestmark = pytest.mark.gpu


def _n_gpus():
    import torch
    return torch.cuda.device_count() if torch.cuda.is_available() else 0


@pytest.mark.parametrize("n_strips", [1, 2, 3])
def test_strips_equal_whole_image(lib, oracle, n_strips):
    """Config 4 in miniature: an image cut into strips (apron rows copied between the strips' buffers,
    labels stitched across the seams) gives exactly the single-image result, CC labels included.
    Several strips may live on one device, so this runs on a 1-GPU box too."""
    if _n_gpus() < 1:
        pytest.skip("no GPU")
    W, H = 160, 150
    img = synth.adversarial_sprite(W, H, 77) if n_strips == 3 else synth.snes_frame(W, H, 78)
    want = oracle.pipeline(img, scale=4, want=("graph_aux", "graph", "labels", "raster"))
    devices = [k % _n_gpus() for k in range(n_strips)]
    with lib.RemasterGroup(devices, W, H, 4) as grp:
        got = grp.remaster_host(img, subdivide=True, want=("rgba", "graph", "graph_aux", "labels"))
        assert np.array_equal(got["graph_aux"], want["graph_aux"])
        assert np.array_equal(got["graph"], want["graph"])
        assert np.array_equal(got["labels"], want["labels"])
        assert np.array_equal(got["rgba"], want["raster"])
        flipped = grp.remaster_host(img, subdivide=True, want=("rgba",), flip_output=True)["rgba"]
        assert np.array_equal(flipped, want["raster"][::-1])


def test_strips_on_two_gpus(lib, oracle):
    if _n_gpus() < 2:
        pytest.skip("needs 2 GPUs")
    W, H = 256, 224
    img = synth.adversarial_sprite(W, H, 79)
    want = oracle.pipeline(img, scale=4, want=("graph", "labels", "raster"))
    with lib.RemasterGroup([0, 1], W, H, 4) as grp:
        got = grp.remaster_host(img, want=("rgba", "graph", "labels"))
    assert np.array_equal(got["graph"], want["graph"])
    assert np.array_equal(got["labels"], want["labels"])
    assert np.array_equal(got["rgba"], want["raster"])


@pytest.mark.parametrize("n_strips", [1, 2, 4])
def test_device_resident_strips(lib, oracle, n_strips):
    """par_group_remaster_device: the strips' own rows are written into their device buffers, halo exchange, kernels and
    the on-device label stitch run without the host, and the own rows of every output equal the single-image result
    (RGBA8 and BGR8).  Strips are spread over the GPUs that exist (all on one device on a 1-GPU box)."""
    import torch
    if _n_gpus() < 1:
        pytest.skip("no GPU")
    W, H, S = 128, 40 * n_strips + 23, 3
    img = synth.adversarial_sprite(W, H, 31) if n_strips != 2 else synth.snes_frame(W, H, 32)
    want = oracle.pipeline(img, scale=S, want=("graph", "labels", "raster"))
    devices = [k % _n_gpus() for k in range(n_strips)]
    with lib.RemasterGroup(devices, W, H, S) as grp:
        strips = grp.strips()
        assert len(strips) == n_strips and strips[0]["own"][0] == 0 and strips[-1]["own"][1] == H
        for st in strips:
            b, e = st["own"]
            lb = st["load"][0]
            st["bgr"].zero_()
            st["bgr"][b - lb:e - lb].copy_(torch.from_numpy(img[b:e]))
        torch.cuda.synchronize()
        for fmt in (lib.OUT_RGBA8, lib.OUT_BGR8):
            wall, dev_ms = grp.remaster_device(subdivide=True, out_format=fmt, want_labels=True)
            assert wall > 0 and dev_ms > 0
            for st in strips:
                b, e = st["own"]
                lb = st["load"][0]
                assert np.array_equal(st["graph"][b - lb:e - lb].cpu().numpy(), want["graph"][b:e])
                assert np.array_equal(st["labels"][b - lb:e - lb].cpu().numpy(), want["labels"][b:e])
                image = st["image"](fmt)[(b - lb) * S:(e - lb) * S].cpu().numpy()
                ref = want["raster"][b * S:e * S]
                assert np.array_equal(image, ref if fmt == lib.OUT_RGBA8 else ref[..., [2, 1, 0]])
        with pytest.raises(lib.RemasterError):
            grp.remaster_device(out_format=lib.OUT_INDEX8)


def test_long_component_across_all_seams(lib, oracle):
    """A one-pixel-wide snake that crosses every seam several times: the stitched labels must still be
    the global minimum index."""
    if _n_gpus() < 1:
        pytest.skip("no GPU")
    W, H = 96, 200
    img = np.zeros((H, W, 3), np.uint8)
    img[...] = (10, 200, 30)
    for x in range(4, W - 4, 8):                      # vertical bars joined alternately at top and bottom
        img[4:H - 4, x] = (250, 20, 20)
        y = H - 5 if (x // 8) % 2 else 4
        img[y, x:x + 9] = (250, 20, 20)
    want = oracle.pipeline(img, want=("labels",))["labels"]
    with lib.RemasterGroup([0, 0, 0, 0], W, H, 2) as grp:
        got = grp.remaster_host(img, want=("labels",))["labels"]
    assert np.array_equal(got, want)
    assert len(np.unique(want)) <= 4


@pytest.mark.parametrize("subdivide", [False, True])
def test_launch_kernel_symbol_matches_reference_semantics(lib, oracle, subdivide):
    """launch_kernel (kernel.cu:286-288): graph_h, edge_count_h, returned triangle list, pos and colorPos."""
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no GPU")
    img = synth.snes_frame(96, 80, 55)
    want = oracle.pipeline(img, subdivide=subdivide, want=("graph", "poly_count", "tri", "ntri"))
    graph, count, diagram, pos, col = lib.launch_kernel(img, subdivide=subdivide, want_vbo=True)
    assert np.array_equal(graph, want["graph"])
    assert np.array_equal(count, want["poly_count"])
    ntri = want["ntri"]
    assert (ntri == count - 2).all()
    m = np.arange(45)[None, :] < 3 * ntri[:, None]
    assert np.array_equal(diagram[m], want["tri"][m])
    assert np.array_equal(pos[m], want["tri"][m])
    assert (pos[~m] == -100.0).all()                      # position_kernel, kernel.cu:130-134
    rgba = np.concatenate([img.reshape(-1, 3)[:, ::-1], np.full((96 * 80, 1), 255, np.uint8)], 1)
    assert np.array_equal(col, np.repeat(rgba[:, None, :], 45, 1))  # color_kernel, kernel.cu:98-101
