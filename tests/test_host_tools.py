"""Host entry points kept from the reference: class Image (Image.h/.cpp) and the program
entry point (main.cpp:166-194: argv[1] = image path; load -> vertical flip -> path -> image out)."""
import os
import subprocess

import numpy as np
import pytest

from pixel_art_remaster_gpu_b200 import synth

cv2 = pytest.importorskip("cv2")


@pytest.fixture(scope="module")
def cli(lib):
    from pixel_art_remaster_gpu_b200 import build as b
    path = b.build_cli()
    assert path and os.path.exists(path)
    return path


def _bgr_top_down(img_bottom_up):
    return np.ascontiguousarray(img_bottom_up[::-1])


@pytest.mark.parametrize("ext", ["png", "ppm", "bmp"])
def test_image_io_round_trip(cli, tmp_path, ext):
    """Image::loadImage -> reverses -> reverses -> saveImage keeps every pixel (odd width: padded rows)."""
    img = _bgr_top_down(synth.snes_frame(37, 21, 3))
    src = str(tmp_path / ("in." + ext))
    assert cv2.imwrite(src, img)
    for out_ext in ("png", "ppm"):
        dst = str(tmp_path / ("out." + out_ext))
        subprocess.run([cli, src, "-o", dst, "--convert-only"], check=True, timeout=60)
        assert np.array_equal(cv2.imread(dst, cv2.IMREAD_COLOR), img)


def test_palette_and_alpha_png_are_read_as_bgr(cli, tmp_path):
    import zlib, struct
    # 4-colour palette PNG, 2 bits per pixel, written by hand (cv2 cannot write palette PNGs)
    w, h = 5, 3
    pal = bytes([255, 0, 0, 0, 255, 0, 0, 0, 255, 10, 20, 30])
    idx = (np.arange(w * h).reshape(h, w) % 4).astype(np.uint8)
    raw = b""
    for y in range(h):
        bits = 0
        for x in range(w):
            bits = (bits << 2) | int(idx[y, x])
        bits <<= 2 * (8 - w)  # pad to 2 bytes
        raw += b"\x00" + struct.pack(">H", bits)

    def chunk(t, d):
        return struct.pack(">I", len(d)) + t + d + struct.pack(">I", zlib.crc32(t + d) & 0xFFFFFFFF)
    png = b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, 2, 3, 0, 0, 0)) + chunk(b"PLTE", pal) + \
        chunk(b"IDAT", zlib.compress(raw)) + chunk(b"IEND", b"")
    src = str(tmp_path / "pal.png")
    open(src, "wb").write(png)
    dst = str(tmp_path / "pal_out.png")
    subprocess.run([cli, src, "-o", dst, "--convert-only"], check=True, timeout=60)
    want = np.frombuffer(pal, np.uint8).reshape(4, 3)[idx][..., ::-1]
    assert np.array_equal(cv2.imread(dst, cv2.IMREAD_COLOR), want)


def test_graph_picture_of_a_given_graph(cli, tmp_path):
    """--draw-graph / --graph-image: the reference's debug view of the similarity graph (printToImage, main.cpp:79-139) —
    20 pixels per source pixel, a black stroke from each pixel's centre (19 + 20 i - 10, rows counted from the bottom)
    half-way towards every linked neighbour, white elsewhere.  The graph comes from a plane as --graph writes it."""
    w, h, sg = 3, 2, 20
    cv2.imwrite(str(tmp_path / "in.png"), np.zeros((h, w, 3), np.uint8))
    g = np.zeros((h, w), np.uint8)          # pipeline orientation: row 0 = bottom scanline
    g[0, 0] = 1 << 4                        # bottom-left pixel: linked to the right (+1, 0)
    g[0, 1] = (1 << 3) | (1 << 2)           # its right neighbour: left (-1, 0) and up-right (+1, +1)
    g[1, 2] = 1 << 5                        # top-right pixel: down-left (-1, -1)
    cv2.imwrite(str(tmp_path / "g.pgm"), np.ascontiguousarray(g[::-1]))   # planes are stored top scanline first
    out = str(tmp_path / "graph.png")
    subprocess.run([cli, str(tmp_path / "in.png"), "--draw-graph", str(tmp_path / "g.pgm"), "--graph-image", out], check=True, timeout=60)
    pic = cv2.imread(out, cv2.IMREAD_COLOR)
    assert pic.shape == (h * sg, w * sg, 3)
    up = pic[::-1, :, 0]                    # rows counted from the bottom, like the drawing
    black = up == 0
    assert set(np.unique(pic)) <= {0, 255}
    c = lambda k: k * sg + sg // 2 - 1      # centre of source pixel k
    assert black[c(0), c(0):c(0) + 11].all()                       # (0,0) -> right, half-way
    assert black[c(0), c(1) - 10:c(1) + 1].all()                   # (1,0) -> left: the two strokes meet
    assert all(black[c(0) + t, c(1) + t] for t in range(11))       # (1,0) -> up-right diagonal
    assert all(black[c(1) - t, c(2) - t] for t in range(11))       # (2,1) -> down-left diagonal: meets it half-way
    assert not black[c(1), c(0) - 3:c(0) + 4].any()                # the unlinked top-left pixel has no stroke
    assert int(black.sum()) == 4 * 11 - 3                          # four strokes, three shared end pixels


def test_missing_file_fails_loudly(cli, tmp_path):
    r = subprocess.run([cli, str(tmp_path / "nope.png")], capture_output=True, text=True)
    assert r.returncode != 0 and "cannot read" in r.stderr


def test_bad_scale_is_refused_before_anything_is_sized(cli, tmp_path):
    src = str(tmp_path / "in.ppm")
    cv2.imwrite(src, np.zeros((8, 8, 3), np.uint8))
    for args in (["-s", "0"], ["-s", "9"], ["-s", "4", "--aa", "4"], ["-s", "2", "--aa", "3"]):
        r = subprocess.run([cli, src, "-o", str(tmp_path / "o.png")] + args, capture_output=True, text=True)
        assert r.returncode == 2 and "unsupported scale" in r.stderr, args


def test_hostile_image_headers_are_refused(cli, tmp_path):
    """Header fields of untrusted files are bounded before anything is allocated or indexed (BMP data offset / height sign,
    PNG IHDR length and size, PNM digits): every one of these ends with an error message, not a crash."""
    import struct
    import zlib

    def bmp(off, w, h, bits=24, size=200):
        hdr = b"BM" + struct.pack("<IHHI", size, 0, 0, off & 0xFFFFFFFF) + struct.pack("<IiiHHIIiiII", 40, w, h, 1, bits, 0, 0, 0, 0, 0, 0)
        return hdr + b"\0" * (size - len(hdr))

    def png_chunk(kind, data):
        return struct.pack(">I", len(data)) + kind + data + struct.pack(">I", zlib.crc32(kind + data))

    sig = bytes([137, 80, 78, 71, 13, 10, 26, 10])
    cases = {
        "neg_offset.bmp": bmp(-64, 4, 4), "int_min_height.bmp": bmp(54, 4, -2 ** 31), "huge.bmp": bmp(54, 2 ** 30, 2 ** 30),
        "offset_past_end.bmp": bmp(1000, 4, 4), "short_ihdr.png": sig + png_chunk(b"IHDR", b"\0" * 8) + png_chunk(b"IEND", b"") + b"\0" * 16,
        "huge.png": sig + png_chunk(b"IHDR", struct.pack(">IIBBBBB", 2 ** 31 - 1, 2 ** 31 - 1, 8, 2, 0, 0, 0)) + png_chunk(b"IDAT", zlib.compress(b"\0")) + png_chunk(b"IEND", b""),
        "no_size.ppm": b"P6\n#only a comment\n", "overflow.ppm": b"P6 99999999999999999999 1 255\n\0\0\0", "letters.ppm": b"P6 a b 255\n",
        "truncated.ppm": b"P6 100 100 255\n\0\0\0",
    }
    for name, data in cases.items():
        path = tmp_path / name
        path.write_bytes(data)
        r = subprocess.run([cli, str(path), "--convert-only", "-o", str(tmp_path / "o.png")], capture_output=True, text=True, timeout=60)
        assert r.returncode == 1 and r.stderr.strip(), (name, r.returncode, r.stderr)


@pytest.mark.gpu
@pytest.mark.parametrize("strips", [0, 2])
def test_cli_image_in_image_out(cli, oracle, tmp_path, strips):
    """remaster_cli in.png -o out.png: equals the oracle's raster of the flipped frame, flipped back."""
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no GPU")
    frame = synth.snes_frame(96, 120, 17)                  # pipeline orientation (row 0 = bottom)
    src = str(tmp_path / "in.png")
    cv2.imwrite(src, _bgr_top_down(frame))
    dst, gpath, gimg = str(tmp_path / "out.png"), str(tmp_path / "graph.png"), str(tmp_path / "graph_picture.png")
    lpath = str(tmp_path / "labels.png")
    cmd = [cli, src, "-o", dst, "-s", "4", "--graph", gpath, "--graph-image", gimg, "--labels", lpath] + (["--strips", str(strips)] if strips else [])
    subprocess.run(cmd, check=True, timeout=300)
    # the CLI pads rows to 4 bytes like IplImage; 96*3 is already aligned, so the oracle sees the same bytes
    want = oracle.pipeline(frame, scale=4, want=("graph", "raster"))
    got = cv2.imread(dst, cv2.IMREAD_COLOR)
    assert np.array_equal(got, want["raster"][::-1, :, 2::-1])   # RGBA bottom-up -> BGR top-down
    assert np.array_equal(cv2.imread(gpath, cv2.IMREAD_GRAYSCALE), want["graph"][::-1])
    # --labels: an RGBA PNG whose R, G, B, A bytes are the little-endian bytes of the int32 label (lossless)
    lab_png = cv2.cvtColor(cv2.imread(lpath, cv2.IMREAD_UNCHANGED), cv2.COLOR_BGRA2RGBA)
    assert lab_png.shape == (120, 96, 4)
    assert np.array_equal(np.ascontiguousarray(lab_png).view("<i4")[..., 0], oracle.cc_labels(want["graph"])[::-1])
    # the graph picture of the same run equals the one drawn from the stored plane
    again = str(tmp_path / "graph_picture_again.png")
    subprocess.run([cli, src, "--draw-graph", gpath, "--graph-image", again], check=True, timeout=60)
    pic = cv2.imread(gimg, cv2.IMREAD_COLOR)
    assert pic.shape == (120 * 20, 96 * 20, 3) and np.array_equal(pic, cv2.imread(again, cv2.IMREAD_COLOR))


@pytest.mark.gpu
def test_cli_raw_video_stream(cli, oracle, tmp_path):
    """remaster_cli size.png --raw-video in.bin --raw-out out.bin: the frame stream the reference's author had wired in
    (main.cpp:67-75: frames of height * widthstep bytes in launch_kernel's layout, some skipped, a bounded number read):
    every remastered frame equals the oracle's raster of its source frame.  50 pixels wide: rows are padded to 152 bytes."""
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no GPU")
    w, h, n, skip, take = 50, 37, 70, 2, 67              # 67 frames: one full batch of 64 and a tail of 3
    frames = synth.snes_stream(n, w, h, first_seed=4242)
    cv2.imwrite(str(tmp_path / "size.png"), _bgr_top_down(frames[0]))
    padded = np.zeros((n, h, 152), np.uint8)
    padded[:, :, :3 * w] = frames.reshape(n, h, 3 * w)
    padded.tofile(str(tmp_path / "in.bin"))
    out = str(tmp_path / "out.bin")
    subprocess.run([cli, str(tmp_path / "size.png"), "--raw-video", str(tmp_path / "in.bin"), "--raw-out", out, "--skip", str(skip),
                    "--frames", str(take), "-s", "2"], check=True, timeout=300)
    got = np.fromfile(out, np.uint8).reshape(take, 2 * h, 2 * w, 4)
    for k in (0, 1, 63, 64, 66):
        src = np.lib.stride_tricks.as_strided(padded[skip + k], shape=(h, w, 3), strides=(152, 3, 1))  # the same padding bytes
        want = oracle.pipeline(src, scale=2, want=("raster",))["raster"]
        assert np.array_equal(got[k], want), k


@pytest.mark.gpu
def test_outlines_host_and_svg(cli, lib, oracle, tmp_path):
    """par_outlines_host (graph -> labels -> border walks -> closed quadratic B-splines, gathered on the host) against the
    oracle's walks and the numpy spline, and `remaster_cli --outlines out.svg`: one filled path per walk, in the colour of
    the walk's start pixel, with the same points (scaled, top scanline first)."""
    import re
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no GPU")
    W, H, S = 48, 40, 4
    frame = synth.snes_frame(W, H, 23)
    want = oracle.pipeline(frame, want=("graph", "labels"))
    walks = oracle.border_walks(want["graph"], want["labels"])
    with lib.Remaster(0, W, H, 1) as c:
        got = c.outlines_host(frame, samples=4)
    assert [s for s, _ in got] == sorted(walks)
    for start, pts in got:
        assert np.array_equal(pts.astype(np.float64), oracle.walk_spline(walks[start], W, 4)), start
    src, svg = str(tmp_path / "in.png"), str(tmp_path / "out.svg")
    cv2.imwrite(src, _bgr_top_down(frame))
    subprocess.run([cli, src, "-o", str(tmp_path / "o.png"), "-s", str(S), "--outlines", svg], check=True, timeout=300)
    text = open(svg).read()
    paths = re.findall(r'<path fill="#([0-9a-f]{6})" d="([^"]*)"/>', text)
    assert len(paths) == len(got) and 'viewBox="0 0 %d %d"' % (W * S, H * S) in text
    for (fill, d), (start, pts) in zip(paths, got):
        b, g, r = frame[start // W, start % W]
        assert fill == "%02x%02x%02x" % (r, g, b)
        xy = np.array([[float(v) for v in m] for m in re.findall(r"[ML]([-0-9.e+]+) ([-0-9.e+]+)", d)])
        assert xy.shape == pts.shape and d.endswith("Z")
        np.testing.assert_allclose(xy, np.stack([pts[:, 0] * S, (H - pts[:, 1]) * S], -1), rtol=2e-4, atol=1e-3)
