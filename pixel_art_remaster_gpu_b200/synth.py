"""Synthetic pixel-art inputs for the BASELINE.json configs (SURVEY.md §8(d) generators G1-G5).

All frames are BGR8, `widthstep == 3*W`, row 0 = bottom scanline (the layout `launch_kernel`
receives after `main.cpp:50-64` has flipped the OpenCV image).  Every generator is a pure function
of (shape, seed): frame k of a stream uses seed `BASE_SEED + k`.
"""
import numpy as np

BASE_SEED = 0xC0FFEE


def _rng(seed):
    return np.random.Generator(np.random.PCG64(int(seed)))


def _tile_bank(rng, n_tiles=64, n_colours=16):
    """64 8x8 tiles of palette indices: half-plane splits, 2x2-blocky blobs, dither and sparse dots."""
    bank = np.zeros((n_tiles, 8, 8), np.uint8)
    yy, xx = np.mgrid[0:8, 0:8]
    for t in range(n_tiles):
        k = int(rng.integers(2, 5))
        cols = rng.choice(n_colours, size=k, replace=False).astype(np.uint8)
        kind = int(rng.integers(0, 5))
        if kind == 0:  # straight or diagonal edge
            a, b = rng.integers(-2, 3, size=2)
            if a == 0 and b == 0:
                a = 1
            c = rng.integers(-4, 12)
            idx = ((a * xx + b * yy) > c).astype(np.uint8)
            if k > 2:
                idx = idx + ((a * xx - b * yy) > c + 3).astype(np.uint8)
        elif kind == 1:  # blocky blobs
            idx = np.kron(rng.integers(0, k, size=(4, 4)), np.ones((2, 2), np.int64)).astype(np.uint8)
        elif kind == 2:  # per-pixel dither
            idx = rng.integers(0, k, size=(8, 8)).astype(np.uint8)
        elif kind == 3:  # flat field with a few single pixels and a diagonal run
            idx = np.zeros((8, 8), np.uint8)
            for _ in range(int(rng.integers(1, 5))):
                idx[rng.integers(0, 8), rng.integers(0, 8)] = rng.integers(1, k)
            d = int(rng.integers(0, 8))
            for s in range(int(rng.integers(2, 8))):
                idx[(d + s) % 8, s] = k - 1
        else:  # diagonal stripes
            w = int(rng.integers(1, 4))
            sgn = 1 if rng.integers(0, 2) else -1
            idx = (((xx + sgn * yy) // w) % k).astype(np.uint8)
        bank[t] = cols[np.minimum(idx, k - 1)]
    return bank


def snes_frame(W=256, H=224, seed=BASE_SEED, n_colours=16):
    """G1/G2: SNES-style frame from a 16-colour palette; returns (H, W, 3) uint8 BGR."""
    rng = _rng(seed)
    palette = rng.integers(0, 256, size=(n_colours, 3), dtype=np.uint8)
    bank = _tile_bank(rng, 64, n_colours)
    th, tw = (H + 7) // 8, (W + 7) // 8
    tmap = rng.integers(0, 64, size=(th, tw))
    # a few horizontally repeated runs of one tile (motifs / flat areas)
    for _ in range(th):
        r, c0 = int(rng.integers(0, th)), int(rng.integers(0, tw))
        tmap[r, c0:c0 + int(rng.integers(2, 9))] = tmap[r, c0]
    idx = bank[tmap].transpose(0, 2, 1, 3).reshape(th * 8, tw * 8)[:H, :W]
    noise = rng.random((H, W)) < 0.05
    idx = np.where(noise, rng.integers(0, n_colours, size=(H, W), dtype=np.uint8), idx)
    return np.ascontiguousarray(palette[idx])


def snes_stream(n_frames, W=256, H=224, first_seed=BASE_SEED):
    """G3: (n_frames, H, W, 3) stack of consecutive-seed G1 frames."""
    out = np.empty((n_frames, H, W, 3), np.uint8)
    for k in range(n_frames):
        out[k] = snes_frame(W, H, first_seed + k)
    return out


def pixel_art_map(W=4096, H=4096, seed=BASE_SEED):
    """G4: one large map built by the G1 tile generator."""
    return snes_frame(W, H, seed)


def adversarial_sprite(W=512, H=448, seed=BASE_SEED):
    """G5: checkerboard dither | diagonal stripes with long anti-diagonal polylines | islands, all
    crossed by long 1-pixel random walks.  Stresses chain walks, the island rule and union-find."""
    rng = _rng(seed)
    # four mutually dissimilar colours + one colour 1 LSB inside the thresholds of colour 0
    cols = np.array([[20, 30, 200], [220, 210, 40], [30, 200, 60], [240, 240, 240], [20, 35, 200]], np.uint8)
    img = np.zeros((H, W), np.uint8)
    yy, xx = np.mgrid[0:H, 0:W]
    a, b = W // 3, 2 * W // 3
    img[:, :a] = ((xx[:, :a] + yy[:, :a]) & 1).astype(np.uint8)            # (i) checkerboard
    img[:, a:b] = ((xx[:, a:b] + yy[:, a:b]) % 3).astype(np.uint8)         # (ii) diagonal stripes
    for _ in range(24):                                                     #      anti-diagonal polylines
        x, y = int(rng.integers(a, b)), int(rng.integers(0, H))
        for s in range(int(rng.integers(40, 120))):
            if a <= x < b and 0 <= y < H:
                img[y, x] = 3
            x += 1
            y -= 1
            if s % 17 == 16:
                y += int(rng.integers(0, 3))
        if a <= x < b - 1 and 1 <= y < H - 1:                               #      T / Y junction at the end
            img[y, x] = 3
            img[y + 1, x] = 3
            img[y - 1, x + 1] = 3
    img[:, b:] = 2                                                          # (iii) islands on a flat field
    n_isl = (W - b) * H // 40
    ys, xs = rng.integers(1, H - 2, n_isl), rng.integers(b + 1, W - 2, n_isl)
    img[ys, xs] = 0
    pair = rng.random(n_isl) < 0.5
    img[ys[pair] + 1, xs[pair] + 1] = 0
    for d in range(0, min(H, W - b) - 2, 1):                                #      a long diagonal line through them
        img[1 + d, b + 1 + d] = 3
    img[:, b + (W - b) // 2:][(yy[:, b + (W - b) // 2:] % 11 == 0)] = 4     #      near-threshold colour bands
    for _ in range(6):                                                      # long 8-connected random walks
        x, y = int(rng.integers(0, W)), int(rng.integers(0, H))
        for _s in range(200):
            img[y % H, x % W] = 3
            x += int(rng.integers(-1, 2))
            y += int(rng.integers(-1, 2))
    return np.ascontiguousarray(cols[img])


def pad_rows(frame, widthstep):
    """Re-lay a (H, W, 3) frame with `widthstep` bytes per row (OpenCV 4-byte row alignment etc.);
    padding bytes are filled with a non-zero pattern so that a kernel ignoring the stride is caught."""
    H, W, _ = frame.shape
    assert widthstep >= 3 * W
    buf = np.full((H, widthstep), 0xA5, np.uint8)
    buf[:, :3 * W] = frame.reshape(H, 3 * W)
    return np.lib.stride_tricks.as_strided(buf, shape=(H, W, 3), strides=(widthstep, 3, 1))
