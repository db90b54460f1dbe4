"""numpy construction of stage A+B (similarity graph with trivial crossings removed) from per-pixel packed-YUV words.

Test helper: builds the expected `graph_aux` of frames far too large for the serial oracle (the 4096 x 4096 frame
that holds all 2^24 colours) from the oracle's own conversion table (`Oracle.yuv_all`), and is itself pinned against
the oracle's stages A and B on small frames (tests/test_golden.py::test_numpy_graph_builder_equals_the_oracle).
Semantics: graph_functions.cu:147-311 (diff; thresholds :14-19), :1211-1274 (crossCheck_4).
"""
import numpy as np

# graph bit e <-> neighbour offset (di, dj)   (graph_functions.cu:162-171)
EDGES = ((-1, 1), (0, 1), (1, 1), (-1, 0), (1, 0), (-1, -1), (0, -1), (1, -1))


def similar(p, q):
    """1 where two arrays of packed YUV words are similar: |dY| <= 5, |dU| <= 7, |dV| <= 6 on the packed fields."""
    p = p.astype(np.int64)
    q = q.astype(np.int64)
    dy = np.abs((p & 0xFF0000) - (q & 0xFF0000))
    du = np.abs((p & 0x00FF00) - (q & 0x00FF00))
    dv = np.abs((p & 0x0000FF) - (q & 0x0000FF))
    return (dy <= 0x50000) & (du <= 0x700) & (dv <= 6)


def graph_aux_from_yuv(yuv):
    """yuv: (H, W) uint32 packed words, row 0 = bottom.  Returns the (H, W) uint8 graph after stages A and B."""
    H, W = yuv.shape
    right = np.zeros((H, W), bool)   # (i,j) ~ (i+1,j)
    up = np.zeros((H, W), bool)      # (i,j) ~ (i,j+1)
    ur = np.zeros((H, W), bool)      # (i,j) ~ (i+1,j+1)   "/"
    ul = np.zeros((H, W), bool)      # (i,j) ~ (i-1,j+1)   "\"
    if W > 1:
        right[:, :-1] = similar(yuv[:, :-1], yuv[:, 1:])
    if H > 1:
        up[:-1, :] = similar(yuv[:-1, :], yuv[1:, :])
    if W > 1 and H > 1:
        ur[:-1, :-1] = similar(yuv[:-1, :-1], yuv[1:, 1:])
        ul[:-1, 1:] = similar(yuv[:-1, 1:], yuv[1:, :-1])
        # stage B: a 2x2 block (lower-left pixel (i,j)) whose four sides are all linked loses both diagonals
        full = right[:-1, :-1] & right[1:, :-1] & up[:-1, :-1] & up[:-1, 1:]
        ur[:-1, :-1] &= ~full
        ul[:-1, 1:] &= ~full
    g = np.zeros((H, W), np.uint8)
    g[:, :] |= (right.astype(np.uint8) << 4)            # bit 4: right
    g[:, 1:] |= (right[:, :-1].astype(np.uint8) << 3)   # bit 3: left
    g[:, :] |= (up.astype(np.uint8) << 1)               # bit 1: up
    g[1:, :] |= (up[:-1, :].astype(np.uint8) << 6)      # bit 6: down
    g[:, :] |= (ur.astype(np.uint8) << 2)               # bit 2: up-right
    g[1:, 1:] |= (ur[:-1, :-1].astype(np.uint8) << 5)   # bit 5: down-left
    g[:, :] |= (ul.astype(np.uint8) << 0)               # bit 0: up-left
    g[1:, :-1] |= (ul[:-1, 1:].astype(np.uint8) << 7)   # bit 7: down-right
    return g


def all_colours_frame():
    """The 4096 x 4096 BGR frame that holds every 24-bit colour once: pixel n (row-major) has bytes (n & 255, n >> 8 & 255, n >> 16)."""
    c = np.arange(1 << 24, dtype=np.uint32).reshape(4096, 4096)
    return np.stack([c & 255, (c >> 8) & 255, c >> 16], -1).astype(np.uint8)
