"""Stage D of the PRODUCT (host-side table built by csrc/cell_table.cpp, exported through
par_cell_from_pattern) against the golden table generated from the reference and against the oracle."""
import os

import numpy as np

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_product_cell_table_equals_reference(lib, oracle):
    z = np.load(os.path.join(GOLD, "cells_4096.npz"))
    hist = {}
    for key in range(4096):
        xy, n = lib.cell_from_pattern(key)
        assert n == int(z["count"][key]), key
        assert np.array_equal(xy * 4, z["verts_q4"][key, : n + 1].astype(np.float32)), key
        node = key & 255
        left = (4 if key & 256 else 0) | (128 if key & 512 else 0)
        right = (1 if key & 1024 else 0) | (32 if key & 2048 else 0)
        oxy, on = oracle.cell_hull(node, left, right)
        assert on == n and np.array_equal(oxy, xy)
        hist[n] = hist.get(n, 0) + 1
    assert hist == {4: 264, 5: 1136, 6: 1544, 7: 896, 8: 256}  # SURVEY App. A.4 probe


def test_product_yuv_word_equals_golden(lib):
    z = np.load(os.path.join(GOLD, "yuv_words.npz"))
    for c, f in zip(z["colour"], z["fused"]):
        assert lib.yuv_word(int(c) & 255, (int(c) >> 8) & 255, int(c) >> 16) == int(f)
