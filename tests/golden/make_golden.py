"""Generates the golden fixtures under tests/golden/ from the reference itself (run in the build
container, where /root/reference exists):

    python tests/golden/make_golden.py

  graph_dumps.npz     the two graph dumps embedded in the reference (kernel.cu:291-293 8x8 "c_pattern",
                      kernel.cu:296 24x24 "alex") and the border walks of alex_png.txt:1-21
  cells_4096.npz      createCellFromPattern for all 4096 pattern keys, from the reference's own code
                      compiled as host C++ (oracle/_ref/libref_host.so): counts + vertices in quarter pixels
  yuv_words.npz       RGBtoYUV of 8192 colours (all greys, all single-channel ramps, seeded random) from the
                      FMA-contracted host build (device-like Y) and from the plain build
  frames_*.npz        four small seeded frames with every stage output of the reference host build:
                      graph_aux, graph, hull/poly counts + vertices, triangle lists
  border_walks_alex.json  what the reference's dead-code walker (cc_functions.cu) returns for the alex dump

The frames themselves are stored too, so the fixtures do not depend on the generator staying stable.
"""
import json
import os
import re
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = "/root/reference"

from oracle.oracle import RefHost, RefHostCC, build  # noqa: E402
from pixel_art_remaster_gpu_b200 import synth  # noqa: E402


def parse_dumps():
    lines = open(os.path.join(REF, "kernel.cu")).read().split("\n")
    c_txt = " ".join(lines[290:293])           # kernel.cu:291-293
    a_txt = lines[295]                          # kernel.cu:296
    c_vals = [int(v) for v in re.findall(r"-?\d+", c_txt.split("{", 1)[1].split("}", 1)[0])]
    a_vals = [int(v) for v in re.findall(r"\(char\)\s*(-?\d+)", a_txt)]
    assert len(c_vals) == 64 and len(a_vals) == 576, (len(c_vals), len(a_vals))
    walks = []
    for ln in open(os.path.join(REF, "alex_png.txt")).read().split("\n")[:21]:
        m = re.match(r"Spline:\s*(\d+)\s*-\s*(.*)", ln)
        if m:
            walks.append([int(v) for v in re.findall(r"\d+", m.group(2))])
    return (np.array(c_vals, np.int64).astype(np.uint8).reshape(8, 8), np.array(a_vals, np.int64).astype(np.uint8).reshape(24, 24), walks)


def main():
    build(ref=True)
    rh, rp = RefHost(fma=True), RefHost(fma=False)
    c_pat, alex, walks = parse_dumps()
    np.savez_compressed(os.path.join(HERE, "graph_dumps.npz"), c_pattern=c_pat, alex=alex,
                        alex_walks=np.array([w + [-1] * (80 - len(w)) for w in walks], np.int32))
    json.dump(RefHostCC().border_walks(alex), open(os.path.join(HERE, "border_walks_alex.json"), "w"))

    cnt = np.zeros(4096, np.int8)
    verts = np.zeros((4096, 9, 2), np.int8)
    for key in range(4096):
        node = key & 255
        left = (4 if key & 256 else 0) | (128 if key & 512 else 0)
        right = (1 if key & 1024 else 0) | (32 if key & 2048 else 0)
        xy, n = rp.cell(node, left, right)
        cnt[key] = n
        verts[key, : n + 1] = np.round(xy * 4).astype(np.int8)
        assert np.array_equal(verts[key, : n + 1] / 4.0, xy)
    np.savez_compressed(os.path.join(HERE, "cells_4096.npz"), count=cnt, verts_q4=verts)

    rng = np.random.default_rng(20261017)
    cols = np.concatenate([np.arange(256) * 0x010101, np.arange(256), np.arange(256) << 8, np.arange(256) << 16,
                           rng.integers(0, 1 << 24, 8192 - 1024)]).astype(np.int64)
    np.savez_compressed(os.path.join(HERE, "yuv_words.npz"), colour=cols.astype(np.uint32),
                        fused=np.array([rh.yuv_word(int(c)) for c in cols], np.uint32),
                        plain=np.array([rp.yuv_word(int(c)) for c in cols], np.uint32))

    frames = {
        "g1_40x30": synth.snes_frame(40, 30, synth.BASE_SEED + 11),
        "g1_64x48": synth.snes_frame(64, 48, synth.BASE_SEED + 12),
        "g5_72x60": synth.adversarial_sprite(72, 60, synth.BASE_SEED + 13),
        "padded_33x21_ws104": synth.pad_rows(synth.snes_frame(33, 21, synth.BASE_SEED + 14), 104),
    }
    for name, img in frames.items():
        Hh, Ww = img.shape[:2]
        ws = img.strides[0]
        raw = np.lib.stride_tricks.as_strided(img, shape=(Hh, ws), strides=(ws, 1)).copy()
        out = {"raw_rows": raw, "width": Ww, "height": Hh, "widthstep": ws}
        for sub in (0, 1):
            r = rh.pipeline(img, bool(sub), ("graph_aux", "graph", "hull", "hull_count", "poly", "poly_count", "tri"))
            # subdivided coordinates are multiples of 1/64: store as int16 in 1/64 units (exact);
            # slots past the vertex count are uninitialised stack in the reference (diagram_functions.cu:301-303)
            r["poly"][np.arange(45)[None, :] >= r["poly_count"][:, None]] = 0
            r["hull"][np.arange(45)[None, :] > r["hull_count"][:, None]] = 0
            assert np.array_equal(np.round(r["poly"] * 64) / 64, r["poly"])
            if sub:
                out.update(graph_aux=r["graph_aux"], graph=r["graph"], hull_count=r["hull_count"].astype(np.int8),
                           hull_q4=np.round(r["hull"][:, :9] * 4).astype(np.int8))
            out["poly_count_sub%d" % sub] = r["poly_count"].astype(np.int8)
            out["poly_q64_sub%d" % sub] = np.round(r["poly"][:, :16] * 64).astype(np.int16)
            ntri = np.maximum(r["poly_count"] - 2, 0)
            tri = r["tri"].copy()
            tri[np.arange(45)[None, :] >= 3 * ntri[:, None]] = 0  # slots past the triangle list are undefined in the reference
            assert np.array_equal(np.round(tri * 64) / 64, tri)
            out["tri_q64_sub%d" % sub] = np.round(tri[:, :42] * 64).astype(np.int32)
        np.savez_compressed(os.path.join(HERE, "frames_%s.npz" % name), **out)
    print("golden fixtures written to", HERE)


if __name__ == "__main__":
    main()
