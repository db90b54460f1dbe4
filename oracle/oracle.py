"""TEST INFRASTRUCTURE ONLY: ctypes bindings for the checkers under oracle/.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module.  The product package never does (tests/test_boundary.py greps for it).

  Oracle      -> oracle/liboracle.so            our plain-C restatement (remaster_oracle.c)
  RefHost     -> oracle/_ref/libref_host*.so    the reference's own .cu files as serial host C++
  RefCuda     -> oracle/_ref/libref_cuda.so     the reference's kernel.cu for sm_100a (GPU box only)
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SLOTS = 45

_u8p = np.ctypeslib.ndpointer(np.uint8, flags="C_CONTIGUOUS")
_i32p = np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS")
_f32p = np.ctypeslib.ndpointer(np.float32, flags="C_CONTIGUOUS")


def build(ref=True):
    """Build the checkers (no-op for the _ref targets when /root/reference is absent)."""
    subprocess.run(["make", "-s", "-C", HERE, "liboracle.so"] + (["ref"] if ref else []), check=True)


def _opt(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class Oracle:
    """Plain-C restatement of the reference path (oracle/remaster_oracle.c)."""

    def __init__(self):
        path = os.path.join(HERE, "liboracle.so")
        if not os.path.exists(path):
            build(ref=False)
        L = self.lib = C.CDLL(path)
        L.orc_yuv_word.restype = C.c_uint32
        L.orc_yuv_word.argtypes = [C.c_int] * 4
        L.orc_dissimilar.argtypes = [C.c_uint32, C.c_uint32]
        L.orc_yuv_all.argtypes = [C.c_int, np.ctypeslib.ndpointer(np.uint32, flags="C_CONTIGUOUS")]
        L.orc_similarity_graph.argtypes = [_u8p, C.c_int, C.c_int, C.c_int, C.c_int, _u8p]
        L.orc_trivial_crossings.argtypes = [_u8p, C.c_int, C.c_int, _u8p]
        L.orc_resolve_crossings.argtypes = [_u8p, C.c_int, C.c_int, _u8p, C.POINTER(C.c_int)]
        L.orc_block_decision.argtypes = [_u8p, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int)]
        L.orc_cell_hull.argtypes = [C.c_uint, C.c_uint, C.c_uint, _f32p]
        L.orc_cells.argtypes = [_u8p, C.c_int, C.c_int, _f32p, _i32p]
        L.orc_subdivide.argtypes = [_u8p, C.c_int, C.c_int, C.c_int, _u8p, _f32p, _i32p, _f32p, _i32p]
        L.orc_ear_clip.argtypes = [_f32p, C.c_int, _f32p]
        L.orc_triangulate.argtypes = [_f32p, _i32p, C.c_int, C.c_int, _f32p, _i32p]
        L.orc_cc_labels.argtypes = [_u8p, C.c_int, C.c_int, _i32p]
        L.orc_border_walks.argtypes = [_u8p, _i32p, C.c_int, C.c_int, _i32p, _i32p, C.c_long]
        L.orc_border_walks.restype = C.c_long
        L.orc_all_dyadic64.argtypes = [_f32p, C.c_long]
        L.orc_raster_triangles.argtypes = [_u8p, C.c_int, C.c_int, C.c_int, C.c_int, _f32p, _i32p, _u8p]
        L.orc_raster_polygons.argtypes = [_u8p, C.c_int, C.c_int, C.c_int, C.c_int, _f32p, _i32p, _u8p]
        L.orc_pipeline.argtypes = [_u8p] + [C.c_int] * 6 + [C.c_void_p] * 10

    # -- single stages ---------------------------------------------------------------------
    def yuv_word(self, b0, b1, b2, fused=True):
        return int(self.lib.orc_yuv_word(int(b0), int(b1), int(b2), int(fused)))

    def yuv_all(self, fused=True):
        out = np.zeros(1 << 24, np.uint32)
        self.lib.orc_yuv_all(int(fused), out)
        return out

    def similarity_graph(self, img, fused=True):
        H, ws = img.shape[0], img.strides[0]
        W = img.shape[1]
        g = np.zeros((H, W), np.uint8)
        self.lib.orc_similarity_graph(_flat(img), W, H, ws, int(fused), g.reshape(-1))
        return g

    def trivial_crossings(self, g):
        H, W = g.shape
        o = np.zeros_like(g)
        self.lib.orc_trivial_crossings(np.ascontiguousarray(g).reshape(-1), W, H, o.reshape(-1))
        return o

    def resolve_crossings(self, aux):
        H, W = aux.shape
        o = np.zeros_like(aux)
        n = C.c_int(0)
        self.lib.orc_resolve_crossings(np.ascontiguousarray(aux).reshape(-1), W, H, o.reshape(-1), C.byref(n))
        return o, n.value

    def block_decision(self, aux, bi, bj):
        H, W = aux.shape
        steps = (C.c_int * 2)()
        d = self.lib.orc_block_decision(np.ascontiguousarray(aux).reshape(-1), W, bi, bj, steps)
        return d, (steps[0], steps[1])

    def cell_hull(self, node, left, right):
        xy = np.zeros(2 * SLOTS, np.float32)
        n = self.lib.orc_cell_hull(node, left, right, xy)
        return xy.reshape(SLOTS, 2)[: n + 1].copy(), n

    def cells(self, g):
        H, W = g.shape
        hull = np.zeros((H * W, SLOTS, 2), np.float32)
        cnt = np.zeros(H * W, np.int32)
        self.lib.orc_cells(np.ascontiguousarray(g).reshape(-1), W, H, hull.reshape(-1), cnt)
        return hull, cnt

    def subdivide(self, img, g, hull, cnt):
        H, W = g.shape
        poly = np.zeros_like(hull)
        pc = np.zeros_like(cnt)
        self.lib.orc_subdivide(_flat(img), W, H, img.strides[0], np.ascontiguousarray(g).reshape(-1),
                               hull.reshape(-1), cnt, poly.reshape(-1), pc)
        return poly, pc

    def triangulate(self, poly, pc, W, H):
        tri = np.zeros_like(poly)
        nt = np.zeros_like(pc)
        self.lib.orc_triangulate(poly.reshape(-1), pc, W, H, tri.reshape(-1), nt)
        return tri, nt

    def cc_labels(self, g):
        H, W = g.shape
        lab = np.zeros((H, W), np.int32)
        self.lib.orc_cc_labels(np.ascontiguousarray(g).reshape(-1), W, H, lab.reshape(-1))
        return lab

    def border_walks(self, g, labels=None):
        """{start node: [nodes]} — the first border walk of every component (dropped walks are absent)."""
        H, W = g.shape
        lab = self.cc_labels(g) if labels is None else labels
        cap = 16 * H * W + 64
        wl = np.zeros(H * W, np.int32)
        nodes = np.zeros(cap, np.int32)
        used = self.lib.orc_border_walks(np.ascontiguousarray(g).reshape(-1), np.ascontiguousarray(lab, np.int32).reshape(-1), W, H, wl, nodes, cap)
        assert used >= 0
        out, pos = {}, 0
        for n in np.nonzero(wl)[0]:
            out[int(n)] = nodes[pos:pos + wl[n]].tolist()
            pos += int(wl[n])
        return out

    @staticmethod
    def walk_spline(nodes, W, samples=4):
        """Closed uniform quadratic B-spline over one walk's node centres (the product's own rule, the reference stops at the
        walk, cc_functions.cu:348-503): (len * samples, 2) float64 points, segment i from the midpoint of nodes i-1, i to
        the midpoint of i, i+1, sampled at t = s / samples."""
        n = np.asarray(nodes, np.int64)
        P = np.stack([n % W + 0.5, n // W + 0.5], -1)
        prev, nxt = np.roll(P, 1, 0), np.roll(P, -1, 0)
        t = (np.arange(samples) / samples)[None, :, None]
        w0, w2 = 0.5 * (1 - t) ** 2, 0.5 * t ** 2
        pts = w0 * prev[:, None] + (1 - w0 - w2) * P[:, None] + w2 * nxt[:, None]
        return pts.reshape(-1, 2)

    def all_dyadic64(self, xy):
        a = np.ascontiguousarray(xy, np.float32).reshape(-1)
        return bool(self.lib.orc_all_dyadic64(a, a.size))

    def raster_triangles(self, img, s, tri, nt):
        H, W = img.shape[:2]
        out = np.zeros((s * H, s * W, 4), np.uint8)
        self.lib.orc_raster_triangles(_flat(img), W, H, img.strides[0], s, tri.reshape(-1), nt, out.reshape(-1))
        return out

    def raster_polygons(self, img, s, poly, pc):
        H, W = img.shape[:2]
        out = np.zeros((s * H, s * W, 4), np.uint8)
        self.lib.orc_raster_polygons(_flat(img), W, H, img.strides[0], s, poly.reshape(-1), pc, out.reshape(-1))
        return out

    # -- whole path ------------------------------------------------------------------------
    def pipeline(self, img, subdivide=True, fused=True, scale=4, want=("graph_aux", "graph")):
        """Run the restated path on one BGR8 frame (H, W, 3) (row 0 = bottom scanline).
        `want` selects outputs among graph_aux graph labels hull hull_count poly poly_count tri ntri raster."""
        H, W = img.shape[:2]
        N = H * W
        bufs = {
            "graph_aux": np.zeros((H, W), np.uint8), "graph": np.zeros((H, W), np.uint8),
            "labels": np.zeros((H, W), np.int32),
            "hull": np.zeros((N, SLOTS, 2), np.float32), "hull_count": np.zeros(N, np.int32),
            "poly": np.zeros((N, SLOTS, 2), np.float32), "poly_count": np.zeros(N, np.int32),
            "tri": np.zeros((N, SLOTS, 2), np.float32), "ntri": np.zeros(N, np.int32),
            "raster": np.zeros((scale * H, scale * W, 4), np.uint8),
        }
        order = ["graph_aux", "graph", "labels", "hull", "hull_count", "poly", "poly_count", "tri", "ntri", "raster"]
        args = [_opt(bufs[k]) if k in want else None for k in order]
        self.lib.orc_pipeline(_flat(img), W, H, img.strides[0], int(subdivide), int(fused), int(scale), *args)
        return {k: bufs[k] for k in order if k in want}


def _flat(img):
    """The frame as the flat byte array the reference indexes (rows `strides[0]` bytes apart)."""
    assert img.dtype == np.uint8 and img.ndim == 3 and img.shape[2] == 3 and img.strides[1] == 3 and img.strides[2] == 1
    H, ws = img.shape[0], img.strides[0]
    base = np.lib.stride_tricks.as_strided(img, shape=(H * ws,), strides=(1,)) if ws != img.shape[1] * 3 else img.reshape(-1)
    return np.ascontiguousarray(base)


class RefHost:
    """The reference's own routines as serial host C++ (oracle/ref_host_driver.cpp)."""

    def __init__(self, fma=True):
        path = os.path.join(HERE, "_ref", "libref_host_fma.so" if fma else "libref_host.so")
        if not os.path.exists(path):
            raise FileNotFoundError(path)
        L = self.lib = C.CDLL(path)
        L.ref_host_rgb_to_yuv.restype = C.c_uint32
        L.ref_host_rgb_to_yuv.argtypes = [C.c_int]
        L.ref_host_cell.argtypes = [C.c_int, C.c_int, C.c_int, _f32p]
        L.ref_host_yuv_all.argtypes = [np.ctypeslib.ndpointer(np.uint32, flags="C_CONTIGUOUS")]
        L.ref_host_pipeline.argtypes = [_u8p] + [C.c_int] * 4 + [C.c_void_p] * 8

    @staticmethod
    def available(fma=True):
        return os.path.exists(os.path.join(HERE, "_ref", "libref_host_fma.so" if fma else "libref_host.so"))

    def yuv_word(self, c):
        return int(self.lib.ref_host_rgb_to_yuv(int(c)))

    def yuv_all(self):
        out = np.zeros(1 << 24, np.uint32)
        self.lib.ref_host_yuv_all(out)
        return out

    def cell(self, node, left, right):
        xy = np.zeros(2 * SLOTS, np.float32)
        n = self.lib.ref_host_cell(node, left, right, xy)
        return xy.reshape(SLOTS, 2)[: n + 1].copy(), n

    def pipeline(self, img, subdivide=True, want=("graph_aux", "graph"), stage_ms=None):
        H, W = img.shape[:2]
        N = H * W
        bufs = {
            "graph_aux": np.zeros((H, W), np.uint8), "graph": np.zeros((H, W), np.uint8),
            "hull": np.zeros((N, SLOTS, 2), np.float32), "hull_count": np.zeros(N, np.int32),
            "poly": np.zeros((N, SLOTS, 2), np.float32), "poly_count": np.zeros(N, np.int32),
            "tri": np.zeros((N, SLOTS, 2), np.float32),
        }
        order = ["graph_aux", "graph", "hull", "hull_count", "poly", "poly_count", "tri"]
        args = [_opt(bufs[k]) if k in want else None for k in order]
        ms = np.zeros(5, np.float64)
        self.lib.ref_host_pipeline(_flat(img), W, H, img.strides[0], int(subdivide), *args, _opt(ms))
        if stage_ms is not None:
            stage_ms[:] = ms
        return {k: bufs[k] for k in order if k in want}


class RefHostCC:
    """The reference's dead-code border walker (cc_functions.cu) as host C++ (oracle/ref_host_cc_driver.cpp)."""

    def __init__(self):
        path = os.path.join(HERE, "_ref", "libref_host_cc.so")
        if not os.path.exists(path):
            raise FileNotFoundError(path)
        self.lib = C.CDLL(path)
        self.lib.ref_host_border_walks.argtypes = [_u8p, C.c_int, C.c_int, _i32p, _i32p, C.c_int]

    @staticmethod
    def available():
        return os.path.exists(os.path.join(HERE, "_ref", "libref_host_cc.so"))

    def border_walks(self, graph):
        H, W = graph.shape
        cap = 4 * H * W + 64
        lst = np.zeros(cap, np.int32)
        sizes = np.zeros(cap, np.int32)
        n = self.lib.ref_host_border_walks(np.ascontiguousarray(graph).reshape(-1), W, H, lst, sizes, cap)
        out, pos = [], 0
        for k in range(n):
            out.append(lst[pos:pos + sizes[k]].tolist())
            pos += sizes[k]
        return out


class RefCuda:
    """The reference's kernel.cu compiled for sm_100a (oracle/ref_cuda_driver.cu).  Needs a GPU."""

    def __init__(self):
        path = os.path.join(HERE, "_ref", "libref_cuda.so")
        if not os.path.exists(path):
            raise FileNotFoundError(path)
        L = self.lib = C.CDLL(path)
        L.ref_cuda_yuv_all.argtypes = [np.ctypeslib.ndpointer(np.uint32, flags="C_CONTIGUOUS")]
        L.ref_cuda_graph_aux.argtypes = [_u8p, C.c_int, C.c_int, C.c_int, _u8p]
        L.ref_cuda_launch.argtypes = [_u8p] + [C.c_int] * 4 + [C.c_void_p] * 5 + [C.POINTER(C.c_double)]
        L.ref_cuda_time_calls.restype = C.c_double
        L.ref_cuda_time_calls.argtypes = [_u8p] + [C.c_int] * 5

    @staticmethod
    def available():
        return os.path.exists(os.path.join(HERE, "_ref", "libref_cuda.so"))

    def yuv_all(self):
        out = np.zeros(1 << 24, np.uint32)
        rc = self.lib.ref_cuda_yuv_all(out)
        assert rc == 0, rc
        return out

    def graph_aux(self, img):
        H, W = img.shape[:2]
        g = np.zeros((H, W), np.uint8)
        assert self.lib.ref_cuda_graph_aux(_flat(img), W, H, img.strides[0], g.reshape(-1)) == 0
        return g

    def launch(self, img, subdivide=True, want=("graph", "edge_count", "diagram")):
        H, W = img.shape[:2]
        N = H * W
        bufs = {
            "graph": np.zeros((H, W), np.uint8), "edge_count": np.zeros(N, np.int32),
            "diagram": np.zeros((N, SLOTS, 2), np.float32), "pos": np.zeros((N, SLOTS, 2), np.float32),
            "color": np.zeros((N, SLOTS, 4), np.uint8),
        }
        order = ["graph", "edge_count", "diagram", "pos", "color"]
        args = [_opt(bufs[k]) if k in want else None for k in order]
        ms = C.c_double(0)
        rc = self.lib.ref_cuda_launch(_flat(img), W, H, img.strides[0], int(subdivide), *args, C.byref(ms))
        assert rc == 0, rc
        out = {k: bufs[k] for k in order if k in want}
        out["wall_ms"] = ms.value
        return out

    def time_calls(self, img, subdivide=True, calls=10):
        H, W = img.shape[:2]
        return float(self.lib.ref_cuda_time_calls(_flat(img), W, H, img.strides[0], int(subdivide), int(calls)))
