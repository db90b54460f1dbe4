"""B200-native pixel-art remaster path — Python host side over the C ABI (include/pixelart_b200.h).

The product is the CUDA library `libpixelart_b200.so` (hand-written sm_100a kernels); this module
only binds it with ctypes and moves pointers around.  torch is used for device memory and streams,
nothing else.  There is NO CPU fallback: importing works anywhere (so the build can be checked on a
CPU box) but creating a `Remaster` context without a B200 raises.

Mirrors the reference's interface for the path:
  reference `launch_kernel(pos, colorPos, time, img_data, w, h, widthstep, edge_count_h, graph_h,
  subdivide)` (kernel.cu:286-288)  ->  `Remaster.remaster(...)` / `launch_kernel(...)` below.
"""
import ctypes as C
import os

import numpy as np

from . import synth  # noqa: F401  (synthetic inputs for the BASELINE configs)

__all__ = ["Remaster", "RemasterGroup", "launch_kernel", "RemasterError", "load_library", "library_path", "cell_from_pattern", "yuv_word",
           "FLAG_SUBDIVIDE", "FLAG_FLIP_OUTPUT", "FLAG_NO_TMA", "CELL_SLOTS", "OUT_RGBA8", "OUT_BGR8", "OUT_INDEX8", "synth"]

HERE = os.path.dirname(os.path.abspath(__file__))
CELL_SLOTS = 45
FLAG_SUBDIVIDE, FLAG_FLIP_OUTPUT, FLAG_NO_TMA, FLAG_DEBUG_WIDE, FLAG_NO_SMOOTH_TABLES, FLAG_AA2, FLAG_AA4 = 1, 2, 4, 8, 16, 32, 64
OUT_RGBA8, OUT_BGR8, OUT_INDEX8 = 0, 1, 2   # par_out_format
_BPP = {OUT_RGBA8: 4, OUT_BGR8: 3, OUT_INDEX8: 1}
_STATUS = {0: "PAR_OK", 1: "PAR_ERR_INVALID", 2: "PAR_ERR_NO_DEVICE", 3: "PAR_ERR_CUDA", 4: "PAR_ERR_CAPACITY"}


class RemasterError(RuntimeError):
    def __init__(self, status, message):
        super().__init__("%s: %s" % (_STATUS.get(status, status), message))
        self.status = status


class ParJob(C.Structure):
    _fields_ = [("bgr", C.c_void_p), ("width", C.c_int), ("height", C.c_int), ("widthstep", C.c_int),
                ("frame_stride", C.c_size_t), ("n_frames", C.c_int), ("scale", C.c_int), ("flags", C.c_uint),
                ("rgba", C.c_void_p), ("graph", C.c_void_p), ("graph_aux", C.c_void_p), ("labels", C.c_void_p),
                ("polygons", C.c_void_p), ("poly_count", C.c_void_p),
                ("out_format", C.c_int), ("palette", C.c_void_p), ("palette_count", C.c_void_p)]


class ParOutlines(C.Structure):
    _fields_ = [("n_walks", C.c_int), ("samples", C.c_int), ("start", C.POINTER(C.c_int32)), ("count", C.POINTER(C.c_int32)),
                ("n_points", C.c_longlong), ("points", C.POINTER(C.c_float))]


class ParStrip(C.Structure):
    _fields_ = [("device", C.c_int), ("own_begin", C.c_int), ("own_end", C.c_int), ("load_begin", C.c_int), ("load_end", C.c_int),
                ("bgr", C.c_void_p), ("image", C.c_void_p), ("graph", C.c_void_p), ("graph_aux", C.c_void_p), ("labels", C.c_void_p)]


class _DeviceBuffer:
    """A raw device allocation of the library seen as an array (torch.as_tensor aliases it through this interface)."""

    def __init__(self, ptr, shape, typestr):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False), "version": 2}


_lib = None


def library_path():
    return os.path.join(HERE, "libpixelart_b200.so")


def load_library():
    """Load the CUDA library; raises if it has not been built (python -m pixel_art_remaster_gpu_b200.build)."""
    global _lib
    if _lib is not None:
        return _lib
    path = library_path()
    if not os.path.exists(path):
        raise ImportError("%s is missing: build it with `python -m pixel_art_remaster_gpu_b200.build` "
                          "(there is no CPU fallback)" % path)
    L = C.CDLL(path)
    P = C.POINTER
    L.par_create.argtypes = [P(C.c_void_p), C.c_int, C.c_int, C.c_int, C.c_int]
    L.par_destroy.argtypes = [C.c_void_p]
    L.par_destroy.restype = None
    L.par_last_error.argtypes = [C.c_void_p]
    L.par_last_error.restype = C.c_char_p
    L.par_device.argtypes = [C.c_void_p]
    L.par_set_stream.argtypes = [C.c_void_p, C.c_void_p]
    L.par_use_own_stream.argtypes = [C.c_void_p]
    L.par_set_sub_batch.argtypes = [C.c_void_p, C.c_int]
    L.par_border_walks.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                   C.c_longlong, C.c_void_p]
    L.par_walk_splines.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_longlong, C.c_int,
                                   C.c_void_p]
    L.par_outlines_host.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, P(ParOutlines)]
    L.par_outlines_free.argtypes = [P(ParOutlines)]
    L.par_outlines_free.restype = None
    L.par_smooth_stats.argtypes = [C.c_void_p, P(C.c_uint64)]
    L.par_profile_enable.argtypes = [C.c_void_p, C.c_int]
    L.par_profile_read.argtypes = [C.c_void_p, P(C.c_double), P(C.c_int)]
    L.par_synchronize.argtypes = [C.c_void_p]
    L.par_launch_count.argtypes = [C.c_void_p]
    L.par_launch_count.restype = C.c_uint64
    for name in ("par_remaster_device", "par_remaster_host", "par_stage_similarity_graph", "par_stage_resolve_crossings",
                 "par_stage_cc_labels", "par_stage_polygons", "par_stage_raster"):
        getattr(L, name).argtypes = [C.c_void_p, P(ParJob)]
    L.par_group_create.argtypes = [P(C.c_void_p), P(C.c_int), C.c_int, C.c_int, C.c_int, C.c_int]
    L.par_group_destroy.argtypes = [C.c_void_p]
    L.par_group_destroy.restype = None
    L.par_group_last_error.argtypes = [C.c_void_p]
    L.par_group_last_error.restype = C.c_char_p
    L.par_group_remaster_host.argtypes = [C.c_void_p, P(ParJob)]
    L.par_group_n_strips.argtypes = [C.c_void_p]
    L.par_group_strip.argtypes = [C.c_void_p, C.c_int, P(ParStrip)]
    L.par_group_remaster_device.argtypes = [C.c_void_p, C.c_uint, C.c_int, C.c_int, C.c_int]
    L.par_group_last_ms.argtypes = [C.c_void_p, P(C.c_double), P(C.c_double)]
    L.launch_kernel.argtypes = [C.c_void_p, C.c_void_p, C.c_float, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_bool]
    L.launch_kernel.restype = C.c_void_p
    L.par_cell_from_pattern.argtypes = [C.c_uint, P(C.c_float)]
    L.par_yuv_word.argtypes = [C.c_int, C.c_int, C.c_int]
    L.par_yuv_word.restype = C.c_uint32
    _lib = L
    return L


def cell_from_pattern(key):
    """Vertices (count+1, 2) of the cell of a 12-bit pattern key and its vertex count (host-side table)."""
    L = load_library()
    buf = (C.c_float * (2 * CELL_SLOTS))()
    n = L.par_cell_from_pattern(int(key), buf)
    if n < 0:
        raise ValueError("bad key %r" % (key,))
    return np.array(buf[: 2 * (n + 1)], np.float32).reshape(n + 1, 2), n


def yuv_word(b0, b1, b2):
    return int(load_library().par_yuv_word(int(b0), int(b1), int(b2)))


def _ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


class Remaster:
    """One remaster context on one GPU (par_context).  Device tensors in, device tensors out."""

    def __init__(self, device=0, max_width=256, max_height=224, max_frames=64):
        import torch
        self._torch = torch
        self.lib = load_library()
        self.handle = C.c_void_p()
        st = self.lib.par_create(C.byref(self.handle), int(device), int(max_width), int(max_height), int(max_frames))
        if st != 0:
            self.handle = C.c_void_p()
            raise RemasterError(st, self.lib.par_last_error(None).decode())
        self.device = torch.device("cuda", int(device))
        self._pinned_stream = None
        self._bind_stream()

    # -- plumbing --------------------------------------------------------------------------
    def close(self):
        h = getattr(self, "handle", None)
        if h is not None and h.value:
            self.lib.par_destroy(h)
            h.value = None

    def __del__(self):
        try:
            self.close()
        except Exception:  # interpreter shutdown
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def use_torch_stream(self, stream=None):
        """Pin all work of this context to one torch stream.  With stream=None (the default state) every call runs on
        torch's CURRENT stream of the context's device at the time of the call, so that outputs allocated under
        `with torch.cuda.stream(s)` are produced on s, and torch.cuda.Event timing and tensor lifetimes see the work."""
        self._pinned_stream = stream
        self._bind_stream()

    def _bind_stream(self):
        s = self._pinned_stream if self._pinned_stream is not None else self._torch.cuda.current_stream(self.device)
        self._check(self.lib.par_set_stream(self.handle, C.c_void_p(s.cuda_stream)))

    def set_sub_batch(self, frames):
        """Run batches in rounds of `frames` frames through all stages (0 = one launch per stage over the whole batch)."""
        self._check(self.lib.par_set_sub_batch(self.handle, int(frames)))

    def synchronize(self):
        self._check(self.lib.par_synchronize(self.handle))

    STAGES = ("similarity_graph", "resolve_crossings", "cc_labels", "polygons", "raster", "palette")

    def profile(self, on=True):
        """Bracket every stage launch with CUDA events on the launching stream (par_profile_enable)."""
        self._check(self.lib.par_profile_enable(self.handle, int(bool(on))))

    def profile_read(self):
        """{stage: (total_ms, launches)} since the last read; synchronizes the stream."""
        ms = (C.c_double * len(self.STAGES))()
        n = (C.c_int * len(self.STAGES))()
        self._check(self.lib.par_profile_read(self.handle, ms, n))
        return {k: (ms[i], n[i]) for i, k in enumerate(self.STAGES)}

    def smooth_stats(self):
        """{'smoothed', 'geometric'}: cells stage E ran on since the context was created, and how many of them
        built and rasterized their polygon instead of using the precomputed smoothing tables."""
        out = (C.c_uint64 * 2)()
        self._check(self.lib.par_smooth_stats(self.handle, out))
        return {"smoothed": int(out[0]), "geometric": int(out[1])}

    @property
    def launch_count(self):
        return int(self.lib.par_launch_count(self.handle))

    def _check(self, st):
        if st != 0:
            raise RemasterError(st, self.lib.par_last_error(self.handle).decode())

    def _job(self, frames, scale, flags, out_format=OUT_RGBA8, **outs):
        t = self._torch
        self._bind_stream()  # (every launch goes to the stream that is current now, unless one was pinned)
        for v in [frames] + list(outs.values()):
            if v is not None and v.is_cuda and v.device != self.device:
                raise ValueError("tensor on %s passed to a context on %s" % (v.device, self.device))
        if frames is not None:
            assert frames.dtype == t.uint8 and frames.dim() == 4 and frames.shape[3] == 3, "frames: (F, H, W, 3) uint8 BGR"
            assert frames.stride(3) == 1 and frames.stride(2) == 3, "pixels must be packed BGR"
            F, H, W = frames.shape[:3]
            ws, fs = frames.stride(1), frames.stride(0)
        else:
            F, H, W = outs["graph"].shape if outs.get("graph") is not None else outs["graph_aux"].shape
            ws, fs = 3 * W, 0
        j = ParJob()
        j.bgr = _ptr(frames)
        j.width, j.height, j.widthstep, j.frame_stride, j.n_frames = W, H, ws, fs if F > 1 else 0, F
        j.scale, j.flags, j.out_format = int(scale), int(flags), int(out_format)
        for k in ("rgba", "graph", "graph_aux", "labels", "polygons", "poly_count", "palette", "palette_count"):
            v = outs.get(k)
            if v is not None:
                assert v.is_contiguous()
            setattr(j, k, _ptr(v))
        return j

    @staticmethod
    def image_shape(F, H, W, scale, out_format):
        """Shape of the output image tensor `rgba` for a format: (F, sH, sW, 4 | 3) or (F, sH, sW) for INDEX8."""
        return (F, scale * H, scale * W) + ((_BPP[out_format],) if out_format != OUT_INDEX8 else ())

    def _alloc(self, F, H, W, scale, want, out_format=OUT_RGBA8, host=False):
        t = self._torch

        def mk(shape, dtype):
            return t.empty(shape, dtype=dtype).pin_memory() if host else t.empty(shape, dtype=dtype, device=self.device)

        o = {}
        if "rgba" in want:
            o["rgba"] = mk(self.image_shape(F, H, W, scale, out_format), t.uint8)
            if out_format == OUT_INDEX8:
                o["palette"] = mk((F, 256), t.int32)   # RGBA8 words
                o["palette_count"] = mk((F,), t.int32)
        if "graph" in want:
            o["graph"] = mk((F, H, W), t.uint8)
        if "graph_aux" in want:
            o["graph_aux"] = mk((F, H, W), t.uint8)
        if "labels" in want:
            o["labels"] = mk((F, H, W), t.int32)
        if "polygons" in want:
            o["polygons"] = mk((F, H * W, CELL_SLOTS, 2), t.float32)
            o["poly_count"] = mk((F, H * W), t.int32)
        return o

    no_tables = False  # set True to bypass the smoothing tables (geometric path for every smoothed cell) on every call of this context
    aa = 1             # samples per output pixel and axis (1 = point sampling, 2 or 4 = anti-aliased: PAR_FLAG_AA2 / AA4)

    def _flags(self, subdivide, flip_output, no_tma):
        return (FLAG_SUBDIVIDE if subdivide else 0) | (FLAG_FLIP_OUTPUT if flip_output else 0) | (FLAG_NO_TMA if no_tma else 0) | \
            (FLAG_NO_SMOOTH_TABLES if self.no_tables else 0) | {1: 0, 2: FLAG_AA2, 4: FLAG_AA4}[self.aa]

    # -- whole path ------------------------------------------------------------------------
    def remaster(self, frames, scale=4, subdivide=True, want=("rgba",), out=None, flip_output=False, no_tma=False, out_format=OUT_RGBA8):
        """frames: CUDA uint8 (F, H, W, 3) BGR, row 0 = bottom scanline.  Returns a dict of CUDA tensors
        for the names in `want` (rgba graph graph_aux labels polygons[+poly_count]); asynchronous.  `out_format` chooses the
        layout of the image `rgba` (OUT_RGBA8, OUT_BGR8, OUT_INDEX8: + `palette` (F, 256) RGBA words, `palette_count` (F,))."""
        F, H, W = frames.shape[:3]
        o = out if out is not None else self._alloc(F, H, W, scale, want, out_format)
        j = self._job(frames, scale, self._flags(subdivide, flip_output, no_tma), out_format, **o)
        self._check(self.lib.par_remaster_device(self.handle, C.byref(j)))
        return o

    def remaster_host(self, frames, scale=4, subdivide=True, want=("rgba",), out=None, flip_output=False, out_format=OUT_RGBA8):
        """Same through HOST buffers (torch CPU tensors, ideally pinned): H2D, kernels, D2H, synchronize."""
        F, H, W = frames.shape[:3]
        assert not frames.is_cuda
        if out is None:
            out = self._alloc(F, H, W, scale, want, out_format, host=True)
        j = self._job(frames, scale, self._flags(subdivide, flip_output, False), out_format, **out)
        self._check(self.lib.par_remaster_host(self.handle, C.byref(j)))
        return out

    @staticmethod
    def expand_indexed(index, palette):
        """(F, sH, sW) palette indices + (F, 256) RGBA words -> (F, sH, sW, 4) uint8 RGBA (numpy, host side; tests and tools)."""
        idx = index.cpu().numpy() if hasattr(index, "cpu") else np.asarray(index)
        pal = (palette.cpu().numpy() if hasattr(palette, "cpu") else np.asarray(palette)).view(np.uint32).reshape(idx.shape[0], 256)
        rgba = np.take_along_axis(pal, idx.reshape(idx.shape[0], -1).astype(np.int64), axis=1)
        return rgba.view(np.uint8).reshape(idx.shape + (4,))

    # -- single stages (parity tests) ------------------------------------------------------
    def similarity_graph(self, frames, no_tma=False):
        F, H, W = frames.shape[:3]
        o = self._alloc(F, H, W, 1, ("graph_aux",))
        j = self._job(frames, 1, self._flags(False, False, no_tma), **o)
        self._check(self.lib.par_stage_similarity_graph(self.handle, C.byref(j)))
        return o["graph_aux"]

    def resolve_crossings(self, graph_aux, no_tma=False):
        F, H, W = graph_aux.shape
        g = self._torch.empty_like(graph_aux)
        j = self._job(None, 1, self._flags(False, False, no_tma), graph_aux=graph_aux, graph=g)
        self._check(self.lib.par_stage_resolve_crossings(self.handle, C.byref(j)))
        return g

    def cc_labels(self, graph):
        lab = self._torch.empty(graph.shape, dtype=self._torch.int32, device=graph.device)
        j = self._job(None, 1, 0, graph=graph, labels=lab)
        self._check(self.lib.par_stage_cc_labels(self.handle, C.byref(j)))
        return lab

    def border_walks(self, graph, labels, capacity_per_frame=None):
        """Border walk of every component (par_border_walks): returns device tensors (walk_len, walk_begin, walk_nodes,
        total).  walk_len / walk_begin are per pixel; walk_nodes is (F, capacity); total is per frame."""
        t = self._torch
        F, H, W = graph.shape
        cap = int(capacity_per_frame) if capacity_per_frame else 4 * H * W
        wl = t.empty((F, H, W), dtype=t.int32, device=graph.device)
        wb = t.empty((F, H, W), dtype=t.int32, device=graph.device)
        nodes = t.empty((F, cap), dtype=t.int32, device=graph.device)
        total = t.empty((F,), dtype=t.int64, device=graph.device)
        self._check(self.lib.par_border_walks(self.handle, graph.data_ptr(), labels.data_ptr(), W, H, F, wl.data_ptr(), wb.data_ptr(),
                                              nodes.data_ptr(), cap, total.data_ptr()))
        return wl, wb, nodes, total

    def walk_splines(self, walk_len, walk_begin, walk_nodes, total, samples=4):
        """Closed uniform quadratic B-spline of every border walk (par_walk_splines): (F, capacity * samples, 2) float32
        device tensor of curve points in source-pixel coordinates; sample (b + i) * samples + s belongs to node i of the
        walk that begins at entry b."""
        t = self._torch
        F, H, W = walk_len.shape
        cap = walk_nodes.shape[1]
        self._bind_stream()
        pts = t.zeros((F, cap * samples, 2), dtype=t.float32, device=walk_len.device)
        self._check(self.lib.par_walk_splines(self.handle, walk_len.data_ptr(), walk_begin.data_ptr(), walk_nodes.data_ptr(), total.data_ptr(),
                                              W, H, F, cap, int(samples), pts.data_ptr()))
        return pts

    def outlines_host(self, image, samples=4):
        """par_outlines_host: numpy uint8 (H, W, 3) BGR frame (row 0 = bottom) -> [(start pixel, (count * samples, 2) float32 curve points)]
        for every component's border walk, in raster order of the start."""
        assert image.dtype == np.uint8 and image.ndim == 3 and image.shape[2] == 3 and image.strides[1:] == (3, 1)
        self._bind_stream()
        o = ParOutlines()
        self._check(self.lib.par_outlines_host(self.handle, image.ctypes.data, image.shape[1], image.shape[0], image.strides[0], int(samples), C.byref(o)))
        try:
            pts = np.ctypeslib.as_array(o.points, shape=(int(o.n_points), 2)).copy() if o.n_points else np.zeros((0, 2), np.float32)
            res, at = [], 0
            for k in range(o.n_walks):
                m = o.count[k] * o.samples
                res.append((int(o.start[k]), pts[at:at + m]))
                at += m
            return res
        finally:
            self.lib.par_outlines_free(C.byref(o))

    @staticmethod
    def walks_as_dict(walk_len, walk_begin, walk_nodes, frame=0):
        """{start node: [nodes]} of one frame, on the host (tests, small inputs)."""
        wl = walk_len[frame].reshape(-1).cpu().numpy()
        wb = walk_begin[frame].reshape(-1).cpu().numpy()
        nodes = walk_nodes[frame].cpu().numpy()
        return {int(n): nodes[wb[n]:wb[n] + wl[n]].tolist() for n in wl.nonzero()[0]}

    def polygons(self, frames, graph, subdivide=True):
        F, H, W = frames.shape[:3]
        o = self._alloc(F, H, W, 1, ("polygons",))
        j = self._job(frames, 1, self._flags(subdivide, False, False), graph=graph, **o)
        self._check(self.lib.par_stage_polygons(self.handle, C.byref(j)))
        return o["polygons"], o["poly_count"]

    def raster(self, frames, graph, scale=4, subdivide=True, flip_output=False, no_tma=False, debug_wide=False, out_format=OUT_RGBA8):
        """Stage D+E+raster alone.  Returns the image; for OUT_INDEX8 the dict {rgba, palette, palette_count}."""
        F, H, W = frames.shape[:3]
        o = self._alloc(F, H, W, scale, ("rgba",), out_format)
        flags = self._flags(subdivide, flip_output, no_tma) | (FLAG_DEBUG_WIDE if debug_wide else 0)
        j = self._job(frames, scale, flags, out_format, graph=graph, **o)
        self._check(self.lib.par_stage_raster(self.handle, C.byref(j)))
        return o["rgba"] if out_format != OUT_INDEX8 else o


class RemasterGroup:
    """One large image cut into horizontal strips over several GPUs (par_group): host image in, host
    results out; apron rows travel between neighbouring devices by peer copies.  `devices` may repeat a
    device index (several strips on one GPU), which is how the single-GPU tests exercise the stitching."""

    def __init__(self, devices, width, height, scale=4):
        self.lib = load_library()
        self.handle = C.c_void_p()
        self.width, self.height, self.scale = int(width), int(height), int(scale)
        arr = (C.c_int * len(devices))(*[int(d) for d in devices])
        st = self.lib.par_group_create(C.byref(self.handle), arr, len(devices), self.width, self.height, self.scale)
        if st != 0:
            self.handle = C.c_void_p()
            raise RemasterError(st, self.lib.par_group_last_error(None).decode())

    def close(self):
        h = getattr(self, "handle", None)
        if h is not None and h.value:
            self.lib.par_group_destroy(h)
            h.value = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def strips(self):
        """The strips as dicts: device, own / load row ranges, and torch tensors ALIASING the strip's device buffers
        (rows load_begin .. load_end): bgr (rows, W, 3), graph / graph_aux (rows, W), labels (rows, W) int32, and
        image(out_format) -> (rows * S, W * S, 4 | 3)."""
        import torch
        out = []
        W, S = self.width, self.scale
        for k in range(self.lib.par_group_n_strips(self.handle)):
            st = ParStrip()
            assert self.lib.par_group_strip(self.handle, k, C.byref(st)) == 0
            rows = st.load_end - st.load_begin
            dev = "cuda:%d" % st.device

            def view(ptr, shape, typestr, dev=dev):
                with torch.cuda.device(dev):
                    return torch.as_tensor(_DeviceBuffer(ptr, shape, typestr), device=dev)

            out.append({"device": st.device, "own": (st.own_begin, st.own_end), "load": (st.load_begin, st.load_end),
                        "bgr": view(st.bgr, (rows, W, 3), "|u1"), "graph": view(st.graph, (rows, W), "|u1"),
                        "graph_aux": view(st.graph_aux, (rows, W), "|u1"), "labels": view(st.labels, (rows, W), "<i4"),
                        "image": (lambda fmt, ptr=st.image, rows=rows, view=view: view(ptr, (rows * S, W * S, _BPP[fmt]), "|u1"))})
        return out

    def remaster_device(self, subdivide=True, out_format=OUT_RGBA8, want_image=True, want_labels=True, flip_output=False):
        """The device-resident tiled path (par_group_remaster_device): every strip's own rows are already in its `bgr`
        buffer (see strips()); results stay on the devices.  Returns (wall_ms, longest strip's device ms)."""
        flags = (FLAG_SUBDIVIDE if subdivide else 0) | (FLAG_FLIP_OUTPUT if flip_output else 0)
        st = self.lib.par_group_remaster_device(self.handle, flags, int(out_format), int(bool(want_image)), int(bool(want_labels)))
        if st != 0:
            raise RemasterError(st, self.lib.par_group_last_error(self.handle).decode())
        wall, dev = C.c_double(), C.c_double()
        self.lib.par_group_last_ms(self.handle, C.byref(wall), C.byref(dev))
        return wall.value, dev.value

    def remaster_host(self, image, subdivide=True, want=("rgba", "graph"), flip_output=False, out=None, out_format=OUT_RGBA8):
        """image: numpy uint8 (H, W, 3) BGR (row stride may exceed 3*W).  Returns numpy arrays; `out` may supply them
        (e.g. views of pinned memory: pageable buffers limit the copies to a few GB/s)."""
        assert image.dtype == np.uint8 and image.shape[:2] == (self.height, self.width) and image.strides[1:] == (3, 1)
        H, W, S = self.height, self.width, self.scale
        shapes = {"rgba": ((S * H, S * W, _BPP[out_format]), np.uint8), "graph": ((H, W), np.uint8), "graph_aux": ((H, W), np.uint8), "labels": ((H, W), np.int32)}
        if out is None:
            out = {k: np.empty(*shapes[k]) for k in ("rgba", "graph", "graph_aux", "labels") if k in want}
        for k, v in out.items():
            assert v.shape == shapes[k][0] and v.dtype == shapes[k][1] and v.flags["C_CONTIGUOUS"], k
        j = ParJob()
        j.bgr = image.ctypes.data
        j.width, j.height, j.widthstep, j.frame_stride, j.n_frames = W, H, image.strides[0], 0, 1
        j.scale = S
        j.out_format = int(out_format)
        j.flags = (FLAG_SUBDIVIDE if subdivide else 0) | (FLAG_FLIP_OUTPUT if flip_output else 0)
        for k, v in out.items():
            setattr(j, k, v.ctypes.data)
        st = self.lib.par_group_remaster_host(self.handle, C.byref(j))
        if st != 0:
            raise RemasterError(st, self.lib.par_group_last_error(self.handle).decode())
        return out


def launch_kernel(img, subdivide=True, want_vbo=False):
    """The reference's entry point (kernel.cu:286-288) through its C symbol, on the current CUDA device.
    img: numpy uint8 (H, W, 3) BGR, row 0 = bottom.  Returns (graph (H,W) uint8, edge_count (H*W,) int32,
    diagram (H*W, 45, 2) float32 triangle list[, pos (H*W,45,2) float32, color (H*W,45,4) uint8])."""
    import torch
    L = load_library()
    H, W = img.shape[:2]
    N = H * W
    ws = img.strides[0]
    flat = np.ascontiguousarray(np.lib.stride_tricks.as_strided(img, shape=(H * ws,), strides=(1,))) if ws != 3 * W else np.ascontiguousarray(img).reshape(-1)
    pos = torch.empty((N, CELL_SLOTS, 2), dtype=torch.float32, device="cuda")
    col = torch.empty((N, CELL_SLOTS, 4), dtype=torch.uint8, device="cuda")
    torch.cuda.synchronize()
    graph = np.zeros((H, W), np.uint8)
    count = np.zeros(N, np.int32)
    ptr = L.launch_kernel(pos.data_ptr(), col.data_ptr(), 0.0, flat.ctypes.data, W, H, ws, count.ctypes.data, graph.ctypes.data, bool(subdivide))
    diagram = np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_float)), shape=(N, CELL_SLOTS, 2)).copy()
    C.CDLL(None).free(C.c_void_p(ptr))  # the caller frees (simpleVBO.cpp:437)
    if want_vbo:
        return graph, count, diagram, pos.cpu().numpy(), col.cpu().numpy()
    return graph, count, diagram
