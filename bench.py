#!/usr/bin/env python
"""Headline benchmark: remastered frames/s at 256x224 -> 4x (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One "step" = one pass of the whole hot path (similarity graph -> crossings -> cells + subdivision ->
direct raster) over this rank's batch of synthetic frames: config C3, a stream of 4096 SNES-style
256x224 16-colour frames per GPU, scale 4, subdivision on, CC labels off (the C1 byte accounting).
Frames are independent, so ranks shard the stream with no data-path collective ("weak" scaling:
every GPU gets its own 4096 frames); torch.distributed is only used for the barrier and the
max-over-ranks reduction of the device time.

Prints ONE JSON line.  `value` is whole-job frames/s with inputs resident in HBM; `e2e` is the same
metric through the host-buffer entry point of the C ABI (pinned host memory, H2D + D2H inside the
timed region, the same 4096-frame step, output as palette indices + palette; RGBA8 beside it);
`roofline` is for the dominant kernel (the rasterizer), timed with CUDA events on the launching
stream inside the same timed region; `cpu_baseline` is the reference's own routines (oracle/_ref,
host build) on this box's cores over a bounded sample.  Also in the line: `sustained` (the same step
back to back for >= 2 s with its clock samples), `strong_scaling` (the one 4096-frame stream split
over the ranks), `tiled` (config C4: one 4096 x 4096 map over all GPUs of the job, device-resident,
checked against one context), `parity` (what each result is pinned by).

--impl reference: the reference's CPU implementation of the path (oracle/_ref/libref_host_fma.so,
the reference's own .cu files compiled as host C++; falls back to the oracle port when that build is
absent) + the oracle's triangle rasterizer standing in for the OpenGL draw, on all host cores.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

W, H, SCALE = 256, 224, 4
FRAMES_PER_GPU = 4096
E2E_FRAMES = 512
ALGO_BYTES_PER_PX = 3 + 1 + 4 * SCALE * SCALE   # BGR in + graph out + RGBA out
RASTER_BYTES_PER_PX = 3 + 1 + 4 * SCALE * SCALE  # raster kernel: colour + final graph in, RGBA out
METRIC = "remastered frames/s at 256x224->4x"


def workload(n_frames=FRAMES_PER_GPU):
    return ("C3: stream of %d synthetic SNES-style 256x224 16-colour BGR8 frames per GPU -> 4x RGBA, "
            "subdivision on, CC labels off" % n_frames)


def profiled_traffic_per_frame():
    """DRAM bytes (read + write) per frame of the raster kernel from the committed ncu capture of a launch over the bench's own
    4096-frame batch (profiles/r4_raster_traffic.json or an earlier round's, written by tools/ncu_traffic.py; round 1's 256-frame capture as a
    fallback); None when neither is there."""
    for name in ("r4_raster_traffic.json", "r3_raster_traffic.json", "r2_raster_traffic.json", "r1_raster_traffic.json"):
        try:
            t = json.load(open(os.path.join(ROOT, "profiles", name)))
            return float(t["dram_bytes_per_frame"]), t.get("source", "profiles/" + name)
        except Exception:
            continue
    return None, None


def measured_peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap,timestamp")
    # The sampler is started BEFORE the warm-up steps: nvidia-smi needs 50-150 ms before its first sample, as long as the
    # whole timed region of a 10-step run.  Samples carry a timestamp; those inside [mark_begin, mark_end] (host clock
    # around the timed region, 30 ms of slack) are the ones reported, the warm-up's (same load) only if none fell inside.

    def __init__(self, device_index):
        self.idx = device_index
        self.proc = None
        self.t0 = self.t1 = None

    def mark_begin(self):
        self.t0 = time.time()

    def mark_end(self):
        self.t1 = time.time()

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "20"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except Exception:
            self.proc.kill()
            out = ""
        import datetime
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        rows = []
        for line in out.splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm, mx = float(f[1]), float(f[2])
            except ValueError:
                continue
            when = None
            if len(f) > 9:
                try:
                    when = datetime.datetime.strptime(f[9], "%Y/%m/%d %H:%M:%S.%f").timestamp()
                except ValueError:
                    pass
            rows.append((when, sm, mx, {name for k, name in enumerate(names) if f[5 + k].lower().startswith("active")}))
        inside = [r for r in rows if r[0] is not None and self.t0 is not None and self.t1 is not None and self.t0 - 0.03 <= r[0] <= self.t1 + 0.03]
        used = inside or rows
        reasons = set()
        for r in used:
            reasons |= r[3]
        return {"sm_mhz": statistics.median([r[1] for r in used]) if used else None, "sm_max_mhz": max(r[2] for r in used) if used else None,
                "reasons": sorted(reasons), "samples": len(used), "samples_in_timed_region": len(inside)}


def host_threads():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def cpu_reference_fps(frames, threads, seconds_budget=20.0):
    """frames/s of the reference's CPU routines (+ oracle raster) over `frames`, `threads` workers."""
    from concurrent.futures import ThreadPoolExecutor
    from oracle.oracle import Oracle, RefHost
    orc = Oracle()
    kind = "port"
    ref = None
    if RefHost.available(fma=True):
        ref = RefHost(fma=True)
        kind = "reference"

    def one(img):
        if ref is not None:
            r = ref.pipeline(img, True, ("graph", "poly_count", "tri"))
            ntri = np.maximum(r["poly_count"] - 2, 0).astype(np.int32)
            orc.raster_triangles(img, SCALE, r["tri"], ntri)
        else:
            orc.pipeline(img, True, True, SCALE, ("graph", "raster"))
        return 1

    one(frames[0])  # warm the page cache / allocator
    t0 = time.perf_counter()
    done = 0
    with ThreadPoolExecutor(threads) as ex:
        # ctypes releases the GIL during the C call, so the workers run on separate cores
        for _ in ex.map(one, frames):
            done += 1
    dt = time.perf_counter() - t0
    return done / dt, kind, done, dt


def time_drop_in(par, img, n_calls):
    """ms per call of the product's `launch_kernel` export (kernel.cu:286-288 signature) on one frame, end to end."""
    import ctypes as C
    import torch
    L = par.load_library()
    Hh, Ww = img.shape[:2]
    N = Hh * Ww
    flat = np.ascontiguousarray(img).reshape(-1)
    pos = torch.empty((N, par.CELL_SLOTS, 2), dtype=torch.float32, device="cuda")
    col = torch.empty((N, par.CELL_SLOTS, 4), dtype=torch.uint8, device="cuda")
    graph = np.zeros((Hh, Ww), np.uint8)
    count = np.zeros(N, np.int32)
    free = C.CDLL(None).free
    free.argtypes = [C.c_void_p]

    def call():
        ptr = L.launch_kernel(pos.data_ptr(), col.data_ptr(), 0.0, flat.ctypes.data, Ww, Hh, 3 * Ww, count.ctypes.data, graph.ctypes.data, True)
        free(C.c_void_p(ptr))

    for _ in range(3):
        call()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n_calls):
        call()
    torch.cuda.synchronize()
    ms = 1e3 * (time.perf_counter() - t0) / n_calls
    return {"value": 1e3 / ms, "unit": "frames/s", "ms_per_call": ms,
            "what": "this library's launch_kernel (the reference's symbol and signature) per frame, end to end: H2D, graph, crossings, "
                    "cells + subdivision + ear clipping into the 45-slot VBO arrays, D2H of the triangle list, graph and counts"}


def bench_config(n_frames=FRAMES_PER_GPU):
    """The `config` object of the JSON line — the same in both arms (ours and --impl reference)."""
    return {"workload": workload(n_frames), "width": W, "height": H, "scale": SCALE, "subdivide": True, "labels": False,
            "frames_per_gpu": n_frames, "algorithmic_bytes_per_frame": ALGO_BYTES_PER_PX * W * H,
            "l2": "inputs (%.0f MB) and outputs (%.1f GB) per step exceed the 126 MB L2; no flush needed"
                  % (n_frames * W * H * 3 / 1e6, n_frames * W * H * SCALE * SCALE * 4 / 1e9)}


PARITY_NOTES = {
    "graph": "similarity graph and crossing decisions bit-exact vs the oracle, which equals the reference's own .cu files compiled here "
             "(host build) and its kernel.cu built for sm_100a on the GPU (tests/test_oracle_vs_reference.py, test_gpu_reference_cuda.py)",
    "polygons": "cell and subdivision vertices equal (exact dyadic arithmetic) vs the same",
    "cc_labels": "reference implementation is dead code (cc_functions.cu is not in its build): parity vs the canonical definition "
                 "(minimum row-major index per component of the final graph) + 1 reference fixture (alex_png.txt, 19 components)",
    "raster": "the reference rasterizes inside the OpenGL driver, so the sampling rule is unpinned by it: parity vs the oracle's restated "
              "GL point-sampling rule painting the REFERENCE'S OWN triangle list (bit-exact, scales 1..8); BGR8 / INDEX8 are the RGBA8 image "
              "in another layout (byte-exact)",
}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    from pixel_art_remaster_gpu_b200 import synth
    from oracle import oracle as orc_mod
    orc_mod.build(ref=os.path.isdir("/root/reference"))
    threads = host_threads()
    uniq = synth.snes_stream(64, W, H)
    probe_fps = cpu_reference_fps(uniq[:max(threads, 8)], threads)[0]
    budget_s = min(20.0, 150.0 / max(args.steps, 1))  # the whole run stays within a few minutes
    per_step = int(min(max(probe_fps * budget_s, threads), 4096))  # bounded sample of the stream per step
    frames = [uniq[k % 64] for k in range(per_step)]
    for _ in range(args.warmup if args.warmup < 2 else 1):
        cpu_reference_fps(frames[:threads], threads)
    t_total, n_total, kind = 0.0, 0, "port"
    for _ in range(args.steps):
        fps, kind, n, dt = cpu_reference_fps(frames, threads)
        t_total += dt
        n_total += n
    value = n_total / t_total
    sample = ("%d frames (64 distinct, cycled) of the C3 stream per step on %d host threads; reference routines A-F "
              "(graph, crossings, cells, subdivision, ear clipping) as host C++%s + oracle triangle raster at 4x"
              % (per_step, threads, "" if kind == "reference" else " [oracle port: oracle/_ref absent]"))
    line = {"metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * t_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u8", "data": "synthetic", "impl": "reference",
            "config": bench_config(args.frames),
            "cpu_baseline": {"value": value, "unit": "frames/s", "cores": threads, "kind": kind, "sample": sample},
            "e2e": {"value": value, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))
    return 0


def bind_to_gpu_numa_node(local):
    """Run this rank (and first-touch its pinned buffers) on the cores next to its GPU: /sys/bus/pci/devices/<bus id>/local_cpulist.
    Returns what was done, for the JSON line."""
    try:
        import torch
        bus = torch.cuda.get_device_properties(local).pci_bus_id
        dom = torch.cuda.get_device_properties(local).pci_domain_id
        dev = torch.cuda.get_device_properties(local).pci_device_id
        path = "/sys/bus/pci/devices/%04x:%02x:%02x.0/" % (dom, bus, dev)
        cpus = open(path + "local_cpulist").read().strip()
        node = open(path + "numa_node").read().strip()
        ids = set()
        for part in cpus.split(","):
            a, _, b = part.partition("-")
            ids.update(range(int(a), int(b or a) + 1))
        allowed = ids & set(os.sched_getaffinity(0))
        if allowed:
            os.sched_setaffinity(0, allowed)
        return {"numa_node": node, "cpus": cpus, "bound": bool(allowed)}
    except Exception as exc:
        return {"bound": False, "why": str(exc)[:120]}


def tiled_map_check(par, synth, torch, n_devices, side):
    """Config C4 under the driver's eyes: one side x side map cut into strips over `n_devices` GPUs (two strips on one GPU at
    N = 1), device-resident (par_group_remaster_device: halo rows by peer copies, kernels, on-device label stitch), compared
    with the same map through ONE context on GPU 0, and timed."""
    S = SCALE
    img = synth.pixel_art_map(side, side, synth.BASE_SEED + 4)
    devices = list(range(n_devices)) if n_devices > 1 else [0, 0]
    res = {"map": "%dx%d synthetic pixel-art map -> %dx RGBA8 + CC labels" % (side, side, S), "n_devices": n_devices, "strips": len(devices)}
    with par.Remaster(0, side, side, 1) as one:
        dev0 = torch.device("cuda", 0)
        whole_in = torch.from_numpy(img[None]).to(dev0)
        want = one.remaster(whole_in, S, True, want=("rgba", "graph", "labels"))
        torch.cuda.synchronize(dev0)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(3):
            one.remaster(whole_in, S, True, out=want)
        b.record()
        torch.cuda.synchronize(dev0)
        res["one_gpu_ms"] = a.elapsed_time(b) / 3
        with par.RemasterGroup(devices, side, side, S) as grp:
            strips = grp.strips()
            for st in strips:
                lo, hi = st["own"]
                lb = st["load"][0]
                st["bgr"].zero_()
                st["bgr"][lo - lb:hi - lb].copy_(torch.from_numpy(img[lo:hi]))
            for d in set(devices):
                torch.cuda.synchronize(d)
            grp.remaster_device(True)  # warm-up (tables, peer mappings)
            walls, devs = [], []
            for _ in range(5):
                w, d = grp.remaster_device(True)
                walls.append(w)
                devs.append(d)
            ok = True
            for st in strips:
                lo, hi = st["own"]
                lb = st["load"][0]
                ok = ok and bool(torch.equal(st["graph"][lo - lb:hi - lb].to(dev0), want["graph"][0, lo:hi]))
                ok = ok and bool(torch.equal(st["labels"][lo - lb:hi - lb].to(dev0), want["labels"][0, lo:hi]))
                ok = ok and bool(torch.equal(st["image"](par.OUT_RGBA8)[(lo - lb) * S:(hi - lb) * S].to(dev0), want["rgba"][0, lo * S:hi * S]))
            res.update({"ok": ok, "ms": min(walls), "strip_device_ms": min(devs),
                        "what": "ms = host wall clock of par_group_remaster_device (enqueue on all devices .. last synchronize), best of 5; "
                                "strip_device_ms = the slowest strip's own stream time (CUDA events); one_gpu_ms = the whole map through one "
                                "context on GPU 0; ok = own rows of every strip (graph, labels, RGBA) equal the one-context result"})
    return res


def run_ours(args):
    import torch
    import torch.distributed as dist
    import pixel_art_remaster_gpu_b200 as par
    from pixel_art_remaster_gpu_b200 import sharding, synth

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product has no CPU path (use --impl reference for the CPU baseline)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    numa = bind_to_gpu_numa_node(local)
    host_group = None
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        host_group = dist.new_group(backend="gloo")  # host-side barriers: waiting ranks must not keep a kernel spinning on their GPU

    def host_barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier(group=host_group)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        t = torch.tensor([float(x)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    n_frames = args.frames
    # this rank's shard of the stream: its own consecutive seeds (no frame is shared between ranks)
    host_frames = torch.from_numpy(synth.snes_stream(n_frames, W, H, first_seed=sharding.stream_seed(rank, n_frames, synth.BASE_SEED)))
    frames = host_frames.to(dev)
    ctx = par.Remaster(local, W, H, n_frames)
    out = {"rgba": torch.empty((n_frames, SCALE * H, SCALE * W, 4), dtype=torch.uint8, device=dev),
           "graph": torch.empty((n_frames, H, W), dtype=torch.uint8, device=dev)}

    def step():
        ctx.remaster(frames, SCALE, True, out=out)

    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    ctx.profile(True)
    ctx.profile_read()
    smooth0 = ctx.smooth_stats()
    launches0 = ctx.launch_count
    barrier()
    clocks.mark_begin()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(args.steps):
        step()
    ev1.record()
    barrier()
    clocks.mark_end()
    clock_info = clocks.stop() if rank == 0 else None
    ms = ev0.elapsed_time(ev1)
    prof = ctx.profile_read()
    ctx.profile(False)
    smooth1 = ctx.smooth_stats()
    launches = ctx.launch_count - launches0
    ms_max = max_over_ranks(ms)
    value = world * n_frames * args.steps / (ms_max * 1e-3)

    # ---- sustained: the same step back to back for >= 2 s, with its own clock samples (the headline's timed region is ~0.1 s)
    sus_clocks = ClockSampler(local)
    if rank == 0:
        sus_clocks.start()
    barrier()
    sus_steps = max(args.steps, int(2200.0 / max(ms_max / args.steps, 1e-3)) + 1)
    sus_clocks.mark_begin()
    s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s0.record()
    for _ in range(sus_steps):
        step()
    s1.record()
    barrier()
    sus_clocks.mark_end()
    sus_info = sus_clocks.stop() if rank == 0 else None
    sus_ms = max_over_ranks(s0.elapsed_time(s1))
    sustained = {"value": world * n_frames * sus_steps / (sus_ms * 1e-3), "unit": "frames/s", "seconds": sus_ms * 1e-3, "steps": sus_steps,
                 "ms_per_step": sus_ms / sus_steps, "clocks": sus_info}

    # ---- strong scaling: the SAME 4096-frame stream cut into contiguous shards (sharding.frame_shard), one per rank
    lo, hi = sharding.frame_shard(n_frames, rank, world)
    shard_out = {"rgba": out["rgba"][:hi - lo], "graph": out["graph"][:hi - lo]}
    shard_in = frames[:hi - lo]  # (synthetic frames of the same generator: which ones a rank holds does not change the work)

    def strong_step():
        ctx.remaster(shard_in, SCALE, True, out=shard_out)

    for _ in range(3):
        strong_step()
    barrier()
    g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    g0.record()
    for _ in range(args.steps):
        strong_step()
    g1.record()
    barrier()
    strong_ms = max_over_ranks(g0.elapsed_time(g1))
    strong = {"value": n_frames * args.steps / (strong_ms * 1e-3), "unit": "frames/s", "frames_total": n_frames, "frames_per_gpu": hi - lo,
              "ms_per_step": strong_ms / args.steps, "scaling": "strong",
              "what": "the 4096-frame stream of config C3 split into contiguous per-GPU shards (sharding.frame_shard), no collective"}
    step()  # (restore the full batch in `out` for the checks below)

    # ---- end to end through the host-buffer entry point of the C ABI (pinned memory, H2D + D2H in the timed region), at the
    # same 4096-frame step.  Headline format: INDEX8 (palette indices + per-frame palette: the same image, losslessly, in a
    # quarter of the bytes — every output pixel is a source colour or black); RGBA8 beside it on a 512-frame step.
    e2e_steps = max(1, min(args.steps, 3))
    pin_in = host_frames.clone().pin_memory()
    pin_idx = ctx._alloc(n_frames, H, W, SCALE, ("rgba", "graph"), par.OUT_INDEX8, host=True)
    ctx.remaster_host(pin_in, SCALE, True, out=pin_idx, out_format=par.OUT_INDEX8)
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        ctx.remaster_host(pin_in, SCALE, True, out=pin_idx, out_format=par.OUT_INDEX8)  # H2D, kernels, D2H; synchronizes
    barrier()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    e2e_value = world * n_frames * e2e_steps / e2e_s
    d2h_idx = sum(v.numel() * v.element_size() for v in pin_idx.values())
    # every frame of the end-to-end result against the device-resident RGBA8 result
    same = bool((pin_idx["palette_count"] <= 256).all()) and bool(torch.equal(pin_idx["graph"], out["graph"].cpu()))
    for c0 in range(0, n_frames, 128):
        c1 = min(c0 + 128, n_frames)
        idx = pin_idx["rgba"][c0:c1].to(dev).view(c1 - c0, -1).long()
        pal = pin_idx["palette"][c0:c1].to(dev)
        same = same and bool(torch.equal(torch.gather(pal, 1, idx), out["rgba"][c0:c1].view(torch.int32).view(c1 - c0, -1)))
        del idx, pal
    # pinned-copy ceiling of this rank with every rank copying at once (the host side is shared)
    peak_buf_d = out["rgba"].view(-1)[: 1 << 30]
    peak_buf_h = torch.empty(1 << 30, dtype=torch.uint8).pin_memory()
    peak_buf_h.copy_(peak_buf_d, non_blocking=True)
    barrier()
    p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    p0.record()
    for _ in range(3):
        peak_buf_h.copy_(peak_buf_d, non_blocking=True)
    p1.record()
    barrier()
    d2h_peak = 3 * (1 << 30) / (max_over_ranks(p0.elapsed_time(p1)) * 1e-3) / 1e9  # GB/s per rank, all ranks busy
    del peak_buf_h
    # RGBA8 beside it (512 frames: 1.9 GB of pinned memory per rank instead of 15 GB)
    n_small = min(E2E_FRAMES, n_frames)
    pin_rgba = ctx._alloc(n_small, H, W, SCALE, ("rgba", "graph"), par.OUT_RGBA8, host=True)
    ctx.remaster_host(pin_in[:n_small], SCALE, True, out=pin_rgba)
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        ctx.remaster_host(pin_in[:n_small], SCALE, True, out=pin_rgba)
    barrier()
    rgba_s = max_over_ranks(time.perf_counter() - t0)
    e2e_rgba = {"value": world * n_small * e2e_steps / rgba_s, "unit": "frames/s", "frames_per_step": n_small,
                "d2h_bytes_per_step": int(sum(v.numel() * v.element_size() for v in pin_rgba.values())),
                "matches_device_path": bool(torch.equal(pin_rgba["rgba"], out["rgba"][:n_small].cpu()))}
    del pin_rgba

    def timed(fn, reps):
        fn()  # (warm-up: the first launch of a kernel variant loads its module)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        a.record()
        for _ in range(reps):
            fn()
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / reps

    # side measurements on rank 0 (not the headline): other frames, smoothing tables off, subdivision off, labels on, other formats
    extras = {}
    if rank == 0:
        n_new = min(512, n_frames)
        novel = torch.from_numpy(synth.snes_stream(n_new, W, H, first_seed=synth.BASE_SEED + 10_000_019)).to(dev)
        o_new = {"rgba": out["rgba"][:n_new], "graph": out["graph"][:n_new]}
        ms_new = timed(lambda: ctx.remaster(novel, SCALE, True, out=o_new), 3)
        extras["other_frames"] = {"value": n_new / (ms_new * 1e-3), "unit": "frames/s", "frames": n_new,
                                  "what": "frames from other seeds than the stream (the tables are content-independent; "
                                          "a small batch, so launch overhead weighs more)"}
        sub = frames[:n_new]
        ctx.no_tables = True
        extras["tables_off"] = {"value": n_new / (timed(lambda: ctx.remaster(sub, SCALE, True, out=o_new), 3) * 1e-3), "unit": "frames/s",
                                "what": "PAR_FLAG_NO_SMOOTH_TABLES: every smoothed cell builds and rasterizes its polygon"}
        ctx.no_tables = False
        extras["subdivide_off"] = {"value": n_new / (timed(lambda: ctx.remaster(sub, SCALE, False, out=o_new), 3) * 1e-3), "unit": "frames/s",
                                   "what": "the reference's default (simpleVBO.cpp:43 subdivide = false): hull cells only"}
        lab = torch.empty((n_new, H, W), dtype=torch.int32, device=dev)
        o_lab = dict(o_new, labels=lab)
        extras["labels_on"] = {"value": n_new / (timed(lambda: ctx.remaster(sub, SCALE, True, out=o_lab), 3) * 1e-3), "unit": "frames/s",
                               "what": "the same path + CC labels (union-find: tile, seam and flatten kernels)"}
        del lab, o_lab
        for fmt, name in ((par.OUT_BGR8, "bgr8"), (par.OUT_INDEX8, "index8")):
            o_fmt = ctx._alloc(n_new, H, W, SCALE, ("rgba", "graph"), fmt)
            extras["device_resident_" + name] = {"value": n_new / (timed(lambda: ctx.remaster(sub, SCALE, True, out=o_fmt, out_format=fmt), 3) * 1e-3),
                                                 "unit": "frames/s", "what": "output left in HBM as " + name.upper()}
            del o_fmt
    # ---- config C4 (one large map over the GPUs of the job), rank 0 drives all devices while the others wait
    # (host-side barriers around it: an NCCL barrier would leave a kernel of the waiting ranks spinning on the very GPUs rank 0
    # is about to use, and two processes' kernels on one GPU are time-sliced)
    tiled = None
    barrier()
    host_barrier()
    if rank == 0 and not args.no_tiled:
        try:
            del novel
        except NameError:
            pass
        torch.cuda.empty_cache()
        try:
            tiled = tiled_map_check(par, synth, torch, world, args.tiled_side)
        except Exception as exc:  # a side check must never break the bench line
            tiled = {"ok": False, "error": str(exc)[:300]}
        torch.cuda.set_device(local)
    host_barrier()
    barrier()

    if rank == 0:
        peak, peak_src = measured_peak_gbs()
        traffic_pf, traffic_src = profiled_traffic_per_frame()
        px = n_frames * W * H
        stage_ms = {k: v[0] / args.steps for k, v in prof.items() if v[1]}  # per step (a stage may take several launches)
        r_ms = stage_ms["raster"]
        raster_gbs = (RASTER_BYTES_PER_PX * px) / (r_ms * 1e-3) / 1e9
        frame_bytes_idx = d2h_idx / n_frames
        line = {
            "metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": bench_config(n_frames),
            "hbm_gbs_algorithmic": value / world * ALGO_BYTES_PER_PX * W * H / 1e9,
            "roofline": {"kernel": "raster_kernel<4> (cells + subdivision + direct raster)", "bound": "hbm",
                         "achieved": raster_gbs, "peak": peak, "unit": "GB/s", "frac": raster_gbs / peak,
                         "traffic": (traffic_pf * n_frames) if traffic_pf else None, "traffic_source": traffic_src,
                         "peak_source": peak_src, "launch_ms": r_ms, "bytes_per_launch": RASTER_BYTES_PER_PX * px,
                         "other_kernels": {k: {"launch_ms": stage_ms[k], "achieved": b * px / (stage_ms[k] * 1e-3) / 1e9,
                                               "frac": b * px / (stage_ms[k] * 1e-3) / 1e9 / peak, "bytes_per_px": b}
                                           for k, b in (("similarity_graph", 4), ("resolve_crossings", 2)) if k in stage_ms}},
            "stage_ms_per_step": stage_ms,
            "e2e": {"value": e2e_value, "unit": "frames/s", "h2d_bytes_per_step": int(pin_in.numel()), "d2h_bytes_per_step": int(d2h_idx),
                    "frames_per_step": n_frames, "steps": e2e_steps, "matches_device_path": same, "frames_compared": n_frames,
                    "output_format": "INDEX8: one palette index per output pixel + the frame's palette (<= 256 RGBA8 entries) + final graph; "
                                     "lossless (every output pixel is a source colour or the black background)",
                    "d2h_gbs": world * d2h_idx * e2e_steps / e2e_s / 1e9,
                    "pinned_d2h_peak_gbs_per_rank": d2h_peak,
                    "frac_of_pinned_copy_peak": (d2h_idx * e2e_steps / e2e_s / 1e9) / d2h_peak,
                    "host_ceiling_fps": world * d2h_peak * 1e9 / frame_bytes_idx,
                    "numa": numa,
                    "rgba8": e2e_rgba},
            "sustained": sustained,
            "strong_scaling": strong,
            "tiled": tiled,
            "parity": PARITY_NOTES,
            "gpu_launches": int(launches), "clocks": clock_info,
            "smoothing": {"cells": smooth1["smoothed"] - smooth0["smoothed"], "geometric_path": smooth1["geometric"] - smooth0["geometric"],
                          "what": "cells stage E ran on in the timed region, and how many of them the precomputed smoothing "
                                  "tables (built at context creation, content-independent) could not express"},
            "variants": extras,
        }
        if world == 1 and not args.no_cpu_baseline:
            threads = host_threads()
            uniq = synth.snes_stream(64, W, H)
            probe_fps = cpu_reference_fps(uniq[:max(threads, 8)], threads)[0]
            n_cpu = int(min(max(probe_fps * 12.0, 64), 4096))  # about 12 s of work on this box
            sample_frames = [uniq[k % 64] for k in range(n_cpu)]
            fps, kind, n, dt = cpu_reference_fps(sample_frames, threads)
            line["cpu_baseline"] = {"value": fps, "unit": "frames/s", "cores": threads, "kind": kind,
                                    "sample": "%d frames (64 distinct, cycled) of the same stream, %d host threads, %.1f s: reference "
                                              "routines A-F as host C++ + oracle triangle raster at 4x" % (n, threads, dt)}
            # the reference's own CUDA build (kernel.cu unmodified, sm_100a) on this GPU, as it is designed to run:
            # one launch_kernel call per frame incl. its per-call cudaMalloc/H2D/D2H/cudaFree.  A reported baseline.
            try:
                from oracle.oracle import RefCuda
                if RefCuda.available():
                    ms_call = RefCuda().time_calls(synth.snes_frame(W, H, synth.BASE_SEED), True, 20)
                    line["reference_cuda_baseline"] = {"value": 1e3 / ms_call, "unit": "frames/s", "ms_per_call": ms_call,
                                                       "what": "reference kernel.cu compiled for sm_100a (-use_fast_math), launch_kernel per frame "
                                                               "end to end as designed (alloc + H2D + 8 kernels + 365 B/px D2H + free); geometry "
                                                               "output only, no rasterization (the reference rasterizes in OpenGL)"}
            except Exception as exc:  # a baseline must never break the bench line
                line["reference_cuda_baseline"] = {"unavailable": str(exc)[:200]}
            # like for like with the line above: OUR export of the same symbol (launch_kernel, same arguments, same outputs
            # incl. the 45-slot VBO arrays and the host triangle list) called once per frame on the same frame
            try:
                line["launch_kernel_drop_in"] = time_drop_in(par, synth.snes_frame(W, H, synth.BASE_SEED), 20)
            except Exception as exc:
                line["launch_kernel_drop_in"] = {"unavailable": str(exc)[:200]}
        print(json.dumps(line))
    barrier()
    ctx.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--frames", type=int, default=FRAMES_PER_GPU, help="frames per GPU per step (default: the C3 stream)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-tiled", action="store_true", help="skip the config-C4 tiled-map check")
    ap.add_argument("--tiled-side", type=int, default=4096, help="side of the square map of the tiled check")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.gpus > 1 and world == 1:
        # convenience: re-launch under torchrun, one rank per GPU
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(args.gpus), "--master-addr", "127.0.0.1",
               "--master-port", str(29500 + os.getpid() % 1000), os.path.abspath(__file__)] + sys.argv[1:]
        return subprocess.call(cmd)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
