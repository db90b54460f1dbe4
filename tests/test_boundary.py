"""The drop-in boundary: the C-ABI library loads, exports every symbol include/pixelart_b200.h declares,
fails loudly without a GPU, and the product never touches the oracle.  No GPU compute here."""
import ctypes
import json
import os
import re
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_functions():
    text = open(os.path.join(ROOT, "include", "pixelart_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    names = set(re.findall(r"\b(par_[a-z_0-9]+|launch_kernel)\s*\(", text))
    return sorted(names)


def test_library_exports_every_declared_symbol(lib):
    names = _declared_functions()
    assert "launch_kernel" in names and "par_remaster_device" in names and len(names) >= 20
    L = ctypes.CDLL(lib.library_path())
    missing = [n for n in names if not hasattr(L, n)]
    assert not missing, "declared in include/pixelart_b200.h but not exported: %s" % missing


def test_reference_entry_point_signature_is_kept(lib):
    """kernel.cu:286-288 — same symbol, C linkage (no C++ mangling)."""
    out = subprocess.run(["nm", "-D", "--defined-only", lib.library_path()], capture_output=True, text=True, check=True).stdout
    assert re.search(r"\bT launch_kernel$", out, flags=re.M)


def test_no_gpu_means_loud_failure(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(lib.RemasterError) as e:
        lib.Remaster(0, 64, 64, 1)
    assert "PAR_ERR_NO_DEVICE" in str(e.value) and "no CPU fallback" in str(e.value)


def test_bad_arguments_are_rejected_without_touching_the_gpu(lib):
    L = lib.load_library()
    assert L.par_remaster_device(None, None) == 1          # PAR_ERR_INVALID
    assert L.par_synchronize(None) == 1
    assert L.par_cell_from_pattern(4096, (ctypes.c_float * 90)()) == -1
    assert L.par_launch_count(None) == 0


def test_product_never_uses_the_oracle():
    """Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline/reference legs may touch oracle/."""
    pkg = os.path.join(ROOT, "pixel_art_remaster_gpu_b200")
    offenders = []
    for d, _dirs, files in os.walk(pkg):
        if os.sep + "build" in d:
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                txt = open(os.path.join(d, f), errors="ignore").read()
                if re.search(r"\boracle\b|liboracle|ref_host|ref_cuda|/root/reference", txt):
                    offenders.append(os.path.relpath(os.path.join(d, f), ROOT))
    assert not offenders, offenders
    lib_path = os.path.join(pkg, "libpixelart_b200.so")
    if os.path.exists(lib_path):
        needed = subprocess.run(["readelf", "-d", lib_path], capture_output=True, text=True).stdout
        assert "oracle" not in needed


def test_reference_arm_of_bench_runs_on_cpu():
    """bench.py --impl reference: the reference's CPU routines on the host cores, one JSON line."""
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=600, check=True).stdout.strip().splitlines()
    line = json.loads(out[-1])
    assert line["impl"] == "reference" and line["unit"] == "frames/s" and line["value"] > 0
    assert line["cpu_baseline"]["kind"] in ("reference", "port") and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["gpu_launches"] == 0
