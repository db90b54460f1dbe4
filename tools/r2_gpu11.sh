#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tiled_and_compat.py tests/test_host_tools.py -m gpu -x -q > gpurun_out/r2k_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2k_pytest.log
tail -3 gpurun_out/r2k_pytest.log
python - <<'PY'
import sys, json, torch
sys.path.insert(0, '.')
import bench, pixel_art_remaster_gpu_b200 as par
from pixel_art_remaster_gpu_b200 import synth
for n in (1, 2):
    print(json.dumps(bench.tiled_map_check(par, synth, torch, n, 4096))[:330])
PY
