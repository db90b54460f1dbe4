#!/bin/bash
# End-of-round measurement on an 8-GPU box: the tiled tests across devices, the tiled map at 1/2/4/8 devices, the bench at 8 (and 4) GPUs.
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo8.txt 2>&1
timeout 600 python -m pytest tests/test_gpu_tiled_and_compat.py -m gpu -x -q 2>&1 | tail -3 > gpurun_out/pytest_gpu8_tiled.txt
python - > gpurun_out/tiled_1_2_4_8.jsonl 2> gpurun_out/tiled_1_2_4_8.err <<'PY'
import sys, json, torch
sys.path.insert(0, '.')
import bench, pixel_art_remaster_gpu_b200 as par
from pixel_art_remaster_gpu_b200 import synth
for n in (1, 2, 4, 8):
    if n <= torch.cuda.device_count():
        print(json.dumps(bench.tiled_map_check(par, synth, torch, n, 4096)), flush=True)
PY
for N in ${BENCH_NS:-8 4}; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_${N}gpu.json 2> gpurun_out/bench_${N}gpu.err
  echo "bench N=$N rc=$?"
done
cat gpurun_out/pytest_gpu8_tiled.txt; cut -c1-260 gpurun_out/tiled_1_2_4_8.jsonl; tail -2 gpurun_out/tiled_1_2_4_8.err
python - <<'PY'
import json
for N in (8, 4):
    try:
        d = json.loads(open('gpurun_out/bench_%dgpu.json' % N).read().strip().splitlines()[-1])
        print(N, 'value', round(d['value']), 'strong', round(d['strong_scaling']['value']), 'e2e', round(d['e2e']['value']), 'ceiling', round(d['e2e']['host_ceiling_fps']),
              'peak/rank', round(d['e2e']['pinned_d2h_peak_gbs_per_rank'], 1), 'tiled', json.dumps(d['tiled'])[:330])
    except Exception as e:
        print(N, 'failed', e)
PY
