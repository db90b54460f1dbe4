#!/bin/bash
# two GPUs: the tiled path across devices (peer copies, on-device stitch), the whole GPU suite, the bench under torchrun
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r2g_gpus.txt; nvidia-smi topo -m >> gpurun_out/r2g_gpus.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2g_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2g_pytest.log
tail -4 gpurun_out/r2g_pytest.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2g_bench_2gpu.json 2> gpurun_out/r2g_bench_2gpu.err; echo "bench rc=$?"
tail -3 gpurun_out/r2g_bench_2gpu.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2g_bench_2gpu.json').read().strip().splitlines()[-1])
for k in ('value','ms_per_step','strong_scaling','tiled'):
    print(k, json.dumps(d.get(k))[:600])
print('e2e', d['e2e']['value'], d['e2e']['pinned_d2h_peak_gbs_per_rank'], d['e2e']['host_ceiling_fps'], d['e2e']['matches_device_path'])
PY
