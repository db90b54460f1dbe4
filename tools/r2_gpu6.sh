#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2f_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2f_pytest.log
tail -4 gpurun_out/r2f_pytest.log
timeout 300 python tools/stage_time.py 4096 > gpurun_out/r2f_stage_time.jsonl 2> gpurun_out/r2f_stage_time.err
cat gpurun_out/r2f_stage_time.jsonl; tail -3 gpurun_out/r2f_stage_time.err
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2f_bench.json 2> gpurun_out/r2f_bench.err; echo "bench rc=$?"
tail -5 gpurun_out/r2f_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2f_bench.json'))
for k in ('value','ms_per_step','stage_ms_per_step','sustained','strong_scaling','tiled','e2e'):
    print(k, json.dumps(d.get(k))[:700])
print('roofline', d['roofline']['frac'], d['roofline']['other_kernels'])
PY
