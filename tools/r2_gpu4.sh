#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2d_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2d_pytest.log
tail -5 gpurun_out/r2d_pytest.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2d_bench.json 2> gpurun_out/r2d_bench.err; echo "bench rc=$?"
python -c "
import json
d=json.load(open('gpurun_out/r2d_bench.json'))
print(d['value'], d['ms_per_step'], d['stage_ms_per_step'], d['roofline']['frac'], d['e2e']['value'])
"
