#!/bin/bash
# Round 2, first GPU session: the whole GPU test suite, the probe, a bench line.
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r2a_gpus.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2a_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2a_pytest.log
tail -5 gpurun_out/r2a_pytest.log
timeout 600 python tools/r2_probe.py 4096 > gpurun_out/r2a_probe.jsonl 2> gpurun_out/r2a_probe.err; echo "probe rc=$?"
tail -20 gpurun_out/r2a_probe.jsonl
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err; echo "bench rc=$?"
tail -c 1500 gpurun_out/r2a_bench.json
