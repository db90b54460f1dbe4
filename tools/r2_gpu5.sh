#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2e_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2e_pytest.log
tail -5 gpurun_out/r2e_pytest.log
for lib in "" pixel_art_remaster_gpu_b200/build/variants/nohead.so pixel_art_remaster_gpu_b200/build/variants/nolut.so ""; do
  PAR_LIB=$lib timeout 300 python tools/stage_time.py 4096 >> gpurun_out/r2e_stage_time.jsonl 2>> gpurun_out/r2e_stage_time.err
done
cat gpurun_out/r2e_stage_time.jsonl; tail -3 gpurun_out/r2e_stage_time.err
