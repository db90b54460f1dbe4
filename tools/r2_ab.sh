#!/bin/bash
# A/B of library variants on the bench stream: tools/r2_ab.sh variant1.so variant2.so ...   ("" = the default build)
mkdir -p gpurun_out
rm -f gpurun_out/ab_stage_time.jsonl
for rep in 1 2; do
  for lib in "$@"; do
    PAR_LIB=$lib timeout 300 python tools/stage_time.py 4096 >> gpurun_out/ab_stage_time.jsonl 2>> gpurun_out/ab_stage_time.err
  done
done
cat gpurun_out/ab_stage_time.jsonl; tail -3 gpurun_out/ab_stage_time.err
