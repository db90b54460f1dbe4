#!/bin/bash
# End-of-round measurement on one B200 (run under gpurun): tests, both bench arms, ncu launch list and full capture.
mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 > gpurun_out/pytest_gpu.txt
python __graft_entry__.py smoke > gpurun_out/smoke.txt 2>&1
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err
python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --frames 1024 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"similarity_graph|resolve_crossings|cc_tile|raster_kernel" -s 4 -c 4 -o gpurun_out/prof_r1 -f python tools/prof_raster.py 256 4 > gpurun_out/ncu_full.log 2>&1
python tools/quick_time.py 256 8 320 240 > gpurun_out/quick_c2.txt 2>&1
python tools/quick_time.py 64 4 512 448 > gpurun_out/quick_c5.txt 2>&1
cat gpurun_out/pytest_gpu.txt gpurun_out/smoke.txt gpurun_out/bench.json
ls -la gpurun_out
