#!/bin/bash
# End-of-round measurement on one B200 (run under gpurun): tests, both bench arms, ncu launch list, full captures, DRAM
# traffic of the raster kernel over the bench's own 4096-frame launch, the other configs, compute-sanitizer.
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 > gpurun_out/pytest_gpu.txt
python __graft_entry__.py smoke > gpurun_out/smoke.txt 2>&1
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err
python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --frames 1024 --no-cpu-baseline --no-tiled > gpurun_out/bench_under_ncu.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"similarity_graph|resolve_crossings|cc_tile|cc_seam|cc_flatten|palette|raster_kernel" -s 6 -c 6 -o gpurun_out/prof_r4 -f python tools/prof_raster.py 256 4 > gpurun_out/ncu_full.log 2>&1
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:raster_kernel -s 1 -c 1 -o gpurun_out/prof_r4_traffic4096 -f python tools/prof_raster.py 4096 4 > gpurun_out/ncu_traffic.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:raster_kernel -s 1 -c 1 -o gpurun_out/prof_r4_s8 -f python tools/prof_raster.py 128 8 > gpurun_out/ncu_s8.log 2>&1
python tools/quick_time.py 256 8 320 240 > gpurun_out/quick_c2.txt 2>&1
python tools/quick_time.py 64 4 512 448 > gpurun_out/quick_c5.txt 2>&1
python tools/stage_time.py 4096 > gpurun_out/stage_time.jsonl 2>&1
( python tools/k4_time.py 2048 4; K4_SPARSE=1 python tools/k4_time.py 2048 4; K4_NOISE=1 python tools/k4_time.py 2048 4; python tools/k4_time.py 512 8 ) > gpurun_out/k4_time.txt 2>&1
( timeout 900 compute-sanitizer --tool memcheck --error-exitcode 1 python tools/sanitize.py; echo "memcheck rc=$?"; timeout 900 compute-sanitizer --tool racecheck --error-exitcode 1 python tools/sanitize.py; echo "racecheck rc=$?" ) > gpurun_out/sanitizer.txt 2>&1
cat gpurun_out/pytest_gpu.txt gpurun_out/smoke.txt; tail -c 600 gpurun_out/bench.json; tail -5 gpurun_out/sanitizer.txt
ls -la gpurun_out
