mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 > gpurun_out/r4b_pytest.txt
for i in 1 2; do
PAR_LIB=pixel_art_remaster_gpu_b200/build/variants/k3base.so python tools/k3_time.py
python tools/k3_time.py
done > gpurun_out/r4b_k3_ab.txt 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none -k regex:"cc_" -c 9 --csv --log-file gpurun_out/r4b_k3_launches.csv python tools/k3_time.py > /dev/null 2>&1
cat gpurun_out/r4b_pytest.txt gpurun_out/r4b_k3_ab.txt; grep -c cc_ gpurun_out/r4b_k3_launches.csv
