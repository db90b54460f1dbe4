mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "label or cc or tiled or config" 2>&1 | tail -3 > gpurun_out/r4f_pytest.txt
for i in 1 2; do
PAR_LIB=pixel_art_remaster_gpu_b200/build/variants/k3prev.so python tools/k3_time.py
python tools/k3_time.py
done > gpurun_out/r4f_k3_ab.txt 2>&1
cat gpurun_out/r4f_pytest.txt gpurun_out/r4f_k3_ab.txt
