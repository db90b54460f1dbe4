mkdir -p gpurun_out
for i in 1 2; do
PAR_LIB=pixel_art_remaster_gpu_b200/build/variants/th16.so python tools/k3_time.py
python tools/k3_time.py
done > gpurun_out/r4a_k3_ab.txt 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"similarity_graph|resolve_crossings|cc_tile|cc_seam|cc_flatten" -s 5 -c 5 -o gpurun_out/r4a_k123 -f python tools/prof_raster.py 256 4 > gpurun_out/r4a_ncu.log 2>&1
cat gpurun_out/r4a_k3_ab.txt; tail -3 gpurun_out/r4a_ncu.log
