mkdir -p gpurun_out
for v in a b c d e; do
PAR_LIB=pixel_art_remaster_gpu_b200/build/variants/$v.so python tools/stage_time.py 4096
done > gpurun_out/r4k_stage.jsonl 2>&1
cat gpurun_out/r4k_stage.jsonl
