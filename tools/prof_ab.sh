mkdir -p gpurun_out
for v in base new; do
  if [ $v = base ]; then lib=pixel_art_remaster_gpu_b200/build/variants/base.so; else lib=""; fi
  for sub in 0 1; do
    PAR_LIB=$lib timeout 300 ncu --set full --clock-control none --import-source on -k regex:raster_kernel -s 2 -c 1 -o gpurun_out/r3f_${v}_sub${sub} -f python tools/prof_k4.py 256 4 $sub > gpurun_out/r3f_ncu_${v}_${sub}.log 2>&1
  done
done
ls -la gpurun_out/r3f_*
