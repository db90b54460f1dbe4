mkdir -p gpurun_out
out=gpurun_out/r3h_ab.txt
: > $out
for v in base "" base ""; do
  if [ -z "$v" ]; then lib=""; else lib=pixel_art_remaster_gpu_b200/build/variants/$v.so; fi
  echo "== ${v:-default}" >> $out
  PAR_LIB=$lib timeout 300 python tools/k4_time.py 2048 4 >> $out 2>&1
done
cat $out
