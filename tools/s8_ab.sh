# A/B of the raster kernel at the scales above 4 (development aid): bash tools/s8_ab.sh out.txt variant ...
mkdir -p gpurun_out
out=gpurun_out/$1; shift
: > $out
for v in "$@"; do
  if [ "$v" = "-" ]; then lib=""; else lib=pixel_art_remaster_gpu_b200/build/variants/$v.so; fi
  for s in 8 6 5 7; do
    echo "== $v scale $s" >> $out
    PAR_LIB=$lib timeout 300 python tools/k4_time.py 512 $s 2>&1 | grep -v smooth >> $out
  done
done
cat $out
