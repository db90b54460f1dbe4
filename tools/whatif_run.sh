# A/B timing of builds of the library on one box (development aid): bash tools/whatif_run.sh out.txt variant ...
# ("" = the shipped library; other names = pixel_art_remaster_gpu_b200/build/variants/<name>.so, tools/build_variant.py)
mkdir -p gpurun_out
out=gpurun_out/$1; shift
: > $out
for v in "$@"; do
  if [ "$v" = "-" ]; then lib=""; else lib=pixel_art_remaster_gpu_b200/build/variants/$v.so; fi
  echo "== $v" >> $out
  PAR_LIB=$lib timeout 300 python tools/k4_time.py 2048 4 2>&1 | grep -v smooth >> $out
  PAR_LIB=$lib K4_SPARSE=1 timeout 300 python tools/k4_time.py 2048 4 2>&1 | grep "sub=True" | sed 's/^/   sparse /' >> $out
done
cat $out
