"""Development aid: another build of the library with extra -D switches on some sources, for A/B timing on the GPU box
(PAR_LIB=pixel_art_remaster_gpu_b200/build/variants/<name>.so python tools/k4_time.py).
Usage: python tools/build_variant.py name file.cu[,file2.cu] -DPAR_K4_SORT=0 ..."""
import os, subprocess, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pixel_art_remaster_gpu_b200 import build as b

name, files, defs = sys.argv[1], sys.argv[2].split(","), sys.argv[3:]
b.build_library()
bd = os.path.join(b.HERE, "build")
vd = os.path.join(bd, "variants")
os.makedirs(vd, exist_ok=True)
objs = []
for s in b.sources():
    base = os.path.basename(s)
    o = os.path.join(bd, base + ".o")
    if base in files:
        o = os.path.join(vd, "%s.%s.o" % (name, base))
        subprocess.run([b._nvcc()] + b.ARCH + b.NVCC_FLAGS + defs + ["-I", os.path.join(b.ROOT, "include"), "-x", "cu", "-c", s, "-o", o], check=True)
    objs.append(o)
out = os.path.join(vd, name + ".so")
subprocess.run([b._nvcc()] + b.ARCH + ["-shared", "-o", out] + objs, check=True)
print(out)
