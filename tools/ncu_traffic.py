"""DRAM traffic of the raster kernel from an ncu report (dram__bytes_read.sum, dram__bytes_write.sum) -> profiles/<name>
(bench.py reads the newest profiles/r*_raster_traffic.json for roofline.traffic).
Usage: python tools/ncu_traffic.py report.ncu-rep frames_in_the_profiled_launch [output name]"""
import csv, io, json, subprocess, sys
rep, frames = sys.argv[1], int(sys.argv[2])
name = sys.argv[3] if len(sys.argv) > 3 else "r2_raster_traffic.json"
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
h, u = rows[0], rows[1]
r = [x for x in rows[2:] if "raster_kernel" in x[h.index("Kernel Name")]][0]
def val(name):
    i = h.index(name)
    scale = {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1.0}[u[i]]
    return float(r[i]) * scale
rd, wr = val("dram__bytes_read.sum"), val("dram__bytes_write.sum")
json.dump({"dram_bytes_read": rd, "dram_bytes_write": wr, "frames": frames, "dram_bytes_per_frame": (rd + wr) / frames,
           "algorithmic_bytes_per_frame": 256 * 224 * 68,
           "source": "ncu, %s, raster_kernel<4> over ONE launch of %d frames of 256x224 (dram__bytes_read.sum + dram__bytes_write.sum)" % (rep.split("/")[-1], frames)},
          open("profiles/" + name, "w"), indent=1)
print(open("profiles/" + name).read())
