"""Aggregate an ncu source-page report of raster_kernel into code regions (by CUDA source line ranges given as
name:first-last ...).  Usage: python tools/ncu_regions.py report.ncu-rep kernel_regex file.cu name:a-b ..."""
import csv, subprocess, sys, io
rep, rx, fn = sys.argv[1], sys.argv[2], sys.argv[3]
regions = []
for spec in sys.argv[4:]:
    name, r = spec.split(":")
    a, b = r.split("-")
    regions.append((name, int(a), int(b)))
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "-k", "regex:" + rx],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
fname, hdr, first_kernel, cur_kernel = None, None, None, None
acc = {}
tot = [0, 0, 0]
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        fname = r[1].split("/")[-1]
        continue
    if r[0] == "Function Name":
        if first_kernel is None:
            first_kernel = r[1]
        cur_kernel = r[1]
        continue
    if r[0] == "Line No":
        hdr = r
        iI, iT, iS = hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed"), hdr.index("# Samples")
        continue
    if hdr is None or cur_kernel != first_kernel or r[0] == "":
        continue
    try:
        n, t, s, ln = int(r[iI]), int(r[iT]), int(r[iS]), int(r[0])
    except (ValueError, IndexError):
        continue
    key = "other:" + fname
    if fname == fn:
        key = "unassigned"
        for name, a, b in regions:
            if a <= ln <= b:
                key = name
                break
    v = acc.setdefault(key, [0, 0, 0])
    v[0] += n; v[1] += t; v[2] += s
    tot[0] += n; tot[1] += t; tot[2] += s
print("total warp instructions %d, thread instructions %d, samples %d" % tuple(tot))
for k, v in sorted(acc.items(), key=lambda kv: -kv[1][0]):
    print("%-28s %5.1f%% instr  %4.1f thr/warp  %5.1f%% samples" % (k, 100.0 * v[0] / tot[0], v[1] / max(v[0], 1), 100.0 * v[2] / max(tot[2], 1)))
