"""Summarise an ncu report per CUDA source line: share of executed warp instructions, threads per warp,
stall samples.  Usage: python tools/ncu_lines.py report.ncu-rep kernel_regex [top_n]"""
import csv, subprocess, sys, io
rep, rx = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "-k", "regex:" + rx],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
fname, lines, hdr, first_kernel, seen_kernel = None, [], None, None, 0
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
    if hdr is None or cur_kernel != first_kernel:
        continue
    if r[0] != "":  # a CUDA source line row (aggregated over its SASS)
        try:
            lines.append((int(r[iI]), int(r[iT]), int(r[iS]), fname, r[0], r[1].strip()))
        except (ValueError, IndexError):
            pass
tot = sum(l[0] for l in lines) or 1
tots = sum(l[2] for l in lines) or 1
print(first_kernel)
print("total warp instructions %d, stall samples %d" % (tot, tots))
for n, t, s, f, ln, src in sorted(lines, reverse=True)[:top]:
    print("%5.1f%% instr  %4.1f thr/warp  %5.1f%% samples  %s:%s  %s" % (100.0 * n / tot, t / max(n, 1), 100.0 * s / tots, f, ln, src[:100]))
