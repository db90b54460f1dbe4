"""Per CUDA source line: shared-memory wavefronts (ideal + excessive) and global L1 tag requests of one kernel in an
ncu --set full report.  Usage: python tools/ncu_lsu.py report.ncu-rep kernel_regex [top_n]"""
import csv, subprocess, sys, io
rep, rx = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "-k", "regex:" + rx], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
fname, hdr, first, cur = None, None, None, None
acc = {}
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        fname = r[1].split("/")[-1]; continue
    if r[0] == "Function Name":
        first = first or r[1]; cur = r[1]; continue
    if r[0] == "Line No":
        hdr = r
        iW, iX, iG, iL2 = hdr.index("L1 Wavefronts Shared"), hdr.index("L1 Wavefronts Shared Excessive"), hdr.index("L1 Tag Requests Global"), hdr.index("L2 Theoretical Sectors Global")
        continue
    if hdr is None or cur != first or r[0] == "":
        continue
    def num(x):
        try: return int(x)
        except ValueError: return 0
    acc[(fname, r[0], r[1].strip()[:90])] = (num(r[iW]), num(r[iX]), num(r[iG]), num(r[iL2]))
tw = sum(v[0] for v in acc.values()); tg = sum(v[2] for v in acc.values())
print("shared wavefronts %d (excessive %d), global tag requests %d" % (tw, sum(v[1] for v in acc.values()), tg))
for k, v in sorted(acc.items(), key=lambda kv: -(kv[1][0] + kv[1][2]))[:top]:
    print("%9d shared (%8d excess) %9d global-tags %9d l2-sectors  %s:%s  %s" % (v[0], v[1], v[2], v[3], k[0], k[1], k[2]))
