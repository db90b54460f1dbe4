"""Golden border walks: the reference's OWN walker (cc_functions.cu, built by oracle/Makefile as oracle/_ref/libref_host_cc.so)
run on small final graphs of synthetic frames.  Only runs where /root/reference exists; writes
tests/golden/border_walks_small.json = [{"name", "W", "H", "graph": [...], "walks": [[...], ...]}, ...].

The reference's walker is dead code: on an island node (no links) getFirstLink returns -1 and c_neighbor_index( ., -1, . )
falls off its switch (cc_functions.cu:107-118, :215-246) — undefined behaviour that crashes the host build — so islands are
linked to a horizontal neighbour before the graph is handed over (the walker takes any symmetric graph), and every graph is
tried in a child process and kept only when the child survives."""
import json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np

CHILD = r"""
import sys, json
sys.path.insert(0, %r)
import numpy as np
from oracle import oracle as om
g = np.array(json.loads(sys.stdin.read()), np.uint8)
print(json.dumps(om.RefHostCC().border_walks(g)))
"""


def without_islands(g):
    g = g.copy()
    H, W = g.shape
    for j in range(H):
        for i in range(W):
            if g[j, i] == 0:
                if i + 1 < W:
                    g[j, i] |= 16
                    g[j, i + 1] |= 8
                else:
                    g[j, i] |= 8
                    g[j, i - 1] |= 16
    return g


def main():
    from oracle import oracle as om
    from pixel_art_remaster_gpu_b200 import synth
    om.build(ref=True)
    orc = om.Oracle()
    cases = []
    makers = [("snes_%d" % k, lambda k=k: synth.snes_frame(28, 24, 100 + k)) for k in range(12)]
    makers += [("adversarial_%d" % k, lambda k=k: synth.adversarial_sprite(96, 80, 20 + k)[8 * k:8 * k + 24, 30:62]) for k in range(6)]
    rng = np.random.default_rng(11)
    pal = rng.integers(0, 256, (2, 3), dtype=np.uint8)
    makers += [("blobs_%d" % k, lambda k=k: np.ascontiguousarray(pal[(rng.random((20, 26)) < 0.5 + 0.08 * k).astype(int)])) for k in range(6)]
    for name, make in makers:
        img = np.ascontiguousarray(make())
        g = without_islands(orc.pipeline(img, True, True, 4, ("graph",))["graph"])
        r = subprocess.run([sys.executable, "-c", CHILD % ROOT], input=json.dumps(g.tolist()), capture_output=True, text=True, timeout=120)
        if r.returncode != 0:
            print("reference walker died on", name, "-> skipped")
            continue
        walks = json.loads(r.stdout.strip().splitlines()[-1])
        cases.append({"name": name, "W": int(g.shape[1]), "H": int(g.shape[0]), "graph": g.reshape(-1).tolist(), "walks": walks})
        print(name, g.shape, len(walks), "walks")
    json.dump(cases, open(os.path.join(ROOT, "tests", "golden", "border_walks_small.json"), "w"))
    print(len(cases), "cases written")


if __name__ == "__main__":
    main()
