#!/usr/bin/env python
"""Render a case with the speculative step on and off and list the pixels that differ
(there must be none; a debugging aid for new fast paths)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import mdz_b200
from views import config2

w, h = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (1920, 1080)
v = config2(w, h, 10000)
out = []
for chunk, cyc in ((32, 0), (-32, 0), (32, 1), (-32, 1)):
    p = mdz_b200.Plan(v, 0)
    p.tune(chunk, 0)
    p.set_cycle_detection(bool(cyc))
    import time
    p.launch(); p.wait(); t0 = time.perf_counter(); p.launch(); p.wait(); dt = time.perf_counter() - t0
    out.append(p.fetch()); p.close()
    print("chunk %d cycle %d: %.2f ms" % (chunk, cyc, dt * 1e3))
for k in (2, 3):
    print("cycle-detect run %d vs full iteration: %d mismatches" % (k, int((out[k] != out[1]).sum())))
bad = np.argwhere(out[0] != out[1])
print("mismatches:", len(bad))
for line, ix in bad[:40]:
    print("line %d ix %d spec %d general %d" % (line, ix, out[0][line, ix], out[1][line, ix]))
