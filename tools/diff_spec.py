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
for chunk in (32, -32):
    p = mdz_b200.Plan(v, 0)
    p.tune(chunk, 0)
    p.launch(); out.append(p.fetch()); p.close()
bad = np.argwhere(out[0] != out[1])
print("mismatches:", len(bad))
for line, ix in bad[:40]:
    print("line %d ix %d spec %d general %d" % (line, ix, out[0][line, ix], out[1][line, ix]))
