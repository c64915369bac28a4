#!/usr/bin/env python
"""Band completion over time for one render (where does a frame's time go?): launches the case, polls the
band flags every 25 ms and prints, per time slice, how many bands were complete and the highest complete row."""
import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import ctypes as C
import numpy as np
import mdz_b200
from mdz_b200 import _native as N
from views import config4m, config4

name = sys.argv[1] if len(sys.argv) > 1 else "mini"
w, h = (int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (3840, 2160)
v = config4m(w, h, 100000) if name == "mini" else config4(w, h, 100000, mode="mpfr")
p = mdz_b200.Plan(v, 0)
p.launch(); p.wait()
flags = np.zeros(p.bands_total() + 1, dtype=np.uint8)
t0 = time.perf_counter()
p.launch()
rows = []
while True:
    n = N.lib.mdzcuda_plan_poll_bands(p.h, flags.ctypes.data_as(C.c_void_p))
    t = time.perf_counter() - t0
    done = np.flatnonzero(flags[:p.bands_total()])
    first_missing = int(np.flatnonzero(flags[:p.bands_total()] == 0)[0]) if n < p.bands_total() else p.bands_total()
    rows.append((t, n, first_missing, int(done.max()) if len(done) else -1))
    if n >= p.bands_total():
        break
    time.sleep(0.025)
p.wait()
print("total %.3f s" % (time.perf_counter() - t0))
step = max(1, len(rows) // 60)
for t, n, fm, mx in rows[::step] + [rows[-1]]:
    print("t=%7.3f  bands done %5d  first missing %5d  highest done %5d" % (t, n, fm, mx))
