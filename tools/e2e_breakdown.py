#!/usr/bin/env python
"""Where the end-to-end time of one render goes (plan create / launch+wait / fetch / destroy)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import mdz_b200
from views import make_view, config2, SEAHORSE
for name, v in (("ld 1920x1080", config2(1920, 1080, 10000)),
                ("mpfr512 960x540", make_view(SEAHORSE[0], SEAHORSE[1], "1e-12", 960, 540, precision=512, depth=10000))):
    for rep in range(3):
        t0 = time.perf_counter(); p = mdz_b200.Plan(v, 0)
        t1 = time.perf_counter(); p.launch(); p.wait()
        t2 = time.perf_counter(); raw = p.fetch()
        t3 = time.perf_counter(); p.close()
        t4 = time.perf_counter()
        print("%-16s rep %d: create %.2f ms, kernel %.2f ms, fetch %.2f ms, destroy %.2f ms" % (
            name, rep, (t1 - t0) * 1e3, (t2 - t1) * 1e3, (t3 - t2) * 1e3, (t4 - t3) * 1e3))
