#!/usr/bin/env python
"""Run one named render case once or a few times (for ncu / quick timing)."""
import argparse
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np
import mdz_b200
from views import make_view, config2, SEAHORSE, deep_embedded_julia, config4, config4m, MINIBROT120


def case(name, scale=1.0):
    w, h = int(960 * scale), int(540 * scale)
    if name == "mini":          # the target view (minibrot in the frame), MPFR 512
        return config4m(w, h, DEPTH or 100000)
    if name == "minigmp":
        return config4m(w, h, DEPTH or 100000, mode="gmp")
    if name == "minioff":       # the same frame moved up by its own height: no interior
        from mdz_b200.mp import Mpfr
        import mpmath
        mpmath.mp.dps = 200
        cy = mpmath.mpf(MINIBROT120[1]) + mpmath.mpf("0.5625e-120")
        return make_view(MINIBROT120[0], mpmath.nstr(cy, 180), "1e-120", w, h, precision=512, depth=DEPTH or 100000)
    if name == "misi":          # round 1's target view: Misiurewicz point, every pixel escapes
        return config4(w, h, DEPTH or 100000, mode="mpfr")
    if name == "ld":
        return config2(2 * w, 2 * h, 10000)
    if name.startswith("cfg2p"):
        # config 2's view (column x = 0, row y = 0 included) in MPFR mode at the given precision
        return make_view("-0.5", "0.0", "4.0", w, h, precision=int(name[5:]), depth=10000)
    if name == "ldshift":
        # config 2 moved by a fraction of a pixel: no column with x exactly 0, no row next to y = 0
        return make_view("-0.4991", "0.0007", "4.0", 2 * w, 2 * h, mode="ld", depth=10000)
    if name.startswith("int"):
        # every pixel inside the main cardioid: `lanes` pixels that all run to depth (per-warp pace at a given occupancy)
        lanes = int(name[3:] or 113664)
        return make_view("-0.1", "0.2", "0.1", 1024, lanes // 1024, mode="ld", depth=10000)
    if name.startswith("mpfr"):
        p = int(name[4:])
        if p == 320:
            return deep_embedded_julia(w, h)
        return make_view(SEAHORSE[0], SEAHORSE[1], "1e-12", w, h, precision=p, depth=10000)
    if name.startswith("gmp"):
        return make_view(SEAHORSE[0], SEAHORSE[1], "1e-12", w, h, mode="gmp", precision=int(name[3:]), depth=10000)
    if name.startswith("sea"):
        return make_view(SEAHORSE[0], SEAHORSE[1], "1e-12", w, h, precision=int(name[3:]), depth=10000)
    if name.startswith("dej"):
        return deep_embedded_julia(w, h, precision=int(name[3:]))
    raise SystemExit("unknown case " + name)


ap = argparse.ArgumentParser()
ap.add_argument("case")
ap.add_argument("--reps", type=int, default=1)
ap.add_argument("--scale", type=float, default=1.0)
ap.add_argument("--chunk", type=int, default=0)
ap.add_argument("--bps", type=int, default=0)
ap.add_argument("--cycle", type=int, default=0, help="1: exact periodicity check on")
ap.add_argument("--park", type=int, default=-1, help="tail compaction: -1 automatic, 0 off, 1 on")
ap.add_argument("--depth", type=int, default=0)
ap.add_argument("--order", type=int, default=0, help="1: bands from the middle outwards")
ap.add_argument("--stride", type=int, default=1, help="render bands 0, stride, 2*stride, ... only (one rank's share of a strong-scaled render)")
a = ap.parse_args()
DEPTH = a.depth
v = case(a.case, a.scale)
plan = mdz_b200.Plan(v, 0, 0, a.stride)
plan.tune(a.chunk, a.bps)
plan.set_cycle_detection(bool(a.cycle))
if a.order:
    plan.set_order(True)
if a.park != -1:
    plan.set_parking(a.park)
for i in range(a.reps):
    t0 = time.perf_counter()
    plan.launch()
    plan.wait()
    dt = time.perf_counter() - t0
    raw = plan.fetch()
    it = int(np.where(raw > 0, raw, v.depth).astype(np.int64).sum())
    import zlib
    print("%s rep %d: %.2f ms, %d iterations, %.3f G it/s, crc %08x, kernel %s" % (a.case, i, dt * 1e3, it, it / dt / 1e9, zlib.crc32(raw.tobytes()), plan.kernel_info()))
