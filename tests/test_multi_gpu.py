"""The library's own multi-device path (mdzcuda_render with ndev > 1, which is also what the rth_* layer
runs): fed plans + the host-side band scheduler.  Needs at least two GPUs -- `gpurun --gpus 2`; on a
one-GPU lease these tests skip (the scheduler's policy and the fed queue's host logic are covered on CPU
by tests/test_band_scheduler_cpu.py, and the fed kernel path on one GPU by test_fed_plan_single_device).
SURVEY 8(e); reference analogue src/render_threads.c:360-393."""
import os
import time

import numpy as np
import pytest

import mdz_b200
from mdz_b200 import _native as N
from views import config2, make_view, SEAHORSE, config4m

pytestmark = pytest.mark.gpu


def ndev():
    return mdz_b200.device_count()


@pytest.mark.parametrize("centre_out", [False, True], ids=["raster", "centre_out_tiles"])
def test_fed_plan_single_device(centre_out):
    """A fed plan on ONE device, driven by hand: bands fed out of order, in several steps, while the kernel
    runs; the result must equal the plain plan's, and the bands must arrive where they belong.  In centre-out
    order every fed chunk is cut into tiles and its middle columns go first."""
    import ctypes as C
    v = make_view(SEAHORSE[0], SEAHORSE[1], "1e-6", 320, 200, precision=128, depth=3000)
    want = mdz_b200.render(v, (0,))
    p = mdz_b200.Plan(v, 0)
    p.set_order(centre_out)
    assert N.lib.mdzcuda_plan_set_fed(p.h, 1)
    p.launch()
    order = list(range(199, -1, -1))            # bottom to top
    for k in range(0, 200, 37):
        chunk = order[k:k + 37]
        arr = (C.c_int * len(chunk))(*chunk)
        assert N.lib.mdzcuda_plan_feed(p.h, arr, len(chunk), 0), mdz_b200.last_error()
        time.sleep(0.002)
    assert N.lib.mdzcuda_plan_feed(p.h, None, 0, 1)
    p.wait()
    got = p.fetch()
    p.close()
    assert np.array_equal(got, want)
    assert mdz_b200.fallback_lines() == 0


@pytest.mark.skipif(ndev() < 2, reason="needs two GPUs")
@pytest.mark.parametrize("view", ["cfg2", "mpfr128", "aa"])
def test_all_devices_dynamic_equals_one_device_and_static(view, monkeypatch):
    v = {"cfg2": config2(1920, 1080, 10000),
         "mpfr128": make_view(SEAHORSE[0], SEAHORSE[1], "1e-12", 960, 540, precision=128, depth=10000),
         "aa": make_view("-0.5", "0.0", "4.0", 400, 225, mode="ld", depth=2000, aa=3)}[view]
    devs = tuple(range(ndev()))
    one = mdz_b200.render(v, (0,))
    dyn = mdz_b200.render(v, devs)
    assert np.array_equal(dyn, one), "%d pixels differ (dynamic scheduler)" % int((dyn != one).sum())
    monkeypatch.setenv("MDZCUDA_SCHED", "static")
    sta = mdz_b200.render(v, devs)
    assert np.array_equal(sta, one), "%d pixels differ (static interleave)" % int((sta != one).sum())
    assert mdz_b200.fallback_lines() == 0


@pytest.mark.skipif(ndev() < 2, reason="needs two GPUs")
def test_busy_device_takes_fewer_bands(monkeypatch):
    """One device has half its SMs taken by another job.  With a fixed interleave it still gets its full share
    and the render waits for it; the scheduler gives it what it can eat."""
    v = make_view(SEAHORSE[0], SEAHORSE[1], "1e-12", 1920, 1080, precision=320, depth=20000)
    n = ndev()
    devs = tuple(range(n))
    mdz_b200.render(v, devs)                                   # contexts, pools, modules
    t0 = time.perf_counter(); want = mdz_b200.render(v, devs); t_all = time.perf_counter() - t0
    t0 = time.perf_counter(); mdz_b200.render(v, devs[:-1]); t_less = time.perf_counter() - t0
    busy = devs[-1]
    hog_ms = int(t_less * 1e3 * 6) + 500

    def timed():
        assert N.lib.mdzcuda_debug_occupy(busy, 74, hog_ms), mdz_b200.last_error()
        t = time.perf_counter(); got = mdz_b200.render(v, devs); dt = time.perf_counter() - t
        # let the occupying kernel finish before the next measurement
        time.sleep(max(0.0, hog_ms * 1e-3 - dt) + 0.05)
        return got, dt
    got, t_dyn = timed()
    assert np.array_equal(got, want)
    monkeypatch.setenv("MDZCUDA_SCHED", "static")
    got, t_sta = timed()
    assert np.array_equal(got, want)
    print("all idle %.1f ms, without the busy device %.1f ms, busy device at half its SMs: scheduler %.1f ms, fixed interleave %.1f ms"
          % (t_all * 1e3, t_less * 1e3, t_dyn * 1e3, t_sta * 1e3))
    assert t_dyn < 1.3 * t_less           # no worse than leaving the busy device out, within 30 %
    assert t_dyn < 0.9 * t_sta            # and clearly better than the fixed split


@pytest.mark.skipif(ndev() < 2, reason="needs two GPUs")
def test_target_view_all_devices_sampled_lines(ref_lib):
    from refpath import ref_render_lines
    v = config4m(3840, 2160, 100000, mode="mpfr", precision=512)
    got = mdz_b200.render(v, tuple(range(ndev())))
    lines = [0, 700, 1079, 2159]
    assert np.array_equal(got[lines], ref_render_lines(ref_lib, v, lines))
    assert (got == 0).mean() > 0.01
