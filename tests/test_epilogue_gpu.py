"""GPU: the fused colour epilogue (palette / interpolation / AA average / pal_offset,
mdz_b200/csrc/colour.cuh) against the image the unmodified reference produced
(tests/golden), and the recolour-only kernel (palette cycling) against the C
oracle's epilogue for every offset step."""
import ctypes as C

import numpy as np
import pytest

import golden_util as G
import mdz_b200
import portpath

pytestmark = pytest.mark.gpu


def oracle_rgb(view, info, raw, pal_offset):
    lib = portpath.load()
    pal = np.zeros(256, dtype=np.uint32)
    pal[:len(info["palette"])] = info["palette"]
    out = np.zeros((view.user_height, view.user_width), dtype=np.uint32)
    rawc = np.ascontiguousarray(raw, dtype=np.int32)
    args = (C.c_double(info["colour_scale"]), int(info["palette_ip"]), pal.ctypes.data_as(C.c_void_p),
            len(info["palette"]), pal_offset)
    if view.aa_factor == 1:
        lib.oracle_palette_apply(rawc.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p),
                                 view.user_width, 0, view.user_height, *args)
    else:
        lib.oracle_do_anti_aliasing(rawc.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p),
                                    view.user_width, view.aa_factor, 0, view.user_height, *args)
    return out


@pytest.mark.parametrize("name", [n for n in G.names() if n != "test_240x180"])
def test_fused_epilogue_reproduces_reference_image(name):
    meta, raw, rgb = G.load(name)
    view, info = G.view_of(meta)
    if info["palette"] is None:
        pytest.skip("fixture without embedded palette")
    plan = mdz_b200.Plan(view)
    plan.set_colour(info["palette"], info["pal_offset"], info["colour_scale"], info["palette_ip"])
    plan.launch()
    got_raw = plan.fetch()
    got_rgb = plan.fetch_rgb()
    plan.close()
    assert np.array_equal(got_raw, raw)
    want = G.packed_rgb(rgb)
    assert np.array_equal(got_rgb, want), "%d of %d pixels differ" % (int((got_rgb != want).sum()), want.size)


@pytest.mark.parametrize("name", ["ship_ld_aa3", "celtic_mpfr128_aa2", "cfg2_ld", "julia_mpfr96"])
def test_palette_cycling_recolour(name):
    meta, raw, rgb = G.load(name)
    view, info = G.view_of(meta)
    plan = mdz_b200.Plan(view)
    plan.set_colour(info["palette"], info["pal_offset"], info["colour_scale"], info["palette_ip"])
    plan.launch()
    plan.wait()
    n = len(info["palette"])
    for off in list(range(0, n, 17)) + [n - 1]:      # palette_rotate_*: pal_offset +-1 mod pal_indexes
        plan.set_colour(info["palette"], off, info["colour_scale"], info["palette_ip"])
        plan.recolour()
        got = plan.fetch_rgb()
        want = oracle_rgb(view, info, raw, off)
        assert np.array_equal(got, want), (name, off, int((got != want).sum()))
    plan.close()


def test_fused_epilogue_across_band_partition():
    meta, raw, rgb = G.load("ship_ld_aa3")
    view, info = G.view_of(meta)
    out = np.zeros((view.user_height, view.user_width), dtype=np.uint32)
    for first in range(3):
        plan = mdz_b200.Plan(view, 0, first, 3)
        plan.set_colour(info["palette"], info["pal_offset"], info["colour_scale"], info["palette_ip"])
        plan.launch()
        plan.fetch_rgb(out)
        plan.close()
    assert np.array_equal(out, G.packed_rgb(rgb))
