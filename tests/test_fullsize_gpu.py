"""GPU parity at BASELINE.json's full sizes.

Where the host cores can render the whole image with the unmodified reference in a
second or two (config 2) the comparison is complete; elsewhere the reference's own line
driver is run on sampled lines of the full-size view (oracle/ref_driver.c:
ref_render_lines) and the rest is covered by size-independent properties: the known
iteration sum, band partitions that must reproduce the whole, the colour epilogue
recomputed on the host from the GPU's own counts, and the periodicity check, which must
change nothing."""
import ctypes as C

import numpy as np
import pytest

import mdz_b200
import portpath
from mdz_b200 import BURNING_SHIP, GENERALIZED_CELTIC
from refpath import ref_render, ref_render_lines
from views import config2, config4, config4m, config5, deep_embedded_julia

pytestmark = pytest.mark.gpu


def iterations(raw, depth):
    return int(np.where(raw > 0, raw, depth).astype(np.int64).sum())


def sample_lines(height, count, aa=1):
    """evenly spread real lines, always including the first, the middle and the last"""
    ls = sorted(set([0, height // 2, height - 1] + [int(k * (height - 1) / max(1, count - 1)) for k in range(count)]))
    return ls


def check_lines(ref_lib, view, got, lines):
    want = ref_render_lines(ref_lib, view, lines)
    bad = np.argwhere(got[lines] != want)
    assert bad.size == 0, "%d mismatching pixels on the sampled lines, first: line %d ix %d got %d want %d" % (
        len(bad), lines[bad[0][0]], bad[0][1], got[lines][tuple(bad[0])], want[tuple(bad[0])])


def test_config2_1920x1080_long_double_complete(ref_lib):
    v = config2(1920, 1080, 10000)
    got = mdz_b200.render(v)
    assert iterations(got, v.depth) == 3482185482
    want, _ = ref_render(ref_lib, v)
    assert np.array_equal(got, want), "%d pixels differ" % int((got != want).sum())
    # the periodicity check finishes the interior early and must not change a pixel
    p = mdz_b200.Plan(v, 0)
    p.set_cycle_detection(True)
    assert np.array_equal(p.run(), want)
    p.close()


def test_config3_deep_embedded_julia_1920x1080_mpfr320(ref_lib):
    v = deep_embedded_julia(1920, 1080)                     # the view its author meant (SURVEY 8d config 3 ii)
    got = mdz_b200.render(v)
    assert (got == 0).mean() < 0.01                         # an embedded Julia set: hardly anything reaches depth
    check_lines(ref_lib, v, got, sample_lines(1080, 10))
    # two interleaved plans (what two GPUs would render) reproduce the whole
    out = np.full_like(got, -1)
    for first in range(2):
        p = mdz_b200.Plan(v, 0, first, 2)
        p.run(out)
        p.close()
    bad = sorted(set(int(r) for r in np.argwhere(out != got)[:, 0]))
    assert not bad, "%d lines differ after the two-plan render: %s" % (len(bad), bad[:40])


def test_config3_as_is_width4_view_1920x1080_mpfr320(ref_lib):
    """BASELINE configs[2] exactly as the reference's cmdline renders gallery/deep_embedded_julia.mdz: an old-style
    file loses its zoom there (SURVEY finding 4), so the render is a width-4 view centred on the file's centre, at
    the file's MPFR-320 -- SURVEY 8(d) config 3 (i).  ~2.5 G pixel-iterations, a sixth of the frame inside."""
    import golden_util as G
    from mdz_b200.mdzfile import view_from_settings
    meta, _, _ = G.load("deep_embedded_julia_asis")
    v, _ = view_from_settings(G.settings_of(meta), 1920, 1080, 1, bug_compatible=True, fixed_re=False)
    assert v.precision == 320 and v.mode == 1
    got = mdz_b200.render(v)
    assert (got == 0).any() and (got > 0).any()
    check_lines(ref_lib, v, got, sample_lines(1080, 10))


@pytest.mark.parametrize("mode", ["gmp", "mpfr"])
def test_config4_3840x2160_512bit_deep_zoom_with_minibrot(ref_lib, mode):
    """BASELINE configs[3] as SURVEY 8(d) specifies it: a 1e-120 wide view at 512 bits, depth 100000, on a
    minibrot nucleus, so that part of the frame reaches maxiter (tests/views.py config4m)."""
    v = config4m(3840, 2160, 100000, mode=mode, precision=512)
    got = mdz_b200.render(v)
    inside = (got == 0).mean()
    assert 0.015 < inside < 0.03                            # the copy of the set: ~2 % of the frame runs to depth
    assert got[got > 0].min() > 4000                        # and nothing escapes early, 1e-120 deep
    check_lines(ref_lib, v, got, [0, 700, 1079, 2159])      # line 700 crosses the minibrot
    if mode == "mpfr":
        p = mdz_b200.Plan(v, 0)
        p.set_cycle_detection(True)
        assert np.array_equal(p.run(), got)
        p.close()


def test_config4_round1_view_every_pixel_escapes(ref_lib):
    """Round 1's config-4 view (on the Misiurewicz point M(23,2), no interior), kept as a second deep view."""
    v = config4(3840, 2160, 100000, mode="mpfr", precision=512)
    got = mdz_b200.render(v)
    assert got.min() > 10000 and got.max() < 100000
    check_lines(ref_lib, v, got, [0, 1079, 2159])


def host_rgb(view, raw, palette, pal_offset, scale, interpolate):
    lib = portpath.load()
    pal = np.zeros(256, dtype=np.uint32)
    pal[:len(palette)] = palette
    out = np.zeros((view.user_height, view.user_width), dtype=np.uint32)
    lib.oracle_do_anti_aliasing(raw.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p),
                                view.user_width, view.aa_factor, 0, view.user_height,
                                C.c_double(scale), int(interpolate), pal.ctypes.data_as(C.c_void_p),
                                len(palette), pal_offset)
    return out


@pytest.mark.parametrize("fractal", [BURNING_SHIP, GENERALIZED_CELTIC], ids=["burning_ship", "generalized_celtic"])
def test_config5_7680x4320_aa3_with_palette_cycling(ref_lib, fractal):
    v = config5(fractal)                                    # 23040 x 12960 supersamples
    rng = np.random.RandomState(5)
    palette = (rng.randint(0, 256, 256) | (rng.randint(0, 256, 256) << 8) | (rng.randint(0, 256, 256) << 16)).astype(np.uint32)
    scale, interpolate = 0.37, True
    plan = mdz_b200.Plan(v, 0)
    plan.set_colour(palette, 0, scale, interpolate)
    raw = plan.run()
    rgb = plan.fetch_rgb()
    assert (raw == 0).any() and (raw > 0).any()
    check_lines(ref_lib, v, raw, sample_lines(v.real_height, 48))
    # fused epilogue == the reference's do_anti_aliasing restated on the host, from the same counts
    assert np.array_equal(rgb, host_rgb(v, raw, palette, 0, scale, interpolate))
    # palette cycling: recolour from the resident counts, no recompute
    for off in (1, 128, 255):
        plan.set_colour(palette, off, scale, interpolate)
        plan.recolour()
        assert np.array_equal(plan.fetch_rgb(), host_rgb(v, raw, palette, off, scale, interpolate)), off
    plan.close()
