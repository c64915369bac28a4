"""GPU parity: iteration counts from libmdzcuda (through its C ABI) against the
unmodified reference hot path (oracle/_ref/libmdzref.so) on the same views.
Bit-exact for the long double and MPFR modes."""
import os

import numpy as np
import pytest

import mdz_b200
from mdz_b200 import (FAMILY_JULIA, MANDELBROT, BURNING_SHIP, GENERALIZED_CELTIC, VARIANT)
from refpath import ref_render
from views import make_view, config2, config4m, SEAHORSE, deep_embedded_julia, honeytrace, gmp_close_path_view

pytestmark = pytest.mark.gpu


def check(view, ref_lib, devices=(0,)):
    got = mdz_b200.render(view, devices)
    want, _ = ref_render(ref_lib, view)
    bad = np.argwhere(got != want)
    assert bad.size == 0, "%d mismatching pixels, first %s: got %d want %d" % (
        len(bad), bad[0], got[tuple(bad[0])], want[tuple(bad[0])])
    return got


def test_config2_small_long_double(ref_lib):
    raw = check(config2(320, 180, 2000), ref_lib)
    assert (raw == 0).any() and (raw > 0).any()


@pytest.mark.parametrize("fractal", [MANDELBROT, BURNING_SHIP, GENERALIZED_CELTIC, VARIANT])
def test_long_double_fractals(ref_lib, fractal):
    check(make_view("-0.5", "-0.3", "3.5", 160, 120, mode="ld", depth=500, fractal=fractal), ref_lib)


@pytest.mark.parametrize("prec", [80, 96, 128, 176, 184, 256, 320, 352, 384, 416, 448, 480, 512, 544, 640, 800, 1024])
def test_mpfr_precisions_seahorse(ref_lib, prec):
    check(make_view(SEAHORSE[0], SEAHORSE[1], "1e-9", 96, 72, precision=prec, depth=1500), ref_lib)


# Above 1024 bits a pixel is rendered by a group of 8, 16 or 32 lanes (coop_kernel.cuh: limbs split over the lanes,
# shuffle products, ballot carries).  The reference accepts any precision from 80 bits up (src/image_info.c:535);
# kernels are instantiated to 8192 bits.  Precisions that fill the lanes' 1024 K bits and that do not.
@pytest.mark.parametrize("prec", [1025, 1100, 1536, 1537, 2048, 3000, 4096, 6144, 8192])
def test_mpfr_wide_precisions_warp_per_pixel(ref_lib, prec):
    w, h = (64, 48) if prec <= 4096 else (40, 30)
    v = make_view(SEAHORSE[0], SEAHORSE[1], "1e-9", w, h, precision=prec, depth=1500)
    p = mdz_b200.Plan(v, 0)
    assert p.kernel_info()["lanes_per_pixel"] == (8 if prec <= 2048 else 16 if prec <= 4096 else 32) and p.kernel_info()["limbs"] == (prec + 31) // 32
    p.close()
    raw = check(v, ref_lib)
    assert (raw > 0).any()


@pytest.mark.parametrize("fractal", [MANDELBROT, BURNING_SHIP, GENERALIZED_CELTIC, VARIANT])
def test_mpfr_2048_fractals_julia_and_antialias(ref_lib, fractal):
    check(make_view("-0.5", "-0.3", "3.5", 64, 48, precision=2048, depth=200, fractal=fractal), ref_lib)
    check(make_view("0", "0", "3.2", 48, 36, precision=2048, depth=150, fractal=fractal, aa=2,
                    family=FAMILY_JULIA, julia=("-0.8", "0.156")), ref_lib)


def test_mpfr_2048_next_to_a_minibrot(ref_lib):
    """The orbit of every pixel returns to ~0 once per period: 200 cancelled bits in two additions, then 400-bit
    exponent gaps (tests/views.py MINIBROT120) -- and part of the frame runs to depth."""
    from views import MINIBROT120
    v = make_view(MINIBROT120[0], MINIBROT120[1], "1e-120", 32, 18, precision=2048, depth=7000)
    raw = check(v, ref_lib)
    assert (raw == 0).any() and (raw > 0).any()


@pytest.mark.parametrize("fractal", [MANDELBROT, BURNING_SHIP, GENERALIZED_CELTIC, VARIANT])
@pytest.mark.parametrize("prec", [80, 128, 320])
def test_mpfr_fractals(ref_lib, fractal, prec):
    check(make_view("-0.5", "-0.3", "3.5", 96, 72, precision=prec, depth=300, fractal=fractal), ref_lib)


@pytest.mark.parametrize("mode,prec", [("ld", 64), ("mpfr", 96), ("mpfr", 256)])
def test_julia(ref_lib, mode, prec):
    check(make_view("0", "0", "3.2", 120, 90, mode=mode, precision=prec, depth=400,
                    family=FAMILY_JULIA, julia=("-0.8", "0.156")), ref_lib)


def test_julia_c_zero_squares_to_underflow(ref_lib):
    # z -> z^2 with c = 0: exponents double every step (DESIGN.md "exponent range")
    check(make_view("0", "0", "3.0", 64, 48, mode="mpfr", precision=96, depth=200,
                    family=FAMILY_JULIA, julia=("0", "0")), ref_lib)


def test_antialias_bands_and_odd_sizes(ref_lib):
    check(make_view("-0.7", "0.0", "3.0", 50, 37, precision=80, depth=200, aa=3), ref_lib)


def test_deep_embedded_julia_320(ref_lib):
    raw = check(deep_embedded_julia(64, 48), ref_lib)
    assert raw.min() > 0


def test_honeytrace_176(ref_lib):
    raw = check(honeytrace(48, 36), ref_lib)
    assert raw.min() >= 3000          # SURVEY 8c known answer: min 3521 at 96x72


def test_zoomed_out_huge_view(ref_lib):
    check(make_view("0", "0", "1e12", 40, 30, precision=128, depth=50), ref_lib)


def test_band_partition_equals_whole(ref_lib):
    v = make_view("-0.5", "0", "3", 64, 48, precision=96, depth=200, aa=2)
    whole = mdz_b200.render(v)
    parts = np.full_like(whole, -1)
    for first in range(3):
        p = mdz_b200.Plan(v, 0, first, 3)
        p.launch(); p.fetch(parts); p.close()
    assert (parts == whole).all()


# ---- GMP mpf mode (fractal_gmp_calculate_line, fractal.c:260-397) ----------------------
def check_gmp(view, ref_lib, threads=None):
    """North-star bar for GMP mode: >= 99.9 % identical, every mismatch reported.
    The observed result is 100 %, so this asserts exact equality and prints the
    mismatch set if that ever stops being true."""
    got = mdz_b200.render(view)
    want, _ = ref_render(ref_lib, view, threads)
    bad = np.argwhere(got != want)
    if bad.size:
        for y, x in bad[:20]:
            print("mismatch at line %d px %d: cuda %d reference %d" % (y, x, got[y, x], want[y, x]))
    assert bad.size == 0, "%d of %d pixels differ (%.4f %%)" % (len(bad), got.size, 100.0 * len(bad) / got.size)
    return got


@pytest.mark.parametrize("prec", [80, 128, 200, 256, 320, 512])
def test_gmp_precisions_seahorse(ref_lib, prec):
    check_gmp(make_view(SEAHORSE[0], SEAHORSE[1], "1e-9", 96, 72, mode="gmp", precision=prec, depth=1500), ref_lib)


@pytest.mark.parametrize("fractal", [MANDELBROT, BURNING_SHIP, GENERALIZED_CELTIC, VARIANT])
@pytest.mark.parametrize("prec", [128, 320])
def test_gmp_fractals(ref_lib, fractal, prec):
    check_gmp(make_view("-0.5", "-0.3", "3.5", 96, 72, mode="gmp", precision=prec, depth=300, fractal=fractal), ref_lib)


# One thread per pixel to 16 limbs (896 bits; mpf_fast.cuh), above that mpf values are held by a group of 16 or 32
# lanes (coop_mpf.cuh): the four shapes at precisions that fill them (P + 2 limbs of T K / 2) and that do not; the
# reference takes any precision (src/image_info.c:535).
@pytest.mark.parametrize("prec", [513, 600, 704, 768, 832, 896, 897, 1024, 1344, 1345, 1856, 1857, 2048, 3904, 4096, 5952, 6000, 8000])
def test_gmp_wide_precisions_lane_groups(ref_lib, prec):
    w, h = (64, 48) if prec <= 4096 else (40, 30)
    v = make_view(SEAHORSE[0], SEAHORSE[1], "1e-9", w, h, mode="gmp", precision=prec, depth=1500)
    p = mdz_b200.Plan(v, 0)
    ki = p.kernel_info()
    p.close()
    assert ki["lanes_per_pixel"] == (1 if prec <= 896 else 8 if prec <= 1856 else 16 if prec <= 3904 else 32) and ki["limbs"] == 2 * ((prec + 127) // 64 + 1)
    raw = check_gmp(v, ref_lib)
    assert (raw > 0).any()


@pytest.mark.parametrize("fractal", [MANDELBROT, BURNING_SHIP, GENERALIZED_CELTIC, VARIANT])
def test_gmp_1024_fractals_julia_and_antialias(ref_lib, fractal):
    check_gmp(make_view("-0.5", "-0.3", "3.5", 64, 48, mode="gmp", precision=1024, depth=300, fractal=fractal), ref_lib)
    check_gmp(make_view("0", "0", "3.2", 48, 36, mode="gmp", precision=1024, depth=300, fractal=fractal, aa=2,
                        family=FAMILY_JULIA, julia=("-0.8", "0.156")), ref_lib, threads=1)


def test_gmp_1024_real_axis_and_minibrot(ref_lib):
    check_gmp(make_view("-0.75", "0.0", "2.5", 48, 36, mode="gmp", precision=1024, depth=400), ref_lib)
    raw = check_gmp(config4m(32, 18, 7000, mode="gmp", precision=1024), ref_lib)
    assert (raw == 0).any() and (raw > 0).any()


@pytest.mark.parametrize("prec", [128, 1024, 1800, 2048, 4096])
def test_gmp_close_subtraction_in_every_pixel(ref_lib, prec):
    """tests/views.py gmp_close_path_view: the first wre2 - wim2 of every pixel is GMP's one-limb-gap subtraction,
    which the lane-group kernels hand to one lane over the shared-memory strip (coop_mpf.cuh cg_sub_close) and the
    one-thread kernels to an out-of-line function -- paths no ordinary view reaches.  More pixels than groups, so
    every strip is used again after it.  (tests/test_coop_vs_gmp.py checks on the CPU that these pixels do take it.)"""
    w, h = (128, 96) if prec <= 2048 else (64, 48)
    raw = check_gmp(gmp_close_path_view(prec, w, h, depth=200), ref_lib, threads=1)
    assert (raw > 0).all()


def test_gmp_beyond_the_kernels_is_refused():
    v = make_view(SEAHORSE[0], SEAHORSE[1], "1e-9", 16, 12, mode="gmp", precision=8001, depth=100)
    assert not mdz_b200.view_supported(v)
    with pytest.raises(mdz_b200.MdzCudaError):
        mdz_b200.Plan(v, 0)


def test_gmp_julia_one_digit_constant(ref_lib):
    # the reference converts the Julia constant per pixel through "%.Re" (fractal.c:341-342),
    # one significant digit on MPFR 4: -0.8 stays -0.8, 0.156 becomes 0.2.  That conversion
    # goes through a static buffer shared by all workers (my_mpfr_to_str.c:6), so the
    # reference itself is only deterministic with ONE thread in this mode: two 8-thread
    # runs of the unmodified reference differ from each other in ~20 % of the pixels.
    check_gmp(make_view("0", "0", "3.2", 96, 72, mode="gmp", precision=128, depth=300,
                        family=FAMILY_JULIA, julia=("-0.8", "0.156")), ref_lib, threads=1)


def test_gmp_real_axis_and_aa(ref_lib):
    check_gmp(make_view("-0.75", "0.0", "2.5", 48, 36, mode="gmp", precision=128, depth=400, aa=2), ref_lib)


def test_gmp_deep_zoom_512(ref_lib):
    # BASELINE configs[3] in miniature: 1e-120 wide window, 512-bit mpf
    cx = "-1.7400623825793399052208441670658256382966417204361718668798624184611829" \
         "1966513096674796282803698889934592955622845248"
    check_gmp(make_view(cx, "0.0281753397792110489924115211443195096875390767429906085704013095958801" 
                        "743240920186385400814658560553615695084486774077", "1e-120", 48, 27,
                        mode="gmp", precision=512, depth=3000), ref_lib)


def test_plan_run_delivers_what_fetch_returns_and_pool_reuse(ref_lib):
    """mdzcuda_plan_run (bands copied to the host while the kernel runs) against
    launch + fetch and against the reference; plans created after a destroy draw their
    buffers from the pool (stale contents must not leak into a result), also after
    mdzcuda_trim, for strided bands and with several plans alive at once."""
    from mdz_b200 import _native
    v = config2(256, 144, 1500)
    want, _ = ref_render(ref_lib, v)
    for rep in range(4):
        p = mdz_b200.Plan(v, 0)
        got = p.run()
        p.launch(); again = p.fetch()
        p.close()
        assert np.array_equal(got, want) and np.array_equal(again, want)
        if rep == 1:
            _native.lib.mdzcuda_trim()
    # a different view of the same size right after: same pooled blocks, new contents
    v2 = make_view("-0.1", "0.8", "0.5", 256, 144, mode="ld", depth=700)
    want2, _ = ref_render(ref_lib, v2)
    assert np.array_equal(mdz_b200.render(v2), want2)
    plans = [mdz_b200.Plan(v2, 0, first, 3) for first in range(3)]
    out = np.full_like(want2, -1)
    for p in plans:
        p.run(out)
    for p in plans:
        p.close()
    assert np.array_equal(out, want2)


CYCLE_VIEWS = [
    ("ld cfg2", lambda: config2(320, 180, 3000)),
    ("ld ship", lambda: make_view("-0.5", "-0.3", "3.5", 160, 120, mode="ld", depth=1500, fractal=BURNING_SHIP)),
    ("ld celtic", lambda: make_view("-0.5", "-0.3", "3.5", 160, 120, mode="ld", depth=1500, fractal=GENERALIZED_CELTIC)),
    ("ld hybrid (parity of the period matters)", lambda: make_view("-0.5", "-0.3", "3.5", 160, 120, mode="ld", depth=1500, fractal=VARIANT)),
    ("ld julia", lambda: make_view("0", "0", "3.2", 120, 90, mode="ld", depth=2000, family=FAMILY_JULIA, julia=("-0.12", "0.74"))),
    ("mpfr80 interior", lambda: make_view("-0.2", "0.1", "1.5", 96, 72, precision=80, depth=2000)),
    ("mpfr128 hybrid", lambda: make_view("-0.5", "-0.3", "3.5", 96, 72, precision=128, depth=800, fractal=VARIANT)),
    ("mpfr320 bulb", lambda: make_view("-1.0", "0.0", "0.6", 64, 48, precision=320, depth=1500)),
    ("mpfr512 cardioid", lambda: make_view("-0.3", "0.2", "0.8", 48, 36, precision=512, depth=1200)),
    ("mpfr1024", lambda: make_view("-0.3", "0.2", "0.8", 24, 18, precision=1024, depth=600)),
]


@pytest.mark.parametrize("name,mk", CYCLE_VIEWS, ids=[c[0] for c in CYCLE_VIEWS])
def test_cycle_detection_changes_nothing(ref_lib, name, mk):
    """The exact periodicity check (mdzcuda_plan_set_cycle_detection) finishes interior
    pixels early; raw_data must equal the reference's, which iterates to depth."""
    v = mk()
    want, _ = ref_render(ref_lib, v)
    assert (want == 0).sum() > 50, "view has no interior: the test would prove nothing"
    for spec in (1, 0):
        p = mdz_b200.Plan(v, 0)
        p.set_cycle_detection(True)
        if not spec:
            p.tune(-16, 0)
        got = p.run()
        p.close()
        bad = np.argwhere(got != want)
        assert bad.size == 0, "%s spec=%d: %d mismatching pixels, first %s: got %d want %d" % (
            name, spec, len(bad), bad[0], got[tuple(bad[0])], want[tuple(bad[0])])


@pytest.mark.parametrize("name,mk", CYCLE_VIEWS, ids=[c[0] for c in CYCLE_VIEWS])
def test_tail_compaction_changes_nothing(ref_lib, name, mk):
    """Parking (mdzcuda_plan_set_parking): the first launch stops when the queue runs dry and
    writes the live pixels' state to HBM, the second resumes them.  Forced on -- these views are
    smaller than the grid, so everything is parked after its first chunk -- raw_data must equal the
    reference's, with and without the periodicity check, whole and as one rank's share."""
    v = mk()
    want, _ = ref_render(ref_lib, v)
    for cyc in (False, True):
        p = mdz_b200.Plan(v, 0)
        p.set_parking(1)
        p.set_cycle_detection(cyc)
        got = p.run()
        p.close()
        bad = np.argwhere(got != want)
        assert bad.size == 0, "%s cyc=%d: %d mismatching pixels, first %s: got %d want %d" % (
            name, cyc, len(bad), bad[0], got[tuple(bad[0])], want[tuple(bad[0])])
    out = np.full_like(want, -1)
    plans = [mdz_b200.Plan(v, 0, first, 2) for first in range(2)]
    for p in plans:
        p.set_parking(1)
        p.run(out)
        p.close()
    assert np.array_equal(out, want)


def test_tail_compaction_is_reported_and_limited_to_four_limbs():
    """mdzcuda_plan_kernels_launched (bench.py's gpu_launches): a parked render is three launches
    (phase 0, the ordering pass, phase 1), any other one.  Parking is compiled into the kernels of
    up to four limbs only (escape_params.cuh kParkMaxLimbs); asking for it on a wider one is a no-op."""
    for kw, park, want in ((dict(mode="ld"), 1, 3), (dict(mode="ld"), 0, 1),
                           (dict(precision=128), 1, 3), (dict(precision=320), 1, 1)):
        p = mdz_b200.Plan(make_view("-0.7", "0.1", "2.5", 96, 72, depth=300, **kw), 0)
        p.set_parking(park)
        assert p.kernels_launched() == 0
        p.run()
        assert p.kernels_launched() == want, (kw, park, p.kernels_launched())
        p.run()
        assert p.kernels_launched() == 2 * want
        p.close()


LEVEL2_VIEWS = [
    # long double mode, pixels on which the plain fast iteration declines for ever: the
    # column x = 0 and the row y = 0 (zero operands), rows next to the real axis (100-bit gaps), fixed
    # points on a diagonal (wre^2 - wim^2 cancels), c = -1 (wre is 0 every other iteration), c = 1/4 + i/8
    # (wre * wim rounds up to a power of two at the fixed point)
    ("axes through the image", lambda: make_view("0", "0", "3.2", 128, 96, mode="ld", depth=3000)),
    ("cfg2 grid (dyadic c)", lambda: config2(480, 270, 3000)),
    ("next to the real axis", lambda: make_view("-0.5", "1e-17", "3.0", 96, 65, mode="ld", depth=2500)),
    ("ship axes", lambda: make_view("0", "0", "3.2", 96, 72, mode="ld", depth=1500, fractal=BURNING_SHIP)),
    ("celtic axes", lambda: make_view("0", "0", "3.2", 96, 72, mode="ld", depth=1500, fractal=GENERALIZED_CELTIC)),
    ("hybrid axes", lambda: make_view("0", "0", "3.2", 96, 72, mode="ld", depth=1500, fractal=VARIANT)),
    ("julia c = -1", lambda: make_view("0", "0", "3.2", 96, 72, mode="ld", depth=2000, family=FAMILY_JULIA, julia=("-1", "0"))),
    ("julia c = 1/4 + i/8", lambda: make_view("0", "0", "3.2", 96, 72, mode="ld", depth=2000, family=FAMILY_JULIA, julia=("0.25", "0.125"))),
]


@pytest.mark.parametrize("name,mk", LEVEL2_VIEWS, ids=[c[0] for c in LEVEL2_VIEWS])
def test_long_double_level_two_pixels_match_reference(ref_lib, name, mk):
    """Views dense in the operands only level 2 of the long double iteration takes
    (ld64_step.cuh add64_core<true> / mul64_core<true>); the warps get there by themselves."""
    v = mk()
    want, _ = ref_render(ref_lib, v)
    for park in (0, 1):
        p = mdz_b200.Plan(v, 0)
        p.set_parking(park)
        got = p.run()
        p.close()
        bad = np.argwhere(got != want)
        assert bad.size == 0, "%s park=%d: %d mismatching pixels, first %s: got %d want %d" % (
            name, park, len(bad), bad[0], got[tuple(bad[0])], want[tuple(bad[0])])


def test_delivery_never_runs_ahead_of_the_kernel(ref_lib):
    """Regression: band flags left in recycled memory (by an earlier plan, or by the same
    plan's previous launch) must not be mistaken for this launch's -- once per ~300 strided
    renders a frame used to be delivered before it was rendered.  Flags now carry the launch
    generation; with the poison hook (conftest.py) an early delivery cannot go unnoticed."""
    v = make_view("-0.743", "0.131", "0.02", 640, 360, precision=128, depth=3000)
    want, _ = ref_render(ref_lib, v)
    p = mdz_b200.Plan(v, 0)
    for rep in range(6):                       # relaunching one plan: its own flags are stale each time
        assert np.array_equal(p.run(), want), rep
        assert p.bands_done() == p.bands_total()
    p.close()
    for rep in range(60):                      # plan after plan on recycled buffers, strided
        first = rep % 2
        out = np.full_like(want, -1)
        q = mdz_b200.Plan(v, 0, first, 2)
        q.run(out)
        q.close()
        assert np.array_equal(out[first::2], want[first::2]), rep
        assert (out[1 - first::2] == -1).all()


def test_cycle_detection_is_invisible_on_random_views():
    """The exact periodicity check is on behind rth_* (rth.cpp), the one place where the drop-in does different work
    from the reference by default: a property test over random views -- every fractal, both families, long double
    and three MPFR precisions, with and without anti-aliasing -- that raw_data is identical with it on and off."""
    import random
    rng = random.Random(20261017)
    modes = [("ld", 64), ("mpfr", 96), ("mpfr", 128), ("mpfr", 256)]
    for k in range(48):
        mode, prec = modes[k % 4]
        fractal = [MANDELBROT, BURNING_SHIP, GENERALIZED_CELTIC, VARIANT][(k // 4) % 4]
        julia = rng.random() < 0.3
        size = 10 ** rng.uniform(-2.5, 0.6)
        # centres that keep part of the set in view: near the boundary of the main body
        cx, cy = rng.uniform(-1.6, 0.4), rng.uniform(-0.9, 0.9)
        kw = dict(mode=mode, precision=max(prec, 80) if mode == "ld" else prec, depth=rng.choice([300, 1000, 3000]),
                  aa=rng.choice([1, 1, 2]), fractal=fractal)
        if julia:
            kw.update(family=FAMILY_JULIA, julia=("%.6f" % rng.uniform(-1.2, 0.4), "%.6f" % rng.uniform(-0.8, 0.8)))
            cx, cy, size = rng.uniform(-0.3, 0.3), rng.uniform(-0.3, 0.3), 10 ** rng.uniform(-0.5, 0.5)
        v = make_view("%.9f" % cx, "%.9f" % cy, "%.9g" % size, 80, 60, **kw)
        a = mdz_b200.Plan(v, 0)
        plain = a.run()
        a.close()
        b = mdz_b200.Plan(v, 0)
        b.set_cycle_detection(True)
        checked = b.run()
        b.close()
        assert np.array_equal(plain, checked), "view %d (%s p%d fractal %d julia %s): %d pixels differ" % (
            k, mode, prec, fractal, julia, int((plain != checked).sum()))


@pytest.mark.parametrize("w,h,aa,first,stride", [(320, 200, 1, 0, 1), (50, 37, 3, 0, 1), (384, 216, 2, 1, 3), (3840, 40, 1, 0, 1), (97, 61, 1, 2, 5)])
def test_centre_out_order_changes_nothing(w, h, aa, first, stride):
    """mdzcuda_plan_set_order: the queue takes the plan's bands cut into tiles, by distance from the image centre,
    instead of in raster order.  Scheduling only -- raw_data must be the same, for whole images, for one rank's
    interleaved share, for widths no tile count divides."""
    v = make_view("-0.6", "0.1", "2.8", w, h, precision=96, depth=400, aa=aa)
    a = mdz_b200.Plan(v, 0, first, stride)
    plain = a.run()
    a.close()
    b = mdz_b200.Plan(v, 0, first, stride)
    b.set_order(centre_out=True)
    ordered = b.run()
    again = b.run()          # the same plan launched twice keeps its table
    b.close()
    assert np.array_equal(plain, ordered) and np.array_equal(plain, again)
