"""CPU-side parity of the kernel's per-pixel logic: pixel_init / pixel_step /
pixel_step_spec + fall-back (mdz_b200/csrc/escape_step.cuh, compiled for the host
by tests/host_emu) against the UNMODIFIED reference per-pixel functions
frac_*_mpfr (oracle/_ref/libmdzref.so, reference src/frac_*.c) on sampled pixels
of several views, with the coordinates built exactly as fractal.c:167-188 does."""
import ctypes as C

import pytest

from mdz_b200 import MANDELBROT, BURNING_SHIP, GENERALIZED_CELTIC, VARIANT
from mdz_b200.mp import Mpfr, MpfrStruct, mpfr, nlimbs64
from views import make_view, deep_embedded_julia, honeytrace, SEAHORSE, MINIBROT120

U = C.POINTER(C.c_uint64)
P = C.POINTER(MpfrStruct)
FRAC = {MANDELBROT: "frac_mandel_mpfr", BURNING_SHIP: "frac_burning_ship_mpfr",
        GENERALIZED_CELTIC: "frac_generalized_celtic_mpfr", VARIANT: "frac_variant_mpfr"}


def coords(view, ix, line):
    p, W = view.precision, view.real_width
    rw, xmin, width = Mpfr(p, W), Mpfr(p, view.xmin), Mpfr(p, view.width)
    t1, x, y = Mpfr(p), Mpfr(p), Mpfr(p)
    mpfr.mpfr_si_div(t1.ref, ix, rw.ref, 0)
    mpfr.mpfr_mul(x.ref, t1.ref, width.ref, 0)
    mpfr.mpfr_add(x.ref, x.ref, xmin.ref, 0)
    mpfr.mpfr_div(t1.ref, width.ref, rw.ref, 0)
    mpfr.mpfr_mul_si(t1.ref, t1.ref, line, 0)
    mpfr.mpfr_sub(y.ref, view.ymax.ref, t1.ref, 0)
    return x, y


def ref_pixel(ref, view, x, y):
    p = view.precision
    fn = getattr(ref, FRAC[view.fractal])
    fn.restype = C.c_long
    fn.argtypes = [C.c_long] + [P] * 8
    bail = Mpfr(p, 4)
    wim, wre, cim, cre = Mpfr(p, y), Mpfr(p, x), Mpfr(p, y), Mpfr(p, x)
    wim2, wre2, t1 = Mpfr(p), Mpfr(p), Mpfr(p)
    mpfr.mpfr_mul(wim2.ref, y.ref, y.ref, 0)
    mpfr.mpfr_mul(wre2.ref, x.ref, x.ref, 0)
    return fn(view.depth, bail.ptr, wim.ptr, wre.ptr, cim.ptr, cre.ptr, wim2.ptr, wre2.ptr, t1.ptr)


def emu_pixel(emu, view, x, y, spec):
    n = nlimbs64(view.precision)
    args = []
    for v in (x, y, x, y):
        s, e, _ = v.parts()
        args += [(C.c_uint64 * n)(*v.limbs()), s, e]
    return emu.emu_pixel(view.precision, view.fractal, view.depth, spec, *args)


@pytest.fixture(scope="module")
def emu(emu_lib):
    emu_lib.emu_pixel.restype = C.c_long
    emu_lib.emu_pixel.argtypes = [C.c_long, C.c_int, C.c_long, C.c_int] + [U, C.c_int, C.c_long] * 4
    return emu_lib


CASES = [
    ("cfg2 p64", lambda: make_view("-0.5", "0.0", "4.0", 192, 108, precision=64, depth=2000), 160),
    ("seahorse p80", lambda: make_view(SEAHORSE[0], SEAHORSE[1], "1e-9", 96, 72, precision=80, depth=1500), 60),
    ("seahorse p512", lambda: make_view(SEAHORSE[0], SEAHORSE[1], "1e-9", 96, 72, precision=512, depth=1500), 16),
    ("deep_embedded_julia p320", lambda: deep_embedded_julia(64, 48), 20),
    ("honeytrace p176", lambda: honeytrace(48, 36, depth=6000), 12),
    ("ship p128", lambda: make_view("-0.5", "-0.3", "3.5", 96, 72, precision=128, depth=300, fractal=BURNING_SHIP), 120),
    ("celtic p96", lambda: make_view("-0.5", "-0.3", "3.5", 96, 72, precision=96, depth=300, fractal=GENERALIZED_CELTIC), 120),
    ("hybrid p184", lambda: make_view("-0.5", "-0.3", "3.5", 96, 72, precision=184, depth=300, fractal=VARIANT), 120),
    ("real axis p96", lambda: make_view("-0.75", "0.0", "2.5", 64, 48, precision=96, depth=500), 64),
    # the row y = 0 takes the real-orbit shortcut of pixel_step (escape_step.cuh) in every fractal type
    ("real axis p64", lambda: make_view("-0.75", "0.0", "2.5", 64, 48, precision=64, depth=2000), 64),
    ("real axis ship p64", lambda: make_view("-0.5", "0.0", "3.5", 64, 48, precision=64, depth=500, fractal=BURNING_SHIP), 64),
    ("real axis celtic p128", lambda: make_view("-0.5", "0.0", "3.5", 64, 48, precision=128, depth=500, fractal=GENERALIZED_CELTIC), 64),
    ("real axis hybrid p320", lambda: make_view("-0.5", "0.0", "3.5", 64, 48, precision=320, depth=500, fractal=VARIANT), 64),
    # next to the period-707 minibrot every orbit returns to ~0 once per period: 200 cancelled bits, then 400-bit gaps
    ("minibrot p512", lambda: make_view(MINIBROT120[0], MINIBROT120[1], "1e-120", 64, 36, precision=512, depth=2200), 6),
    ("minibrot p448 ship", lambda: make_view(MINIBROT120[0], MINIBROT120[1], "1e-120", 64, 36, precision=448, depth=1500, fractal=BURNING_SHIP), 6),
    ("seahorse p384 celtic", lambda: make_view(SEAHORSE[0], SEAHORSE[1], "1e-9", 96, 72, precision=384, depth=600, fractal=GENERALIZED_CELTIC), 16),
    ("seahorse p352 hybrid", lambda: make_view(SEAHORSE[0], SEAHORSE[1], "1e-9", 96, 72, precision=352, depth=600, fractal=VARIANT), 16),
]


@pytest.mark.parametrize("name,mk,count", CASES, ids=[c[0] for c in CASES])
@pytest.mark.parametrize("spec", [0, 1, 2, 3, 4, 5], ids=["general", "speculative", "speculative-smem-checkpoint",
                                                     "speculative-wide", "speculative-wide-smem-checkpoint", "hybrid"])
def test_pixels_match_reference(emu, ref_lib, name, mk, count, spec):
    view = mk()
    W, H = view.real_width, view.real_height
    for k in range(count):
        ix, line = (k * 37 + 5) % W, (k * 53 + (H // 2 if k % 4 == 0 else 3)) % H
        x, y = coords(view, ix, line)
        assert emu_pixel(emu, view, x, y, spec) == ref_pixel(ref_lib, view, x, y), (name, ix, line)


@pytest.mark.parametrize("fractal", [MANDELBROT, BURNING_SHIP, GENERALIZED_CELTIC, VARIANT])
@pytest.mark.parametrize("prec", [64, 96, 320])
def test_real_axis_shortcut_matches_reference(emu, ref_lib, fractal, prec):
    """y exactly 0: pixel_step's real-orbit shortcut (escape_step.cuh: wre = RN(wre2 + c_re),
    wre2 = RN(wre^2), wim stays 0) against the reference's nine-call loop body on points of the
    real axis -- interior, the chaotic part of the antenna, the tip at -2, and outside."""
    view = make_view("-0.75", "0.0", "2.5", 64, 48, precision=prec, depth=3000, fractal=fractal)
    zero = Mpfr(prec, 0)
    for xs in ("-2", "-1.9999", "-1.75", "-1.5436890126920763", "-1.25", "-1", "-0.75", "-0.1", "0", "0.2",
               "0.25", "0.2500001", "0.3", "1", "2", "-2.0000001"):
        x = Mpfr(prec, xs)
        want = ref_pixel(ref_lib, view, x, zero)
        for spec in (0, 1, 3, 5):
            assert emu_pixel(emu, view, x, zero, spec) == want, (fractal, prec, xs, spec)


# ---- GMP mpf mode: mpf_sf.cuh gmp_pixel_* vs the reference's frac_*_gmp -------------
from mdz_b200.mp import (MpfStruct, Mpf, mpf_init2, mpf_set, mpf_set_si, mpf_mul, mpf_add, mpf_sub,
                         mpf_div, mpf_ui_div, mpf_mul_ui)

GFRAC = {MANDELBROT: "frac_mandel_gmp", BURNING_SHIP: "frac_burning_ship_gmp",
         GENERALIZED_CELTIC: "frac_generalized_celtic_gmp", VARIANT: "frac_variant_gmp"}
GP = C.POINTER(MpfStruct)


def gmp_coords(view, ix, line):
    """fractal.c:299-328 with the real libgmp."""
    p = view.precision
    rw, xmin, width, t1, x, y = (Mpf(p) for _ in range(6))
    mpf_set_si(rw.ref, view.real_width)
    mpf_set(xmin.ref, view.gxmin.ref)
    mpf_set(width.ref, view.gwidth.ref)
    mpf_ui_div(t1.ref, ix, rw.ref)
    mpf_mul(x.ref, t1.ref, width.ref)
    mpf_add(x.ref, x.ref, xmin.ref)
    mpf_div(t1.ref, width.ref, rw.ref)
    mpf_mul_ui(t1.ref, t1.ref, line)
    mpf_sub(y.ref, view.gymax.ref, t1.ref)
    return x, y


def gmp_ref_pixel(ref, view, x, y):
    p = view.precision
    fn = getattr(ref, GFRAC[view.fractal])
    fn.restype = C.c_long
    fn.argtypes = [C.c_long] + [GP] * 8
    vals = [Mpf(p) for _ in range(8)]       # bail wim wre cim cre wim2 wre2 t1
    bail, wim, wre, cim, cre, wim2, wre2, t1 = vals
    mpf_set_si(bail.ref, 4)
    for dst, src in ((wim, y), (wre, x), (cim, y), (cre, x)):
        mpf_set(dst.ref, src.ref)
    mpf_mul(wim2.ref, y.ref, y.ref)
    mpf_mul(wre2.ref, x.ref, x.ref)
    return fn(view.depth, bail.ptr, wim.ptr, wre.ptr, cim.ptr, cre.ptr, wim2.ptr, wre2.ptr, t1.ptr)


def fixed_mpf(v, nl):
    sg, e, limbs = v.parts()
    out = [0] * nl
    for i, w in enumerate(limbs):
        out[nl - len(limbs) + i] = w
    return (C.c_uint64 * nl)(*out), e, sg


GCASES = [
    ("seahorse gmp 128", lambda: make_view(SEAHORSE[0], SEAHORSE[1], "1e-9", 96, 72, mode="gmp", precision=128, depth=1500), 60),
    ("seahorse gmp 512", lambda: make_view(SEAHORSE[0], SEAHORSE[1], "1e-9", 96, 72, mode="gmp", precision=512, depth=1500), 16),
    ("full set gmp 80", lambda: make_view("-0.5", "0.0", "4.0", 96, 72, mode="gmp", precision=80, depth=400), 200),
    ("ship gmp 256", lambda: make_view("-0.5", "-0.3", "3.5", 96, 72, mode="gmp", precision=256, depth=300, fractal=BURNING_SHIP), 100),
    ("celtic gmp 320", lambda: make_view("-0.5", "-0.3", "3.5", 96, 72, mode="gmp", precision=320, depth=300, fractal=GENERALIZED_CELTIC), 100),
    ("hybrid gmp 128", lambda: make_view("-0.5", "-0.3", "3.5", 96, 72, mode="gmp", precision=128, depth=300, fractal=VARIANT), 100),
    ("real axis gmp 128", lambda: make_view("-0.75", "0.0", "2.5", 64, 48, mode="gmp", precision=128, depth=500), 64),
    ("seahorse gmp 896", lambda: make_view(SEAHORSE[0], SEAHORSE[1], "1e-9", 96, 72, mode="gmp", precision=896, depth=1500), 8),
    ("celtic gmp 640", lambda: make_view("-0.5", "-0.3", "3.5", 96, 72, mode="gmp", precision=640, depth=300, fractal=GENERALIZED_CELTIC), 40),
]


@pytest.mark.parametrize("impl", ["emu_gmp_pixel", "emu_gmpf_pixel"], ids=["clear", "fast"])
@pytest.mark.parametrize("name,mk,count", GCASES, ids=[c[0] for c in GCASES])
def test_gmp_pixels_match_reference(emu_lib, ref_lib, name, mk, count, impl):
    fn = getattr(emu_lib, impl)
    fn.restype = C.c_long
    fn.argtypes = [C.c_int, C.c_int, C.c_long] + [U, C.c_long, C.c_int] * 4
    view = mk()
    nl = (max(53, view.precision) + 127) // 64 + 1
    W, H = view.real_width, view.real_height
    for k in range(count):
        ix, line = (k * 37 + 5) % W, (k * 53 + (H // 2 if k % 4 == 0 else 3)) % H
        x, y = gmp_coords(view, ix, line)
        args = []
        for v in (x, y, x, y):
            args += list(fixed_mpf(v, nl))
        got = fn(nl, view.fractal, view.depth, *args)
        assert got == gmp_ref_pixel(ref_lib, view, x, y), (name, ix, line)
