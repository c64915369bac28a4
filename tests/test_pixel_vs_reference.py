"""CPU-side parity of the kernel's per-pixel logic: pixel_init / pixel_step /
pixel_step_spec + fall-back (mdz_b200/csrc/escape_step.cuh, compiled for the host
by tests/host_emu) against the UNMODIFIED reference per-pixel functions
frac_*_mpfr (oracle/_ref/libmdzref.so, reference src/frac_*.c) on sampled pixels
of several views, with the coordinates built exactly as fractal.c:167-188 does."""
import ctypes as C

import pytest

from mdz_b200 import MANDELBROT, BURNING_SHIP, GENERALIZED_CELTIC, VARIANT
from mdz_b200.mp import Mpfr, MpfrStruct, mpfr, nlimbs64
from views import make_view, deep_embedded_julia, honeytrace, SEAHORSE

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
]


@pytest.mark.parametrize("name,mk,count", CASES, ids=[c[0] for c in CASES])
@pytest.mark.parametrize("spec", [0, 1], ids=["general", "speculative"])
def test_pixels_match_reference(emu, ref_lib, name, mk, count, spec):
    view = mk()
    W, H = view.real_width, view.real_height
    for k in range(count):
        ix, line = (k * 37 + 5) % W, (k * 53 + (H // 2 if k % 4 == 0 else 3)) % H
        x, y = coords(view, ix, line)
        assert emu_pixel(emu, view, x, y, spec) == ref_pixel(ref_lib, view, x, y), (name, ix, line)
