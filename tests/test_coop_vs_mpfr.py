"""Differential test of the warp-cooperative arithmetic (mdz_b200/csrc/coop_ops.cuh: one warp per value,
limbs split over the lanes, shuffle-broadcast products, ballot carry resolution -- the kernels for precisions
above 1024 bits), compiled for the host by tests/host_emu/coop_emu.cpp with a warp emulated as 32 lanes in
lock step, against the real libmpfr.so.6: mul, sqr, add, sub, the sign-specialised adds, the comparison
with 4 and the escape test; then whole pixels against the reference's own frac_*_mpfr functions
(oracle/_ref/libmdzref.so, reference src/frac_*.c).  Bit-exact (sign, exponent, every mantissa bit)."""
import ctypes as C
import os
import random
import subprocess

import pytest

from mdz_b200 import MANDELBROT, BURNING_SHIP, GENERALIZED_CELTIC, VARIANT
from mdz_b200.mp import Mpfr, mpfr, nlimbs64
from test_arith_vs_mpfr import rand_pair, rand_mant, mpfr_op, U64P
from test_pixel_vs_reference import coords, ref_pixel
from views import make_view, SEAHORSE, MINIBROT120

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
# (K limbs per lane, T lanes per value: T K limbs = 32 T K bits of working width), precisions that fill it and that do not;
# 8 x 6, 8 x 8, 16 x 8, 32 x 6, 32 x 8 are what the kernels use, the others exercise the same code at other shapes
CASES = [(4, 16, 2048), (4, 16, 1025), (4, 16, 1100), (4, 16, 2047), (4, 16, 1984), (8, 16, 4096), (8, 16, 2049), (8, 16, 3000),
         (6, 32, 6144), (6, 32, 5000), (8, 32, 8192), (8, 32, 7000), (2, 32, 2048), (4, 32, 4096), (2, 16, 1024), (2, 16, 700), (8, 8, 2048), (8, 8, 1537), (6, 8, 1536), (6, 8, 1025), (6, 8, 1100)]


@pytest.fixture(scope="module")
def coop():
    src = os.path.join(ROOT, "tests", "host_emu", "coop_emu.cpp")
    out = os.path.join(ROOT, "tests", "host_emu", "libcoopemu.so")
    deps = [src] + [os.path.join(ROOT, "mdz_b200", "csrc", f) for f in ("coop_ops.cuh", "limb_ops.cuh", "mpfr_sf.cuh", "mp_convert.h")]
    if not os.path.exists(out) or any(os.path.getmtime(d) > os.path.getmtime(out) for d in deps):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-o", out, src])
    lib = C.CDLL(out)
    lib.coop_binop.argtypes = [C.c_int, C.c_int, C.c_int, C.c_long, U64P, C.c_int, C.c_long, U64P, C.c_int, C.c_long, U64P,
                               C.POINTER(C.c_int), C.POINTER(C.c_long)]
    lib.coop_pixel.restype = C.c_long
    lib.coop_pixel.argtypes = [C.c_int, C.c_int, C.c_long, C.c_int, C.c_long] + [U64P, C.c_int, C.c_long] * 4
    return lib


def coop_op(lib, KT, op, prec, a, b):
    K, T = KT
    n = nlimbs64(prec)
    al, bl = (C.c_uint64 * n)(*a.limbs()), (C.c_uint64 * n)(*b.limbs())
    rl, rs, re_ = (C.c_uint64 * n)(), C.c_int(), C.c_long()
    sa, ea, _ = a.parts()
    sb, eb, _ = b.parts()
    assert lib.coop_binop(K, T, op, prec, al, sa, ea, bl, sb, eb, rl, C.byref(rs), C.byref(re_))
    assert rs.value != 99, "bits below the precision are set"
    assert rs.value != 98, "the groups of the warp disagree"
    if op in (6, 11):
        return rs.value
    if rs.value == 0:
        return (0, 0, 0)
    full = 0
    for i in range(n):
        full |= rl[i] << (64 * i)
    return (rs.value, re_.value, full >> (64 * n - prec))


@pytest.mark.parametrize("K,T,prec", CASES)
def test_ops_match_libmpfr(coop, K, T, prec):
    rng = random.Random(5000 + prec + T)
    K = (K, T)
    for _ in range(400 if K[0] * T <= 128 else 200):
        a, b = rand_pair(rng, prec)
        for op, name in ((0, "mul"), (1, "sqr"), (2, "add"), (3, "sub")):
            assert coop_op(coop, K, op, prec, a, b) == mpfr_op(name, prec, a, b), (name, a.parts(), b.parts())
        sa, ea, ma = a.parts()
        sb, eb, mb = b.parts()
        a2 = Mpfr(prec).set_parts(abs(sa), ea, ma)
        b2 = Mpfr(prec).set_parts(abs(sb), eb, mb)
        for op, name in ((4, "sub"), (5, "add")):
            assert coop_op(coop, K, op, prec, a2, b2) == mpfr_op(name, prec, a2, b2), (name, a2.parts(), b2.parts())


@pytest.mark.parametrize("K,prec", [((4, 16), 2048), ((4, 16), 1100), ((8, 16), 4096), ((2, 32), 2048)])
def test_structured_carries_across_lanes(coop, K, prec):
    """Operands that make carries / borrows run through many lanes: all-ones runs, single low bits, sums that
    wrap to a power of two, differences that cancel down to one bit."""
    rng = random.Random(99 + prec)
    top = 1 << (prec - 1)
    ones = (1 << prec) - 1
    mants = [top, ones, top | 1, ones ^ 1, ones ^ (1 << (prec // 2)), top | (1 << (prec // 2)), top | ((1 << (prec // 2)) - 1),
             ones ^ ((1 << 37) - 1), top | (1 << 32), top | (1 << 31), ones ^ (1 << 32)]
    for ma in mants:
        for mb in mants:
            for de in (0, 1, -1, 2, 31, 32, 33, 64, prec - 1, prec, prec + 1, prec + 2, -prec, 1023, 1024, 1025):
                a = Mpfr(prec).set_parts(1, 0, ma)
                b = Mpfr(prec).set_parts(rng.choice([1, -1]), de, mb)
                for op, name in ((0, "mul"), (2, "add"), (3, "sub")):
                    assert coop_op(coop, K, op, prec, a, b) == mpfr_op(name, prec, a, b), (name, a.parts(), b.parts())


def test_greater_than_4_and_escape(coop):
    K, prec = (4, 16), 2048
    four = Mpfr(prec, 4)
    rng = random.Random(7)
    for _ in range(400):
        a = Mpfr(prec).set_parts(rng.choice([1, 1, 1, -1, 0]), rng.randrange(1, 6), rand_mant(rng, prec))
        k = rng.randrange(10)
        if k == 0:
            a = Mpfr(prec).set_parts(1, 3, 1 << (prec - 1))            # exactly 4
        if k == 1:
            a = Mpfr(prec).set_parts(1, 3, (1 << (prec - 1)) | 1)      # 4 + ulp
        if k == 2:
            a = Mpfr(prec).set_parts(1, 3, (1 << (prec - 1)) | (1 << 64))
        assert coop_op(coop, K, 6, prec, a, a) == (1 if mpfr.mpfr_greater_p(a.ref, four.ref) else 0), a.parts()
        # RN(a + b) > 4 for non-negative a, b around 2
        b = Mpfr(prec).set_parts(1, rng.randrange(0, 4), rand_mant(rng, prec))
        a2 = Mpfr(prec).set_parts(1, rng.randrange(0, 4), rand_mant(rng, prec))
        s = Mpfr(prec)
        mpfr.mpfr_add(s.ref, a2.ref, b.ref, 0)
        assert coop_op(coop, K, 11, prec, a2, b) == (1 if mpfr.mpfr_greater_p(s.ref, four.ref) else 0)


def coop_pixel(lib, KT, view, x, y):
    K, T = KT
    n = nlimbs64(view.precision)
    args = []
    for v in (x, y, x, y):
        s, e, _ = v.parts()
        args += [(C.c_uint64 * n)(*v.limbs()), s, e]
    return lib.coop_pixel(K, T, view.precision, view.fractal, view.depth, *args)


@pytest.mark.parametrize("K,prec", [((8, 8), 2048), ((6, 8), 1100), ((6, 8), 1536), ((4, 16), 2048), ((8, 16), 4096), ((6, 32), 6144)])
@pytest.mark.parametrize("fractal", [MANDELBROT, BURNING_SHIP, GENERALIZED_CELTIC, VARIANT])
def test_pixels_match_reference(coop, ref_lib, K, prec, fractal):
    views = [make_view(SEAHORSE[0], SEAHORSE[1], "1e-12", 64, 36, precision=prec, depth=400 if prec <= 2048 else 150, fractal=fractal),
             make_view("-0.5", "0.0", "4.0", 64, 36, precision=prec, depth=120, fractal=fractal)]
    if fractal == MANDELBROT and prec in (2048, 1536):
        # next to the period-707 minibrot: the orbit returns to ~0 (200 cancelled bits, then 400-bit gaps)
        views.append(make_view(MINIBROT120[0], MINIBROT120[1], "1e-120", 64, 36, precision=prec, depth=1500))
    for view in views:
        pts = [(3, 2), (40, 17), (63, 35), (32, 18)] if view.depth < 1000 else [(10, 5)]
        for ix, line in pts:
            x, y = coords(view, ix, line)
            assert coop_pixel(coop, K, view, x, y) == ref_pixel(ref_lib, view, x, y), (ix, line)
