"""Differential test of the lane-group GMP mpf arithmetic (mdz_b200/csrc/coop_mpf.cuh: the kernels for mpf
precisions above 512 bits), compiled for the host by tests/host_emu/coop_gmp_emu.cpp with a warp emulated as 32
lanes in lock step, against the real libgmp.so.10: mpf_mul, mpf_mul_ui(., 2), mpf_add, mpf_sub, mpf_cmp(., 4) on
the structured operands of test_arith_vs_gmp.py (nearly equal values, the one-limb-gap "close" pattern, zeros);
then whole pixels against the reference's own frac_*_gmp functions (oracle/_ref/libmdzref.so).  Values are
compared exactly."""
import ctypes as C
import os
import random
import subprocess

import pytest

from mdz_b200 import MANDELBROT, BURNING_SHIP, GENERALIZED_CELTIC, VARIANT
from mdz_b200.mp import mpf_mul, mpf_mul_ui, mpf_add, mpf_sub, mpf_cmp
from test_arith_vs_gmp import G, U64, canon, prec_limbs, rand_pair
from test_pixel_vs_reference import gmp_coords, gmp_ref_pixel, fixed_mpf
from views import make_view, SEAHORSE, gmp_close_path_view

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
# (K words per lane, T lanes per value, mpf precision in bits): a value of P + 1 limbs needs (P + 2) limbs of T K / 2.
# 8 x 6 (to 1344 bits), 8 x 8 (1856), 16 x 8 (3904), 32 x 6 (5952), 32 x 8 (8000) are what the kernels use
CASES = [(4, 16, 576), (4, 16, 1024), (4, 16, 1856), (8, 16, 1857), (8, 16, 2048), (8, 16, 3904),
         (6, 32, 4096), (6, 32, 5952), (8, 32, 6000), (8, 32, 8000), (4, 32, 2048), (4, 16, 128), (8, 8, 1856), (8, 8, 1345), (6, 8, 1344), (6, 8, 897), (6, 8, 1024)]


@pytest.fixture(scope="module")
def coop():
    src = os.path.join(ROOT, "tests", "host_emu", "coop_gmp_emu.cpp")
    out = os.path.join(ROOT, "tests", "host_emu", "libcoopgmpemu.so")
    deps = [src] + [os.path.join(ROOT, "mdz_b200", "csrc", f) for f in ("coop_mpf.cuh", "coop_ops.cuh", "limb_ops.cuh", "mpfr_sf.cuh")]
    if not os.path.exists(out) or any(os.path.getmtime(d) > os.path.getmtime(out) for d in deps):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-o", out, src])
    lib = C.CDLL(out)
    P64 = C.POINTER(U64)
    lib.coop_gmp_op.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, P64, C.c_long, C.c_int, P64, C.c_long, C.c_int,
                                P64, C.POINTER(C.c_long), C.POINTER(C.c_int)]
    lib.coop_gmp_pixel.restype = C.c_long
    lib.coop_gmp_close_calls.restype = C.c_long
    lib.coop_gmp_pixel.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_long] + [P64, C.c_long, C.c_int] * 4
    return lib


def coop_op(lib, KT, op, a, b):
    K, T = KT
    nl = a.P + 1
    al, ae, as_ = a.fixed()
    bl, be, bs = b.fixed()
    rl, re_, rs = (U64 * nl)(), C.c_long(), C.c_int()
    assert lib.coop_gmp_op(K, T, op, nl, (U64 * nl)(*al), ae, as_, (U64 * nl)(*bl), be, bs, rl, C.byref(re_), C.byref(rs))
    assert rs.value != 99, "words below the value's limbs are set"
    assert rs.value != 98, "the groups of the warp disagree"
    assert rs.value != 97, "a zero with words set"
    assert rs.value != 96, "the shared-memory strip was not left in its resting state"
    if op == 4:
        return rs.value
    return canon(rs.value, re_.value, list(rl))


@pytest.mark.parametrize("K,T,p", CASES, ids=["%dx%d-%d" % (t, k, p) for k, t, p in CASES])
def test_ops_match_libgmp(coop, K, T, p):
    P = prec_limbs(p)
    assert T * K // 2 >= P + 2
    rng = random.Random(4100 + p + K)
    four = G(P, 1, 1, [4])
    close0 = coop.coop_gmp_close_calls()
    for _ in range(1200 if P > 40 else 2500):
        a, b = rand_pair(rng, P)
        r = G(P)
        for op, fn in ((0, mpf_mul), (2, mpf_add), (3, mpf_sub)):
            fn(r.ref, a.ref, b.ref)
            assert coop_op(coop, (K, T), op, a, b) == r.value(), (op, a.fixed(), b.fixed())
        mpf_mul(r.ref, a.ref, a.ref)
        assert coop_op(coop, (K, T), 0, a, a) == r.value(), ("sqr", a.fixed())
        mpf_mul_ui(r.ref, a.ref, 2)
        assert coop_op(coop, (K, T), 1, a, a) == r.value(), ("mul2", a.fixed())
        c = G(P, 1, rng.choice([0, 1, 1, 1, 2]), [rng.choice([0, 1, rng.getrandbits(64)]), rng.choice([3, 4, 4, 5])])
        assert coop_op(coop, (K, T), 4, c, c) == (1 if mpf_cmp(c.ref, four.ref) > 0 else 0)
    assert coop.coop_gmp_close_calls() - close0 > 20, "the one-limb-gap subtraction was not exercised"


PIXEL_CASES = [
    ("seahorse gmp 1024", 4, 16, lambda: make_view(SEAHORSE[0], SEAHORSE[1], "1e-9", 96, 72, mode="gmp", precision=1024, depth=1500), 10),
    ("seahorse gmp 2048", 8, 16, lambda: make_view(SEAHORSE[0], SEAHORSE[1], "1e-9", 96, 72, mode="gmp", precision=2048, depth=1500), 5),
    ("full set gmp 600", 4, 16, lambda: make_view("-0.5", "0.0", "4.0", 96, 72, mode="gmp", precision=600, depth=300), 60),
    ("ship gmp 1024", 4, 16, lambda: make_view("-0.5", "-0.3", "3.5", 96, 72, mode="gmp", precision=1024, depth=200, fractal=BURNING_SHIP), 30),
    ("celtic gmp 1024", 4, 16, lambda: make_view("-0.5", "-0.3", "3.5", 96, 72, mode="gmp", precision=1024, depth=200, fractal=GENERALIZED_CELTIC), 30),
    ("hybrid gmp 4096", 6, 32, lambda: make_view("-0.5", "-0.3", "3.5", 96, 72, mode="gmp", precision=4096, depth=100, fractal=VARIANT), 8),
    ("real axis gmp 1024", 4, 16, lambda: make_view("-0.75", "0.0", "2.5", 64, 48, mode="gmp", precision=1024, depth=300), 24),
]


@pytest.mark.parametrize("name,K,T,mk,count", PIXEL_CASES, ids=[c[0] for c in PIXEL_CASES])
def test_pixels_match_reference(coop, ref_lib, name, K, T, mk, count):
    view = mk()
    nl = (max(53, view.precision) + 127) // 64 + 1
    W, H = view.real_width, view.real_height
    seen = set()
    for k in range(count):
        ix, line = (k * 37 + 5) % W, (k * 53 + (H // 2 if k % 4 == 0 else 3)) % H
        x, y = gmp_coords(view, ix, line)
        args = []
        for v in (x, y, x, y):
            args += list(fixed_mpf(v, nl))
        got = coop.coop_gmp_pixel(K, T, nl, view.fractal, view.depth, *args)
        assert got == gmp_ref_pixel(ref_lib, view, x, y), (name, ix, line)
        seen.add(got)
    assert len(seen) > 2, "every sampled pixel gave the same count"


@pytest.mark.parametrize("K,T,prec", [(6, 8, 1024), (8, 8, 1800), (8, 16, 2048), (6, 32, 4096)])
def test_close_subtraction_inside_a_pixel(coop, ref_lib, K, T, prec):
    """tests/views.py gmp_close_path_view: every pixel's first wre2 - wim2 takes the one-limb-gap path (counted), and
    the pixel still comes out as the reference's frac_mandel_gmp computes it."""
    view = gmp_close_path_view(prec, 8, 6)
    nl = (max(53, prec) + 127) // 64 + 1
    from mdz_b200.mp import Mpf, mpf_set_str
    cx, cy = Mpf(prec, "-1"), Mpf(prec, "0.3")
    for ix, line in ((0, 0), (3, 0), (0, 4), (7, 5)):
        x, y = gmp_coords(view, ix, line)
        before = coop.coop_gmp_close_calls()
        args = []
        for v in (x, y, cx, cy):
            args += list(fixed_mpf(v, nl))
        got = coop.coop_gmp_pixel(K, T, nl, view.fractal, view.depth, *args)
        assert coop.coop_gmp_close_calls() > before, "the pixel did not reach the close subtraction"
        assert got == gmp_ref_pixel_julia(ref_lib, view, x, y, cx, cy) and got > 0


def gmp_ref_pixel_julia(ref, view, x, y, cx, cy):
    """frac_*_gmp on z0 = (x, y) with the constant (cx, cy): fractal.c:331-342's set-up"""
    from test_pixel_vs_reference import GFRAC, GP
    from mdz_b200.mp import Mpf, mpf_set, mpf_set_si, mpf_mul
    p = view.precision
    fn = getattr(ref, GFRAC[view.fractal])
    fn.restype = C.c_long
    fn.argtypes = [C.c_long] + [GP] * 8
    bail, wim, wre, cim, cre, wim2, wre2, t1 = (Mpf(p) for _ in range(8))
    mpf_set_si(bail.ref, 4)
    for dst, src in ((wim, y), (wre, x), (cim, cy), (cre, cx)):
        mpf_set(dst.ref, src.ref)
    mpf_mul(wim2.ref, y.ref, y.ref)
    mpf_mul(wre2.ref, x.ref, x.ref)
    return fn(view.depth, bail.ptr, wim.ptr, wre.ptr, cim.ptr, cre.ptr, wim2.ptr, wre2.ptr, t1.ptr)
