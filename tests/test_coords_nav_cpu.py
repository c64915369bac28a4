"""View navigation and precision advice (mdz_b200/coords.py: Coords) side by side with the
UNMODIFIED reference's `struct coords` and coords_* functions (oracle/_ref/libmdzref.so,
reference src/coords.c): after every operation every field must be bit-identical."""
import ctypes as C
import random

import pytest

from mdz_b200.coords import Coords
from mdz_b200.mp import Mpfr, MpfrStruct, EXP_ZERO


class RefCoords(C.Structure):
    """struct coords, reference src/coords.h:28-57."""
    _fields_ = [("img_width", C.c_int), ("img_height", C.c_int),
                ("xmin", MpfrStruct), ("xmax", MpfrStruct), ("ymin", MpfrStruct), ("ymax", MpfrStruct),
                ("width", MpfrStruct), ("height", MpfrStruct), ("cx", MpfrStruct), ("cy", MpfrStruct),
                ("size", C.POINTER(MpfrStruct)), ("_size", MpfrStruct),
                ("aspect", C.c_double), ("precision", C.c_long), ("recommend", C.c_long),
                ("gmp_precision", C.c_ulong),
                ("init_cx", C.c_double), ("init_cy", C.c_double), ("init_size", C.c_double)]


def parts(s):
    """An mpfr_t of the reference as (sign, exp, mantissa), like Mpfr.parts()."""
    if s.exp == EXP_ZERO + 1:
        return "nan"
    if s.exp == EXP_ZERO:
        return (0, 0, 0)
    n = (s.prec + 63) // 64
    full = 0
    for i in range(n):
        full |= s.d[i] << (64 * i)
    return (1 if s.sign > 0 else -1, s.exp, full >> (64 * n - s.prec))


def ref_state(rc):
    out = {f: parts(getattr(rc, f)) for f in Coords.FIELDS}
    out.update(precision=rc.precision, recommend=rc.recommend, aspect=rc.aspect)
    return out


@pytest.fixture(scope="module")
def ref(ref_lib):
    P = C.POINTER(RefCoords)
    ref_lib.coords_new.restype = P
    ref_lib.coords_new.argtypes = [C.c_int, C.c_int, C.c_double, C.c_double, C.c_double]
    for name, args in (("coords_reset", []), ("coords_center_to_rect", []), ("coords_rect_to_center", []),
                       ("coords_zoom", [C.c_double]), ("coords_zoom_to", [C.c_int] * 3),
                       ("coords_center_to", [C.c_int] * 2), ("coords_reposition", [C.c_int] * 4),
                       ("coords_set_precision", [C.c_long]), ("coords_set", [C.c_int] * 2),
                       ("coords_size", [C.POINTER(MpfrStruct)]), ("coords_to", [C.POINTER(MpfrStruct)] * 2),
                       ("coords_set_rect", [C.POINTER(MpfrStruct)] * 3)):
        fn = getattr(ref_lib, name)
        fn.argtypes = [P] + args
        fn.restype = None
    ref_lib.coords_calculate_precision.argtypes = [P]
    ref_lib.coords_calculate_precision.restype = C.c_int
    return ref_lib


@pytest.mark.parametrize("w,h", [(1920, 1080), (240, 180), (480, 640), (512, 512)])
def test_navigation_matches_reference_field_by_field(ref, w, h):
    rng = random.Random(w * 10007 + h)
    rc = ref.coords_new(w, h, -0.5, 0.0, 4.0)
    mine = Coords(w, h, -0.5, 0.0, 4.0)
    ref.coords_reset(rc)
    assert mine.reset() == rc.contents.recommend
    ref.coords_center_to_rect(rc)
    mine.center_to_rect()
    assert mine.state() == ref_state(rc.contents)
    for step in range(120):
        op = rng.choice(["zoom_to", "zoom_to", "zoom", "reposition", "center_to", "precision", "resize", "size", "to", "rect"])
        if op == "zoom_to":
            a = (rng.randrange(w), rng.randrange(h), rng.randrange(1, w))
            ref.coords_zoom_to(rc, *a); mine.zoom_to(*a)
        elif op == "zoom":
            z = rng.choice([0.5, 2.0, 0.1, 1.25, 1.0 / 3.0])
            ref.coords_zoom(rc, z); mine.zoom(z)
        elif op == "reposition":
            a = [rng.randrange(-w, 2 * w), rng.randrange(-h, 2 * h), rng.randrange(w), rng.randrange(h)]
            ref.coords_reposition(rc, *a); mine.reposition(*a)
            ref.coords_center_to_rect(rc); mine.center_to_rect()
        elif op == "center_to":
            a = (rng.randrange(w), rng.randrange(h))
            ref.coords_center_to(rc, *a); mine.center_to(*a)
            ref.coords_center_to_rect(rc); mine.center_to_rect()
        elif op == "precision":
            # what MDZ does once the advice exceeds the precision in use (image_info.c:228-233 via the GUI)
            p = max(rc.contents.recommend + rng.randrange(0, 40), rng.choice([53, 80, 96, 130]))
            ref.coords_set_precision(rc, p); mine.set_precision(p)
        elif op == "resize":
            nw, nh = rng.choice([(w, h), (h, w), (640, 480), (333, 777)])
            ref.coords_set(rc, nw, nh); mine.set(nw, nh)
            ref.coords_center_to_rect(rc); mine.center_to_rect()
            ref.coords_set(rc, w, h); mine.set(w, h)
            ref.coords_center_to_rect(rc); mine.center_to_rect()
        elif op == "size":
            v = Mpfr(mine.precision, "%.17g" % (rng.random() * 3 + 1e-3))
            ref.coords_size(rc, v.ptr); assert mine.set_size(v) == rc.contents.recommend
            ref.coords_center_to_rect(rc); mine.center_to_rect()
        elif op == "to":
            a, b = Mpfr(mine.precision, "%.17g" % rng.uniform(-2, 1)), Mpfr(mine.precision, "%.17g" % rng.uniform(-1, 1))
            ref.coords_to(rc, a.ptr, b.ptr); mine.to(a, b)
            ref.coords_center_to_rect(rc); mine.center_to_rect()
        else:
            x0 = rng.uniform(-2, 0.5)
            a, b, c = (Mpfr(mine.precision, "%.17g" % x0), Mpfr(mine.precision, "%.17g" % (x0 + rng.uniform(1e-6, 1))),
                       Mpfr(mine.precision, "%.17g" % rng.uniform(-1, 1)))
            ref.coords_set_rect(rc, a.ptr, b.ptr, c.ptr); mine.set_rect(a, b, c)
        assert mine.state() == ref_state(rc.contents), (step, op)
        assert ref.coords_calculate_precision(rc) == mine.calculate_precision(), (step, op)


def test_precision_advice_grows_with_depth_and_picks_limbs(ref):
    """Zooming by 10 adds log2(10) bits to the advice; the limb count follows it (long double while
    64 bits do, then ceil(p / 32) words for MPFR, 2 * (P + 1) for GMP mpf)."""
    rc = ref.coords_new(1920, 1080, -0.75, 0.1, 4.0)
    mine = Coords(1920, 1080, -0.75, 0.1, 4.0)
    ref.coords_reset(rc); mine.reset()
    ref.coords_center_to_rect(rc); mine.center_to_rect()
    seen = []
    for k in range(60):
        ref.coords_zoom(rc, 0.1); mine.zoom(0.1)
        assert mine.recommend == rc.contents.recommend
        if mine.recommend > mine.precision:
            ref.coords_set_precision(rc, mine.recommend); mine.set_precision(mine.recommend)
        seen.append((mine.recommend, mine.limb_count(), mine.limb_count("gmp")))
    assert seen[0][0] in (15, 16)                                   # 12 bits at size 4, 1920 wide; + log2(10)
    assert all(2 <= b[0] - a[0] <= 5 for a, b in zip(seen, seen[1:]))
    assert abs(seen[-1][0] - (12 + 60 * 3.3219)) < 3
    for rec, n, g in seen:
        assert n == (2 if rec <= 64 else (rec + 31) // 32)
        assert g == (2 if rec <= 64 else 2 * ((max(53, rec) + 127) // 64 + 1))
