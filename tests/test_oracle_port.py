"""Pin the plain-C oracle (oracle/mdz_oracle.c): its arithmetic against libmpfr
operation by operation, and its renders against the unmodified reference."""
import ctypes as C
import random

import numpy as np
import pytest

import portpath
from mdz_b200 import FAMILY_JULIA, BURNING_SHIP, GENERALIZED_CELTIC, VARIANT
from mdz_b200.mp import Mpfr, mpfr
from refpath import ref_render
from test_arith_vs_mpfr import rand_pair
from views import make_view, config2, SEAHORSE, deep_embedded_julia


@pytest.mark.parametrize("prec", [64, 80, 113, 176, 320, 512])
def test_oracle_arithmetic_matches_libmpfr(prec):
    lib = portpath.load()
    rng = random.Random(77 + prec)
    n = (prec + 63) // 64
    for _ in range(600):
        a, b = rand_pair(rng, prec)
        keep = []
        for op, name in ((0, "mul"), (1, "add"), (2, "sub"), (3, "div")):
            if name == "div" and b.parts()[0] == 0:
                continue
            want = Mpfr(prec)
            getattr(mpfr, "mpfr_" + name)(want.ref, a.ref, b.ref, 0)
            rl, rs, re_ = (C.c_uint64 * n)(), C.c_int(), C.c_long()
            assert lib.oracle_mpfr_op(op, prec, portpath.to_num(a, keep), portpath.to_num(b, keep),
                                      rl, C.byref(rs), C.byref(re_))
            full = 0
            for i in range(n):
                full |= rl[i] << (64 * i)
            got = (0, 0, 0) if rs.value == 0 else (rs.value, re_.value, full >> (64 * n - prec))
            assert got == want.parts(), (name, a.parts(), b.parts())


VIEWS = [
    ("cfg2 long double", lambda: config2(96, 54, 1500)),
    ("ship long double", lambda: make_view("-0.5", "-0.3", "3.5", 64, 48, mode="ld", depth=300, fractal=BURNING_SHIP)),
    ("seahorse mpfr 80", lambda: make_view(SEAHORSE[0], SEAHORSE[1], "1e-9", 40, 30, precision=80, depth=800)),
    ("celtic mpfr 128", lambda: make_view("-0.5", "-0.3", "3.5", 48, 36, precision=128, depth=200, fractal=GENERALIZED_CELTIC)),
    ("hybrid mpfr 184", lambda: make_view("-0.5", "-0.3", "3.5", 48, 36, precision=184, depth=200, fractal=VARIANT)),
    ("julia mpfr 96", lambda: make_view("0", "0", "3.2", 48, 36, precision=96, depth=300, family=FAMILY_JULIA, julia=("-0.8", "0.156"))),
    ("deep_embedded_julia 320", lambda: deep_embedded_julia(24, 18)),
]


@pytest.mark.parametrize("name,mk", VIEWS, ids=[v[0] for v in VIEWS])
def test_oracle_render_matches_reference(ref_lib, name, mk):
    view = mk()
    want, _ = ref_render(ref_lib, view)
    got = portpath.port_render(view)
    assert np.array_equal(got, want), name


# ---- GMP mpf mode of the C oracle -------------------------------------------------------
@pytest.mark.parametrize("p", [80, 128, 256, 320, 512])
def test_oracle_mpf_arithmetic_matches_libgmp(p):
    from test_arith_vs_gmp import G, prec_limbs, rand_pair as grand_pair, canon
    from mdz_b200.mp import mpf_mul, mpf_mul_ui, mpf_add, mpf_sub, mpf_cmp
    lib = portpath.load()
    PM = C.POINTER(portpath.OracleMpf)
    lib.oracle_mpf_op.argtypes = [C.c_int, C.c_int, PM, PM, C.POINTER(C.c_uint64), C.POINTER(C.c_int),
                                  C.POINTER(C.c_long), C.POINTER(C.c_int)]
    P = prec_limbs(p)
    rng = random.Random(3100 + p)

    def as_oracle(g, keep):
        n = abs(g.s.size)
        arr = (C.c_uint64 * max(1, n))(*[g.buf[i] for i in range(n)])
        keep.append(arr)
        return C.pointer(portpath.OracleMpf(0 if n == 0 else (1 if g.s.size > 0 else -1), g.s.exp, n,
                                            C.cast(arr, C.POINTER(C.c_uint64))))

    for _ in range(3000):
        a, b = grand_pair(rng, P)
        keep = []
        oa, ob = as_oracle(a, keep), as_oracle(b, keep)
        r = G(P)
        for op, fn in ((0, mpf_mul), (2, mpf_add), (3, mpf_sub), (1, None)):
            if fn is None:
                mpf_mul_ui(r.ref, a.ref, 2)
            else:
                fn(r.ref, a.ref, b.ref)
            rl, rn, re_, rs = (C.c_uint64 * 48)(), C.c_int(), C.c_long(), C.c_int()
            assert lib.oracle_mpf_op(op, P, oa, ob, rl, C.byref(rn), C.byref(re_), C.byref(rs))
            got = canon(rs.value, re_.value, [rl[i] for i in range(rn.value)])
            assert got == r.value(), (op, a.fixed(), b.fixed())


GVIEWS = [
    ("seahorse gmp 128", lambda: make_view(SEAHORSE[0], SEAHORSE[1], "1e-9", 40, 30, mode="gmp", precision=128, depth=800), None),
    ("ship gmp 256", lambda: make_view("-0.5", "-0.3", "3.5", 48, 36, mode="gmp", precision=256, depth=200, fractal=BURNING_SHIP), None),
    ("hybrid gmp 80", lambda: make_view("-0.5", "-0.3", "3.5", 48, 36, mode="gmp", precision=80, depth=200, fractal=VARIANT), None),
    ("julia gmp 128 (1 thread)", lambda: make_view("0", "0", "3.2", 48, 36, mode="gmp", precision=128, depth=300, family=FAMILY_JULIA, julia=("-0.8", "0.156")), 1),
    ("seahorse gmp 512", lambda: make_view(SEAHORSE[0], SEAHORSE[1], "1e-9", 24, 18, mode="gmp", precision=512, depth=800), None),
]


@pytest.mark.parametrize("name,mk,threads", GVIEWS, ids=[v[0] for v in GVIEWS])
def test_oracle_gmp_render_matches_reference(ref_lib, name, mk, threads):
    view = mk()
    want, _ = ref_render(ref_lib, view, threads)
    got = portpath.port_render(view)
    assert np.array_equal(got, want), name
