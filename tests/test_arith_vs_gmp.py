"""Differential test of the GMP-mpf-faithful limb algorithms (mdz_b200/csrc/mpf_sf.cuh,
host build) against the real libgmp.so.10 (GMP 6.3.0): mpf_mul, mpf_mul_ui(.,2),
mpf_add, mpf_sub and mpf_cmp(.,4) on structured random operands built by writing
the __mpf_struct fields directly.  Values are compared
exactly (sign, limb exponent, limbs with trailing zero limbs ignored)."""
import ctypes as C
import random

import pytest

from mdz_b200.mp import MpfStruct, mpf_mul, mpf_mul_ui, mpf_add, mpf_sub, mpf_cmp, mpf_init2, mpf_set_si

U64 = C.c_uint64
M64 = (1 << 64) - 1


def prec_limbs(p):
    return (max(53, p) + 127) // 64


class G:
    """mpf value with Python-owned limb storage (prec+1 limbs as GMP allocates)."""

    def __init__(self, P, sign=0, exp=0, limbs=()):
        self.P = P
        self.buf = (U64 * (P + 2))()
        self.s = MpfStruct(P, 0, 0, C.cast(self.buf, C.POINTER(U64)))
        self.set(sign, exp, limbs)

    def set(self, sign, exp, limbs):
        limbs = list(limbs)
        assert len(limbs) <= self.P + 1
        for i, w in enumerate(limbs):
            self.buf[i] = w
        self.s.size = len(limbs) * (1 if sign >= 0 else -1)
        self.s.exp = exp if limbs else 0

    @property
    def ref(self):
        return C.byref(self.s)

    def value(self):
        n = abs(self.s.size)
        l = [self.buf[i] for i in range(n)]
        return canon(0 if n == 0 else (1 if self.s.size > 0 else -1), self.s.exp, l)

    def fixed(self):
        """-> (limbs top-aligned in P+1, exp, sign) for the emulation."""
        n = abs(self.s.size)
        nl = self.P + 1
        out = [0] * nl
        for i in range(n):
            out[nl - n + i] = self.buf[i]
        return out, self.s.exp, (0 if n == 0 else (1 if self.s.size > 0 else -1))


def canon(sign, exp, limbs):
    """value-canonical form: strip trailing (low) zero limbs."""
    limbs = list(limbs)
    if sign == 0 or not any(limbs):
        return (0, 0, ())
    while limbs and limbs[0] == 0:
        limbs.pop(0)
    assert limbs[-1] != 0
    return (sign, exp, tuple(limbs))


def strip_low(limbs):
    limbs = list(limbs)
    while len(limbs) > 1 and limbs[0] == 0 and random.random() < 0.5:
        limbs.pop(0)
    return limbs


def rand_limb(rng):
    k = rng.randrange(8)
    if k == 0:
        return 0
    if k == 1:
        return 1
    if k == 2:
        return 1 << 63
    if k == 3:
        return M64
    if k == 4:
        return rng.getrandbits(rng.randrange(1, 64))
    return rng.getrandbits(64)


def rand_val(rng, P, exp0):
    n = rng.randrange(1, P + 2)
    limbs = [rand_limb(rng) for _ in range(n)]
    if limbs[-1] == 0:
        limbs[-1] = rng.choice([1, 2, M64, rng.getrandbits(64) | 1])
    return G(P, rng.choice([1, -1]), exp0 + rng.randrange(-6, 7), limbs)


def rand_pair(rng, P):
    a = rand_val(rng, P, 0)
    if rng.randrange(2):
        # nearly equal: same exponent / off by one, shared top limbs
        n = abs(a.s.size)
        limbs = [a.buf[i] for i in range(n)]
        k = rng.randrange(0, n)
        for i in range(k + 1):
            if rng.randrange(2):
                limbs[i] = rand_limb(rng)
        if rng.randrange(3) == 0 and limbs[-1] > 1:
            limbs[-1] += rng.choice([-1, 1]) if limbs[-1] < M64 else -1
        if limbs[-1] == 0:
            limbs[-1] = 1
        b = G(P, rng.choice([1, -1]), a.s.exp + rng.choice([0, 0, 0, 1, -1]), limbs)
    else:
        b = rand_val(rng, P, a.s.exp)
    if rng.randrange(12) == 0:
        # the one-limb-gap "close" pattern of mpf_sub: 1:0:0.. against ff..ff:ff..:x
        e0 = rng.randrange(-3, 4)
        nz = rng.randrange(0, P + 1)                      # zero limbs below u's top
        ul = [rand_limb(rng) for _ in range(P - nz)] + [0] * nz + [1]
        ul = ul[-(P + 1):]
        nf = rng.randrange(1, P + 2)                      # leading ff limbs of v
        vl = [rand_limb(rng) for _ in range(P + 1 - nf)] + [M64] * nf
        sg = rng.choice([1, -1])
        a = G(P, sg, e0 + 1, strip_low(ul))
        b = G(P, sg if rng.randrange(4) else -sg, e0, strip_low(vl))
        if rng.randrange(2):
            a, b = b, a
    if rng.randrange(40) == 0:
        a = G(P)
    if rng.randrange(40) == 0:
        b = G(P)
    return a, b


@pytest.fixture(scope="module")
def emu(emu_lib):
    P64 = C.POINTER(U64)
    for fn in (emu_lib.emu_gmp_op, emu_lib.emu_gmpf_op):
        fn.argtypes = [C.c_int, C.c_int, P64, C.c_long, C.c_int, P64, C.c_long, C.c_int,
                       P64, C.POINTER(C.c_long), C.POINTER(C.c_int)]
    return emu_lib


def emu_op(emu, op, a, b, fast=False):
    nl = a.P + 1
    al, ae, as_ = a.fixed()
    bl, be, bs = b.fixed()
    rl, re_, rs = (U64 * nl)(), C.c_long(), C.c_int()
    fn = emu.emu_gmpf_op if fast else emu.emu_gmp_op
    assert fn(op, nl, (U64 * nl)(*al), ae, as_, (U64 * nl)(*bl), be, bs, rl, C.byref(re_), C.byref(rs))
    if op == 4:
        return rs.value
    return canon(rs.value, re_.value, list(rl))


@pytest.mark.parametrize("p", [80, 128, 200, 256, 320, 512, 1024])
def test_ops_match_libgmp(emu, p):
    P = prec_limbs(p)
    rng = random.Random(900 + p)
    four = G(P, 1, 1, [4])
    for _ in range(6000):
        a, b = rand_pair(rng, P)
        r = G(P)
        for op, fn in ((0, mpf_mul), (2, mpf_add), (3, mpf_sub)):
            fn(r.ref, a.ref, b.ref)
            assert emu_op(emu, op, a, b) == r.value(), (op, a.fixed(), b.fixed())
        mpf_mul(r.ref, a.ref, a.ref)
        assert emu_op(emu, 0, a, a) == r.value(), ("sqr", a.fixed())
        mpf_mul_ui(r.ref, a.ref, 2)
        assert emu_op(emu, 1, a, a) == r.value(), ("mul2", a.fixed())
        # compare against 4 on a small positive value
        c = G(P, 1, rng.choice([0, 1, 1, 1, 2]), [rng.choice([0, 1, rng.getrandbits(64)]), rng.choice([3, 4, 4, 5])])
        assert emu_op(emu, 4, c, c) == (1 if mpf_cmp(c.ref, four.ref) > 0 else 0)


@pytest.mark.parametrize("p", [80, 128, 200, 256, 320, 512, 576, 640, 704, 768, 832, 896, 1024])
def test_fast_ops_match_libgmp(emu, p):
    """mpf_fast.cuh (the register / IMAD.WIDE implementation the kernel uses)."""
    P = prec_limbs(p)
    rng = random.Random(1700 + p)
    four = G(P, 1, 1, [4])
    for _ in range(6000):
        a, b = rand_pair(rng, P)
        r = G(P)
        for op, fn in ((0, mpf_mul), (2, mpf_add), (3, mpf_sub)):
            fn(r.ref, a.ref, b.ref)
            assert emu_op(emu, op, a, b, True) == r.value(), (op, a.fixed(), b.fixed())
        mpf_mul(r.ref, a.ref, a.ref)
        assert emu_op(emu, 5, a, a, True) == r.value(), ("sqr", a.fixed())
        mpf_mul_ui(r.ref, a.ref, 2)
        assert emu_op(emu, 1, a, a, True) == r.value(), ("mul2", a.fixed())
        c = G(P, 1, rng.choice([0, 1, 1, 1, 2]), [rng.choice([0, 1, rng.getrandbits(64)]), rng.choice([3, 4, 4, 5])])
        assert emu_op(emu, 4, c, c, True) == (1 if mpf_cmp(c.ref, four.ref) > 0 else 0)


@pytest.mark.parametrize("p", [80, 128, 320, 512, 640, 896])
def test_fast_products_next_to_truncation_boundary(emu, p):
    """gf_mul forms only the high columns of the product and falls back to the full
    product when the guard word is within the truncation error of wrapping.  Build
    operands whose exact product has that guard word a few units from 0xffffffff with
    random words below, so the carry into the kept limbs really depends on what was
    left out; and squares likewise (2-adic square roots)."""
    from test_arith_vs_mpfr import _sqrt_mod_2k
    P = prec_limbs(p)
    M = 2 * P                      # 32-bit words of an operand
    rng = random.Random(5100 + p)
    k = 32 * (M - 4)               # product bits below position M-4
    for trial in range(1500):
        # The kept columns hold (true product - omitted part); the omitted part is worth up to
        # ~M units of word M-6.  When the TRUE product has word M-5 == 0 and a tiny word M-6,
        # the truncated sum borrows through both and the limbs above come out one short.
        delta = rng.randrange(0, 3 * M + 8)
        w5 = 0 if rng.randrange(4) else rng.choice([1, 0xFFFFFFFF, 0xFFFFFFFE])
        low = rng.getrandbits(32 * (M - 6)) if rng.randrange(4) else 0
        L = (w5 << (32 * (M - 5))) | (delta << (32 * (M - 6))) | low
        if trial % 3 == 0:
            L = (L & ~7) | 1
            a = _sqrt_mod_2k(L % (1 << k), k)
            a |= rng.getrandbits(32 * 4) << k
            b = a
        else:
            a = rng.getrandbits(32 * M) | 1
            b = (L * pow(a, -1, 1 << k)) % (1 << k)
            b |= rng.getrandbits(32 * 4) << k
        # small top limbs make the product's top limb zero (GMP drops it: the kept window
        # then reaches two words further down, right next to the guard word)
        topmask = (1 << (64 * (P - 1))) - 1
        if P >= 4 and rng.randrange(3):
            a = (a & topmask) | (rng.randrange(1, 1 << 20) << (64 * (P - 1)))
            b = (b & topmask) | (rng.randrange(1, 1 << 20) << (64 * (P - 1)))
        if a >> (64 * (P - 1)) == 0:
            a |= 1 << (64 * P - 7)
        if b >> (64 * (P - 1)) == 0:
            b |= 1 << (64 * P - 7)
        if trial % 3 == 0:
            b = a
        def limbs(v):
            return [rng.getrandbits(64)] + [(v >> (64 * i)) & M64 for i in range(P)]
        A = G(P, 1, rng.randrange(-2, 3), limbs(a))
        B = A if trial % 3 == 0 else G(P, rng.choice([1, -1]), rng.randrange(-2, 3), limbs(b))
        r = G(P)
        mpf_mul(r.ref, A.ref, B.ref)
        if trial % 3 == 0:
            assert emu_op(emu, 5, A, A, True) == r.value(), hex(a)
        else:
            assert emu_op(emu, 0, A, B, True) == r.value(), (hex(a), hex(b))
