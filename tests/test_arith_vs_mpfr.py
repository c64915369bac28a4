"""Differential test of the limb algorithms (mdz_b200/csrc/mpfr_sf.cuh, compiled
for the host by tests/host_emu with the PTX carry primitives emulated) against
the real libmpfr.so.6: mul, sqr, add, sub and the two sign-specialised adds, at
the precisions the gallery uses plus odd ones, with zeros, exact ties, deep
cancellation, and exponent gaps beyond 2p.  Bit-exact (sign, exponent, every
mantissa bit)."""
import ctypes as C
import random

import pytest

from mdz_b200.mp import Mpfr, mpfr, nlimbs64

U64P = C.POINTER(C.c_uint64)
PRECS = [64, 65, 80, 95, 96, 97, 113, 128, 176, 184, 256, 320, 511, 512, 544, 640, 777, 1000, 1024]


def rand_mant(rng, prec):
    kind = rng.randrange(8)
    top = 1 << (prec - 1)
    if kind == 0:
        return top
    if kind == 1:
        return (1 << prec) - 1
    if kind == 2:
        return top | rng.getrandbits(min(prec - 1, 8))
    if kind == 3:
        k = rng.randrange(1, prec)
        return top | ((rng.getrandbits(k) << (prec - 1 - k)) & (top - 1))
    if kind == 4:
        m, bit, pos = 0, 1, prec
        while pos > 0:
            run = min(rng.randrange(1, 40), pos)
            if bit:
                m |= ((1 << run) - 1) << (pos - run)
            pos -= run
            bit ^= 1
        return m | top
    return top | rng.getrandbits(prec - 1)


def rand_pair(rng, prec):
    ma = rand_mant(rng, prec)
    k = rng.randrange(10)
    if k == 0:
        mb = (ma ^ rng.getrandbits(rng.randrange(1, 12))) | (1 << (prec - 1))
    elif k == 1:
        mb = ma
    else:
        mb = rand_mant(rng, prec)
    ea = rng.randrange(-6, 6)
    g = rng.randrange(12)
    if g < 5:
        eb = ea + rng.randrange(-2, 3)
    elif g < 9:
        eb = ea + rng.randrange(-40, 41)
    elif g < 11:
        eb = ea + rng.randrange(-2 * prec - 70, 2 * prec + 70)
    else:
        eb = ea + rng.choice([-1, 1]) * (prec + rng.randrange(-3, 4))
    sa, sb = rng.choice([1, -1]), rng.choice([1, -1])
    if rng.randrange(60) == 0:
        sa = 0
    if rng.randrange(60) == 0:
        sb = 0
    return Mpfr(prec).set_parts(sa, ea, ma), Mpfr(prec).set_parts(sb, eb, mb)


def mpfr_op(name, prec, a, b):
    r = Mpfr(prec)
    if name == "sqr":
        mpfr.mpfr_sqr(r.ref, a.ref, 0)
    else:
        getattr(mpfr, "mpfr_" + name)(r.ref, a.ref, b.ref, 0)
    return r.parts()


def emu_op(emu, op, prec, a, b):
    n = nlimbs64(prec)
    al, bl = (C.c_uint64 * n)(*a.limbs()), (C.c_uint64 * n)(*b.limbs())
    rl, rs, re_ = (C.c_uint64 * n)(), C.c_int(), C.c_long()
    sa, ea, _ = a.parts()
    sb, eb, _ = b.parts()
    assert emu.emu_binop(op, prec, al, sa, ea, bl, sb, eb, rl, C.byref(rs), C.byref(re_))
    if rs.value == 0:
        return (0, 0, 0)
    full = 0
    for i in range(n):
        full |= rl[i] << (64 * i)
    return (rs.value, re_.value, full >> (64 * n - prec))


@pytest.fixture(scope="module")
def emu(emu_lib):
    emu_lib.emu_binop.argtypes = [C.c_int, C.c_long, U64P, C.c_int, C.c_long,
                                  U64P, C.c_int, C.c_long, U64P,
                                  C.POINTER(C.c_int), C.POINTER(C.c_long)]
    return emu_lib


@pytest.mark.parametrize("prec", PRECS)
def test_ops_match_libmpfr(emu, prec):
    rng = random.Random(1000 + prec)
    for _ in range(1500):
        a, b = rand_pair(rng, prec)
        for op, name in ((0, "mul"), (1, "sqr"), (2, "add"), (3, "sub")):
            assert emu_op(emu, op, prec, a, b) == mpfr_op(name, prec, a, b), (name, a.parts(), b.parts())
        sa, ea, ma = a.parts()
        sb, eb, mb = b.parts()
        a2 = Mpfr(prec).set_parts(abs(sa), ea, ma)
        b2 = Mpfr(prec).set_parts(abs(sb), eb, mb)
        for op, name in ((4, "sub"), (5, "add")):
            assert emu_op(emu, op, prec, a2, b2) == mpfr_op(name, prec, a2, b2), (name, a2.parts(), b2.parts())


def test_greater_than_4(emu):
    prec = 80
    four = Mpfr(prec, 4)
    rng = random.Random(7)
    for _ in range(2000):
        a = Mpfr(prec).set_parts(rng.choice([1, 1, 1, -1, 0]), rng.randrange(1, 6),
                                 rand_mant(rng, prec))
        if rng.randrange(10) == 0:
            a = Mpfr(prec).set_parts(1, 3, 1 << (prec - 1))      # exactly 4
        if rng.randrange(10) == 0:
            a = Mpfr(prec).set_parts(1, 3, (1 << (prec - 1)) | 1)  # 4 + ulp
        rs = C.c_int()
        n = nlimbs64(prec)
        al = (C.c_uint64 * n)(*a.limbs())
        sa, ea, _ = a.parts()
        emu.emu_binop(6, prec, al, sa, ea, al, sa, ea, (C.c_uint64 * n)(), C.byref(rs), C.byref(C.c_long()))
        assert rs.value == (1 if mpfr.mpfr_greater_p(a.ref, four.ref) else 0), a.parts()


def _sqrt_mod_2k(L, k):
    """x with x*x == L (mod 2^k) for L == 1 (mod 8), by lifting one bit at a time."""
    x = 1
    for b in range(3, k):
        if (x * x - L) >> b & 1:
            x += 1 << (b - 1)
    assert (x * x - L) % (1 << k) == 0
    return x


@pytest.mark.parametrize("prec", [64, 80, 128, 176, 320, 512])
def test_products_next_to_rounding_boundaries(emu, prec):
    """The fast path rounds from a truncated (high-part) product and falls back
    to the full product when the decision is within the truncation error of a
    boundary.  Build operands whose exact 2p-bit product has the bits below the
    rounding position a few guard-limb units from wrapping / from zero, with
    random low limbs, so the truncated sum really differs from the true one."""
    rng = random.Random(4242 + prec)
    n32 = (prec + 31) // 32
    unit = 1 << max(0, 32 * (n32 - 1) - (32 * n32 - prec) - 1)   # ~ one guard-limb ulp in product bits
    four = 0
    for trial in range(1200):
        w = prec - 1 - rng.randrange(2)            # width of the field below the round bit (sh = 0 / 1)
        delta = rng.randrange(0, 5 * n32 + 12)
        lowrand = rng.getrandbits(max(1, unit.bit_length() - 1)) if rng.randrange(4) else 0
        if rng.randrange(2):
            B = ((1 << w) - delta * unit - lowrand) % (1 << w)      # about to wrap
        else:
            B = (delta * unit + lowrand) % (1 << w)                  # just above zero
        L = B | (rng.getrandbits(1) << w)                           # the round bit itself
        if trial % 3 == 0:
            # square: a*a == L (mod 2^(p-1)) needs L == 1 (mod 8)
            L = (L & ~7) | 1
            a = _sqrt_mod_2k(L % (1 << (prec - 1)), prec - 1)
            if rng.randrange(2):
                a = (-a) % (1 << (prec - 1))
            a |= 1 << (prec - 1)
            b = a
        else:
            a = rand_mant(rng, prec) | 1
            b = (L * pow(a, -1, 1 << prec)) % (1 << (prec - 1)) | (1 << (prec - 1))
        A = Mpfr(prec).set_parts(1, rng.randrange(-3, 3), a)
        Bv = Mpfr(prec).set_parts(rng.choice([1, -1]), rng.randrange(-3, 3), b)
        assert emu_op(emu, 0, prec, A, Bv) == mpfr_op("mul", prec, A, Bv), (hex(a), hex(b))
        if a == b:
            assert emu_op(emu, 1, prec, A, A) == mpfr_op("sqr", prec, A, A), hex(a)
            four += 1
    assert four > 100


# ---- the p = 64 fast operations of ld64_step.cuh (long double mode) -----------------------
def test_ld64_fast_ops_match_libmpfr_or_decline(emu):
    """mul64_spec / add64_spec either give exactly libmpfr's result at precision 64 or
    raise their `rare` flag (the kernel then redoes the iteration with the general code);
    and they do not decline inside the domain they claim: gaps <= 62, fewer than 31
    cancelled bits, non-zero operands, no carry out of the rounding increment.  The level-2
    addition (add64_spec<true>) also takes gaps >= 66, an exactly zero operand on one side, and
    up to 62 cancelled bits."""
    emu.emu_ld64_op.argtypes = [C.c_int, C.c_uint64, C.c_int, C.c_long, C.c_uint64, C.c_int, C.c_long,
                                U64P, C.POINTER(C.c_int), C.POINTER(C.c_long)]
    rng = random.Random(640064)
    prec = 64
    declined = {0: 0, 2: 0, 3: 0, 4: 0, 5: 0, 8: 0}
    for k in range(40000):
        a, b = rand_pair(rng, prec)
        if k % 7 == 0:          # near-cancellation with the exponents one apart
            sa, ea, ma = a.parts()
            if sa:
                b = Mpfr(prec).set_parts(sa, ea + rng.choice([-1, 0, 1]), (ma ^ rng.getrandbits(rng.randrange(1, 64))) | (1 << 63))
        if k % 11 == 0:         # product within a few units of a power of two: the increment can carry out
            sa, ea, ma = a.parts()
            if sa:
                q = (1 << 127) // ma + rng.randrange(-2, 3)
                q = min(max(q, 1 << 63), (1 << 64) - 1)
                b = Mpfr(prec).set_parts(rng.choice([1, -1]), rng.randrange(-3, 3), q)
        if k % 5 == 0:          # gaps around and beyond the frame: 58 .. 140 bits, either way
            sa, ea, ma = a.parts()
            sb, eb, mb = b.parts()
            if sa and sb:
                if k % 15 == 0:     # a power of two: a quarter of its last place can decide a difference
                    a = Mpfr(prec).set_parts(sa, ea, 1 << 63)
                b = Mpfr(prec).set_parts(sb, ea + rng.choice([-1, 1]) * rng.randrange(58, 141), mb)
        if k % 9 == 0:          # deep cancellation: 20 .. 70 leading bits in common
            sa, ea, ma = a.parts()
            if sa:
                keep = rng.randrange(20, 71)
                mb = ma ^ rng.getrandbits(max(64 - keep, 1)) if keep < 64 else ma
                b = Mpfr(prec).set_parts(rng.choice([1, -1]), ea, mb | (1 << 63))
        sa, ea, ma = a.parts()
        sb, eb, mb = b.parts()
        if sa and sb:           # the dedicated difference of two non-negative values (wre^2 - wim^2)
            pa, pb = Mpfr(prec).set_parts(1, ea, ma), Mpfr(prec).set_parts(1, eb, mb)
            want = mpfr_op("sub", prec, pa, pb)
            for op in (6, 7):
                rm, rs, re_ = C.c_uint64(), C.c_int(), C.c_long()
                rare = emu.emu_ld64_op(op, ma, 1, ea, mb, 1, eb, C.byref(rm), C.byref(rs), C.byref(re_))
                if not rare:
                    got = (rs.value, re_.value, rm.value) if rm.value else (0, 0, 0)
                    assert got == want, ("subpos", op, pa.parts(), pb.parts(), want)
                else:
                    gap = abs(ea - eb)
                    lost = max(ea, eb) - want[1] if want[0] else 999
                    if op == 6:
                        assert gap > 62 or lost >= 31 or want[2] == 1 << 63, ("subpos declined", pa.parts(), pb.parts(), want)
                    else:
                        assert 64 <= gap <= 65 or 63 <= lost < 999, ("subpos level 2 declined", pa.parts(), pb.parts(), want)
        for op, name in ((0, "mul"), (2, "add"), (3, "sub"), (4, "add"), (5, "sub"), (8, "mul")):
            rm, rs, re_ = C.c_uint64(), C.c_int(), C.c_long()
            rare = emu.emu_ld64_op(op, ma, sa, ea, mb, sb, eb, C.byref(rm), C.byref(rs), C.byref(re_))
            want = mpfr_op(name, prec, a, b)
            if rare:
                declined[op] += 1
                big = max(ea if sa else -10**9, eb if sb else -10**9)
                if op == 8:
                    assert False, ("level 2 product declined", a.parts(), b.parts(), want)
                if op >= 4:
                    # level 2: zero operands are fine, so is any gap but 64 and 65, an exact zero result
                    # and a rounding carry-out; the one excuse left is 63 or more cancelled bits
                    covered = sa == 0 or sb == 0 or not 64 <= abs(ea - eb) <= 65
                    covered = covered and (want[0] == 0 or want[1] > big - 63)
                    assert not covered, ("level 2 " + name, a.parts(), b.parts(), want)
                    continue
                covered = sa != 0 and sb != 0 and (op == 0 or abs(ea - eb) <= 62) and want[0] != 0
                if covered and op == 0:
                    covered = want[2] != 1 << 63
                if covered and op != 0:
                    # the only excuses left: >= 31 bits cancelled, or the rounding carried out
                    covered = want[1] > big - 31 and want[2] != 1 << 63
                assert not covered, (name, a.parts(), b.parts(), want)
            else:
                got = (rs.value, re_.value, rm.value) if rm.value else (0, 0, 0)
                assert got == want, (name, op, a.parts(), b.parts(), want)
    assert declined[8] == 0
    assert declined[0] < 2500 and declined[2] < 16000 and declined[3] < 16000, declined
    assert declined[4] < declined[2] // 3 and declined[5] < declined[3] // 3, declined


# ---- the wide-gap speculative addition (mpfr_sf.cuh: fadd_spec_wide) ----------------------
@pytest.mark.parametrize("prec", [80, 96, 128, 176, 256, 320, 511, 512])
def test_wide_gap_speculative_add_matches_libmpfr_or_declines(emu, prec):
    """fadd_spec_wide gives libmpfr's result or declines, and inside its domain (gap <= 126
    bits, no zero operand, fewer than 63 cancelled bits) it may decline only for a rounding
    carry out of the lowest limb."""
    rng = random.Random(7000 + prec)
    declined = covered_declines = 0
    for k in range(4000):
        a, b = rand_pair(rng, prec)
        sa, ea, ma = a.parts()
        sb, eb, mb = b.parts()
        if k % 3 == 0 and sa and sb:             # medium gaps: the cases this variant exists for
            eb = ea + rng.choice([-1, 1]) * rng.randrange(25, 131)
            b = Mpfr(prec).set_parts(sb, eb, mb)
        a2 = Mpfr(prec).set_parts(abs(sa), ea, ma)
        b2 = Mpfr(prec).set_parts(abs(sb), eb, mb)
        for op, name, x, y in ((7, "add", a, b), (8, "sub", a, b), (9, "sub", a2, b2), (10, "add", a2, b2)):
            n = nlimbs64(prec)
            al, bl = (C.c_uint64 * n)(*x.limbs()), (C.c_uint64 * n)(*y.limbs())
            rl, rs, re_ = (C.c_uint64 * n)(), C.c_int(), C.c_long()
            sx, ex, _ = x.parts()
            sy, ey, _ = y.parts()
            rc = emu.emu_binop(op, prec, al, sx, ex, bl, sy, ey, rl, C.byref(rs), C.byref(re_))
            want = mpfr_op(name, prec, x, y)
            if rc == 2:
                declined += 1
                R = 32 * ((prec + 31) // 32) - prec
                top = lambda v: v.parts()[2] >> (prec - 32)
                eff_sub = (name == "add") != (sx == sy)
                inside = sx != 0 and sy != 0 and abs(ex - ey) <= 126 and want[0] != 0
                if inside and eff_sub:
                    inside = want[1] >= max(ex, ey) - 61
                if inside:
                    inside = ((want[2] << R) & 0xffffffff) != 0     # else: the increment may have left limb 0
                covered_declines += inside
                continue
            assert rc == 1
            full = 0
            for i in range(n):
                full |= rl[i] << (64 * i)
            got = (0, 0, 0) if rs.value == 0 else (rs.value, re_.value, full >> (64 * n - prec))
            assert got == want, (name, op, x.parts(), y.parts(), got, want)
    assert covered_declines == 0, (declined, covered_declines)


@pytest.mark.parametrize("prec", [64, 80, 128, 320, 512])
def test_escape_test_matches_mpfr_around_four(emu, prec):
    """escaped(): RN(a + b) > 4 for squares a, b >= 0 -- decided from the top limbs unless the
    sum is within 2^-27 of 4 (mpfr_sf.cuh: escape_precheck) -- against mpfr_add +
    mpfr_greater_p, with sums crowded around 4 at every distance from 2^-(p+2) to 1."""
    rng = random.Random(4000 + prec)
    four = Mpfr(prec, 4)
    n = nlimbs64(prec)
    for k in range(6000):
        # a in [0, 4), b = 4 - a + delta with |delta| from far below an ulp to O(1)
        ea = rng.choice([3, 2, 2, 1, 0, -5, -40])
        a = Mpfr(prec).set_parts(1, min(ea, 2) if ea == 3 else ea, rand_mant(rng, prec))
        b = Mpfr(prec)
        mpfr.mpfr_sub(b.ref, four.ref, a.ref, 0)
        kind = k % 4
        if kind != 0:
            d = Mpfr(prec).set_parts(rng.choice([1, -1]), 3 - rng.randrange(0, prec + 3), rand_mant(rng, prec) if kind == 1 else 1 << (prec - 1))
            mpfr.mpfr_add(b.ref, b.ref, d.ref, 0)
        sb, eb, mb = b.parts()
        if sb <= 0 or eb > 3:
            continue
        sa, ea2, _ = a.parts()
        s = Mpfr(prec)
        mpfr.mpfr_add(s.ref, a.ref, b.ref, 0)
        want = 1 if mpfr.mpfr_greater_p(s.ref, four.ref) else 0
        rs = C.c_int()
        al, bl = (C.c_uint64 * n)(*a.limbs()), (C.c_uint64 * n)(*b.limbs())
        emu.emu_binop(11, prec, al, sa, ea2, bl, sb, eb, (C.c_uint64 * n)(), C.byref(rs), C.byref(C.c_long()))
        assert rs.value == want, (a.parts(), b.parts())
