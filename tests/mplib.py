"""ctypes access to the real libmpfr.so.6 / libgmp.so.10 for differential tests.

TEST INFRASTRUCTURE ONLY.  The box ships the runtime libraries without headers,
so the struct layouts are declared here (public ABI, see include/mdz_mp_abi.h).
Values are built by writing the struct fields directly.
"""
import ctypes as C
import random

LONG_MIN = -(1 << 63)
EXP_ZERO = LONG_MIN + 1

mpfr = C.CDLL("libmpfr.so.6")
gmp = C.CDLL("libgmp.so.10")


class MpfrStruct(C.Structure):
    _fields_ = [("prec", C.c_long), ("sign", C.c_int), ("exp", C.c_long),
                ("d", C.POINTER(C.c_uint64))]


class MpfStruct(C.Structure):
    _fields_ = [("prec", C.c_int), ("size", C.c_int), ("exp", C.c_long),
                ("d", C.POINTER(C.c_uint64))]


for name in ("mpfr_mul", "mpfr_add", "mpfr_sub", "mpfr_div"):
    getattr(mpfr, name).argtypes = [C.POINTER(MpfrStruct)] * 3 + [C.c_int]
mpfr.mpfr_sqr.argtypes = [C.POINTER(MpfrStruct)] * 2 + [C.c_int]
mpfr.mpfr_init2.argtypes = [C.POINTER(MpfrStruct), C.c_long]
mpfr.mpfr_clear.argtypes = [C.POINTER(MpfrStruct)]
mpfr.mpfr_set_str.argtypes = [C.POINTER(MpfrStruct), C.c_char_p, C.c_int, C.c_int]
mpfr.mpfr_set.argtypes = [C.POINTER(MpfrStruct), C.POINTER(MpfrStruct), C.c_int]
mpfr.mpfr_set_si.argtypes = [C.POINTER(MpfrStruct), C.c_long, C.c_int]
mpfr.mpfr_set_d.argtypes = [C.POINTER(MpfrStruct), C.c_double, C.c_int]
mpfr.mpfr_get_d.argtypes = [C.POINTER(MpfrStruct), C.c_int]
mpfr.mpfr_get_d.restype = C.c_double
mpfr.mpfr_get_version.restype = C.c_char_p


def nlimbs64(prec):
    return (prec + 63) // 64


class Mpfr:
    """An mpfr_t owned by Python (limb storage is a ctypes array we keep alive)."""

    def __init__(self, prec):
        self.prec = prec
        self.n = nlimbs64(prec)
        # MPFR keeps an allocation-size word in front of the limbs; mpfr_set_prec
        # is never called on these, so a bare array is enough for arithmetic.
        self.buf = (C.c_uint64 * (self.n + 1))()
        self.buf[0] = self.n
        self.s = MpfrStruct(prec, 1, EXP_ZERO,
                            C.cast(C.byref(self.buf, 8), C.POINTER(C.c_uint64)))

    @property
    def ref(self):
        return C.byref(self.s)

    def set_parts(self, sign, exp, mant):
        """sign +-1 (0 = zero); mant is a prec-bit int with the top bit set."""
        if sign == 0 or mant == 0:
            self.s.sign = 1
            self.s.exp = EXP_ZERO
            return self
        assert mant >> (self.prec - 1) == 1, "mantissa not normalised"
        full = mant << (64 * self.n - self.prec)
        for i in range(self.n):
            self.buf[1 + i] = (full >> (64 * i)) & 0xFFFFFFFFFFFFFFFF
        self.s.sign = 1 if sign > 0 else -1
        self.s.exp = exp
        return self

    def parts(self):
        """-> (sign, exp, mant) with sign 0 for zero."""
        if self.s.exp == EXP_ZERO:
            return (0, 0, 0)
        assert self.s.exp > LONG_MIN + 3, "NaN/Inf"
        full = 0
        for i in range(self.n):
            full |= self.buf[1 + i] << (64 * i)
        return (1 if self.s.sign > 0 else -1, self.s.exp,
                full >> (64 * self.n - self.prec))

    def limbs(self):
        return [self.buf[1 + i] for i in range(self.n)]

    def set_str(self, text, base=10):
        mpfr.mpfr_set_str(self.ref, text.encode(), base, 0)
        return self

    def set_d(self, v):
        mpfr.mpfr_set_d(self.ref, float(v), 0)
        return self

    def to_float(self):
        return mpfr.mpfr_get_d(self.ref, 0)

    def to_fraction_parts(self):
        s, e, m = self.parts()
        return s, m, e - self.prec   # value = s * m * 2^(e-prec)


def rand_mant(rng, prec):
    """Structured random prec-bit mantissa with the top bit set."""
    kind = rng.randrange(8)
    top = 1 << (prec - 1)
    if kind == 0:
        return top                                   # minimal
    if kind == 1:
        return (1 << prec) - 1                        # all ones
    if kind == 2:
        return top | rng.getrandbits(min(prec - 1, 8))          # few low bits
    if kind == 3:
        k = rng.randrange(1, prec)
        return top | (rng.getrandbits(k) << (prec - 1 - k) if k < prec - 1 else rng.getrandbits(prec - 1))
    if kind == 4:
        # long runs of ones / zeros
        m = 0
        bit = 1
        pos = prec
        while pos > 0:
            run = rng.randrange(1, 40)
            run = min(run, pos)
            if bit:
                m |= ((1 << run) - 1) << (pos - run)
            pos -= run
            bit ^= 1
        return m | top
    return top | rng.getrandbits(prec - 1)


def mpfr_op(name, prec, a, b=None):
    r = Mpfr(prec)
    if name == "sqr":
        mpfr.mpfr_sqr(r.ref, a.ref, 0)
    else:
        getattr(mpfr, "mpfr_" + name)(r.ref, a.ref, b.ref, 0)
    return r
