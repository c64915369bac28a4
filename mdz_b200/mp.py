"""ctypes views of the public GMP / MPFR value structs (include/mdz_mp_abi.h).

MDZ's host API hands `mpfr_t` / `mpf_t` values across the render boundary
(reference src/image_info.h:65-66), so the Python host mirror needs to build
them.  The box ships libmpfr.so.6 / libgmp.so.10 without headers; layouts are
the documented public ABI.
"""
import ctypes as C

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


_P = C.POINTER(MpfrStruct)
_G = C.POINTER(MpfStruct)
for _n in ("mpfr_mul", "mpfr_add", "mpfr_sub", "mpfr_div"):
    getattr(mpfr, _n).argtypes = [_P, _P, _P, C.c_int]
mpfr.mpfr_sqr.argtypes = [_P, _P, C.c_int]
mpfr.mpfr_set.argtypes = [_P, _P, C.c_int]
mpfr.mpfr_abs.argtypes = [_P, _P, C.c_int]
mpfr.mpfr_neg.argtypes = [_P, _P, C.c_int]
mpfr.mpfr_set_str.argtypes = [_P, C.c_char_p, C.c_int, C.c_int]
mpfr.mpfr_set_si.argtypes = [_P, C.c_long, C.c_int]
mpfr.mpfr_set_d.argtypes = [_P, C.c_double, C.c_int]
mpfr.mpfr_get_d.argtypes = [_P, C.c_int]
mpfr.mpfr_get_d.restype = C.c_double
mpfr.mpfr_div_d.argtypes = [_P, _P, C.c_double, C.c_int]
mpfr.mpfr_mul_d.argtypes = [_P, _P, C.c_double, C.c_int]
mpfr.mpfr_div_ui.argtypes = [_P, _P, C.c_ulong, C.c_int]
mpfr.mpfr_mul_si.argtypes = [_P, _P, C.c_long, C.c_int]
mpfr.mpfr_si_div.argtypes = [_P, C.c_long, _P, C.c_int]
mpfr.mpfr_greater_p.argtypes = [_P, _P]
mpfr.mpfr_cmp.argtypes = [_P, _P]
mpfr.mpfr_get_version.restype = C.c_char_p
mpfr.mpfr_get_str.argtypes = [C.c_char_p, C.POINTER(C.c_long), C.c_int, C.c_size_t, _P, C.c_int]
mpfr.mpfr_get_str.restype = C.c_void_p
mpfr.mpfr_free_str.argtypes = [C.c_void_p]
mpfr.mpfr_free_str.restype = None
mpfr.mpfr_set_nan.argtypes = [_P]
mpfr.mpfr_set_nan.restype = None
mpfr.mpfr_free_str.argtypes = [C.c_void_p]

def _gf(name, argtypes, restype=None):
    fn = getattr(gmp, "__gmpf_" + name)     # exported names carry the __gmpf_ prefix
    fn.argtypes = argtypes
    fn.restype = restype
    return fn


mpf_init2 = _gf("init2", [_G, C.c_ulong])
mpf_clear = _gf("clear", [_G])
mpf_set_str = _gf("set_str", [_G, C.c_char_p, C.c_int], C.c_int)
mpf_set = _gf("set", [_G, _G])
mpf_set_si = _gf("set_si", [_G, C.c_long])
mpf_get_d = _gf("get_d", [_G], C.c_double)
mpf_add = _gf("add", [_G, _G, _G])
mpf_sub = _gf("sub", [_G, _G, _G])
mpf_mul = _gf("mul", [_G, _G, _G])
mpf_div = _gf("div", [_G, _G, _G])
mpf_mul_ui = _gf("mul_ui", [_G, _G, C.c_ulong])
mpf_ui_div = _gf("ui_div", [_G, C.c_ulong, _G])
mpf_abs = _gf("abs", [_G, _G])
mpf_cmp = _gf("cmp", [_G, _G], C.c_int)


def nlimbs64(prec):
    return (prec + 63) // 64


class Mpfr:
    """An mpfr_t whose limb storage is owned by Python.

    MPFR keeps an allocation-size word in front of the limbs; precision is
    never changed on these objects, so a fixed array is enough for arithmetic.
    """

    def __init__(self, prec, value=None):
        self.prec = int(prec)
        self.n = nlimbs64(self.prec)
        self.buf = (C.c_uint64 * (self.n + 1))()
        self.buf[0] = self.n
        self.s = MpfrStruct(self.prec, 1, EXP_ZERO,
                            C.cast(C.byref(self.buf, 8), C.POINTER(C.c_uint64)))
        if value is not None:
            if isinstance(value, str):
                self.set_str(value)
            elif isinstance(value, Mpfr):
                mpfr.mpfr_set(self.ref, value.ref, 0)
            elif isinstance(value, int):
                mpfr.mpfr_set_si(self.ref, value, 0)
            else:
                self.set_d(value)

    @property
    def ref(self):
        return C.byref(self.s)

    @property
    def ptr(self):
        return C.pointer(self.s)

    def set_parts(self, sign, exp, mant):
        """sign +-1 (0 = zero); mant is a prec-bit int with its top bit set."""
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
        """-> (sign, exp, mant); sign 0 for zero."""
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
        if mpfr.mpfr_set_str(self.ref, text.encode(), base, 0) != 0:
            raise ValueError("mpfr_set_str failed on %r" % text)
        return self

    def out_str(self):
        """What mpfr_out_str(fd, 10, 0, x, GMP_RNDN) prints: all the digits needed to read the
        value back exactly, one digit before the point, exponent always present."""
        if self.s.exp == EXP_ZERO + 1:                     # NaN (mpfr.h: __MPFR_EXP_NAN)
            return "@NaN@"
        e = C.c_long()
        raw = mpfr.mpfr_get_str(None, C.byref(e), 10, 0, self.ref, 0)
        digits = C.cast(raw, C.c_char_p).value.decode()
        mpfr.mpfr_free_str(C.c_void_p(raw))
        sign = ""
        if digits.startswith("-"):
            sign, digits = "-", digits[1:]
        if self.s.exp == EXP_ZERO:                          # zero: mpfr_get_str gives "0000..." with e = 0
            return sign + digits[0] + "." + digits[1:] + "e0"
        return "%s%s.%se%d" % (sign, digits[0], digits[1:], e.value - 1)

    def set_nan(self):
        mpfr.mpfr_set_nan(self.ref)
        return self

    def set_d(self, v):
        mpfr.mpfr_set_d(self.ref, float(v), 0)
        return self

    def to_float(self):
        return mpfr.mpfr_get_d(self.ref, 0)

    def to_json(self):
        s, e, m = self.parts()
        return {"prec": self.prec, "sign": s, "exp": e, "mant": hex(m)}

    @staticmethod
    def from_json(d):
        return Mpfr(d["prec"]).set_parts(d["sign"], d["exp"], int(d["mant"], 16))

    def __repr__(self):
        return "Mpfr(%d, %r)" % (self.prec, self.to_float())


class Mpf:
    """An mpf_t initialised by libgmp (mpf_init2), freed on collection."""

    def __init__(self, prec_bits, text=None):
        self.s = MpfStruct()
        mpf_init2(C.byref(self.s), int(prec_bits))
        self._live = True
        if text is not None:
            if mpf_set_str(C.byref(self.s), str(text).encode(), 10) != 0:
                raise ValueError("mpf_set_str failed on %r" % text)

    @property
    def ref(self):
        return C.byref(self.s)

    @property
    def ptr(self):
        return C.pointer(self.s)

    def parts(self):
        """-> (sign, exp_limbs, [limbs LS first])."""
        n = abs(self.s.size)
        return (0 if n == 0 else (1 if self.s.size > 0 else -1), self.s.exp,
                [self.s.d[i] for i in range(n)])

    def value_parts(self):
        """-> (sign, integer, shift) with value = sign * integer * 2^shift, trailing zero limbs stripped."""
        sg, e, l = self.parts()
        v = 0
        for i, w in enumerate(l):
            v |= w << (64 * i)
        if v == 0:
            return (0, 0, 0)
        sh = 64 * (e - len(l))
        while v & 1 == 0:
            v >>= 1
            sh += 1
        return (sg, v, sh)

    def to_float(self):
        return mpf_get_d(self.ref)

    def __del__(self):
        if getattr(self, "_live", False):
            mpf_clear(C.byref(self.s))
            self._live = False
