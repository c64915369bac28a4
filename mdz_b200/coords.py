"""Centre/size -> view rectangle, as MDZ's host does it before a render.

Host-side, O(1) per view.  Follows reference src/coords.c:204-227 (coords
precision), :255-262 (coords_set), :265-302 (coords_center_to_rect), :305-324
(coords_get_rect / coords_get_rect_gmp), :13-18 + src/my_mpfr_to_str.c
(MPFR -> mpf through decimal text) and src/image_info.c:261-267 (img->xmin..
kept at max(P, 80) bits), using the same libmpfr / libgmp calls.
"""
import ctypes as C

from .mp import Mpfr, Mpf, MpfrStruct, EXP_ZERO, mpfr, gmp

DEFAULT_PRECISION = 80          # coords.h:10


def coords_precision(precision):
    """coords_set_precision (coords.c:204-216)."""
    p = max(int(precision), DEFAULT_PRECISION)
    gmp_prec = 64 * ((max(53, p) + 127) // 64) - 64      # mpf_get_prec(mpf_init2(p))
    return max(p, gmp_prec)


def center_to_rect(cx, cy, size, img_width, img_height, precision):
    """-> (xmin, xmax, ymax, width) as Mpfr at max(precision, 80) bits.

    cx, cy, size: decimal strings (or Mpfr); img_width/img_height give the
    aspect (coords_set, coords.c:259: aspect = (double)w / h).
    """
    cp = coords_precision(precision)
    aspect = float(img_width) / float(img_height)
    c_cx, c_cy, c_size = Mpfr(cp, cx), Mpfr(cp, cy), Mpfr(cp, size)
    width, height, tmp = Mpfr(cp), Mpfr(cp), Mpfr(cp)
    xmin, xmax, ymin, ymax = Mpfr(cp), Mpfr(cp), Mpfr(cp), Mpfr(cp)
    if aspect > 1.0:
        mpfr.mpfr_set(width.ref, c_size.ref, 0)                 # *c->size = _size
        mpfr.mpfr_div_d(height.ref, width.ref, aspect, 0)
        mpfr.mpfr_div_ui(tmp.ref, width.ref, 2, 0)
        mpfr.mpfr_sub(xmin.ref, c_cx.ref, tmp.ref, 0)
        mpfr.mpfr_add(xmax.ref, xmin.ref, width.ref, 0)
        mpfr.mpfr_div_d(tmp.ref, tmp.ref, aspect, 0)
        mpfr.mpfr_sub(ymin.ref, c_cy.ref, tmp.ref, 0)
        mpfr.mpfr_add(ymax.ref, ymin.ref, height.ref, 0)
    else:
        mpfr.mpfr_set(height.ref, c_size.ref, 0)
        mpfr.mpfr_mul_d(width.ref, height.ref, aspect, 0)
        mpfr.mpfr_div_ui(tmp.ref, height.ref, 2, 0)
        mpfr.mpfr_sub(ymin.ref, c_cy.ref, tmp.ref, 0)
        mpfr.mpfr_add(ymax.ref, ymin.ref, height.ref, 0)
        mpfr.mpfr_mul_d(tmp.ref, tmp.ref, aspect, 0)
        mpfr.mpfr_sub(xmin.ref, c_cx.ref, tmp.ref, 0)
        mpfr.mpfr_add(xmax.ref, xmin.ref, width.ref, 0)
    ip = max(int(precision), DEFAULT_PRECISION)                 # image_info.c:261-267
    return tuple(Mpfr(ip, v) for v in (xmin, xmax, ymax, width)) + ((xmin, xmax, ymax, width),)


MAX_DP = 4096                  # my_mpfr_to_str.h


def mpfr_to_decimal(v, fixed_re=True):
    """my_mpfr_to_str (my_mpfr_to_str.c:68): mpfr_snprintf with "%.Re".  MDZ was
    written for MPFR 2.3-3.0; with MPFR 4 that literal prints ONE significant
    digit (SURVEY finding 3).  fixed_re=True uses "%Re" (all digits), which is
    what the oracle's mdz_fixre build does; fixed_re=False is bug-compatible."""
    buf = C.create_string_buffer(MAX_DP + 1)
    fmt = b"%Re" if fixed_re else b"%.Re"
    mpfr.mpfr_snprintf(buf, C.c_size_t(MAX_DP), fmt, v.ref)
    return buf.value.decode()


def rect_to_gmp(rect_coords, precision, fixed_re=True):
    """coords_get_rect_gmp: (gxmin, gymax, gwidth) as Mpf at `precision` bits,
    converted from the coords-precision rect through decimal text."""
    xmin, xmax, ymax, width = rect_coords
    ip = max(int(precision), DEFAULT_PRECISION)
    return tuple(Mpf(ip, mpfr_to_decimal(v, fixed_re)) for v in (xmin, ymax, width))


# ---------------------------------------------------------------------------
# View navigation and precision advice (reference src/coords.c:151-190, :343-456):
# the host-side state MDZ keeps per image -- centre, size, rect, precision -- and the
# operations its GUI and cmdline drive it with.  Same libmpfr calls in the same order,
# so every field is bit-identical to the reference's `struct coords` (tests/
# test_coords_nav_cpu.py runs both side by side).  O(1) per user action; nothing here
# touches the GPU.  The result feeds center_to_rect / the render boundary above.
# ---------------------------------------------------------------------------
for _n, _a in (("mpfr_div_si", [C.POINTER(MpfrStruct)] * 2 + [C.c_long, C.c_int]),
               ("mpfr_log2", [C.POINTER(MpfrStruct)] * 2 + [C.c_int]),
               ("mpfr_get_si", [C.POINTER(MpfrStruct), C.c_int])):
    getattr(mpfr, _n).argtypes = _a
mpfr.mpfr_get_si.restype = C.c_long
mpfr.mpfr_log2.restype = C.c_int


class Coords:
    """`struct coords` (coords.h:28-57) with the operations of coords.c."""

    FIELDS = ("xmin", "xmax", "ymin", "ymax", "width", "height", "cx", "cy", "_size")

    def __init__(self, img_width, img_height, init_cx=-0.5, init_cy=0.0, init_size=4.0):
        """coords_new (coords.c:76-107): everything at 80 bits, only _size is set."""
        self.precision = DEFAULT_PRECISION
        self.recommend = 0
        self.gmp_precision = 0
        for f in self.FIELDS:
            setattr(self, f, Mpfr(self.precision).set_nan())
        self.set(img_width, img_height)
        self.init_cx, self.init_cy, self.init_size = float(init_cx), float(init_cy), float(init_size)
        self._size.set_d(self.init_size)

    # *c->size aliases width or height (coords.c:228, :259)
    @property
    def size(self):
        return self.width if self.aspect > 1.0 else self.height

    def set(self, img_width, img_height):
        """coords_set (coords.c:254-261)."""
        self.img_width, self.img_height = int(img_width), int(img_height)
        self.aspect = float(self.img_width) / float(self.img_height)
        mpfr.mpfr_set(self.size.ref, self._size.ref, 0)

    def set_precision(self, precision):
        """coords_set_precision (coords.c:204-231): at least 80 bits, at least what mpf_init2
        of the same request really gives; values are carried over (precision_change, :458-468)."""
        p = coords_precision(precision)
        self.gmp_precision = 64 * ((max(53, max(int(precision), DEFAULT_PRECISION)) + 127) // 64) - 64
        for f in self.FIELDS:
            old = getattr(self, f)
            new = Mpfr(p)
            mpfr.mpfr_set(new.ref, old.ref, 0)          # tmp = RN_p(x); x = tmp
            setattr(self, f, new)
        self.precision = p

    def reset(self):
        """coords_reset (coords.c:190-199)."""
        self.set_precision(DEFAULT_PRECISION)
        self.cx.set_d(self.init_cx)
        self.cy.set_d(self.init_cy)
        self._size.set_d(self.init_size)
        mpfr.mpfr_set(self.size.ref, self._size.ref, 0)
        return self.calculate_precision()

    def calculate_precision(self):
        """coords_calculate_precision (coords.c:151-187): log2(4 / pixel size), rounded to the
        nearest integer by mpfr_get_si, plus one when mpfr_log2's ternary value says the
        logarithm itself was rounded down."""
        p = self.precision
        tmp, bail, px, prec = Mpfr(p), Mpfr(p), Mpfr(p), Mpfr(p)
        bail.set_d(4.0)
        mpfr.mpfr_div_si(px.ref, self.width.ref, self.img_width, 0)
        mpfr.mpfr_div(tmp.ref, bail.ref, px.ref, 0)
        tern = mpfr.mpfr_log2(prec.ref, tmp.ref, 0)
        r = int(mpfr.mpfr_get_si(prec.ref, 0))
        if tern < 0:
            r += 1
        self.recommend = r
        return r

    def rect_to_center(self):
        """coords_rect_to_center (coords.c:234-251)."""
        mpfr.mpfr_sub(self.width.ref, self.xmax.ref, self.xmin.ref, 0)
        mpfr.mpfr_div_d(self.height.ref, self.width.ref, self.aspect, 0)
        mpfr.mpfr_sub(self.ymin.ref, self.ymax.ref, self.height.ref, 0)
        mpfr.mpfr_add(self.cx.ref, self.xmin.ref, self.xmax.ref, 0)
        mpfr.mpfr_div_ui(self.cx.ref, self.cx.ref, 2, 0)
        mpfr.mpfr_add(self.cy.ref, self.ymin.ref, self.ymax.ref, 0)
        mpfr.mpfr_div_ui(self.cy.ref, self.cy.ref, 2, 0)

    def center_to_rect(self):
        """coords_center_to_rect (coords.c:265-302)."""
        tmp = Mpfr(self.precision)
        if self.aspect > 1.0:
            mpfr.mpfr_div_d(self.height.ref, self.width.ref, self.aspect, 0)
            mpfr.mpfr_div_ui(tmp.ref, self.width.ref, 2, 0)
            mpfr.mpfr_sub(self.xmin.ref, self.cx.ref, tmp.ref, 0)
            mpfr.mpfr_add(self.xmax.ref, self.xmin.ref, self.width.ref, 0)
            mpfr.mpfr_div_d(tmp.ref, tmp.ref, self.aspect, 0)
            mpfr.mpfr_sub(self.ymin.ref, self.cy.ref, tmp.ref, 0)
            mpfr.mpfr_add(self.ymax.ref, self.ymin.ref, self.height.ref, 0)
        else:
            mpfr.mpfr_mul_d(self.width.ref, self.height.ref, self.aspect, 0)
            mpfr.mpfr_div_ui(tmp.ref, self.height.ref, 2, 0)
            mpfr.mpfr_sub(self.ymin.ref, self.cy.ref, tmp.ref, 0)
            mpfr.mpfr_add(self.ymax.ref, self.ymin.ref, self.height.ref, 0)
            mpfr.mpfr_mul_d(tmp.ref, tmp.ref, self.aspect, 0)
            mpfr.mpfr_sub(self.xmin.ref, self.cx.ref, tmp.ref, 0)
            mpfr.mpfr_add(self.xmax.ref, self.xmin.ref, self.width.ref, 0)

    def set_rect(self, xmin, xmax, ymax):
        """coords_set_rect (coords.c:327-333)."""
        mpfr.mpfr_set(self.xmin.ref, xmin.ref, 0)
        mpfr.mpfr_set(self.xmax.ref, xmax.ref, 0)
        mpfr.mpfr_set(self.ymax.ref, ymax.ref, 0)
        self.rect_to_center()

    def to(self, cx, cy):
        """coords_to (coords.c:336-340)."""
        mpfr.mpfr_set(self.cx.ref, cx.ref, 0)
        mpfr.mpfr_set(self.cy.ref, cy.ref, 0)

    def set_size(self, size):
        """coords_size (coords.c:367-372)."""
        mpfr.mpfr_set(self._size.ref, size.ref, 0)
        mpfr.mpfr_set(self.size.ref, size.ref, 0)
        return self.calculate_precision()

    def pixel_to_coord(self, x=None, y=None):
        """coords_pixel_to_coord (coords.c:432-455), in place on Mpfr values holding pixel
        indices; the temporaries take x's precision, as in the reference (so x is required)."""
        tp = x.prec
        tmp, tmp2 = Mpfr(tp), Mpfr(tp)
        if y is not None:
            mpfr.mpfr_div_si(tmp2.ref, self.width.ref, self.img_width, 0)
            mpfr.mpfr_mul(tmp.ref, tmp2.ref, y.ref, 0)
            mpfr.mpfr_sub(y.ref, self.ymax.ref, tmp.ref, 0)
        if x is not None:
            mpfr.mpfr_div_si(tmp2.ref, x.ref, self.img_width, 0)
            mpfr.mpfr_mul(tmp.ref, tmp2.ref, self.width.ref, 0)
            mpfr.mpfr_add(x.ref, tmp.ref, self.xmin.ref, 0)

    def center_to(self, px, py):
        """coords_center_to (coords.c:375-392)."""
        cx, cy = Mpfr(self.precision, int(px)), Mpfr(self.precision, int(py))
        self.pixel_to_coord(cx, cy)
        mpfr.mpfr_set(self.cx.ref, cx.ref, 0)
        mpfr.mpfr_set(self.cy.ref, cy.ref, 0)

    def zoom(self, multiplier):
        """coords_zoom (coords.c:343-356)."""
        mpfr.mpfr_mul_d(self._size.ref, self._size.ref, float(multiplier), 0)
        mpfr.mpfr_set(self.size.ref, self._size.ref, 0)
        if self.aspect > 1.0:
            mpfr.mpfr_div_d(self.height.ref, self.width.ref, self.aspect, 0)
        else:
            mpfr.mpfr_mul_d(self.width.ref, self.height.ref, self.aspect, 0)
        self.calculate_precision()
        self.center_to_rect()

    def zoom_to(self, px, py, pw):
        """coords_zoom_to (coords.c:359-364): centre on a pixel, shrink to a box pw pixels wide."""
        z = float(pw) / self.img_width
        self.center_to(px, py)
        self.zoom(z)

    def reposition(self, p1x, p1y, p2x, p2y):
        """coords_reposition (coords.c:395-429): move the centre by the vector p1 -> p2."""
        p = self.precision
        a, b, c, d = Mpfr(p, int(p1x)), Mpfr(p, int(p1y)), Mpfr(p, int(p2x)), Mpfr(p, int(p2y))
        self.pixel_to_coord(a, b)
        self.pixel_to_coord(c, d)
        pxd, pyd = Mpfr(p), Mpfr(p)
        mpfr.mpfr_sub(pxd.ref, c.ref, a.ref, 0)
        mpfr.mpfr_sub(pyd.ref, d.ref, b.ref, 0)
        mpfr.mpfr_add(self.cx.ref, self.cx.ref, pxd.ref, 0)
        mpfr.mpfr_add(self.cy.ref, self.cy.ref, pyd.ref, 0)

    def limb_count(self, mode="mpfr"):
        """The smallest kernel this view needs: 32-bit limbs for the recommended precision
        (SURVEY 8f-4).  Long double (2 limbs) while 64 bits are enough."""
        need = max(self.recommend, 1)
        if need <= 64:
            return 2
        if mode == "gmp":
            return 2 * ((max(53, need) + 127) // 64 + 1)
        return (need + 31) // 32

    def state(self):
        """Every numeric field as comparable data (tests)."""
        out = {f: getattr(self, f).parts() if getattr(self, f).s.exp != EXP_ZERO + 1 else "nan" for f in self.FIELDS}
        out.update(precision=self.precision, recommend=self.recommend, aspect=self.aspect)
        return out
