"""Centre/size -> view rectangle, as MDZ's host does it before a render.

Host-side, O(1) per view.  Follows reference src/coords.c:204-227 (coords
precision), :255-262 (coords_set), :265-302 (coords_center_to_rect), :305-324
(coords_get_rect / coords_get_rect_gmp), :13-18 + src/my_mpfr_to_str.c
(MPFR -> mpf through decimal text) and src/image_info.c:261-267 (img->xmin..
kept at max(P, 80) bits), using the same libmpfr / libgmp calls.
"""
import ctypes as C

from .mp import Mpfr, Mpf, mpfr, gmp

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
