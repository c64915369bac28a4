"""ctypes entry into oracle/liboracle.so -- the plain-C restatement of the hot
path (oracle/mdz_oracle.c).  TEST INFRASTRUCTURE ONLY."""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "oracle", "liboracle.so")


class OracleNum(C.Structure):
    _fields_ = [("prec", C.c_long), ("sign", C.c_int), ("exp", C.c_long),
                ("limbs", C.POINTER(C.c_uint64))]


_lib = None


def load():
    global _lib
    if _lib is None:
        src = os.path.join(ROOT, "oracle", "mdz_oracle.c")
        if not os.path.exists(SO) or os.path.getmtime(src) > os.path.getmtime(SO):
            subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "liboracle.so"],
                                  stdout=subprocess.DEVNULL)
        _lib = C.CDLL(SO)
        P = C.POINTER(OracleNum)
        _lib.oracle_render.argtypes = [C.c_int, C.c_long, C.c_int, C.c_int, C.c_long, C.c_int, C.c_int,
                                       P, P, P, P, P, P, C.c_int, C.c_void_p]
        _lib.oracle_mpfr_op.argtypes = [C.c_int, C.c_long, P, P, C.POINTER(C.c_uint64),
                                        C.POINTER(C.c_int), C.POINTER(C.c_long)]
    return _lib


def to_num(v, keep):
    """mdz_b200.mp.Mpfr -> OracleNum (keeps the limb array alive in `keep`)."""
    if v is None:
        return None
    s, e, _ = v.parts()
    arr = (C.c_uint64 * v.n)(*v.limbs())
    keep.append(arr)
    return C.pointer(OracleNum(v.prec, s, e, C.cast(arr, C.POINTER(C.c_uint64))))


def port_render(view, threads=None):
    lib = load()
    if view.mode == 2:
        return port_render_gmp(view, threads)
    threads = threads or os.cpu_count() or 1
    out = np.full((view.real_height, view.real_width), -1, dtype=np.int32)
    keep = []
    ok = lib.oracle_render(view.mode, view.precision, view.family, view.fractal, view.depth,
                           view.real_width, view.real_height,
                           to_num(view.xmin, keep), to_num(view.xmax, keep), to_num(view.ymax, keep),
                           to_num(view.width, keep), to_num(view.julia_re, keep), to_num(view.julia_im, keep),
                           threads, out.ctypes.data_as(C.c_void_p))
    assert ok == 1, "oracle_render refused the view"
    return out


class OracleMpf(C.Structure):
    _fields_ = [("sign", C.c_int), ("exp", C.c_long), ("n", C.c_int), ("limbs", C.POINTER(C.c_uint64))]


def to_mpf(v, keep):
    """mdz_b200.mp.Mpf -> OracleMpf."""
    sg, e, limbs = v.parts()
    arr = (C.c_uint64 * max(1, len(limbs)))(*limbs)
    keep.append(arr)
    return OracleMpf(sg, e, len(limbs), C.cast(arr, C.POINTER(C.c_uint64)))


def gmp_tables(view):
    """fractal_gmp_calculate_line's set-up (fractal.c:299-328) with libgmp itself: x per column, y per line."""
    from mdz_b200.mp import Mpf, mpf_set, mpf_set_si, mpf_mul, mpf_add, mpf_sub, mpf_div, mpf_ui_div, mpf_mul_ui
    p = view.precision
    rw, xmin, width, t1 = Mpf(p), Mpf(p), Mpf(p), Mpf(p)
    mpf_set_si(rw.ref, view.real_width)
    mpf_set(xmin.ref, view.gxmin.ref)
    mpf_set(width.ref, view.gwidth.ref)
    xs, ys = [], []
    for ix in range(view.real_width):
        x = Mpf(p)
        mpf_ui_div(t1.ref, ix, rw.ref)
        mpf_mul(x.ref, t1.ref, width.ref)
        mpf_add(x.ref, x.ref, xmin.ref)
        xs.append(x)
    for line in range(view.real_height):
        y = Mpf(p)
        mpf_div(t1.ref, width.ref, rw.ref)
        mpf_mul_ui(t1.ref, t1.ref, line)
        mpf_sub(y.ref, view.gymax.ref, t1.ref)
        ys.append(y)
    return xs, ys


def port_render_gmp(view, threads=None):
    from mdz_b200.mp import Mpf
    from mdz_b200.coords import mpfr_to_decimal
    lib = load()
    lib.oracle_render_gmp.argtypes = [C.c_int, C.c_int, C.c_int, C.c_long, C.c_int, C.c_int,
                                      C.POINTER(OracleMpf), C.POINTER(OracleMpf),
                                      C.POINTER(OracleMpf), C.POINTER(OracleMpf), C.c_void_p]
    xs, ys = gmp_tables(view)
    keep = []
    xa = (OracleMpf * len(xs))(*[to_mpf(v, keep) for v in xs])
    ya = (OracleMpf * len(ys))(*[to_mpf(v, keep) for v in ys])
    jre = jim = None
    if view.family == 1:
        # mpfr_to_gmp through "%.Re" (fractal.c:341-342, my_mpfr_to_str.c:68)
        # the constant as the host's conversion leaves it: the view's own mpf copy ("%Re" hosts), else "%.Re" (stock)
        gre = getattr(view, "gjulia_re", None) or Mpf(view.precision, mpfr_to_decimal(view.julia_re, False))
        gim = getattr(view, "gjulia_im", None) or Mpf(view.precision, mpfr_to_decimal(view.julia_im, False))
        jre = C.pointer(to_mpf(gre, keep))
        jim = C.pointer(to_mpf(gim, keep))
    P = (max(53, view.precision) + 127) // 64
    out = np.full((view.real_height, view.real_width), -1, dtype=np.int32)
    ok = lib.oracle_render_gmp(P, view.family, view.fractal, view.depth, view.real_width, view.real_height,
                               xa, ya, jre, jim, out.ctypes.data_as(C.c_void_p))
    assert ok == 1
    return out
