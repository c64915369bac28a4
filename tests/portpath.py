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


def port_render_gmp(view, threads=None):
    raise NotImplementedError("GMP mpf mode of the C oracle")
