"""ctypes entry into oracle/_ref/libmdzref.so -- the UNMODIFIED reference hot
path (fractal.c, frac_*.c, render_threads.c) behind oracle/ref_driver.c.
TEST INFRASTRUCTURE ONLY."""
import ctypes as C
import os
import subprocess

import numpy as np

from mdz_b200.mp import MpfrStruct, MpfStruct

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "oracle", "_ref", "libmdzref.so")


def load():
    if not os.path.exists(SO) and os.path.exists("/root/reference/src/fractal.c"):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "ref"],
                              stdout=subprocess.DEVNULL)
    if not os.path.exists(SO):
        return None
    lib = C.CDLL(SO)
    P, G = C.POINTER(MpfrStruct), C.POINTER(MpfStruct)
    lib.ref_render.argtypes = [C.c_int, C.c_long, C.c_int, C.c_int, C.c_long,
                               C.c_int, C.c_int, C.c_int,
                               P, P, P, P, G, G, G, P, P,
                               C.c_int, C.c_void_p, C.POINTER(C.c_double)]
    lib.ref_render.restype = C.c_int
    return lib


def ref_render(lib, view, threads=None):
    """Render an mdz_b200.ImageView with the reference's own pool -> (raw, seconds)."""
    threads = threads or os.cpu_count() or 2
    out = np.full((view.real_height, view.real_width), -1, dtype=np.int32)
    secs = C.c_double()

    def p(v):
        return v.ptr if v is not None else None
    ok = lib.ref_render(view.mode, view.precision, view.family, view.fractal, view.depth,
                        view.user_width, view.user_height, view.aa_factor,
                        p(view.xmin), p(view.xmax), p(view.ymax), p(view.width),
                        p(view.gxmin), p(view.gymax), p(view.gwidth),
                        p(view.julia_re), p(view.julia_im),
                        threads, out.ctypes.data_as(C.c_void_p), C.byref(secs))
    assert ok == 1
    return out, secs.value
