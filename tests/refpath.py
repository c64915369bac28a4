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
    if hasattr(lib, "ref_render_lines"):
        lib.ref_render_lines.argtypes = [C.c_int, C.c_long, C.c_int, C.c_int, C.c_long,
                                         C.c_int, C.c_int, C.c_int,
                                         P, P, P, P, G, G, G, P, P,
                                         C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        lib.ref_render_lines.restype = C.c_int
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


def ref_render_lines(lib, view, lines, threads=None):
    """The reference's own line driver on selected real lines of a view of any size
    -> int32 [len(lines)][real_width] (oracle/ref_driver.c: ref_render_lines)."""
    threads = threads or os.cpu_count() or 2
    lines = np.ascontiguousarray(np.asarray(lines, dtype=np.int32))
    out = np.full((len(lines), view.real_width), -1, dtype=np.int32)

    def p(v):
        return v.ptr if v is not None else None
    ok = lib.ref_render_lines(view.mode, view.precision, view.family, view.fractal, view.depth,
                              view.user_width, view.user_height, view.aa_factor,
                              p(view.xmin), p(view.xmax), p(view.ymax), p(view.width),
                              p(view.gxmin), p(view.gymax), p(view.gwidth),
                              p(view.julia_re), p(view.julia_im),
                              lines.ctypes.data_as(C.c_void_p), len(lines), threads,
                              out.ctypes.data_as(C.c_void_p))
    assert ok == 1
    return out
