"""Loader for the in-tree CUDA library (mdz_b200/libmdzcuda.so).

There is no fallback: if the library has not been built (run
`python -c "import __graft_entry__ as g; g.build()"` or `make -C mdz_b200/csrc`)
importing this module raises, and every render call needs a CUDA device.
"""
import ctypes as C
import os

from .mp import MpfrStruct, MpfStruct

_HERE = os.path.dirname(os.path.abspath(__file__))
# MDZCUDA_LIB: another build of the same library (A/B measurements of kernel variants, tools/)
LIB_PATH = os.environ.get("MDZCUDA_LIB") or os.path.join(_HERE, "libmdzcuda.so")


class MdzCudaError(RuntimeError):
    pass


class View(C.Structure):
    """struct mdzcuda_view (include/mdzcuda.h)."""
    _fields_ = [
        ("mode", C.c_int), ("precision", C.c_long),
        ("family", C.c_int), ("fractal", C.c_int), ("depth", C.c_long),
        ("real_width", C.c_int), ("real_height", C.c_int), ("aa_factor", C.c_int),
        ("xmin", C.POINTER(MpfrStruct)), ("xmax", C.POINTER(MpfrStruct)),
        ("ymax", C.POINTER(MpfrStruct)), ("width", C.POINTER(MpfrStruct)),
        ("gxmin", C.POINTER(MpfStruct)), ("gymax", C.POINTER(MpfStruct)),
        ("gwidth", C.POINTER(MpfStruct)),
        ("julia_re", C.POINTER(MpfrStruct)), ("julia_im", C.POINTER(MpfrStruct)),
        ("gjulia_re", C.POINTER(MpfStruct)), ("gjulia_im", C.POINTER(MpfStruct)),
    ]


class Colour(C.Structure):
    """struct mdzcuda_colour (include/mdzcuda.h)."""
    _fields_ = [("palette", C.POINTER(C.c_uint32)), ("pal_indexes", C.c_int), ("pal_offset", C.c_int),
                ("colour_scale", C.c_double), ("palette_ip", C.c_int)]


class KernelInfo(C.Structure):
    _fields_ = [(n, C.c_int) for n in (
        "limbs", "regs_per_thread", "local_bytes", "shared_bytes",
        "block_threads", "blocks_per_sm", "grid_blocks", "sm_count", "lanes_per_pixel")]


# every symbol include/mdzcuda.h declares: name -> (restype, argtypes)
SYMBOLS = {
    "mdzcuda_last_error": (C.c_char_p, []),
    "mdzcuda_device_count": (C.c_int, []),
    "mdzcuda_view_supported": (C.c_int, [C.POINTER(View)]),
    "mdzcuda_fallback_lines": (C.c_long, []),
    "mdzcuda_plan_set_fed": (C.c_int, [C.c_void_p, C.c_int]),
    "mdzcuda_plan_feed": (C.c_int, [C.c_void_p, C.POINTER(C.c_int), C.c_int, C.c_int]),
    "mdzcuda_plan_backlog": (C.c_longlong, [C.c_void_p]),
    "mdzcuda_plan_stream": (C.c_void_p, [C.c_void_p]),
    "mdzcuda_plan_set_order": (C.c_int, [C.c_void_p, C.c_int]),
    "mdzcuda_plan_create": (C.c_void_p, [C.POINTER(View), C.c_int, C.c_int, C.c_int]),
    "mdzcuda_plan_tune": (C.c_int, [C.c_void_p, C.c_int, C.c_int]),
    "mdzcuda_plan_set_cycle_detection": (C.c_int, [C.c_void_p, C.c_int]),
    "mdzcuda_plan_set_parking": (C.c_int, [C.c_void_p, C.c_int]),
    "mdzcuda_plan_kernels_launched": (C.c_int, [C.c_void_p]),
    "mdzcuda_plan_launch": (C.c_int, [C.c_void_p, C.c_void_p]),
    "mdzcuda_plan_wait": (C.c_int, [C.c_void_p]),
    "mdzcuda_plan_cancel": (C.c_int, [C.c_void_p]),
    "mdzcuda_plan_bands_done": (C.c_int, [C.c_void_p]),
    "mdzcuda_plan_bands_total": (C.c_int, [C.c_void_p]),
    "mdzcuda_plan_fetch": (C.c_int, [C.c_void_p, C.c_void_p]),
    "mdzcuda_plan_poll_bands": (C.c_int, [C.c_void_p, C.c_void_p]),
    "mdzcuda_plan_fetch_bands": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int]),
    "mdzcuda_plan_set_colour": (C.c_int, [C.c_void_p, C.POINTER(Colour)]),
    "mdzcuda_plan_recolour": (C.c_int, [C.c_void_p, C.c_void_p]),
    "mdzcuda_plan_fetch_rgb": (C.c_int, [C.c_void_p, C.c_void_p]),
    "mdzcuda_plan_device_raw": (C.c_void_p, [C.c_void_p]),
    "mdzcuda_plan_local_lines": (C.c_int, [C.c_void_p]),
    "mdzcuda_plan_kernel_info": (C.c_int, [C.c_void_p, C.POINTER(KernelInfo)]),
    "mdzcuda_plan_run": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "mdzcuda_plan_destroy": (None, [C.c_void_p]),
    "mdzcuda_trim": (None, []),
    "mdzcuda_render": (C.c_int, [C.POINTER(View), C.c_void_p, C.c_int, C.POINTER(C.c_int)]),
    "mdzcuda_imad_peak": (C.c_double, [C.c_int, C.c_int]),
    "mdzcuda_imad32_peak": (C.c_double, [C.c_int, C.c_int]),
    "mdzcuda_debug_occupy": (C.c_int, [C.c_int, C.c_int, C.c_int]),
}


def load():
    if not os.path.exists(LIB_PATH):
        raise MdzCudaError(
            "libmdzcuda.so is not built (%s missing); there is no CPU fallback. "
            "Build it with __graft_entry__.build()." % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SYMBOLS.items():
        if os.environ.get("MDZCUDA_LIB") and not hasattr(lib, name):
            continue        # an older build under A/B comparison; the in-tree library must export everything
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    return lib


class _Lib:
    """The library, mapped at the first call rather than at import: host-only users of this package
    (the coords / .mdz / palette mirror, and bench.py's reference arm, which must not map any of the
    product's code) never touch it.  A missing build still fails loudly, at the first use."""
    _h = None

    def __getattr__(self, name):
        if _Lib._h is None:
            _Lib._h = load()
        return getattr(_Lib._h, name)


lib = _Lib()


def last_error():
    return (lib.mdzcuda_last_error() or b"").decode()
