"""ImageView / Plan / render: the render-time slice of MDZ's image_info.

ImageView carries exactly the fields the reference's line drivers read from
`image_info` when a render starts (reference src/image_info.h:63-122, read in
src/fractal.c:29-117 / :120-257 / :260-397), under the same names.
"""
import ctypes as C

import numpy as np

from . import _native as _l
from .mp import Mpfr, Mpf

MODE_LD, MODE_MPFR, MODE_GMP = 0, 1, 2
FAMILY_MANDEL, FAMILY_JULIA = 0, 1
MANDELBROT, BURNING_SHIP, GENERALIZED_CELTIC, VARIANT = 0, 1, 2, 3


class ImageView:
    """Render-time view.  xmin/xmax/ymax/width are Mpfr (as coords_get_rect
    leaves them in img->xmin.., reference render.c:34-35); gxmin/gymax/gwidth are
    Mpf (coords_get_rect_gmp, render.c:36-37); julia_re/julia_im are Mpfr."""

    def __init__(self, *, use_multi_prec=False, use_rounding=True, precision=80,
                 family=FAMILY_MANDEL, fractal=MANDELBROT, depth=300,
                 user_width=480, user_height=360, aa_factor=1,
                 xmin=None, xmax=None, ymax=None, width=None,
                 gxmin=None, gymax=None, gwidth=None,
                 julia_re=None, julia_im=None, gjulia_re=None, gjulia_im=None):
        self.use_multi_prec = use_multi_prec
        self.use_rounding = use_rounding
        self.precision = int(precision)
        self.family = family
        self.fractal = fractal
        self.depth = int(depth)
        self.user_width = int(user_width)
        self.user_height = int(user_height)
        self.aa_factor = max(1, int(aa_factor))
        self.xmin, self.xmax, self.ymax, self.width = xmin, xmax, ymax, width
        self.gxmin, self.gymax, self.gwidth = gxmin, gymax, gwidth
        self.julia_re, self.julia_im = julia_re, julia_im
        self.gjulia_re, self.gjulia_im = gjulia_re, gjulia_im      # optional mpf copy of the constant (GMP mode)

    # image_info.c:129-130
    @property
    def real_width(self):
        return self.user_width * self.aa_factor

    @property
    def real_height(self):
        return self.user_height * self.aa_factor

    @property
    def mode(self):
        # the callback image_info_set_multi_prec installs (image_info.c:238-248)
        if not self.use_multi_prec:
            return MODE_LD
        return MODE_MPFR if self.use_rounding else MODE_GMP

    def c_view(self):
        v = _l.View()
        v.mode = self.mode
        v.precision = self.precision
        v.family = self.family
        v.fractal = self.fractal
        v.depth = self.depth
        v.real_width = self.real_width
        v.real_height = self.real_height
        v.aa_factor = self.aa_factor
        for name in ("xmin", "xmax", "ymax", "width", "julia_re", "julia_im"):
            val = getattr(self, name)
            if val is not None:
                setattr(v, name, val.ptr)
        for name in ("gxmin", "gymax", "gwidth", "gjulia_re", "gjulia_im"):
            val = getattr(self, name)
            if val is not None:
                setattr(v, name, val.ptr)
        return v


def device_count():
    return _l.lib.mdzcuda_device_count()


def fallback_lines():
    """Lines the rth_* layer rendered through the host's line callback instead of the GPU (0 on a healthy run)."""
    return int(_l.lib.mdzcuda_fallback_lines())


def view_supported(view):
    """True when a GPU kernel exists for the view's mode and precision."""
    cv = view.c_view()
    return bool(_l.lib.mdzcuda_view_supported(C.byref(cv)))


def imad_peak(device=0, ms=200, wide=True):
    """Measured multiply peak, ops/s: wide=True is IMAD.WIDE.U32.X carry chains
    (32x32->64 MACs, the roofline unit), wide=False the 32-bit IMAD issue rate."""
    fn = _l.lib.mdzcuda_imad_peak if wide else _l.lib.mdzcuda_imad32_peak
    r = fn(device, ms)
    if r <= 0:
        raise _l.MdzCudaError("imad_peak: " + _l.last_error())
    return r


class Plan:
    """One device's share of a render (bands band_first, +band_stride, ...)."""

    def __init__(self, view, device=0, band_first=0, band_stride=1):
        self.view = view
        self._cv = view.c_view()
        self.h = _l.lib.mdzcuda_plan_create(C.byref(self._cv), device, band_first, band_stride)
        if not self.h:
            raise _l.MdzCudaError("plan_create: " + _l.last_error())
        self.device = device
        self.band_first, self.band_stride = band_first, band_stride

    def _chk(self, ok, what):
        if not ok:
            raise _l.MdzCudaError(what + ": " + _l.last_error())

    def tune(self, chunk_iters=0, blocks_per_sm=0):
        self._chk(_l.lib.mdzcuda_plan_tune(self.h, chunk_iters, blocks_per_sm), "plan_tune")

    def set_cycle_detection(self, on=True):
        """Exact periodicity check: same raw_data, interior pixels finish early."""
        self._chk(_l.lib.mdzcuda_plan_set_cycle_detection(self.h, 1 if on else 0), "plan_set_cycle_detection")

    def set_parking(self, mode=-1):
        """Tail compaction: -1 automatic, 0 off, 1 on.  Same raw_data either way."""
        self._chk(_l.lib.mdzcuda_plan_set_parking(self.h, mode), "plan_set_parking")

    def set_order(self, centre_out=True):
        """Sequence in which the queue visits the plan's bands: raster (default) or from the middle outwards."""
        self._chk(_l.lib.mdzcuda_plan_set_order(self.h, 1 if centre_out else 0), "plan_set_order")

    def launch(self, stream=None):
        self._chk(_l.lib.mdzcuda_plan_launch(self.h, C.c_void_p(stream or 0)), "plan_launch")

    def wait(self):
        self._chk(_l.lib.mdzcuda_plan_wait(self.h), "plan_wait")

    def cancel(self):
        self._chk(_l.lib.mdzcuda_plan_cancel(self.h), "plan_cancel")

    def bands_done(self):
        return _l.lib.mdzcuda_plan_bands_done(self.h)

    def bands_total(self):
        return _l.lib.mdzcuda_plan_bands_total(self.h)

    def local_lines(self):
        return _l.lib.mdzcuda_plan_local_lines(self.h)

    def device_raw(self):
        return _l.lib.mdzcuda_plan_device_raw(self.h)

    def fetch(self, out=None):
        v = self.view
        if out is None:
            out = np.full((v.real_height, v.real_width), -1, dtype=np.int32)
        assert out.dtype == np.int32 and out.flags["C_CONTIGUOUS"]
        assert out.shape == (v.real_height, v.real_width)
        self._chk(_l.lib.mdzcuda_plan_fetch(self.h, out.ctypes.data_as(C.c_void_p)), "plan_fetch")
        return out

    def run(self, out=None, stream=None):
        """launch + deliver: bands are copied to `out` while the kernel runs."""
        v = self.view
        if out is None:
            out = np.full((v.real_height, v.real_width), -1, dtype=np.int32)
        assert out.dtype == np.int32 and out.flags["C_CONTIGUOUS"]
        assert out.shape == (v.real_height, v.real_width)
        self._chk(_l.lib.mdzcuda_plan_run(self.h, C.c_void_p(stream or 0), out.ctypes.data_as(C.c_void_p)), "plan_run")
        return out

    def set_colour(self, palette, pal_offset=0, colour_scale=1.0, palette_ip=False):
        """Switch the fused colour epilogue on (palette: packed R|G<<8|B<<16 ints)."""
        pal = np.ascontiguousarray(np.asarray(palette, dtype=np.uint32))
        self._pal = pal
        c = _l.Colour(pal.ctypes.data_as(C.POINTER(C.c_uint32)), len(pal), int(pal_offset),
                      float(colour_scale), 1 if palette_ip else 0)
        self._chk(_l.lib.mdzcuda_plan_set_colour(self.h, C.byref(c)), "plan_set_colour")

    def recolour(self, stream=None):
        self._chk(_l.lib.mdzcuda_plan_recolour(self.h, C.c_void_p(stream or 0)), "plan_recolour")

    def fetch_rgb(self, out=None):
        v = self.view
        if out is None:
            out = np.zeros((v.user_height, v.user_width), dtype=np.uint32)
        assert out.dtype == np.uint32 and out.flags["C_CONTIGUOUS"] and out.shape == (v.user_height, v.user_width)
        self._chk(_l.lib.mdzcuda_plan_fetch_rgb(self.h, out.ctypes.data_as(C.c_void_p)), "plan_fetch_rgb")
        return out

    def kernels_launched(self):
        """Kernels this plan has launched so far (3 per render with tail compaction, else 1)."""
        return int(_l.lib.mdzcuda_plan_kernels_launched(self.h))

    def kernel_info(self):
        ki = _l.KernelInfo()
        self._chk(_l.lib.mdzcuda_plan_kernel_info(self.h, C.byref(ki)), "plan_kernel_info")
        return {n: getattr(ki, n) for n, _ in ki._fields_}

    def close(self):
        if self.h:
            _l.lib.mdzcuda_plan_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def render(view, devices=(0,)):
    """Host view in, host raw_data out (img->raw_data layout), over `devices`."""
    out = np.full((view.real_height, view.real_width), -1, dtype=np.int32)
    cv = view.c_view()
    devs = (C.c_int * len(devices))(*devices)
    ok = _l.lib.mdzcuda_render(C.byref(cv), out.ctypes.data_as(C.c_void_p), len(devices), devs)
    if not ok:
        raise _l.MdzCudaError("render: " + _l.last_error())
    return out
