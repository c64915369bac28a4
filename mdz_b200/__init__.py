"""mdz_b200 -- host-side mirror of MDZ's render interface over libmdzcuda.

The product is the CUDA library (mdz_b200/csrc, built in-tree as
mdz_b200/libmdzcuda.so) behind MDZ's own C API; this package is the thin
Python binding used by the tests, bench.py and the GTK-free harness.
"""
from ._native import lib, last_error, MdzCudaError, LIB_PATH  # noqa: F401
from .render import (ImageView, Plan, render, MODE_LD, MODE_MPFR, MODE_GMP,  # noqa: F401
                     FAMILY_MANDEL, FAMILY_JULIA, MANDELBROT, BURNING_SHIP,
                     GENERALIZED_CELTIC, VARIANT, imad_peak, device_count, fallback_lines, view_supported)
