"""CPU-side checks of the C ABI: the library loads, exports every symbol the
headers declare, and fails loudly (no fallback) when there is no CUDA device."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols(header):
    text = open(os.path.join(ROOT, "include", header)).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b((?:mdzcuda|rth)_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    import mdz_b200
    lib = C.CDLL(mdz_b200.LIB_PATH)
    names = declared_symbols("mdzcuda.h")
    assert len(names) >= 15
    for n in names:
        assert hasattr(lib, n), n
    rth = os.path.join(ROOT, "include", "mdz_rth.h")
    if os.path.exists(rth):
        for n in declared_symbols("mdz_rth.h"):
            assert hasattr(lib, n), n


def test_python_binding_covers_header():
    from mdz_b200 import _native
    assert sorted(_native.SYMBOLS) == declared_symbols("mdzcuda.h")


def test_no_cpu_fallback_without_device():
    import mdz_b200
    from views import config2
    if mdz_b200.device_count() > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(mdz_b200.MdzCudaError):
        mdz_b200.render(config2(16, 9, 10))
    assert "CUDA" in mdz_b200.last_error() or "device" in mdz_b200.last_error()


def test_bad_arguments_are_rejected():
    import mdz_b200
    from views import config2
    v = config2(16, 9, 10)
    v.depth = 0
    with pytest.raises(mdz_b200.MdzCudaError):
        mdz_b200.Plan(v)


def test_null_plan_is_an_error_not_a_crash():
    """Every plan entry point checks its handle: 0 (or -1 for the launch counter) and an
    error text, as the reference's pool returns 0 on failure (src/render_threads.c:103-183)."""
    from mdz_b200 import _native
    lib = _native.lib
    assert lib.mdzcuda_plan_kernels_launched(None) == -1
    assert lib.mdzcuda_plan_set_parking(None, 1) == 0
    assert b"null plan" in lib.mdzcuda_last_error()


def test_view_supported_states_the_precision_limits():
    """mdzcuda_view_supported answers without a device: MDZ admits 80..99999999 bits (src/image_info.c:535), kernels
    exist for MPFR 33..8192 bits and GMP mpf to 8000 bits; everything else goes to the host's line callback behind
    rth_*  and is refused by the plain entry points with a reason."""
    import mdz_b200
    from views import make_view, SEAHORSE
    def ok(mode, prec):
        return mdz_b200.view_supported(make_view(SEAHORSE[0], SEAHORSE[1], "1e-9", 16, 12, mode=mode, precision=prec, depth=10))
    for prec in (80, 1024, 1025, 1536, 1537, 2048, 2049, 4096, 4097, 6144, 6145, 8192):
        assert ok("mpfr", prec), prec
    assert not ok("mpfr", 8193) and "8192" in mdz_b200.last_error()
    assert not ok("mpfr", 99999999)
    for prec in (80, 512, 513, 896, 897, 1344, 1345, 1856, 1857, 3904, 3905, 5952, 5953, 8000):
        assert ok("gmp", prec), prec
    assert not ok("gmp", 8001) and "8000" in mdz_b200.last_error()
    assert ok("ld", 80)
