import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


# Test hook of libmdzcuda (mdzcuda.cu, mdzcuda_plan_launch): every launch first fills its
# iteration buffer with 0x7f7f7f7f, so a band handed to the host before the kernel has
# completed it is a deterministic mismatch instead of plausible stale data.
os.environ.setdefault("MDZCUDA_DEBUG_POISON", "1")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_collection_modifyitems(config, items):
    """A stuck test (a rendezvous that never completes, a pool that never signals) fails after ten
    minutes instead of holding the whole run; needs pytest-timeout, a no-op without it."""
    if not config.pluginmanager.hasplugin("timeout"):
        return
    for item in items:
        if item.get_closest_marker("timeout") is None:
            item.add_marker(pytest.mark.timeout(600))


@pytest.fixture(scope="session")
def emu_lib():
    """Host build of the device arithmetic headers (tests/host_emu): test-only."""
    import ctypes as C
    src = os.path.join(ROOT, "tests", "host_emu", "emu.cpp")
    out = os.path.join(ROOT, "tests", "host_emu", "libmdzemu.so")
    deps = [src] + [os.path.join(ROOT, "mdz_b200", "csrc", f)
                    for f in ("limb_ops.cuh", "mpfr_sf.cuh", "escape_step.cuh", "ld64_step.cuh", "mp_convert.h", "mpf_sf.cuh", "mpf_fast.cuh")]
    deps = [d for d in deps if os.path.exists(d)]
    if not os.path.exists(out) or any(os.path.getmtime(d) > os.path.getmtime(out) for d in deps):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-o", out, src])
    return C.CDLL(out)


@pytest.fixture(scope="session")
def ref_lib():
    """The unmodified reference hot path (oracle/_ref/libmdzref.so).  Built in the
    container from /root/reference; on the GPU box the prebuilt file is used."""
    import refpath
    lib = refpath.load()
    if lib is None:
        pytest.skip("oracle/_ref/libmdzref.so not available")
    return lib


@pytest.fixture(autouse=True)
def _no_host_fallback(request):
    """Parity evidence must come from the CUDA kernels: after every GPU test the library's count of lines
    rendered through the host's line callback (mdzcuda_fallback_lines, rth.cpp) has to be zero."""
    yield
    if request.node.get_closest_marker("gpu") is not None:
        import mdz_b200
        assert mdz_b200.fallback_lines() == 0, "the rth_* layer fell back to the host callback"
