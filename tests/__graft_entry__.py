"""Driver entry points: build() compiles everything, smoke() runs one tiny render
on cuda:0 and checks it against the oracle."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.abspath(__file__))


def build():
    """Compile libmdzcuda.so for sm_100a (nvcc -gencode arch=compute_100a,code=sm_100a
    -lineinfo, see mdz_b200/csrc/Makefile), the oracle's C restatement, and -- when
    /root/reference is present -- the unmodified reference into oracle/_ref/.
    Building the checker is not using it: the product never loads oracle/."""
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "mdz_b200", "csrc")])
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "all"])
    sys.path.insert(0, ROOT)
    import mdz_b200  # noqa: F401  (raises if the library did not build / load)


def smoke():
    """One small MPFR-128 Mandelbrot render on cuda:0 through the C ABI, compared
    bit-for-bit with the oracle (the unmodified reference if oracle/_ref is there,
    else the C restatement)."""
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import numpy as np
    import mdz_b200
    from views import make_view, config2
    import oracle_check

    if mdz_b200.device_count() < 1:
        raise RuntimeError("smoke() needs a CUDA device: " + mdz_b200.last_error())
    for view in (make_view("-0.7", "0.1", "2.5", 96, 72, precision=128, depth=400),
                 config2(128, 72, 1000)):
        got = mdz_b200.render(view, (0,))
        want, kind = oracle_check.oracle_render(view)
        if not np.array_equal(got, want):
            raise AssertionError("smoke: %d pixels differ from the %s oracle"
                                 % (int((got != want).sum()), kind))
        iters = int(np.where(got > 0, got, view.depth).sum())
        print("smoke ok: %dx%d mode=%d p=%d, %d pixel-iterations, bit-exact vs %s oracle"
              % (view.real_width, view.real_height, view.mode, view.precision, iters, kind))


if __name__ == "__main__":
    build()
    if len(sys.argv) > 1 and sys.argv[1] == "smoke":
        smoke()
