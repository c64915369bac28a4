"""The Julia preview as a caller of the boundary (reference main_gui.c:28-29, 533-599, 786-793):
160x90, 2x2 anti-aliasing, restarted two hundred times a couple of milliseconds apart through rth_*, each
restart arriving while the previous frame may still be rendering.  Every frame that completed before the next
restart must be the reference's own render of that frame's constant -- a frame cut short leaves nothing behind,
and nothing of an abandoned frame leaks into the next."""
import math
import os
import subprocess

import numpy as np
import pytest

from mdz_b200 import ImageView, FAMILY_JULIA, MANDELBROT
from mdz_b200.mp import Mpfr
from refpath import ref_render_lines

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def build(tmp_path):
    exe = str(tmp_path / "rth_preview")
    subprocess.check_call(["gcc", "-std=gnu99", "-O1", "-o", exe, os.path.join(ROOT, "tests", "host_emu", "rth_preview.c"),
                           "-L" + os.path.join(ROOT, "mdz_b200"), "-lmdzcuda", "-Wl,-rpath," + os.path.join(ROOT, "mdz_b200"),
                           "-l:libmpfr.so.6", "-l:libgmp.so.10", "-lpthread", "-lm"])
    return exe


def preview_view(prec_arg, i):
    prec = prec_arg or 80

    def num(v):
        return Mpfr(prec).set_d(v)
    return ImageView(use_multi_prec=prec_arg != 0, use_rounding=True, precision=prec, family=FAMILY_JULIA, fractal=MANDELBROT,
                     depth=300, user_width=160, user_height=90, aa_factor=2,
                     xmin=num(-1.625), xmax=num(1.625), ymax=num(0.9140625), width=num(3.25),
                     julia_re=num(-0.8 + 0.25 * math.cos(0.1 * i)), julia_im=num(0.156 + 0.25 * math.sin(0.13 * i)))


@pytest.mark.parametrize("prec_arg", [0, 128], ids=["long_double", "mpfr128"])
def test_restarted_preview_frames_equal_the_reference(ref_lib, tmp_path, prec_arg):
    exe = build(tmp_path)
    out = str(tmp_path / "frames.bin")
    r = subprocess.run([exe, str(prec_arg), "200", "2500", out], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and r.stdout.startswith("OK "), r.stdout + r.stderr
    assert "libmdzcuda" not in r.stderr, r.stderr[-1000:]          # rendered by the CUDA kernels, not the host callback
    blob = np.fromfile(out, dtype=np.int32)
    rec = 1 + 320 * 180
    assert blob.size % rec == 0
    frames = blob.reshape(-1, rec)
    completed = len(frames)
    assert completed >= 20, r.stdout          # most of the slow events must have finished a frame
    lines = list(range(180))
    for f in frames:
        i = int(f[0])
        want = ref_render_lines(ref_lib, preview_view(prec_arg, i), lines)
        got = f[1:].reshape(180, 320)
        assert np.array_equal(got, want), "frame %d: %d pixels differ" % (i, int((got != want).sum()))
    print("%s: %d of 200 frames completed, all equal to the reference" % (prec_arg or "long double", completed))
