"""The drop-in's host path, on a box WITHOUT a GPU: oracle/_ref/mdz_cuda is the unmodified reference
(everything except src/render_threads.c) linked against libmdzcuda.so.  With no CUDA device the rth_*
layer must render with the line callback MDZ installed (SURVEY 8b "keep the callback as the CPU
fallback"; reference src/render_threads.c:375-377), say so on stderr, and reproduce the golden
fixtures the stock reference produced -- instead of reporting a cleared raw_data as an image.
(On the GPU box the same binary is run by tests/test_dropin_gpu.py, which asserts the opposite:
no fallback line.)"""
import os
import subprocess

import numpy as np
import pytest

import golden_util as G

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "oracle", "_ref")


def run_cmdline(exe, meta, tmp_path, env=None, threads=4):
    src = tmp_path / (meta["name"] + ".mdz")
    src.write_text(meta["mdz_text"])
    out = tmp_path / (meta["name"] + ".ppm")
    r = subprocess.run([exe, "-l", str(src), "-w", str(meta["width"]), "-h", str(meta["height"]),
                        "-A", str(meta["aa"]), "-t", str(threads), "-R", str(out)],
                       stdout=subprocess.DEVNULL, stderr=subprocess.PIPE, text=True, cwd=str(tmp_path), timeout=600,
                       env=dict(os.environ, **(env or {})))
    assert r.returncode == 0, r.stderr[-2000:]
    blob = open(str(out) + ".raw", "rb").read()
    hdr, rest = blob.split(b"\n", 1)
    _, rw, rh, _, _ = hdr.split()
    rw, rh = int(rw), int(rh)
    raw = np.frombuffer(rest[:rw * rh * 4], dtype=np.int32).reshape(rh, rw)
    ppm = open(str(out), "rb").read().split(b"\n", 3)
    rgb = np.frombuffer(ppm[3], dtype=np.uint8).reshape(meta["height"], meta["width"], 3)
    return raw, rgb, r.stderr


# one fixture per line driver (long double / MPFR / GMP), one with anti-aliasing
@pytest.mark.parametrize("name", ["cfg2_ld", "test_240x180", "space_pad_fixre", "celtic_mpfr128_aa2"])
def test_reference_cmdline_falls_back_to_its_own_callback(name, tmp_path):
    import mdz_b200
    meta, raw, rgb = G.load(name)
    exe = os.path.join(REF, "mdz_cuda_fixre" if meta["fixre"] else "mdz_cuda")
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref/mdz_cuda not built")
    env = {}
    if mdz_b200.device_count() > 0:
        env["MDZCUDA_FORCE_HOST"] = "1"      # on a GPU box: exercise the same path through the test hook
    got_raw, got_rgb, err = run_cmdline(exe, meta, tmp_path, env)
    assert "host's own line callback" in err, err[-500:]
    assert np.array_equal(got_raw, raw), "%d raw pixels differ" % int((got_raw != raw).sum())
    _, info = G.view_of(meta)
    if info["palette"] is not None:
        assert np.array_equal(got_rgb, rgb)
