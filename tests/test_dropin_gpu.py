"""The drop-in boundary on a GPU:
 1. oracle/_ref/mdz_cuda is the UNMODIFIED reference (main.c, cmdline.c, render.c,
    image_info.c, coords.c, palette.c ... everything except src/render_threads.c)
    linked against libmdzcuda.so.  Its cmdline render must reproduce the golden
    fixtures the stock reference produced: raw_data bit-exact and, since the
    colouring is the reference's own code running on our raw_data, the image too.
 2. tests/host_emu/rth_protocol.c drives start / stop / restart-while-rendering /
    quit on the rth_* API."""
import os
import subprocess

import numpy as np
import pytest

import golden_util as G

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "oracle", "_ref")


def run_cmdline(exe, meta, tmp_path):
    src = tmp_path / (meta["name"] + ".mdz")
    src.write_text(meta["mdz_text"])
    out = tmp_path / (meta["name"] + ".ppm")
    r = subprocess.run([exe, "-l", str(src), "-w", str(meta["width"]), "-h", str(meta["height"]),
                        "-A", str(meta["aa"]), "-t", "4", "-R", str(out)],
                       check=True, stdout=subprocess.DEVNULL, stderr=subprocess.PIPE, text=True, cwd=str(tmp_path), timeout=300)
    # the image has to come from the CUDA kernels, not from the host callback the drop-in keeps as its fallback
    assert "libmdzcuda" not in r.stderr, r.stderr[-1000:]
    blob = open(str(out) + ".raw", "rb").read()
    hdr, rest = blob.split(b"\n", 1)
    _, rw, rh, _, _ = hdr.split()
    rw, rh = int(rw), int(rh)
    raw = np.frombuffer(rest[:rw * rh * 4], dtype=np.int32).reshape(rh, rw)
    ppm = open(str(out), "rb").read().split(b"\n", 3)
    rgb = np.frombuffer(ppm[3], dtype=np.uint8).reshape(meta["height"], meta["width"], 3)
    return raw, rgb


@pytest.mark.parametrize("name", G.names())
def test_reference_cmdline_on_libmdzcuda(name, tmp_path):
    meta, raw, rgb = G.load(name)
    exe = os.path.join(REF, "mdz_cuda_fixre" if meta["fixre"] else "mdz_cuda")
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref/mdz_cuda not built")
    view, info = G.view_of(meta)
    got_raw, got_rgb = run_cmdline(exe, meta, tmp_path)
    assert np.array_equal(got_raw, raw), "%d raw pixels differ" % int((got_raw != raw).sum())
    if info["palette"] is not None:          # without an embedded palette MDZ seeds rand() from the clock
        assert np.array_equal(got_rgb, rgb)


def test_rth_protocol(tmp_path):
    exe = str(tmp_path / "rth_protocol")
    subprocess.check_call(["gcc", "-std=gnu99", "-O1", "-o", exe,
                           os.path.join(ROOT, "tests", "host_emu", "rth_protocol.c"),
                           "-L" + os.path.join(ROOT, "mdz_b200"), "-lmdzcuda",
                           "-Wl,-rpath," + os.path.join(ROOT, "mdz_b200"),
                           "-l:libmpfr.so.6", "-l:libgmp.so.10", "-lpthread"])
    outs = []
    for prec in (128, 64):
        r = subprocess.run([exe, str(prec)], capture_output=True, text=True, timeout=300)
        assert r.returncode == 0 and r.stdout.startswith("OK "), r.stdout + r.stderr
        assert "libmdzcuda" not in r.stderr, r.stderr[-1000:]
        outs.append(r.stdout)
    # the same view rendered twice must give the same checksum
    r2 = subprocess.run([exe, "128"], capture_output=True, text=True, timeout=300)
    assert r2.stdout == outs[0]
