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
    # a host built with the "%Re" fix says so: the library converts the Julia constant of GMP mode itself (include/mdzcuda.h)
    env = dict(os.environ, MDZCUDA_RE_FORMAT="full") if meta.get("fixre") else None
    cmd = [exe, "-l", str(src), "-w", str(meta["width"]), "-h", str(meta["height"]), "-A", str(meta["aa"]), "-t", "4", "-R", str(out)]
    for attempt in range(3):
        try:
            r = subprocess.run(cmd, env=env, check=True, stdout=subprocess.DEVNULL, stderr=subprocess.PIPE, text=True,
                               cwd=str(tmp_path), timeout=120)
            break
        except subprocess.TimeoutExpired:
            # the stock binary's pool can miss a wake-up and wait for ever (the reference's BUGS:1-6); ours must not
            if attempt == 2 or os.path.basename(exe).startswith("mdz_cuda"):
                raise
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


def wide_meta(prec, w, h, depth):
    """gallery/test.mdz's settings (tests/golden/test_240x180.npz) at another precision, size and depth"""
    meta, _, _ = G.load("test_240x180")
    m = dict(meta)
    text = meta["mdz_text"].replace("precision 80\n", "precision %d\n" % prec).replace("depth 3000\n", "depth %d\n" % depth)
    assert text != meta["mdz_text"]
    m.update(mdz_text=text, width=w, height=h, precision=prec, name="wide%d" % prec)
    return m


def test_precision_2048_file_renders_on_the_gpu(tmp_path):
    """A `precision 2048` settings file through the unmodified reference's cmdline: the stock binary (its own
    pthread pool) and the one linked against libmdzcuda must write the same raw_data and the same image, and the
    drop-in must have rendered it with the CUDA kernels (no fallback line on stderr)."""
    stock, ours = os.path.join(REF, "mdz"), os.path.join(REF, "mdz_cuda")
    if not (os.path.exists(stock) and os.path.exists(ours)):
        pytest.skip("oracle/_ref binaries not built")
    meta = wide_meta(2048, 96, 72, 400)
    (tmp_path / "a").mkdir(); (tmp_path / "b").mkdir()
    want_raw, want_rgb = run_cmdline(stock, meta, tmp_path / "a")
    got_raw, got_rgb = run_cmdline(ours, meta, tmp_path / "b")
    assert np.array_equal(got_raw, want_raw), "%d raw pixels differ" % int((got_raw != want_raw).sum())
    assert np.array_equal(got_rgb, want_rgb)


def test_gmp_precision_1024_file_renders_on_the_gpu(tmp_path):
    """gallery/space_pad.mdz (GMP mpf mode) at `precision 1024`: above the one-thread mpf kernels, so the drop-in
    renders it with the lane-group kernels (coop_mpf.cuh) -- same raw_data as the stock binary, no fallback line
    (run_cmdline asserts that stderr does not mention libmdzcuda)."""
    stock, ours = os.path.join(REF, "mdz"), os.path.join(REF, "mdz_cuda")
    if not (os.path.exists(stock) and os.path.exists(ours)):
        pytest.skip("oracle/_ref binaries not built")
    meta, _, _ = G.load("space_pad_asis")
    m = dict(meta)
    text = meta["mdz_text"].replace("precision 128\n", "precision 1024\n")
    assert text != meta["mdz_text"] and "multi-rounding no" in text
    m.update(mdz_text=text, width=96, height=72, precision=1024, name="gmpwide1024")
    (tmp_path / "a").mkdir(); (tmp_path / "b").mkdir()
    want_raw, want_rgb = run_cmdline(stock, m, tmp_path / "a")
    got_raw, got_rgb = run_cmdline(ours, m, tmp_path / "b")
    assert np.array_equal(got_raw, want_raw), "%d raw pixels differ" % int((got_raw != want_raw).sum())
    assert (got_raw > 0).any()


def test_precision_beyond_the_kernels_falls_back_to_the_host_callback(tmp_path):
    """`precision 16384` is a legal setting (src/image_info.c:535: 80..99999999) with no GPU kernel: the drop-in must
    render it with MDZ's own line callback -- correct image, one line on stderr -- not report a blank one as done."""
    stock, ours = os.path.join(REF, "mdz"), os.path.join(REF, "mdz_cuda")
    if not (os.path.exists(stock) and os.path.exists(ours)):
        pytest.skip("oracle/_ref binaries not built")
    meta = wide_meta(16384, 40, 30, 200)
    (tmp_path / "a").mkdir(); (tmp_path / "b").mkdir()
    want_raw, _ = run_cmdline(stock, meta, tmp_path / "a")
    src = tmp_path / "b" / (meta["name"] + ".mdz")
    src.write_text(meta["mdz_text"])
    out = tmp_path / "b" / "o.ppm"
    r = subprocess.run([ours, "-l", str(src), "-w", "40", "-h", "30", "-A", "1", "-t", "4", "-R", str(out)],
                       check=True, stdout=subprocess.DEVNULL, stderr=subprocess.PIPE, text=True, cwd=str(tmp_path / "b"), timeout=600)
    assert "no GPU kernel" in r.stderr and "host's own line callback" in r.stderr, r.stderr[-500:]
    blob = open(str(out) + ".raw", "rb").read()
    hdr, rest = blob.split(b"\n", 1)
    raw = np.frombuffer(rest[:40 * 30 * 4], dtype=np.int32).reshape(30, 40)
    assert np.array_equal(raw, want_raw)


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
