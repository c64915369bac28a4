"""The .mdz writer (mdz_b200/mdzfile.py: settings_text / save_mdz) against what the
UNMODIFIED reference cmdline binary writes itself: `mdz -l file -w W -h H -L log` puts
image_info_save_settings' block (reference src/image_info.c:346-419) at the head of the log.
Every gallery file, both format versions; then a write -> read round trip."""
import glob
import os
import subprocess

import pytest

from mdz_b200.mdzfile import load_mdz, save_mdz, settings_text

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "oracle", "_ref", "mdz")
GALLERY = "/root/reference/gallery"


def gallery_files():
    return sorted(glob.glob(os.path.join(GALLERY, "*.mdz")))


@pytest.mark.skipif(not (os.path.exists(REF) and os.path.isdir(GALLERY)),
                    reason="needs the reference checkout and oracle/_ref/mdz (this container)")
@pytest.mark.parametrize("path", gallery_files(), ids=[os.path.basename(p) for p in gallery_files()])
def test_settings_block_equals_the_reference_log(path, tmp_path):
    log, png = str(tmp_path / "log"), str(tmp_path / "o.png")
    r = subprocess.run([REF, "-l", path, "-w", "40", "-h", "24", "-R", png, "-L", log],
                       capture_output=True, text=True, cwd=str(tmp_path), timeout=120)
    if not os.path.exists(log) or os.path.getsize(log) == 0:
        pytest.skip("reference could not render this file here: " + (r.stderr or r.stdout)[-200:])
    want = open(log).read()
    want = want[:want.index("render-time")]
    got = settings_text(load_mdz(path), 40, 24)
    assert got == want


def test_write_read_round_trip(tmp_path):
    golden = os.path.join(ROOT, "tests", "golden")
    import golden_util as G
    for name in G.names():
        meta, _, _ = G.load(name)
        s = G.settings_of(meta)
        out = str(tmp_path / (name + ".mdz"))
        save_mdz(out, s)
        t = load_mdz(out)
        # the writer always writes the new format; an old-style file's corners come back as centre/size
        assert (t.family, t.fractal, t.depth, t.precision) == (s.family, s.fractal, s.depth, s.precision)
        assert (t.use_multi_prec, t.palette_ip, t.pal_offset) == (s.use_multi_prec, s.palette_ip, s.pal_offset)
        assert t.colour_scale == s.colour_scale and t.palette == s.palette and t.rnd == s.rnd
        if s.center is not None:
            assert settings_text(t) == settings_text(s)
