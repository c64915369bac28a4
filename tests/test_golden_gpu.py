"""GPU: libmdzcuda through its C ABI against the golden fixtures produced by the
unmodified reference cmdline (tests/golden/make_golden.py).  Bit-exact."""
import numpy as np
import pytest

import golden_util as G
import mdz_b200

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", G.names())
def test_cuda_reproduces_reference_raw(name):
    meta, raw, rgb = G.load(name)
    view, info = G.view_of(meta)
    try:
        got = mdz_b200.render(view)
    except mdz_b200.MdzCudaError as ex:
        if "GMP" in str(ex):
            pytest.skip(str(ex))
        raise
    bad = int((got != raw).sum())
    assert bad == 0, "%d of %d pixels differ" % (bad, raw.size)


@pytest.mark.parametrize("name", ["test", "ship_ld_aa3", "celtic_mpfr128_aa2"])
def test_cmdline_harness_png_equals_reference_image(name, tmp_path):
    """`python -m mdz_b200 -l file -w W -h H -A n -R out.png`: decoded PNG == the image
    the reference's cmdline render produced."""
    from mdz_b200.__main__ import main
    from mdz_b200.png import read_png_rgb8
    meta, raw, rgb = G.load(name)
    src = tmp_path / (name + ".mdz")
    src.write_text(meta["mdz_text"])
    out = str(tmp_path / "out.png")
    assert main(["-l", str(src), "-w", str(meta["width"]), "-h", str(meta["height"]),
                 "-A", str(meta["aa"]), "-R", out, "--gpus", "1"]) == 0
    assert np.array_equal(read_png_rgb8(out), rgb)
