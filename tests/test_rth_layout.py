"""include/mdz_rth.h mirrors the memory layout of the reference's image_info and
rthdata (src/image_info.h:63-122, src/render_threads.h:20-32).  The reference's
offsets were printed by oracle/layout_probe.c compiled against the reference's
own headers (tests/golden/ref_layout.txt); when oracle/_ref/layout_probe is
present it is re-run as well."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def mine(tmp_path):
    exe = str(tmp_path / "layout_mine")
    subprocess.check_call(["gcc", "-std=gnu99", "-o", exe,
                           os.path.join(ROOT, "tests", "host_emu", "layout_mine.c")])
    return subprocess.check_output([exe]).decode()


def test_layout_matches_recorded_reference_layout(tmp_path):
    want = open(os.path.join(ROOT, "tests", "golden", "ref_layout.txt")).read()
    assert mine(tmp_path) == want


def test_layout_matches_live_reference_headers(tmp_path):
    probe = os.path.join(ROOT, "oracle", "_ref", "layout_probe")
    if not os.path.exists(probe):
        import pytest
        pytest.skip("oracle/_ref/layout_probe not built")
    assert mine(tmp_path) == subprocess.check_output([probe]).decode()
