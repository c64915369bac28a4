"""CPU checks against the golden fixtures produced by the unmodified reference
cmdline (`mdz -l file -w W -h H -A n -R out`): the harness' .mdz reader and
coords maths reproduce the rect, the C oracle reproduces raw_data, and the
oracle's colour epilogue reproduces the decoded image."""
import ctypes as C

import numpy as np
import pytest

import golden_util as G
import portpath


@pytest.mark.parametrize("name", G.names())
def test_harness_rect_equals_reference_rect(name):
    meta, raw, rgb = G.load(name)
    view, info = G.view_of(meta)
    want = G.golden_rect(meta)
    got = [view.xmin, view.xmax, view.ymax, view.width]
    for g, w, label in zip(got, want, ("xmin", "xmax", "ymax", "width")):
        assert g.parts() == w.parts(), (name, label)
    assert (view.real_height, view.real_width) == raw.shape
    assert view.depth == meta["depth"] and view.precision == meta["precision"]


@pytest.mark.parametrize("name", [n for n in G.names() if n != "test_240x180"])
def test_c_oracle_reproduces_reference_raw(name):
    meta, raw, rgb = G.load(name)
    view, info = G.view_of(meta)
    try:
        got = portpath.port_render(view)
    except NotImplementedError as ex:
        pytest.skip(str(ex))
    assert np.array_equal(got, raw), "%d pixels differ" % int((got != raw).sum())


@pytest.mark.parametrize("name", G.names())
def test_c_oracle_epilogue_reproduces_reference_image(name):
    meta, raw, rgb = G.load(name)
    view, info = G.view_of(meta)
    if info["palette"] is None:
        pytest.skip("fixture without embedded palette")
    lib = portpath.load()
    pal = np.zeros(256, dtype=np.uint32)
    pal[:len(info["palette"])] = info["palette"]
    out = np.zeros((view.user_height, view.user_width), dtype=np.uint32)
    rawc = np.ascontiguousarray(raw, dtype=np.int32)
    args = (C.c_double(info["colour_scale"]), int(info["palette_ip"]), pal.ctypes.data_as(C.c_void_p),
            len(info["palette"]), info["pal_offset"])
    if view.aa_factor == 1:
        lib.oracle_palette_apply(rawc.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p),
                                 view.user_width, 0, view.user_height, *args)
    else:
        lib.oracle_do_anti_aliasing(rawc.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p),
                                    view.user_width, view.aa_factor, 0, view.user_height, *args)
    assert np.array_equal(out, G.packed_rgb(rgb)), "%d pixels differ" % int((out != G.packed_rgb(rgb)).sum())
