"""Palette pipeline (mdz_b200/palette.py) entry for entry against the UNMODIFIED reference
palette code (oracle/_ref/libmdzpal.so = reference src/palette.c + src/globals.c): .map
reading and writing, palette_randomize under the same libc rand() stream, every channel
function, rotation."""
import ctypes as C
import os
import random

import pytest

from mdz_b200.palette import Palette, PF_EX_RG, PF_INV_B

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "oracle", "_ref", "libmdzpal.so")
pytestmark = pytest.mark.skipif(not os.path.exists(LIB), reason="oracle/_ref/libmdzpal.so not built (make -C oracle)")


class RandomPalette(C.Structure):       # src/random_palette.h:20-32
    _fields_ = [(n, C.c_double) for n in ("r_strength", "g_strength", "b_strength", "r_bands", "g_bands", "b_bands")] + \
               [(n, C.c_int) for n in ("offset", "stripe", "spread")]


class FunctionPalette(C.Structure):     # src/random_palette.h:35-42
    _fields_ = [(n, C.c_int) for n in ("func", "offset", "stripe", "spread")]


@pytest.fixture(scope="module")
def ref():
    lib = C.CDLL(LIB)
    lib.palette_init()                  # allocates the global table (no default.map on this box: returns 0)
    lib.palette_load.argtypes = [C.c_char_p]
    lib.palette_save.argtypes = [C.c_char_p]
    return lib


def ref_colours(lib):
    n = C.c_int.in_dll(lib, "pal_indexes").value
    tab = C.POINTER(C.c_uint32).in_dll(lib, "palette")
    return [tab[i] for i in range(n)]


def write_map(path, rng, n, junk=False):
    with open(path, "w") as f:
        for i in range(n):
            f.write(" %d %d %d%s\n" % (rng.randrange(256), rng.randrange(256), rng.randrange(256),
                                       "  trailing words" if junk and i % 7 == 3 else ""))
        if junk:
            f.write("not a colour\n 1 2 3\n")


@pytest.mark.parametrize("n,junk", [(256, False), (256, True), (17, True), (300, False), (1, False)])
def test_map_files_read_and_written_like_the_reference(ref, tmp_path, n, junk):
    rng = random.Random(n * 2 + junk)
    src = str(tmp_path / "in.map")
    write_map(src, rng, n, junk)
    assert ref.palette_load(src.encode()) == 1
    mine = Palette.load(src)
    assert mine.colours == ref_colours(ref)
    out_ref, out_mine = str(tmp_path / "ref.map"), str(tmp_path / "mine.map")
    assert ref.palette_save(out_ref.encode()) == 1
    mine.save(out_mine)
    assert open(out_ref).read() == open(out_mine).read()


def test_empty_file_is_refused(ref, tmp_path):
    p = str(tmp_path / "empty.map")
    open(p, "w").write("nothing here\n")
    assert ref.palette_load(p.encode()) == 0
    assert Palette.load(p) is None


@pytest.mark.parametrize("seed", range(12))
def test_randomize_matches_reference_under_the_same_rand_stream(ref, tmp_path, seed):
    rng = random.Random(1000 + seed)
    n = rng.choice([256, 256, 64, 100])
    src = str(tmp_path / "in.map")
    write_map(src, rng, n)
    assert ref.palette_load(src.encode()) == 1
    mine = Palette.load(src)
    libc = C.CDLL(None)
    for rnd in range(4):                       # repeated application, as the GUI's "randomize" button
        rp = RandomPalette(rng.choice([0.0, 0.3, 1.0, rng.random()]), rng.random(), rng.random(),
                           rng.choice([0.01, 0.05, 0.5, 1.0]), rng.random() * 0.3, rng.random(),
                           rng.choice([0, 0, 10]), rng.choice([1, 2, 8]), rng.choice([1, 1, 3]))
        args = (rp.r_strength, rp.g_strength, rp.b_strength, rp.r_bands, rp.g_bands, rp.b_bands,
                rp.offset, rp.stripe, rp.spread)
        libc.srand(seed * 10 + rnd)
        ref.palette_randomize(C.byref(rp))
        libc.srand(seed * 10 + rnd)
        mine.randomize(*args)
        assert mine.colours == ref_colours(ref), (seed, rnd)


@pytest.mark.parametrize("func", range(PF_EX_RG, PF_INV_B + 1))
def test_channel_functions_match_reference(ref, tmp_path, func):
    rng = random.Random(50 + func)
    src = str(tmp_path / "in.map")
    write_map(src, rng, 256)
    assert ref.palette_load(src.encode()) == 1
    mine = Palette.load(src)
    for offset, stripe, spread in ((0, 1, 1), (5, 4, 2), (100, 16, 16), (0, 3, 1)):
        fp = FunctionPalette(func, offset, stripe, spread)
        ref.palette_apply_func(C.byref(fp))
        mine.apply_func(func, offset, stripe, spread)
        assert mine.colours == ref_colours(ref)


def test_rotation_and_shift_match_reference(ref, tmp_path):
    src = str(tmp_path / "in.map")
    write_map(src, random.Random(3), 200)
    assert ref.palette_load(src.encode()) == 1
    mine = Palette.load(src)
    off = C.c_int.in_dll(ref, "pal_offset")
    off.value = 0
    rng = random.Random(4)
    for _ in range(700):
        k = rng.randrange(3)
        if k == 0:
            ref.palette_rotate_forward(); mine.rotate_forward()
        elif k == 1:
            ref.palette_rotate_backward(); mine.rotate_backward()
        else:
            v = rng.randrange(-450, 450)
            ref.palette_shift(v); mine.shift(v)
        assert mine.offset == off.value
