"""MDZ's palette pipeline on the host: .map files, the random and channel-function
generators, rotation ("palette cycling").

Restates reference src/palette.c:142-168 (palette_read), :190-203 (palette_write),
:212-329 (palette_randomize), :332-376 (palette_apply_func), :379-400 (rotate / shift)
with the same integer and double arithmetic, so that a palette made here is entry for
entry the one MDZ would have made (tests/test_palette_cpu.py runs the unmodified
reference beside it).  A palette is a list of up to 256 packed colours 0x00BBGGRR
(src/palette.h:16-19, little endian); `offset` is MDZ's global pal_offset.  The colours
reach the GPU through Plan.set_colour / Plan.recolour (the fused epilogue and the
recolour-only kernel, mdz_b200/csrc/colour.cuh); nothing in this file computes pixels.
"""
import ctypes as C
import re

# palette function ids (src/random_palette.h:5-17)
PF_EX_RG, PF_EX_GB, PF_EX_BR, PF_ROT_RGB, PF_INV_RGB, PF_INV_R, PF_INV_G, PF_INV_B = range(8)

_LINE = re.compile(r"\s*([+-]?\d+)\s+([+-]?\d+)\s+([+-]?\d+)")       # sscanf(buf, " %d %d %d", ...)
_M32 = 0xFFFFFFFF
RAND_MAX = 2147483647


def rgb(r, g, b):
    """RGB(r,g,b) (palette.h:16) on 32-bit operands: out-of-range channels spill exactly as in C."""
    return ((r & _M32) | ((g << 8) & _M32) | ((b << 16) & _M32)) & _M32


def red(x):
    return x & 0xff


def green(x):
    return (x >> 8) & 0xff


def blue(x):
    return (x >> 16) & 0xff


def _trunc(x):
    """double -> int conversion of C (towards zero)."""
    return int(x)


def _cmod(a, b):
    """C's % on ints (truncated division)."""
    return a - b * _trunc(a / b)


class Palette:
    def __init__(self, colours=None, offset=0):
        self.colours = list(colours) if colours is not None else [0] * 256
        self.offset = offset

    @property
    def indexes(self):
        return len(self.colours)          # pal_indexes

    # ---- files (palette_read / palette_write) -----------------------------------
    @classmethod
    def parse(cls, lines):
        """Up to 256 lines " R G B"; stops at the first line that does not start with three
        integers.  Returns None when there is not a single entry (palette_read -> 0)."""
        out = []
        for ln in lines:
            if len(out) >= 256:
                break
            m = _LINE.match(ln)
            if not m:
                break
            out.append(rgb(*(int(v) for v in m.groups())))
        return cls(out) if out else None

    @classmethod
    def load(cls, path):
        with open(path, "r", errors="replace") as f:
            return cls.parse(f)

    def text(self):
        return "".join(" %d %d %d\n" % (red(c), green(c), blue(c)) for c in self.colours)

    def save(self, path):
        with open(path, "w") as f:
            f.write(self.text())

    # ---- generators ---------------------------------------------------------------
    def randomize(self, r_strength, g_strength, b_strength, r_bands, g_bands, b_bands,
                  offset=0, stripe=1, spread=1, rand=None):
        """palette_randomize: per channel, `bands * indexes` random levels joined by straight
        lines are mixed into the existing colour with weight `strength`.  `rand` yields what C's
        rand() would (default: this process's libc rand(), so srand(seed) reproduces MDZ)."""
        if rand is None:
            libc = C.CDLL(None)
            libc.rand.restype = C.c_int
            rand = libc.rand
        n = self.indexes
        st = [s if s != 0 else 0.01 for s in (r_strength, g_strength, b_strength)]
        hs = [1 - s for s in st]
        hsm = [128 * s for s in st]
        cnt = [max(_trunc(n * b), 1) for b in (r_bands, g_bands, b_bands)]
        rnd = []
        for ch in range(3):                                   # all of red's draws, then green's, then blue's
            lv = [_trunc(hsm[ch] - 255 * (float(rand()) / RAND_MAX) * st[ch]) for _ in range(cnt[ch])]
            lv.append(lv[0])
            rnd.append(lv)
        bcsize = [float(c) / n for c in cnt]
        band = [0.0, 0.0, 0.0]
        for i in range(n):
            ch_val = []
            for ch, get in enumerate((red, green, blue)):
                bnd = _trunc(band[ch])
                bdif = band[ch] - bnd
                difb = 1 - bdif
                mix = _trunc(rnd[ch][bnd] * difb + rnd[ch][bnd + 1] * bdif)
                ch_val.append(_trunc(hsm[ch] + get(self.colours[i]) * hs[ch] + mix))
                band[ch] += bcsize[ch]
            if i >= offset and _cmod(i + offset, stripe) < spread:
                self.colours[i] = rgb(*ch_val)
        return self

    def apply_func(self, func, offset=0, stripe=1, spread=1):
        """palette_apply_func: exchange / rotate / invert channels on the selected stripes."""
        for i in range(self.indexes):
            r, g, b = red(self.colours[i]), green(self.colours[i]), blue(self.colours[i])
            if i >= offset and _cmod(i + offset, stripe) < spread:
                new = {PF_EX_RG: (g, r, b), PF_EX_GB: (r, b, g), PF_EX_BR: (b, g, r), PF_ROT_RGB: (g, b, r),
                       PF_INV_RGB: (255 - r, 255 - g, 255 - b), PF_INV_R: (255 - r, g, b),
                       PF_INV_G: (r, 255 - g, b), PF_INV_B: (r, g, 255 - b)}.get(func)
                if new is not None:
                    self.colours[i] = rgb(*new)
        return self

    # ---- cycling --------------------------------------------------------------------
    def rotate_backward(self):
        self.offset -= 1
        if self.offset < 0:
            self.offset = self.indexes - 1

    def rotate_forward(self):
        self.offset += 1
        if self.offset == self.indexes:
            self.offset = 0

    def shift(self, offset):
        self.offset = _cmod(self.indexes + offset, self.indexes)
