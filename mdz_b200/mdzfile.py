"""Reader for MDZ's `.mdz` settings files and the cmdline path from file to view.

GTK-free restatement of the host side that feeds the render pool:
  - file format: reference src/mdzfileio.c:15-64 (line cleaning, `#` comments),
    src/image_info.c:456-700 (key order, both format versions),
    src/palette.c:142-168 (embedded palette: up to 256 " R G B" lines);
  - cmdline semantics: src/main.c:27-63 (init_misc: -w/-h/-A override, then
    image_info_set -> coords_set), src/coords.c:255-262, :327-333, :366-371.

The old-style quirk is reproduced on purpose (bug_compatible=True): a file that
gives xmin/xmax/ymax never updates coords' private `_size`, so the second
image_info_set resets the width to the initial 4.0 (SURVEY finding 4).
"""
import re

from .coords import mpfr_to_decimal, center_to_rect, coords_precision, rect_to_gmp
from .mp import Mpfr, mpfr
from .render import (ImageView, FAMILY_MANDEL, FAMILY_JULIA, MANDELBROT, BURNING_SHIP,
                     GENERALIZED_CELTIC, VARIANT)

FAMILY_STR = ["mandelbrot", "julia"]                                       # fractal.c:11-16
FRACTAL_STR = ["mandelbrot", "burning ship", "generalized celtic", "mandel-celtic hybrid"]  # fractal.c:19-26
DEFAULT_WIDTH = 480                                                        # image_info.h:22


class MdzFileError(ValueError):
    pass


class MdzSettings:
    def __init__(self):
        self.version = (0, 0, 8)
        self.family = FAMILY_MANDEL
        self.fractal = MANDELBROT
        self.depth = 300
        self.aspect = 4.0 / 3.0
        self.colour_scale = 1.0
        self.palette_ip = False
        self.use_multi_prec = False
        self.use_rounding = True
        self.precision = 80
        self.center = None        # (cx, cy, size) decimal strings, or
        self.rect = None          # (xmin, xmax, ymax) decimal strings
        self.julia = None         # (re, im) decimal strings
        self.pal_offset = 0
        self.palette = None       # list of packed R | G<<8 | B<<16 (palette.h:16)
        self.palette_file = None
        # random-palette parameters (random_palette.h:20-31; defaults of image_info.c:290-299)
        self.rnd = dict(r_strength=1.0, r_bands=0.05, g_strength=1.0, g_bands=0.08,
                        b_strength=1.0, b_bands=0.2, offset=0, stripe=1, spread=1)


def _clean(line):
    # mdzfileio.c:15-35: tabs -> spaces, cut at the first control character
    out = []
    for ch in line:
        if ch == "\t":
            ch = " "
        if ord(ch) < 32:
            break
        out.append(ch)
    return "".join(out).strip()


def load_mdz(path):
    with open(path, "r", errors="replace") as f:
        raw_lines = f.readlines()
    lines = []
    for ln in raw_lines:
        c = _clean(ln)
        if c and not c.startswith("#"):
            lines.append(c)
    if not lines or not lines[0].startswith("mdz fractal settings"):
        raise MdzFileError("not an mdz settings file")
    s = MdzSettings()
    m = re.match(r"mdz fractal settings\s+(\d+)\.(\d+)\.(\d+)", lines[0])
    if m:
        s.version = tuple(int(x) for x in m.groups())
    new_style = not (s.version[0] == 0 and s.version[1] == 0)
    i = 1
    kv = {}
    order = []
    pal_start = None
    while i < len(lines):
        ln = lines[i]
        if ln == "settings":
            i += 1
            continue
        if ln == "palette":
            pal_start = i + 1
            break
        key, _, val = ln.partition(" ")
        # multi-word values ("burning ship") stay intact
        kv[key] = val.strip()
        order.append(key)
        i += 1

    def need(key):
        if key not in kv:
            raise MdzFileError("missing %s setting" % key)
        return kv[key]

    def index(key, table):
        v = need(key)
        if v not in table:
            raise MdzFileError("error in %s setting: %r" % (key, v))
        return table.index(v)

    if new_style:
        s.family = index("family", FAMILY_STR)
        s.fractal = index("fractal", FRACTAL_STR)
    else:
        s.family = index("fractal", FAMILY_STR)
        s.fractal = MANDELBROT
    s.depth = int(need("depth"))
    if not (1 <= s.depth <= 2147483647):
        raise MdzFileError("depth out of range")
    s.aspect = float(need("aspect"))
    s.colour_scale = float(need("colour-scale"))
    s.palette_ip = index("colour-interpolate", ["no", "yes"]) == 1
    if new_style:
        s.use_multi_prec = index("multi-precision", ["no", "yes"]) == 1
        s.use_rounding = index("multi-rounding", ["no", "yes"]) == 1
    else:
        s.use_multi_prec = index("mpfr", ["no", "yes"]) == 1
        s.use_rounding = True                     # image_info.c:519
    s.precision = int(need("precision"))
    if not (80 <= s.precision <= 99999999):
        raise MdzFileError("precision out of range")
    if "cx" in kv:
        s.center = (need("cx"), need("cy"), need("size"))
    elif "xmin" in kv:
        s.rect = (need("xmin"), need("xmax"), need("ymax"))
    else:
        raise MdzFileError("error in coordinates setting")
    if s.family == FAMILY_JULIA:
        s.julia = (need("julia-real"), need("julia-imag"))
    s.pal_offset = int(kv.get("palette-offset", "0"))
    for key, name, conv in (("r-strength", "r_strength", float), ("r-bands", "r_bands", float),
                            ("g-strength", "g_strength", float), ("g-bands", "g_bands", float),
                            ("b-strength", "b_strength", float), ("b-bands", "b_bands", float),
                            ("rnd-offset", "offset", int), ("rnd-stripe", "stripe", int), ("rnd-spread", "spread", int)):
        if key in kv:
            s.rnd[name] = conv(kv[key])
    if pal_start is not None and pal_start < len(lines):
        if lines[pal_start] == "data":
            pal = []
            for ln in lines[pal_start + 1:]:
                if len(pal) >= 256:
                    break
                parts = ln.split()
                if len(parts) != 3:
                    break
                try:
                    r, g, b = (int(x) for x in parts)
                except ValueError:
                    break
                pal.append((r & 0xFFFFFFFF) | ((g << 8) & 0xFFFFFFFF) | ((b << 16) & 0xFFFFFFFF))
            s.palette = pal
        elif lines[pal_start].startswith("file "):
            s.palette_file = lines[pal_start][5:].strip()
    return s


def view_from_settings(s, width=None, height=None, aa=1, aspect_opt=0.0,
                       bug_compatible=True, fixed_re=True):
    """What `mdz -l file -w W -h H -A aa -R out.png` hands to the render pool.

    Returns (ImageView, info) where info carries colour_scale / palette_ip /
    pal_offset / palette for the colour epilogue.
    """
    P = s.precision
    cp = coords_precision(P)
    # ---- load time: image is 480 x (480/aspect) (image_info.c:665-669) ----
    w0 = DEFAULT_WIDTH
    h0 = int(w0 / s.aspect)
    aspect0 = float(w0) / h0
    if s.center is not None:
        cx, cy = Mpfr(cp, Mpfr(P, s.center[0])), Mpfr(cp, Mpfr(P, s.center[1]))
        size = Mpfr(cp, Mpfr(P, s.center[2]))
    else:
        # coords_set_rect -> coords_rect_to_center (coords.c:229-249, :327-333)
        xmin, xmax, ymax = (Mpfr(cp, Mpfr(P, t)) for t in s.rect)
        wd, ht, ymin, cx, cy = Mpfr(cp), Mpfr(cp), Mpfr(cp), Mpfr(cp), Mpfr(cp)
        mpfr.mpfr_sub(wd.ref, xmax.ref, xmin.ref, 0)
        mpfr.mpfr_div_d(ht.ref, wd.ref, aspect0, 0)
        mpfr.mpfr_sub(ymin.ref, ymax.ref, ht.ref, 0)
        mpfr.mpfr_add(cx.ref, xmin.ref, xmax.ref, 0)
        mpfr.mpfr_div_ui(cx.ref, cx.ref, 2, 0)
        mpfr.mpfr_add(cy.ref, ymin.ref, ymax.ref, 0)
        mpfr.mpfr_div_ui(cy.ref, cy.ref, 2, 0)
        if bug_compatible:
            size = Mpfr(cp, 4.0)                  # `_size` never left its initial value
        else:
            size = wd if aspect0 > 1.0 else ht
    # ---- init_misc (main.c:27-63): the command line overrides the size ----
    if not width and not height:
        width = w0
        height = int(w0 / aspect_opt) if aspect_opt else h0
    elif not width:
        width = int(height * (aspect_opt if aspect_opt else aspect0))
    elif not height:
        height = int(width / (aspect_opt if aspect_opt else aspect0))
    aa = max(1, aa)
    xmin, xmax, ymax, wdt, crect = center_to_rect(cx, cy, size, width, height, P)
    view = ImageView(use_multi_prec=s.use_multi_prec, use_rounding=s.use_rounding, precision=P,
                     family=s.family, fractal=s.fractal, depth=s.depth,
                     user_width=width, user_height=height, aa_factor=aa,
                     xmin=xmin, xmax=xmax, ymax=ymax, width=wdt)
    if s.use_multi_prec and not s.use_rounding:
        view.gxmin, view.gymax, view.gwidth = rect_to_gmp(crect, P, fixed_re)
    if s.family == FAMILY_JULIA:
        ip = max(P, 80)
        # image_info.c:703-704: rounded into img->u.julia at the image precision
        # (c_im keeps its previous precision, image_info.c:271-272 -- 80 bits by default)
        view.julia_re = Mpfr(ip, Mpfr(P, s.julia[0]))
        view.julia_im = Mpfr(ip if not bug_compatible else 80, Mpfr(P, s.julia[1]))
        if view.mode == 2 and fixed_re:
            # GMP mode converts the constant to mpf through decimal text per pixel (fractal.c:341-342, coords.c:13-18);
            # a host with the "%Re" fix hands the library the mpf values it would have got (include/mdzcuda.h: gjulia_*)
            from .mp import Mpf
            view.gjulia_re = Mpf(P, mpfr_to_decimal(view.julia_re, True))       # mpf_init2(c_re, img->precision), fractal.c:286
            view.gjulia_im = Mpf(P, mpfr_to_decimal(view.julia_im, True))
    info = dict(colour_scale=s.colour_scale, palette_ip=s.palette_ip, pal_offset=s.pal_offset,
                palette=s.palette, palette_file=s.palette_file)
    return view, info


def view_from_mdz(path, width=None, height=None, aa=1, **kw):
    return view_from_settings(load_mdz(path), width, height, aa, **kw)


# ---------------------------------------------------------------------------
# writer
# ---------------------------------------------------------------------------
FILE_HEADER = "mdz fractal settings"            # image_info.c:20
VERSION = "0.1.2"                                # src/Makefile:1


def settings_text(s, width=None, height=None, aspect_opt=0.0, bug_compatible=True, rect=None):
    """The block image_info_save_settings writes (reference src/image_info.c:346-419) for
    settings `s` as the cmdline path holds them after init_misc: what `mdz -l file -w W -h H
    -L log` puts at the head of its log.  The centre and size are printed at the coords
    precision with every digit (mpfr_out_str base 10, n = 0); the active reference (centre or
    corners) is the uncommented one.  On the cmdline path the log is written before the first
    coords_get_rect, so img->xmin/xmax/ymax are still NaN (render.c:21-22 precedes :34-37);
    pass rect=(xmin, xmax, ymax) as Mpfr to write the values the GUI's save would."""
    P = s.precision
    cp = coords_precision(P)
    w0 = DEFAULT_WIDTH
    h0 = int(w0 / s.aspect)
    aspect0 = float(w0) / h0
    if s.center is not None:
        cx, cy = Mpfr(cp, Mpfr(P, s.center[0])), Mpfr(cp, Mpfr(P, s.center[1]))
        size = Mpfr(cp, Mpfr(P, s.center[2]))
    else:
        xmin, xmax, ymax = (Mpfr(cp, Mpfr(P, t)) for t in s.rect)
        wd, ht, ymin, cx, cy = Mpfr(cp), Mpfr(cp), Mpfr(cp), Mpfr(cp), Mpfr(cp)
        mpfr.mpfr_sub(wd.ref, xmax.ref, xmin.ref, 0)
        mpfr.mpfr_div_d(ht.ref, wd.ref, aspect0, 0)
        mpfr.mpfr_sub(ymin.ref, ymax.ref, ht.ref, 0)
        mpfr.mpfr_add(cx.ref, xmin.ref, xmax.ref, 0)
        mpfr.mpfr_div_ui(cx.ref, cx.ref, 2, 0)
        mpfr.mpfr_add(cy.ref, ymin.ref, ymax.ref, 0)
        mpfr.mpfr_div_ui(cy.ref, cy.ref, 2, 0)
        size = Mpfr(cp, 4.0) if bug_compatible else (wd if aspect0 > 1.0 else ht)
    if not width and not height:
        width, height = w0, (int(w0 / aspect_opt) if aspect_opt else h0)
    elif not width:
        width = int(height * (aspect_opt if aspect_opt else aspect0))
    elif not height:
        height = int(width / (aspect_opt if aspect_opt else aspect0))
    aspect = float(width) / height                                  # image_info_set (image_info.c:133)
    center = "" if s.center is not None else "#"                    # img->ui_ref_center
    corner = "#" if s.center is not None else ""
    ip = max(P, 80)
    r = rect if rect is not None else [Mpfr(ip).set_nan() for _ in range(3)]
    yn = lambda b: "yes" if b else "no"
    out = ["# http://jwm-art.net/mdz/", "settings",
           "family %s" % FAMILY_STR[s.family], "fractal %s" % FRACTAL_STR[s.fractal],
           "depth %d" % s.depth, "aspect %0.20f" % aspect, "colour-scale %0.20f" % s.colour_scale,
           "colour-interpolate %s" % yn(s.palette_ip), "multi-precision %s" % yn(s.use_multi_prec),
           "multi-rounding %s" % yn(s.use_rounding), "precision %d" % P,
           "%scx %s" % (center, cx.out_str()), "%scy %s" % (center, cy.out_str()),
           "%ssize %s" % (center, size.out_str()),
           "%sxmin %s" % (corner, r[0].out_str()), "%sxmax %s" % (corner, r[1].out_str()),
           "%symax %s" % (corner, r[2].out_str())]
    if s.family == FAMILY_JULIA:
        # image_info.c:703-704 / :271-272: c_re at the image precision, c_im keeps 80 bits
        out.append("julia-real %s" % Mpfr(ip, Mpfr(P, s.julia[0])).out_str())
        out.append("julia-imag %s" % Mpfr(ip if not bug_compatible else 80, Mpfr(P, s.julia[1])).out_str())
    out.append("palette-offset %d" % s.pal_offset)
    q = s.rnd
    out += ["r-strength %f" % q["r_strength"], "r-bands %f" % q["r_bands"],
            "g-strength %f" % q["g_strength"], "g-bands %f" % q["g_bands"],
            "b-strength %f" % q["b_strength"], "b-bands %f" % q["b_bands"],
            "rnd-offset %d" % q["offset"], "rnd-stripe %d" % q["stripe"], "rnd-spread %d" % q["spread"]]
    return "\n".join(out) + "\n"


def save_mdz(path, s, **kw):
    """A complete settings file (image_info_f_save_all, image_info.c:323-343): header,
    settings block, palette (embedded data or the file it came from)."""
    text = "%s %s\n" % (FILE_HEADER, VERSION) + settings_text(s, **kw) + "palette\n"
    if s.palette_file:
        text += "file %s\n" % s.palette_file
    else:
        text += "data\n"
        for c in (s.palette or []):
            text += " %d %d %d\n" % (c & 0xff, (c >> 8) & 0xff, (c >> 16) & 0xff)      # palette_write, palette.c:190-203
    with open(path, "w") as f:
        f.write(text)
