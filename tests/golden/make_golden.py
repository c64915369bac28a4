#!/usr/bin/env python
"""Generate tests/golden/*.npz by running the UNMODIFIED reference here.

Run in the build container (needs /root/reference and oracle/_ref/, built by
`make -C oracle ref`).  Each fixture is what `mdz -l FILE -w W -h H [-A n] -R out`
produced: the rect the hot path saw (exact hex from mpfr_out_str), raw_data and
the coloured image, plus the library versions that produced them.  GMP-mode
files are rendered with the oracle's mdz_fixre build ("%.Re" -> "%Re", SURVEY
finding 3) and marked so.
"""
import ctypes as C
import json
import os
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
GALLERY = "/root/reference/gallery"
sys.path.insert(0, ROOT)

PALETTE = "".join(" %d %d %d\n" % ((i * 7) % 256, (i * 13 + 40) % 256, (255 - i * 3) % 256) for i in range(256))


def synthetic(name, body):
    return name, "mdz fractal settings 0.1.0\nsettings\n" + body + "palette\ndata\n" + PALETTE


SYNTH = [
    synthetic("cfg2_ld", "family mandelbrot\nfractal mandelbrot\ndepth 2000\naspect 1.77777777777777767909\n"
              "colour-scale 0.5\ncolour-interpolate no\nmulti-precision no\nmulti-rounding no\nprecision 80\n"
              "cx -0.5\ncy 0.0\nsize 4.0\n"),
    synthetic("ship_ld_aa3", "family mandelbrot\nfractal burning ship\ndepth 400\naspect 1.33333333333333325932\n"
              "colour-scale 0.37\ncolour-interpolate yes\nmulti-precision no\nmulti-rounding no\nprecision 80\n"
              "cx -0.5\ncy -0.5\nsize 4.0\npalette-offset 37\n"),
    synthetic("celtic_mpfr128_aa2", "family mandelbrot\nfractal generalized celtic\ndepth 300\naspect 1.0\n"
              "colour-scale 1.25\ncolour-interpolate no\nmulti-precision yes\nmulti-rounding yes\nprecision 128\n"
              "cx -0.5\ncy 0.0\nsize 4.0\npalette-offset 5\n"),
    # Julia + GMP mpf: the reference converts the constant per pixel through decimal text (fractal.c:341-342,
    # my_mpfr_to_str.c:68), one digit with "%.Re" on MPFR 4 -- stock build -- and all digits in the "%Re" build
    synthetic("julia_gmp128_asis", "family julia\nfractal mandelbrot\ndepth 300\naspect 1.33333333333333325932\n"
              "colour-scale 0.8\ncolour-interpolate no\nmulti-precision yes\nmulti-rounding no\nprecision 128\n"
              "cx 0.0\ncy 0.0\nsize 3.2\njulia-real -0.8\njulia-imag 0.156\n"),
    synthetic("julia_gmp128_fixre", "family julia\nfractal mandelbrot\ndepth 300\naspect 1.33333333333333325932\n"
              "colour-scale 0.8\ncolour-interpolate no\nmulti-precision yes\nmulti-rounding no\nprecision 128\n"
              "cx 0.0\ncy 0.0\nsize 3.2\njulia-real -0.8\njulia-imag 0.156\n"),
    synthetic("julia_mpfr96", "family julia\nfractal mandel-celtic hybrid\ndepth 300\naspect 1.33333333333333325932\n"
              "colour-scale 0.8\ncolour-interpolate yes\nmulti-precision yes\nmulti-rounding yes\nprecision 96\n"
              "cx 0.0\ncy 0.0\nsize 3.2\njulia-real -0.8\njulia-imag 0.156\n"),
]

# (fixture name, source, width, height, aa)
RUNS = [
    ("test", GALLERY + "/test.mdz", 60, 45, 1),                       # BASELINE configs[0] (240x180) scaled down
    ("test_240x180", GALLERY + "/test.mdz", 240, 180, 1),             # BASELINE configs[0] itself
    ("deep_embedded_julia_asis", GALLERY + "/deep_embedded_julia.mdz", 48, 36, 1),
    ("honeytrace", GALLERY + "/honeytrace.mdz", 32, 24, 1),
    ("satellite", GALLERY + "/satellite.mdz", 40, 30, 1),
    ("subdudehue", GALLERY + "/subdudehue.mdz", 40, 30, 1),
    ("mdz_banner", GALLERY + "/mdz_banner.mdz", 40, 30, 1),
    ("confirm", GALLERY + "/confirm.mdz", 32, 24, 1),
    ("polyp", GALLERY + "/polyp.mdz", 32, 24, 2),
    ("space_pad_fixre", GALLERY + "/space_pad.mdz", 32, 32, 1),
    ("floral_fixre", GALLERY + "/floral.mdz", 32, 32, 1),
    # GMP mode exactly as the stock reference behaves with this box's MPFR 4.2.1: the rect keeps one decimal digit
    # (SURVEY finding 3) and the view collapses; the drop-in inherits that and must reproduce it
    ("space_pad_asis", GALLERY + "/space_pad.mdz", 32, 32, 1),
] + [(n, None, w, h, a) for (n, _), (w, h, a) in zip(SYNTH, [(96, 54, 1), (40, 30, 3), (36, 36, 2), (48, 36, 1), (48, 36, 1), (48, 36, 1)])]


def run_one(name, src, w, h, aa, tmp):
    if src is None:
        text = dict(SYNTH)[name]
        src = os.path.join(tmp, name + ".mdz")
        open(src, "w").write(text)
    else:
        text = open(src).read()
        dst = os.path.join(tmp, name + ".mdz")
        open(dst, "w").write(text)
        src = dst
    exe = os.path.join(ROOT, "oracle", "_ref", "mdz_fixre" if name.endswith("_fixre") else "mdz")
    out = os.path.join(tmp, name + ".ppm")
    threads = "1" if name.startswith("julia_gmp") else "8"     # Julia + GMP converts through a static buffer shared by the workers
    subprocess.run([exe, "-l", src, "-w", str(w), "-h", str(h), "-A", str(aa), "-t", threads, "-R", out],
                   check=True, stdout=subprocess.DEVNULL, cwd=tmp)
    blob = open(out + ".raw", "rb").read()
    hdr, rest = blob.split(b"\n", 1)
    _, rw, rh, raa, depth = hdr.split()
    rw, rh = int(rw), int(rh)
    raw = np.frombuffer(rest[:rw * rh * 4], dtype=np.int32).reshape(rh, rw).copy()
    tail = rest[rw * rh * 4:].decode().split("\n")
    prec = int(tail[1].split()[2])
    rect = [t.strip() for t in tail[2:6]]
    ppm = open(out, "rb").read()
    parts = ppm.split(b"\n", 3)
    rgb = np.frombuffer(parts[3], dtype=np.uint8).reshape(h, w, 3).copy()
    meta = dict(name=name, width=w, height=h, aa=aa, depth=int(depth), precision=prec,
                rect_hex=rect, fixre=name.endswith("_fixre"), mdz_text=text)
    return raw, rgb, meta


def main():
    mpfr = C.CDLL("libmpfr.so.6")
    mpfr.mpfr_get_version.restype = C.c_char_p
    gmp_ver = C.c_char_p.in_dll(C.CDLL("libgmp.so.10"), "__gmp_version").value.decode()
    versions = dict(mpfr=mpfr.mpfr_get_version().decode(), gmp=gmp_ver,
                    reference="jwm-art-net/MDZ 0.1.2, src compiled unmodified with -O3 (oracle/Makefile)")
    with tempfile.TemporaryDirectory() as tmp:
        for name, src, w, h, aa in RUNS:
            raw, rgb, meta = run_one(name, src, w, h, aa, tmp)
            meta["versions"] = versions
            np.savez_compressed(os.path.join(HERE, name + ".npz"), raw=raw, rgb=rgb,
                                meta=np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8))
            it = int(np.where(raw > 0, raw, meta["depth"]).astype(np.int64).sum())
            print("%-28s %4dx%-4d aa%d p%-4d iterations %10d max %6d inside %.4f" % (
                name, w, h, aa, meta["precision"], it, raw.max(), (raw == 0).mean()))


if __name__ == "__main__":
    main()
