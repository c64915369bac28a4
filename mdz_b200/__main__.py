"""`python -m mdz_b200` -- MDZ's non-interactive render (reference src/cmdline.c:15-29,
src/render.c:10-105) without GTK: same flags, .mdz settings files, PNG output, the
render done by libmdzcuda with the colour epilogue fused into the kernel.

  python -m mdz_b200 -l gallery/honeytrace.mdz -w 1920 -h 1080 -A 2 -R out.png [-L log]
"""
import argparse
import sys
import time

import numpy as np

from . import Plan, device_count, last_error
from .mdzfile import load_mdz, view_from_settings, settings_text
from .png import write_png


def load_map(path):
    """.map palette: up to 256 lines " R G B" (reference src/palette.c:142-168)."""
    from .palette import Palette
    pal = Palette.load(path)
    return pal.colours if pal is not None else None


def main(argv=None):
    ap = argparse.ArgumentParser(prog="mdz_b200", add_help=False)
    ap.add_argument("--help", action="help")
    ap.add_argument("-l", "--load-settings", required=True)
    ap.add_argument("-P", "--load-palette")
    ap.add_argument("-L", "--log-file")
    ap.add_argument("-w", "--width", type=int, default=0)
    ap.add_argument("-h", "--height", type=int, default=0)
    ap.add_argument("-a", "--aspect-ratio", type=float, default=0.0)
    ap.add_argument("-R", "--render", required=True)
    ap.add_argument("-A", "--anti-alias", type=int, default=1)
    ap.add_argument("-t", "--threads", type=int, default=0, help="accepted for compatibility; GPUs do the work")
    ap.add_argument("--gpus", type=int, default=0, help="number of GPUs (default: all visible)")
    ap.add_argument("--fix-re", action="store_true",
                    help='GMP mode: convert the rect with "%%Re" instead of the reference\'s "%%.Re" (SURVEY finding 3)')
    a = ap.parse_args(argv)

    s = load_mdz(a.load_settings)
    view, col = view_from_settings(s, a.width, a.height, a.anti_alias, a.aspect_ratio,
                                   bug_compatible=True, fixed_re=a.fix_re)
    palette = load_map(a.load_palette) if a.load_palette else col["palette"]
    if palette is None and col["palette_file"]:
        palette = load_map(col["palette_file"])
    if palette is None:
        raise SystemExit("no palette: embed one in the settings file or pass -P file.map "
                         "(MDZ would seed a random one from the clock)")
    ndev = a.gpus or device_count()
    if ndev < 1:
        raise SystemExit("no CUDA device: " + last_error())
    print("calculating...")
    t0 = time.perf_counter()
    plans = [Plan(view, d, d, ndev) for d in range(ndev)]
    for p in plans:
        p.set_colour(palette, col["pal_offset"], col["colour_scale"], col["palette_ip"])
        p.launch()
    rgb = np.zeros((view.user_height, view.user_width), dtype=np.uint32)
    raw = np.full((view.real_height, view.real_width), -1, dtype=np.int32)
    for p in plans:
        p.fetch(raw)
        p.fetch_rgb(rgb)
        p.close()
    dt = time.perf_counter() - t0
    print("%4d of %4d lines done [time taken: %.3f]" % (view.user_height, view.user_height, dt))
    write_png(a.render, rgb)
    if a.log_file:
        with (sys.stdout if a.log_file == "-" else open(a.log_file, "w")) as f:
            # render.c:21-22 writes the settings block before the render, :98-102 the rest after it
            f.write(settings_text(s, a.width, a.height, a.aspect_ratio))
            f.write("render-time %.3fs\nsaved-image %s\n" % (dt, a.render))
    return 0


if __name__ == "__main__":
    sys.exit(main())
