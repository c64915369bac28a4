"""Load tests/golden/*.npz (made by tests/golden/make_golden.py from the
unmodified reference) and rebuild the view the way the GTK-free harness does."""
import glob
import json
import os
import tempfile

import numpy as np

from mdz_b200.mdzfile import load_mdz, view_from_settings
from mdz_b200.mp import Mpfr

HERE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def names():
    return sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(HERE, "*.npz")))


def load(name):
    z = np.load(os.path.join(HERE, name + ".npz"))
    meta = json.loads(bytes(z["meta"]).decode())
    return meta, z["raw"], z["rgb"]


def settings_of(meta):
    with tempfile.NamedTemporaryFile("w", suffix=".mdz", delete=False) as f:
        f.write(meta["mdz_text"])
        path = f.name
    try:
        return load_mdz(path)
    finally:
        os.unlink(path)


def view_of(meta):
    s = settings_of(meta)
    return view_from_settings(s, meta["width"], meta["height"], meta["aa"],
                              bug_compatible=True, fixed_re=meta["fixre"])


def golden_rect(meta):
    ip = max(meta["precision"], 80)
    return [Mpfr(ip).set_str(t, 16) for t in meta["rect_hex"]]


def packed_rgb(rgb):
    """HxWx3 uint8 -> the reference's packed guint32 (R | G<<8 | B<<16)."""
    r = rgb.astype(np.uint32)
    return r[..., 0] | (r[..., 1] << 8) | (r[..., 2] << 16)
