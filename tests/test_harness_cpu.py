"""GTK-free harness pieces that need no GPU: the PNG writer round-trips."""
import numpy as np

from mdz_b200.png import write_png, read_png_rgb8, rgb_bytes


def test_png_roundtrip(tmp_path):
    rng = np.random.default_rng(3)
    rgb = rng.integers(0, 1 << 24, size=(37, 53), dtype=np.uint32)
    path = str(tmp_path / "x.png")
    write_png(path, rgb)
    assert np.array_equal(read_png_rgb8(path), rgb_bytes(rgb))
    # readable by an independent decoder as well
    from PIL import Image
    assert np.array_equal(np.asarray(Image.open(path).convert("RGB")), rgb_bytes(rgb))
