"""RGB8, non-interlaced PNG writer on zlib (reference src/my_png.c:28-109 does the
same through libpng, which this box does not have): IHDR 8-bit colour type 2,
no interlace, filler byte of the packed guint32 stripped (png_set_filler AFTER)."""
import struct
import zlib

import numpy as np


def _chunk(tag, data):
    return struct.pack(">I", len(data)) + tag + data + struct.pack(">I", zlib.crc32(tag + data) & 0xFFFFFFFF)


def rgb_bytes(rgb_data):
    """packed R | G<<8 | B<<16 (palette.h:16) [H][W] uint32 -> [H][W][3] uint8."""
    a = np.ascontiguousarray(rgb_data, dtype=np.uint32)
    return np.stack([a & 0xFF, (a >> 8) & 0xFF, (a >> 16) & 0xFF], axis=-1).astype(np.uint8)


def write_png(path, rgb_data, level=6):
    px = rgb_bytes(rgb_data)
    h, w, _ = px.shape
    rows = np.concatenate([np.zeros((h, 1), dtype=np.uint8), px.reshape(h, w * 3)], axis=1)   # filter type 0
    body = (b"\x89PNG\r\n\x1a\n" + _chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, 8, 2, 0, 0, 0))
            + _chunk(b"IDAT", zlib.compress(rows.tobytes(), level)) + _chunk(b"IEND", b""))
    with open(path, "wb") as f:
        f.write(body)


def read_png_rgb8(path):
    """Minimal reader for files written by write_png (tests)."""
    data = open(path, "rb").read()
    assert data[:8] == b"\x89PNG\r\n\x1a\n"
    pos, idat, w, h = 8, b"", 0, 0
    while pos < len(data):
        n, tag = struct.unpack(">I4s", data[pos:pos + 8])
        chunk = data[pos + 8:pos + 8 + n]
        if tag == b"IHDR":
            w, h, depth, ctype = struct.unpack(">IIBB", chunk[:10])
            assert (depth, ctype) == (8, 2)
        elif tag == b"IDAT":
            idat += chunk
        pos += 12 + n
    raw = np.frombuffer(zlib.decompress(idat), dtype=np.uint8).reshape(h, 1 + 3 * w)
    assert (raw[:, 0] == 0).all()
    return raw[:, 1:].reshape(h, w, 3)
