"""Host half of the compressed read-back (csrc/readback_decode.cu): run-length rows -> pixels, against a numpy encoder of the same
format (one word per run: B | G << 8 | R << 16 | (length - 1) << 24, runs cut at 256-pixel boundaries, a count word per row with
bit 31 = raw row).  The GPU half (picture_rle_kernel) is checked against the raw read-back in tests/test_fused_gpu.py."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _encode_row(row):
    w = len(row)
    x = np.arange(w)
    start = np.ones(w, bool)
    start[1:] = row[1:] != row[:-1]
    start |= (x & 255) == 0
    pos = np.flatnonzero(start)
    lens = np.diff(np.append(pos, w))
    return (row[pos] & 0xFFFFFF) | ((lens - 1).astype(np.uint32) << 24)


@pytest.fixture(scope="module")
def decoder(tmp_path_factory):
    so = tmp_path_factory.mktemp("rle") / "libdecode.so"
    subprocess.run(["g++", "-O3", "-mavx2", "-shared", "-fPIC", "-x", "c++", os.path.join(ROOT, "square_wheel_b200", "csrc", "readback_decode.cu"), "-o", str(so)], check=True)
    lib = C.CDLL(str(so))
    f = lib._ZN3sqw19decode_picture_rowsEPKjS1_jjjjPjm
    f.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_void_p, C.c_size_t]
    f.restype = None
    return f


@pytest.mark.parametrize("width, height", [(3840, 24), (1027, 7), (255, 3), (17, 5), (8, 2), (1, 1)])
def test_decoder_expands_run_length_rows(decoder, width, height):
    r = np.random.default_rng(width * 31 + height)
    for kind in ("runs", "noise", "flat", "raw rows"):
        if kind == "noise":
            pic = r.integers(0, 1 << 24, (height, width), dtype=np.uint32)
        elif kind == "flat":
            pic = np.full((height, width), 0x123456, np.uint32)
        else:
            lens = r.integers(1, 40, (height, width))  # run lengths 1..39 with a few long ones
            lens[r.random((height, width)) < 0.02] = 700
            pic = np.zeros((height, width), np.uint32)
            for y in range(height):
                vals = r.integers(0, 1 << 24, width, dtype=np.uint32)
                pic[y] = np.repeat(vals, lens[y])[:width]
        pic |= np.uint32(0xFF000000)
        rle = np.zeros((height, width), np.uint32)
        counts = np.zeros(height, np.uint32)
        for y in range(height):
            if kind == "raw rows" and y % 2 == 0:
                rle[y] = pic[y]
                counts[y] = 0x80000000 | width
            else:
                e = _encode_row(pic[y])
                rle[y, : len(e)] = e
                counts[y] = len(e)
        pitch = width + 3
        out = np.full((height + 1, pitch), 0x5A5A5A5A, np.uint32)
        decoder(rle.ctypes.data, counts.ctypes.data, width, width, 0, height, out.ctypes.data, pitch)
        assert np.array_equal(out[:height, :width], pic), f"{kind} {width}x{height}"
        assert (out[:height, width:] == 0x5A5A5A5A).all() and (out[height] == 0x5A5A5A5A).all(), f"{kind}: wrote outside the rows"
