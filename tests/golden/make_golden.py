"""
Generates tests/golden/ref_conv_golden.npz by RUNNING THE REFERENCE ITSELF
(cpp/int8conv/refConv2DForward.hpp, compiled unmodified into oracle/_ref/libref_conv.so).

Run in the authoring container (where /root/reference exists):
    make -C oracle ref && python tests/golden/make_golden.py

The reference holds no golden vectors of its own; these fixtures are its outputs on
seeded inputs so that the oracle stays pinned on machines where the reference cannot be rebuilt.
Inputs are regenerated from the seed by the tests; only the int32 outputs (and a CRC of the inputs)
are stored.
"""
import os
import sys
import zlib

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
from oracle import oracle  # noqa: E402

# every instantiated shape except the slow ones (config 1 in full = minutes; 16x16 crop ~20 s)
SLOW = {(1, 64, 58, 58, 64, 56, 56, 3, 3), (1, 64, 18, 18, 64, 16, 16, 3, 3)}


def inputs(shape, style, seed):
    b, ic, ih, iw, oc, oh, ow, kh, kw = shape
    rng = np.random.default_rng(seed)
    if style == "ref":   # {0,1}-valued, check.cu:43-44,69-75
        x = rng.integers(0, 2, size=(b, ic, ih, iw), dtype=np.int8)
        w = rng.integers(0, 2, size=(oc, ic, kh, kw), dtype=np.int8)
    elif style == "extreme":  # worst-case magnitudes
        x = rng.choice(np.array([-128, 127], dtype=np.int8), size=(b, ic, ih, iw))
        w = rng.choice(np.array([-128, 127], dtype=np.int8), size=(oc, ic, kh, kw))
    else:
        x = rng.integers(-128, 128, size=(b, ic, ih, iw), dtype=np.int8)
        w = rng.integers(-128, 128, size=(oc, ic, kh, kw), dtype=np.int8)
    return x, w


def main():
    out = {}
    idx = 0
    for shape in oracle.ref_shapes():
        if shape in SLOW:
            continue
        for style in ("full", "ref", "extreme"):
            seed = 7000 + idx
            x, w = inputs(shape, style, seed)
            y = oracle.ref_conv2d_forward(x, w)
            key = f"case{idx:02d}"
            out[key + "_shape"] = np.array(shape, dtype=np.int32)
            out[key + "_style"] = np.array(style)
            out[key + "_seed"] = np.array(seed, dtype=np.int64)
            out[key + "_crc"] = np.array(zlib.crc32(x.tobytes() + w.tobytes()), dtype=np.int64)
            out[key + "_y"] = y
            print(key, shape, style, "sum", int(y.astype(np.int64).sum()))
            idx += 1
    path = os.path.join(os.path.dirname(__file__), "ref_conv_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
