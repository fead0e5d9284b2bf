#!/usr/bin/env python
"""Generates tests/golden/ref_images.npz from the reference's own test fixtures (run in the build container, where
/root/reference exists; the GPU box only sees the committed .npz):

    python tests/golden/make_golden_ref_images.py

Per image of /root/reference/testdata/{dog,rsu1,rsu2}.jpg (the inputs README.md:34-55 of the reference uses):
  <name>_jpg      the file's bytes (camera / encoder-produced baseline JPEGs, 4:2:2, no restart markers — Huffman and
                  quantisation tables this image's Pillow did not write)
  <name>_sha256   SHA-256 of the pixels the reference's decode lines (server/detector.py:128-133: PIL -> np.array) give
  <name>_probe    those pixels at [::52, ::52] (human-readable spot check; the survey's probe of dog.jpg pixel (0,0) =
                  (116, 134, 76) -> input (0.4549, 0.5255, 0.2980))
"""
import hashlib
import io
import os
import sys

import numpy as np
from PIL import Image

REF = os.environ.get("FASTDET_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    out = {}
    for name in ("dog", "rsu1", "rsu2"):
        with open(os.path.join(REF, "testdata", name + ".jpg"), "rb") as fp:
            data = fp.read()
        img = Image.open(io.BytesIO(data))
        assert img.size == (416, 416) and img.mode == "RGB"
        px = np.array(img)
        out[name + "_jpg"] = np.frombuffer(data, np.uint8)
        out[name + "_sha256"] = np.frombuffer(hashlib.sha256(px.tobytes()).digest(), np.uint8)
        out[name + "_probe"] = px[::52, ::52].copy()
        print(name, len(data), "bytes; pixel (0,0) =", px[0, 0].tolist())
    np.savez(os.path.join(HERE, "ref_images.npz"), **out)
    return 0


if __name__ == "__main__":
    sys.exit(main())
