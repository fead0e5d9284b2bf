"""Generates tests/golden/jpeg.npz: small baseline JPEG streams written by Pillow / libjpeg-turbo in this image and the
pixels the reference's decode lines (server/detector.py:128-133: Image.open + np.array) return for them.  The fixture
travels to the GPU box, so the device decoder is pinned to THESE outputs even if the box's Pillow were different.

    python tests/golden/make_golden_jpeg.py
"""
import io
import os
import sys

import numpy as np
from PIL import Image

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import ref_jpeg  # noqa: E402


def picture(h, w, seed):
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:h, 0:w]
    img = np.stack([127 + 100 * np.sin(xx / 7.0 + seed) * np.cos(yy / 11.0), 127 + 120 * np.sin((xx + yy) / 5.0),
                    (xx * 3 + yy * 5 + seed * 17) % 256], -1).astype(np.float64)
    img += rng.normal(0, 25, img.shape)
    img[h // 4:h // 2, w // 4:w // 2] = rng.integers(0, 255, 3)
    return np.clip(img, 0, 255).astype(np.uint8)


def main():
    out = {}
    cases = [(64, 80, 0, 75, 0), (64, 80, 1, 75, 0), (64, 80, 2, 75, 0), (37, 29, 2, 90, 0), (37, 29, 1, 30, 2),
             (48, 48, 2, 100, 3), (17, 131, 0, 50, 1), (96, 96, 2, 5, 0)]
    for i, (h, w, sub, q, rst) in enumerate(cases):
        buf = io.BytesIO()
        kw = dict(restart_marker_blocks=rst) if rst else {}
        Image.fromarray(picture(h, w, i)).save(buf, 'JPEG', quality=q, subsampling=sub, **kw)
        data = buf.getvalue()
        out[f'jpeg{i}'] = np.frombuffer(data, np.uint8)
        out[f'rgb{i}'] = ref_jpeg.decode_reference(data)
        out[f'meta{i}'] = np.array([h, w, sub, q, rst])
    # damaged-but-decodable streams whose coefficients leave the range where libjpeg's C text and its SIMD code agree
    # (16-bit lane wrap, saturating packs, DC-only shortcut): Pillow's answer for them is part of the contract too
    import warnings
    from fastdet_b200 import _native
    rng = np.random.default_rng(2024)
    kept = 0
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        while kept < 8:
            buf = io.BytesIO()
            Image.fromarray(picture(40, 56, kept)).save(buf, 'JPEG', quality=int(rng.integers(30, 95)), subsampling=int(rng.integers(0, 3)))
            d = bytearray(buf.getvalue())
            for _ in range(int(rng.integers(1, 4))):
                pos = int(rng.integers(2, len(d)))
                d[pos] = int(rng.integers(0, 256))
            d = bytes(d)
            try:
                _native.jpeg_coefficients(d)
                rgb = ref_jpeg.decode_reference(d)
            except Exception:
                continue
            clean = ref_jpeg.decode_reference(buf.getvalue())
            if np.abs(rgb.astype(int) - clean.astype(int)).max() < 200:
                continue  # keep the ones where the damage drives samples into the clamps
            out[f'damaged_jpeg{kept}'] = np.frombuffer(d, np.uint8)
            out[f'damaged_rgb{kept}'] = rgb
            kept += 1
    out['damaged_count'] = np.array(kept)
    out['count'] = np.array(len(cases))
    np.savez_compressed(os.path.join(HERE, 'jpeg.npz'), **out)
    print('wrote', os.path.join(HERE, 'jpeg.npz'), sum(v.nbytes for v in out.values()), 'bytes raw')


if __name__ == '__main__':
    main()
