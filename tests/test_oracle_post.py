"""The oracle's pre/post restatement replayed against fixtures produced by the reference module itself
(tests/golden/make_golden.py).  Bit-exact: same Python-double arithmetic in the same order."""
import glob
import io
import os

import numpy as np
import pytest

from oracle import ref_post


def _heads(z):
    outs = []
    i = 0
    while f"head{i}" in z:
        m = z[f"head{i}"]
        h, w, a, s = m.shape
        f = (m.astype(np.float32) / np.float32(64.0)).reshape(h, w, a * s)
        outs.append(np.ascontiguousarray(f.transpose(2, 0, 1))[None])
        i += 1
    return outs


POST = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "post_*.npz")))


@pytest.mark.parametrize("path", POST, ids=[os.path.basename(p)[:-4] for p in POST])
@pytest.mark.parametrize("fast", [False, True])
def test_post_matches_reference(path, fast):
    z = np.load(path)
    outs = _heads(z)
    nc, thr = int(z["num_classes"]), float(z["threshold"])
    # candidates: same set, same order, same doubles
    cands = []
    first = 0
    dec = ref_post.decode_head_fast if fast else ref_post.decode_head
    for anchors, out in zip(ref_post.ANCHORS[len(outs)], outs):
        m = np.ascontiguousarray(out[0].transpose(1, 2, 0))
        got = dec(anchors, m, nc, (416, 416), thr, first)
        first += m.shape[0] * m.shape[1] * 3
        cands.extend(got)
    ref_c = z["candidates"]
    assert len(cands) == len(ref_c)
    if len(cands):
        assert [c[0] for c in cands] == sorted(c[0] for c in cands)  # insertion order
        got = np.array([c[1:] for c in cands], np.float64)
        assert np.array_equal(got, ref_c)
    results, _, _ = ref_post.detect_from_heads(outs, 0, nc, (416, 416), thr, fast=fast)
    ref_r = z["results"]
    assert len(results) == len(ref_r)
    if len(results):
        assert np.array_equal(np.array(results, np.float64), ref_r)


def test_known_answers(golden_dir):
    z = np.load(os.path.join(golden_dir, "kat.npz"))
    cands = [(i, int(r[0]), r[1], r[2], r[3], r[4], r[5]) for i, r in enumerate(z["nms_in"])]
    kept = ref_post.soft_nms(cands, 0.1)
    got = np.array([[c[1], c[2], c[3], c[4], c[5], c[6]] for _, c in kept])
    assert np.array_equal(got, z["nms_out"])
    assert [c[0] for _, c in kept] == [0, 2]  # B decays below threshold, different-class C survives
    assert ref_post.overlap((0, 0, 2, 2), (0, 0, 1, 1)) == float(z["iou_big_small"]) == 0.25
    assert ref_post.overlap((0, 0, 1, 1), (0, 0, 2, 2)) == float(z["iou_small_big"]) == 1.0
    assert ref_post.overlap((.1, .1, .2, .2), (.6, .6, .1, .1)) == float(z["iou_disjoint"]) == 0
    assert [ref_post.logistic(v) for v in z["sigmoid_in"]] == list(z["sigmoid_out"])
    with pytest.raises(OverflowError):
        ref_post.logistic(-800.0)
    with pytest.raises(KeyError):
        ref_post.detect_from_heads([np.zeros((1, 255, 2, 2), np.float32)], 0, 80, (416, 416), 0.1)


def test_preprocess_and_perform(golden_dir):
    from PIL import Image

    z = np.load(os.path.join(golden_dir, "pre.npz"))
    lut = (np.arange(256, dtype=np.uint8).reshape(1, 1, 256, 1).repeat(3, 3))
    a = ref_post.normalise(lut.reshape(1, 256, 3))
    assert a.dtype == np.float32 and np.array_equal(a[0, 0, 0], z["lut"])
    img = np.array(Image.open(io.BytesIO(z["png"].tobytes())))
    assert img.shape == (416, 416, 3)
    a = ref_post.normalise(img)
    assert a.shape == (1, 3, 416, 416)
    assert float(a.astype(np.float64).sum()) == float(z["input_sum"])
    assert np.array_equal(a[0, :, ::52, ::52], z["input_probe"])
    results, _, _ = ref_post.detect_from_heads(_heads(z), 0, 80, (416, 416), 0.1)
    assert np.array_equal(np.array(results, np.float64), z["results"])


def test_letterbox_identity_and_shape():
    rng = np.random.default_rng(0)
    src = rng.integers(0, 256, size=(416, 416, 3), dtype=np.uint8)
    out, geom = ref_post.letterbox_u8(src, 416, 416)
    assert np.array_equal(out, src) and geom == (0, 0, 416, 416)
    src = rng.integers(0, 256, size=(480, 640, 3), dtype=np.uint8)
    out, (ox, oy, nw, nh) = ref_post.letterbox_u8(src, 416, 416)
    assert (nw, nh) == (416, 312) and (ox, oy) == (0, 52)
    assert (out[:oy] == 128).all() and (out[oy + nh:] == 128).all()
