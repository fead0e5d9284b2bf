#!/usr/bin/env python
"""Generates tests/golden/*.npz by running the REFERENCE's own code (server/detector.py) in the build
container.  The reference cannot travel to the GPU box, so its outputs are committed as fixtures.

    python tests/golden/make_golden.py            # needs /root/reference (read-only)

What is recorded (all produced by reference functions, none by this repo's oracle):
  post_*.npz    head tensors (int16, value = logit*64, exactly representable in f32) + the tuples
                ONNXDetector.process_yolo / soft_nms / the tail of perform() return for them
  kat.npz       small known-answer cases: Soft-NMS triple, overlap asymmetry, DummyDetector tuple
  pre.npz       normalise LUT (k/255 -> f32 for k = 0..255) taken from the reference expression, and
                perform() run end-to-end on an in-memory PNG with a stub ``model.run``
"""
import io
import os
import sys

import numpy as np

REF = os.environ.get("FASTDET_REFERENCE", "/root/reference")
sys.path.insert(0, os.path.join(REF, "server"))
import detector as ref  # noqa: E402  (the reference module)

HERE = os.path.dirname(os.path.abspath(__file__))


def make_heads(rng, shapes, num_classes, n_hot, obj_lo=-7.0):
    """Sparse-candidate head maps quantised to 1/64: background objectness ~ obj_lo, `n_hot` hot boxes
    placed in clusters so that Soft-NMS actually has overlaps to decay."""
    span = 5 + num_classes
    heads = []
    for (h, w) in shapes:
        m = rng.normal(0.0, 1.0, size=(h, w, 3, span))
        m[..., 4] = obj_lo + rng.normal(0.0, 0.5, size=(h, w, 3))
        heads.append(m)
    for _ in range(n_hot):
        hi = int(rng.integers(len(shapes)))
        h, w = shapes[hi]
        gy, gx = int(rng.integers(h)), int(rng.integers(w))
        for dy, dx, k in ((0, 0, int(rng.integers(3))), (0, min(1, w - 1 - gx), int(rng.integers(3))),
                          (min(1, h - 1 - gy), 0, int(rng.integers(3)))):
            if rng.random() < 0.7:
                cell = heads[hi][gy + dy, gx + dx, k]
                cell[4] = rng.normal(1.5, 1.5)
                cell[5 + int(rng.integers(num_classes))] += rng.normal(3.0, 1.0)
                cell[2:4] = rng.normal(0.0, 0.6, size=2)
    q = [np.clip(np.rint(m * 64), -32000, 32000).astype(np.int16) for m in heads]
    return q


def heads_to_nchw(q):
    """int16 [h,w,3,span] -> float32 [1, 3*span, h, w] exactly as an ONNX output would be laid out."""
    outs = []
    for m in q:
        h, w, a, s = m.shape
        f = (m.astype(np.float32) / np.float32(64.0)).reshape(h, w, a * s)
        outs.append(np.ascontiguousarray(f.transpose(2, 0, 1))[None])
    return outs


class StubModel:
    def __init__(self, outs):
        self.outs = outs
        self.last_input = None

    def run(self, names, feeds):
        self.last_input = feeds["input"]
        return self.outs


def new_detector(num_classes):
    det = ref.ONNXDetector.__new__(ref.ONNXDetector)
    ref.Detector.__init__(det, num_classes=num_classes)
    import logging
    det.logger = logging.getLogger("golden")
    det.mode, det.path = None, "<stub>"
    return det


def run_reference_post(outs, num_classes, threshold):
    """The body of ONNXDetector.perform after model.run (detector.py:136-144), using reference functions."""
    det = new_detector(num_classes)
    aas = det.ANCHORS[len(outs)]
    objs = []
    per_head = []
    for anchors, output in zip(aas, outs):
        o = output.transpose(0, 2, 3, 1)
        found = det.process_yolo(anchors, o[0], threshold=threshold)
        per_head.append(len(found))
        objs.extend(found)
    cand = np.array([[o.klass, o.conf, *o.bbox] for o in objs], np.float64).reshape(-1, 6)
    kept = ref.soft_nms(objs, threshold=threshold)
    (width, height) = det.image_size
    results = [(o.klass, o.conf, o.bbox[0] * width, o.bbox[1] * height, o.bbox[2] * width, o.bbox[3] * height)
               for o in kept]
    return cand, np.array(results, np.float64).reshape(-1, 6), np.array(per_head)


def main():
    rng = np.random.default_rng(20261018)
    cases = {
        "post_tiny80": dict(shapes=[(13, 13), (26, 26)], nc=80, n_hot=30, thr=0.1),
        "post_rsu9": dict(shapes=[(13, 13), (26, 26), (52, 52)], nc=9, n_hot=40, thr=0.1),
        "post_full80_small": dict(shapes=[(4, 4), (8, 8), (16, 16)], nc=80, n_hot=25, thr=0.3),
        "post_dense": dict(shapes=[(6, 6), (12, 12)], nc=9, n_hot=150, thr=0.05),
        "post_empty": dict(shapes=[(13, 13), (26, 26)], nc=80, n_hot=0, thr=0.1),
    }
    for name, c in cases.items():
        q = make_heads(rng, c["shapes"], c["nc"], c["n_hot"])
        outs = heads_to_nchw(q)
        cand, results, per_head = run_reference_post(outs, c["nc"], c["thr"])
        print(f"{name}: {len(cand)} candidates -> {len(results)} kept")
        np.savez_compressed(os.path.join(HERE, name + ".npz"), num_classes=c["nc"], threshold=c["thr"],
                            candidates=cand, results=results, per_head=per_head,
                            **{f"head{i}": m for i, m in enumerate(q)})

    # known-answer cases
    A = ref.YOLOObject(1, .9, (.1, .1, .2, .2))
    B = ref.YOLOObject(1, .8, (.11, .11, .2, .2))
    C = ref.YOLOObject(2, .5, (.6, .6, .1, .1))
    kept = ref.soft_nms([A, B, C], 0.1)
    big = ref.YOLOObject(1, .9, (0., 0., 2., 2.))
    small = ref.YOLOObject(1, .9, (0., 0., 1., 1.))
    dummy = ref.DummyDetector().perform(b"")
    np.savez_compressed(
        os.path.join(HERE, "kat.npz"),
        nms_in=np.array([[o.klass, o.conf, *o.bbox] for o in (A, B, C)]),
        nms_out=np.array([[o.klass, o.conf, *o.bbox] for o in kept]),
        iou_big_small=big.get_iou(small.bbox), iou_small_big=small.get_iou(big.bbox),
        iou_disjoint=float(A.get_iou(C.bbox)),
        sigmoid_in=np.array([-30.0, -4.5, -1.0, 0.0, 0.25, 3.0, 20.0]),
        sigmoid_out=np.array([ref.sigmoid(v) for v in (-30.0, -4.5, -1.0, 0.0, 0.25, 3.0, 20.0)]),
        dummy=np.array(dummy, np.float64))

    # preprocess: the reference expression on every u8 value, and perform() end-to-end on a PNG
    lut = (np.arange(256, dtype=np.uint8).reshape(1, 16, 16, 1).repeat(3, axis=3) / 255).astype(np.float32)
    lut = lut[0, :, :, 0].reshape(256)
    from PIL import Image
    yy, xx = np.mgrid[0:416, 0:416]
    img = np.stack([(xx * 255 // 415), (yy * 255 // 415), ((xx // 32 + yy // 32) % 2) * 200 + 20], axis=2).astype(np.uint8)
    buf = io.BytesIO()
    Image.fromarray(img, "RGB").save(buf, format="PNG", optimize=True)
    png = buf.getvalue()
    q = make_heads(rng, [(13, 13), (26, 26)], 80, 12)
    outs = heads_to_nchw(q)
    det = new_detector(80)
    det.model = StubModel(outs)
    results = det.perform(png, threshold=0.1)
    a = det.model.last_input
    assert a.shape == (1, 3, 416, 416) and a.dtype == np.float32
    np.savez_compressed(os.path.join(HERE, "pre.npz"), lut=lut, png=np.frombuffer(png, np.uint8),
                        input_sum=np.float64(a.astype(np.float64).sum()),
                        input_probe=a[0, :, ::52, ::52].copy(), results=np.array(results, np.float64).reshape(-1, 6),
                        **{f"head{i}": m for i, m in enumerate(q)})
    print("pre: png", len(png), "bytes,", len(results), "results")
    try:
        det.perform(png[:50] + png[60:], threshold=0.1)
    except Exception as e:  # documents the exception family for undecodable bytes
        print("bad bytes ->", type(e).__mro__[:3])


if __name__ == "__main__":
    main()
