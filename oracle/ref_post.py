"""oracle/ref_post.py — CPU restatement of the reference's pre- and post-processing.

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's CPU-baseline
legs, never by the product path (fastdet_b200/ fails loudly without its CUDA library instead).

Follows reference server/detector.py:
  normalise()        :133-134   u8 HWC -> (x/255 in float64) -> float32 -> NCHW
  decode_head()      :148-166   sigmoid/anchor decode with the two threshold tests, arg-max on raw logits
  overlap()          :15-22, :38-42   area(sel ∩ other) / area(sel)  — asymmetric, NOT IoU
  soft_nms()         :45-59     class-agnostic Gaussian Soft-NMS, sequential arg-max + decay
  detect()           :136-144   head i pairs with ANCHORS[n_heads][i]; results in network-input pixels
Pinned by tests/golden/*.npz, which were produced by importing the reference module itself in the build
container (tests/golden/make_golden.py); tests/test_oracle_post.py replays them bit-for-bit.

Candidates are plain tuples ``(index, klass, conf, x, y, w, h)``; ``index`` is the insertion order the
reference's list/dict iteration gives (head 0 first, then row, column, anchor), which is what breaks ties.
"""
from __future__ import annotations

import math
from typing import List, Sequence, Tuple

import numpy as np

# reference server/detector.py:96-106 — keyed by the number of graph outputs, coarsest grid first
ANCHORS = {
    3: (((116, 90), (156, 198), (373, 326)),
        ((30, 61), (62, 45), (59, 119)),
        ((10, 13), (16, 30), (33, 23))),
    2: (((81, 82), (135, 169), (344, 319)),
        ((10, 14), (23, 27), (37, 58))),
}

Cand = Tuple[int, int, float, float, float, float, float]


def logistic(v: float) -> float:
    # detector.py:12-13; math.exp raises OverflowError below about -709 exactly like the reference
    return 1 / (1 + math.exp(-v))


def normalise(frame_u8: np.ndarray) -> np.ndarray:
    """[H,W,3] u8 -> [1,3,H,W] f32, values are the correctly rounded k/255 (detector.py:133-134)."""
    h, w, _ = frame_u8.shape
    a = (frame_u8.reshape(1, h, w, 3) / 255).astype(np.float32)
    return a.transpose(0, 3, 1, 2)


def decode_head(anchors, m: np.ndarray, num_classes: int, net_wh: Tuple[int, int], threshold: float,
                first_index: int = 0) -> List[Cand]:
    """m: [rows, cols, 3*(5+nc)] float32 (one frame, one head, channels last — detector.py:139)."""
    net_w, net_h = net_wh
    rows, cols, _ = m.shape
    span = 5 + num_classes
    out: List[Cand] = []
    idx = first_index
    for gy in range(rows):
        for gx in range(cols):
            cell = m[gy, gx]
            for k, (aw, ah) in enumerate(anchors):
                base = span * k
                this = idx
                idx += 1
                score = logistic(cell[base + 4])
                if score < threshold:
                    continue
                cx = (gx + logistic(cell[base + 0])) / cols
                cy = (gy + logistic(cell[base + 1])) / rows
                bw = aw * math.exp(cell[base + 2]) / net_w
                bh = ah * math.exp(cell[base + 3]) / net_h
                best = int(np.argmax(cell[base + 5:base + 5 + num_classes]))  # first maximum on ties
                score *= logistic(cell[base + 5 + best])
                if score < threshold:
                    continue
                out.append((this, best + 1, score, cx - bw / 2, cy - bh / 2, bw, bh))
    return out


def decode_head_fast(anchors, m: np.ndarray, num_classes: int, net_wh, threshold: float,
                     first_index: int = 0) -> List[Cand]:
    """Same result as decode_head; a numpy pre-filter on objectness skips the cells that cannot pass,
    then the scalar code above's arithmetic runs on the survivors only."""
    net_w, net_h = net_wh
    rows, cols, _ = m.shape
    span = 5 + num_classes
    mm = m.reshape(rows, cols, 3, span)
    obj = mm[..., 4].astype(np.float64)
    # sigmoid(v) >= thr  <=>  v >= logit(thr); keep a safety margin and let the exact test decide
    if 0.0 < threshold < 1.0:
        cut = math.log(threshold / (1 - threshold)) - 1e-6
        mask = obj >= cut
    else:
        mask = np.ones_like(obj, bool)
    out: List[Cand] = []
    for gy, gx, k in zip(*np.nonzero(mask)):
        gy, gx, k = int(gy), int(gx), int(k)
        v = mm[gy, gx, k]
        aw, ah = anchors[k]
        score = logistic(v[4])
        if score < threshold:
            continue
        cx = (gx + logistic(v[0])) / cols
        cy = (gy + logistic(v[1])) / rows
        bw = aw * math.exp(v[2]) / net_w
        bh = ah * math.exp(v[3]) / net_h
        best = int(np.argmax(v[5:5 + num_classes]))
        score *= logistic(v[5 + best])
        if score < threshold:
            continue
        out.append((first_index + (gy * cols + gx) * 3 + k, best + 1, score, cx - bw / 2, cy - bh / 2, bw, bh))
    return out


def overlap(sel: Sequence[float], other: Sequence[float]) -> float:
    """area(sel ∩ other) / area(sel); 0 when disjoint (detector.py:15-22, 38-42).  Boxes are (x,y,w,h)."""
    x0, y0, w0, h0 = sel
    x1, y1, w1, h1 = other
    iw = min(x0 + w0, x1 + w1) - max(x0, x1)
    ih = min(y0 + h0, y1 + h1) - max(y0, y1)
    if iw <= 0 or ih <= 0:
        return 0
    return (iw * ih) / (w0 * h0)


def soft_nms(cands: List[Cand], threshold: float, leftovers: dict = None) -> List[Tuple[float, Cand]]:
    """Returns [(decayed_score_at_selection, candidate)] in the reference's output order.

    detector.py:45-59: repeatedly take the arg-max of the current scores (first in insertion order on ties,
    strict '<' at :51), stop when it is below threshold, decay every remaining score by
    exp(-3 * overlap(selected, other)**2).  The final ``sort(reverse=True)`` (:58) orders by the decayed
    score; selection order is already non-increasing, so only exact ties could reorder — there the
    reference raises TypeError (YOLOObject is unorderable), and this oracle keeps selection order."""
    live = [(c, c[2]) for c in cands]
    picked: List[Tuple[float, Cand]] = []
    while cands:
        top, top_i = -1, -1
        for i, (_, s) in enumerate(live):
            if top < s:
                top, top_i = s, i
        if top < threshold:
            break
        sel = live.pop(top_i)[0]
        picked.append((top, sel))
        box = sel[3:7]
        live = [(c, s * math.exp(-3 * (overlap(box, c[3:7]) ** 2))) for (c, s) in live]
    if leftovers is not None:  # test aid: final (decayed) score of every candidate that was never selected
        leftovers.update({c[0]: s for (c, s) in live})
    return picked


def detect_from_heads(heads: Sequence[np.ndarray], frame: int, num_classes: int, net_wh, threshold: float,
                      fast: bool = True, leftovers: dict = None):
    """heads: graph outputs [N,C,H,W] float32 in graph order.  Returns the reference's result tuples
    ``(klass, conf, x, y, w, h)`` for one frame (detector.py:136-144) plus the candidate indices kept."""
    net_w, net_h = net_wh
    anchor_sets = ANCHORS[len(heads)]  # KeyError for anything but 2 or 3 outputs, like the reference
    cands: List[Cand] = []
    first = 0
    dec = decode_head_fast if fast else decode_head
    for anchors, out in zip(anchor_sets, heads):
        m = np.ascontiguousarray(out[frame].transpose(1, 2, 0))
        cands.extend(dec(anchors, m, num_classes, net_wh, threshold, first))
        first += m.shape[0] * m.shape[1] * 3
    kept = soft_nms(cands, threshold, leftovers)
    results = [(c[1], c[2], c[3] * net_w, c[4] * net_h, c[5] * net_w, c[6] * net_h) for (_, c) in kept]
    return results, [c[0] for (_, c) in kept], [s for (s, _) in kept]


def box_iou(a: Sequence[float], b: Sequence[float]) -> float:
    """True IoU of two (x, y, w, h) boxes (the parity metric of BASELINE.json's north_star; NOT the reference's overlap)."""
    iw = min(a[0] + a[2], b[0] + b[2]) - max(a[0], b[0])
    ih = min(a[1] + a[3], b[1] + b[3]) - max(a[1], b[1])
    if iw <= 0 or ih <= 0:
        return 0.0
    return iw * ih / (a[2] * a[3] + b[2] * b[3] - iw * ih)


def detection_spread(heads_a, heads_b, num_classes: int, net_wh, threshold: float):
    """Test aid: IoU and |dconf| of the detections two sets of head tensors (one frame each) give at the same anchor box
    with the same class — how far two evaluations of the same network are apart in the spec's own metric."""
    wa, ia, _ = detect_from_heads(heads_a, 0, num_classes, net_wh, threshold)
    wb, ib, _ = detect_from_heads(heads_b, 0, num_classes, net_wh, threshold)
    db = dict(zip(ib, wb))
    ious, dconfs = [], []
    for box, w in zip(ia, wa):
        if box in db and db[box][0] == w[0]:
            ious.append(box_iou(w[2:], db[box][2:]))
            dconfs.append(abs(w[1] - db[box][1]))
    return ious, dconfs


def letterbox_u8(src: np.ndarray, net_w: int, net_h: int, fill: int = 128) -> Tuple[np.ndarray, Tuple[int, int, int, int]]:
    """Integer-exact restatement of the repo's letterbox kernel (an EXTENSION — the reference server rejects
    frames that are not already net-sized, detector.py:131-132; parity for this function is pinned only
    against this restatement).  Aspect-preserving bilinear resize in 16.16 fixed point, centred, grey fill.
    Returns (frame[net_h,net_w,3] u8, (off_x, off_y, new_w, new_h))."""
    sh, sw, _ = src.shape
    if sw * net_h >= sh * net_w:  # width-limited
        new_w, new_h = net_w, max(1, (sh * net_w + sw // 2) // sw)
    else:
        new_h, new_w = net_h, max(1, (sw * net_h + sh // 2) // sh)
    off_x, off_y = (net_w - new_w) // 2, (net_h - new_h) // 2
    out = np.full((net_h, net_w, 3), fill, np.uint8)
    s = src.astype(np.int64)
    # half-pixel centres: sx = (dx + 0.5) * sw / new_w - 0.5, in 16.16 fixed point
    def axis(n_dst, n_src):
        d = np.arange(n_dst, dtype=np.int64)
        pos = ((2 * d + 1) * n_src * 65536) // (2 * n_dst) - 32768
        pos = np.clip(pos, 0, (n_src - 1) * 65536)
        i0 = pos >> 16
        fr = pos & 0xFFFF
        i1 = np.minimum(i0 + 1, n_src - 1)
        return i0, i1, fr
    x0, x1, fx = axis(new_w, sw)
    y0, y1, fy = axis(new_h, sh)
    fx = fx[None, :, None]
    fy = fy[:, None, None]
    top = s[y0][:, x0] * (65536 - fx) + s[y0][:, x1] * fx
    bot = s[y1][:, x0] * (65536 - fx) + s[y1][:, x1] * fx
    val = (top * (65536 - fy) + bot * fy + (1 << 31)) >> 32
    out[off_y:off_y + new_h, off_x:off_x + new_w] = val.astype(np.uint8)
    return out, (off_x, off_y, new_w, new_h)
