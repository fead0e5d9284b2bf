"""Shared comparison helper for result lists that came through different batch sizes.

A frame's tuples do not depend on the batch it rode in beyond the fp32 re-association the kernel forms of different
batch sizes imply (split-K parts): same boxes and classes, scores within the spec's 1e-2.  Boxes are matched
by class and position (within 1.5 px), not by rounded coordinates — a coordinate that sits on x.5 must not split a match."""


def same_detections(got, want, threshold, max_unmatched=2, dconf=1e-2, dpix=1.5):
    """got / want: reference-style tuples (klass, conf, x, y, w, h).  Solid boxes only (conf >= threshold + 1e-2): a
    near-threshold one may come or go.  Returns True when all but max_unmatched solid boxes of either side have a partner."""
    g = [t for t in got if t[1] >= threshold + 1e-2]
    w = [t for t in want if t[1] >= threshold + 1e-2]

    def partner(t, pool):
        for u in pool:
            if u[0] == t[0] and abs(u[2] - t[2]) <= dpix and abs(u[3] - t[3]) <= dpix and abs(u[1] - t[1]) <= dconf:
                return True
        return False

    miss_g = sum(not partner(t, want) for t in g)
    miss_w = sum(not partner(t, got) for t in w)
    return miss_g <= max_unmatched and miss_w <= max_unmatched
