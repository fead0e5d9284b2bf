#!/usr/bin/env python
"""Developer tool: batch-1 (and small-batch) device latency of the forward pass under tiling / split-K options."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fastdet_b200 import _native, modelgen

onnx = modelgen.build_onnx("full", 80, 416, 2)
frames = np.ascontiguousarray(np.stack([modelgen.synthetic_frame(100 + i, 416) for i in range(8)]))
base = dict(latency_bn=1, split_k_min_kb=32, split_k_max=4)
combos = [dict(), dict(latency_bn=0), dict(latency_bn=128), dict(latency_bn=64), dict(split_k=0), dict(pdl=0), dict(graph=0)]
for c in combos:
    opts = dict(base, pdl=1, graph=1)
    opts.update(c)
    for k, v in opts.items():
        _native.set_option(k, v)
    m = _native.Model(onnx, 80, (416, 416), device=0)
    out = []
    for n in (1, 2, 4, 8):
        m.preprocess(frames[:n], n, (416, 416))
        out.append(min(m.time_forward(n, 50) for _ in range(3)))
    ms = m.time_layers(1, 20)
    top = np.argsort(-ms)[:6]
    print(f"{c or 'default'}: forward bs1/2/4/8 {out[0]:.4f} {out[1]:.4f} {out[2]:.4f} {out[3]:.4f} ms; layers alone bs1 sum {ms.sum():.3f} ms; slowest " +
          " ".join(f"{i}:{ms[i]*1e3:.0f}us" for i in top), flush=True)
    m.close()
for k, v in dict(base, pdl=1, graph=1).items():
    _native.set_option(k, v)
