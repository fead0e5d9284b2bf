#!/usr/bin/env python
"""Developer tool: A/B of the halo-patch kernel's plane skew and ring depth on one box, one process (layer-alone times)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fastdet_b200 import _native, modelgen

onnx = modelgen.build_onnx("full", 80, 416, 2)
frames = np.ascontiguousarray(np.stack([modelgen.synthetic_frame(100 + i, 416) for i in range(8)])[np.arange(64) % 8])
for rep in range(2):
    for skew, slots in [(0, 8), (1, 8), (0, 0), (1, 0)]:
        _native.set_option("halo_skew", skew)
        _native.set_option("halo_slots", slots)
        m = _native.Model(onnx, 80, (416, 416), device=0)
        m.preprocess(frames, 64, (416, 416))
        ms = m.time_layers(64, 10)
        fw = min(m.time_forward(64, 20) for _ in range(2))
        print(f"skew={skew} slots={slots or 'max'}: halo layers conv2 {ms[1]:.4f} conv4 {ms[3]:.4f} conv7 {ms[6]:.4f} conv9 {ms[8]:.4f}; "
              f"layers 0-8 {ms[:9].sum():.4f}; all {ms.sum():.4f}; forward {fw:.4f} ms", flush=True)
        m.close()
_native.set_option("halo_skew", 1)
_native.set_option("halo_slots", 0)
