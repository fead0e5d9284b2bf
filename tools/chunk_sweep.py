#!/usr/bin/env python
"""Developer tool: device time of the forward pass (captured graph) under different chunking options."""
import argparse, json, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fastdet_b200 import _native, modelgen

ap = argparse.ArgumentParser()
ap.add_argument("--size", type=int, default=416); ap.add_argument("--batch", type=int, default=64)
ap.add_argument("--reps", type=int, default=20)
ap.add_argument("--frames", default="-1,1,2,4,8,0")
ap.add_argument("--interleave", type=int, default=0)
a = ap.parse_args()
onnx = modelgen.build_onnx("full", 80, a.size, 2)
frames = np.stack([modelgen.synthetic_frame(100 + i, a.size) for i in range(8)])[np.arange(a.batch) % 8]
_native.set_option("chunk_interleave", a.interleave)
for cf in [int(x) for x in a.frames.split(",")]:
    _native.set_option("chunk_frames", cf)
    m = _native.Model(onnx, 80, (a.size, a.size), device=0)
    m.preprocess(np.ascontiguousarray(frames), a.batch, (a.size, a.size))
    t = min(m.time_forward(a.batch, a.reps) for _ in range(3))
    info = m.exec_info(a.batch)
    ms = m.time_layers(a.batch, 5)
    print(f"chunk_frames={cf:3d}: forward {t:.4f} ms; chunks {[e['chunk_frames'] for e in info[:10]]}; layers 0-8: "
          + " ".join(f"{x:.3f}" for x in ms[:9]) + f" = {ms[:9].sum():.3f} ms; all layers {ms.sum():.3f} ms", flush=True)
    m.close()
_native.set_option("chunk_frames", 0)
