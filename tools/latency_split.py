#!/usr/bin/env python
"""Developer tool: where the batch-1 latency of fd_detect goes (H2D, forward graph, decode + Soft-NMS, D2H)."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fastdet_b200 import _native, modelgen
m = _native.Model(modelgen.build_onnx("full", 80, 416, 2), 80, (416, 416), device=0)
frame = torch.from_numpy(modelgen.synthetic_frame(100, 416)[None]).pin_memory()
st = torch.cuda.Stream(); torch.cuda.set_stream(st); sp = st.cuda_stream
ev = [torch.cuda.Event(enable_timing=True) for _ in range(5)]
acc = np.zeros(4); wall = []
for i in range(120):
    t0 = time.perf_counter()
    ev[0].record(st)
    m.preprocess(frame.data_ptr(), 1, (416, 416), stream=sp); ev[1].record(st)
    m.forward(1, stream=sp); ev[2].record(st)
    m.postprocess(1, 0.1, max_det=256, stream=sp); ev[3].record(st)
    d, c, t = m.fetch(1, stream=sp)
    wall.append((time.perf_counter() - t0) * 1e3)
    if i >= 20:
        acc += [ev[k].elapsed_time(ev[k + 1]) for k in range(3)] + [0]
acc /= 100
print(f"device ms: H2D {acc[0]:.3f}  forward {acc[1]:.3f}  decode+nms+D2H {acc[2]:.3f}   wall p50 {np.percentile(wall[20:], 50):.3f} ms, detections {c[0]}")
