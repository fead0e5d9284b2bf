import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.getcwd())
from fastdet_b200 import _native, modelgen
m = _native.Model(modelgen.build_onnx("full", 80, 416, 2), 80, (416, 416), device=0)
n = 64
frames = np.stack([modelgen.synthetic_frame(100 + i, 416) for i in range(4)])[np.arange(n) % 4]
m.preprocess(np.ascontiguousarray(frames), n, (416, 416)); m.forward(n)
st = torch.cuda.Stream(); torch.cuda.set_stream(st); sp = st.cuda_stream
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for i in range(5): m.postprocess(n, 0.1, max_det=256, stream=sp)
torch.cuda.synchronize()
e0.record(st)
for i in range(50): m.postprocess(n, 0.1, max_det=256, stream=sp)
e1.record(st); torch.cuda.synchronize()
d, c, t = m.fetch(n, stream=sp)
print("postprocess (decode + soft-nms + D2H) ms:", e0.elapsed_time(e1) / 50, "detections", int(c.sum()))
