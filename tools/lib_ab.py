#!/usr/bin/env python
"""Developer tool: forward-pass time of two builds of the library on one box (FASTDET_LIB selects the build; one process each)."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CODE = r'''
import os, sys, numpy as np
sys.path.insert(0, %r)
from fastdet_b200 import _native, modelgen
onnx = modelgen.build_onnx("full", 80, 416, 2)
frames = np.ascontiguousarray(np.stack([modelgen.synthetic_frame(100 + i, 416) for i in range(8)])[np.arange(64) %% 8])
m = _native.Model(onnx, 80, (416, 416), device=0)
m.preprocess(frames, 64, (416, 416))
print(os.path.basename(os.environ.get("FASTDET_LIB", "HEAD")), " ".join("%%.4f" %% m.time_forward(64, 20) for _ in range(4)), flush=True)
''' % ROOT
libs = [None] + [os.path.join(ROOT, "abtest", f) for f in sorted(os.listdir(os.path.join(ROOT, "abtest"))) if f.endswith(".so")]
for rep in range(3):
    for lib in libs:
        env = dict(os.environ)
        if lib:
            env["FASTDET_LIB"] = lib
        subprocess.run([sys.executable, "-c", CODE], env=env)
