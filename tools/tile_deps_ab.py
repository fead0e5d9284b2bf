#!/usr/bin/env python
"""Developer tool: forward-pass time with and without tile-level dependencies (one process, one box)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fastdet_b200 import _native, modelgen
for size, batch in ((416, 64),):
    onnx = modelgen.build_onnx("full", 80, size, 2)
    frames = np.ascontiguousarray(np.stack([modelgen.synthetic_frame(100 + i, size) for i in range(8)])[np.arange(batch) % 8])
    for rep in range(2):
        for td, mm in ((0, 0), (15, 12000), (13, 45000), (15, 45000), (11, 45000), (13, 180000)):
            _native.set_option("tile_deps", td)
            _native.set_option("tile_deps_max_m", mm)
            m = _native.Model(onnx, 80, (size, size), device=0)
            m.preprocess(frames, batch, (size, size))
            t = min(m.time_forward(batch, 20) for _ in range(3))
            linked = sum(e["tile_linked"] for e in m.exec_info(batch))
            print(f"{size} bs{batch} tile_deps={td} max_m={mm}: forward {t:.4f} ms ({linked} linked layers)", flush=True)
            m.close()
_native.set_option("tile_deps", 0)
