#!/usr/bin/env python
"""Developer tool: head error against the fp32 oracle with the fused stem (default) and with the two-kernel path (option
stem=0) on the same frames, and how far the two plans' second-layer outputs are apart."""
import argparse, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fastdet_b200 import _native, modelgen
from oracle import ref_graph, ref_post

ap = argparse.ArgumentParser()
ap.add_argument("--arch", default="rsu"); ap.add_argument("--classes", type=int, default=9)
ap.add_argument("--size", type=int, default=416); ap.add_argument("--batch", type=int, default=3)
ap.add_argument("--seed", type=int, default=3); ap.add_argument("--first", type=int, default=120)
a = ap.parse_args()
data = modelgen.build_onnx(a.arch, a.classes, a.size, a.seed)
frames = np.stack([modelgen.synthetic_frame(a.first + i, a.size) for i in range(a.batch)])
x = np.concatenate([ref_post.normalise(f) for f in frames])
exe = ref_graph.GraphExecutor(data)
want = exe.run(x)
outs = {}
for stem in (1, 0):
    with _native.option("stem", stem):
        m = _native.Model(data, a.classes, (a.size, a.size), device=0)
        m.preprocess(frames, a.batch, (a.size, a.size)); m.forward(a.batch)
        got = m.heads(a.batch)
        L = m.layers()
        ref1 = exe.run(x, keep=[L[1]["out_name"]])[L[1]["out_name"]]
        l1 = m.layer_output(1, a.batch)
        outs[stem] = l1
        print("stem", stem, [e["kernel_name"] for e in m.exec_info(a.batch)[:2]],
              "head err / max|ref|:", [round(float(np.abs(g - r).max() / np.abs(r).max()), 5) for g, r in zip(got, want)],
              "layer1 rms rel", float(np.sqrt(np.mean((l1 - ref1) ** 2)) / np.sqrt(np.mean(ref1 ** 2))), "max", float(np.abs(l1 - ref1).max()))
        m.close()
d = np.abs(outs[1] - outs[0])
print("layer1 stem vs two-kernel: differing", float((d > 0).mean()), "max", float(d.max()))
