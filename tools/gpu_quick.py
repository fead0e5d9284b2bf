#!/usr/bin/env python
"""Developer tool: head-tensor error of the GPU path against oracle outputs precomputed on the build box
(.devcache/ref_*.npz), so no CPU oracle time is spent on the GPU box.  Falls back to computing the oracle
when the regenerated model differs bit-wise (different CPU -> different calibration)."""
import hashlib, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fastdet_b200 import _native, modelgen
from oracle import ref_graph, ref_post

for arch, nc, seed in (("tiny", 80, 1), ("full", 80, 2), ("rsu", 9, 3)):
    data = modelgen.build_onnx(arch, nc, 416, seed)
    z = np.load(f".devcache/ref_{arch}.npz")
    frames = np.stack([modelgen.synthetic_frame(100 + i) for i in range(2)])
    same = hashlib.sha256(data).hexdigest() == str(z["sha"])
    if same:
        ref = [z[f"h{i}"] for i in range(len([k for k in z.files if k.startswith("h")]))]
    else:
        t = time.time()
        ref = ref_graph.GraphExecutor(data).run(np.concatenate([ref_post.normalise(f) for f in frames]))
        print(arch, "model bytes differ from the build box; oracle recomputed in", round(time.time() - t, 1), "s")
    m = _native.Model(data, nc, (416, 416), device=0)
    m.preprocess(frames, 2, (416, 416)); m.forward(2)
    got = m.heads(2)
    print(arch, "same_bytes", same, "max_rel", [f"{np.abs(g - r).max() / np.abs(r).max():.2e}" for g, r in zip(got, ref)],
          "rms_rel", [f"{np.sqrt(np.mean((g - r) ** 2)) / np.sqrt(np.mean(r ** 2)):.2e}" for g, r in zip(got, ref)])
    m.postprocess(2, 0.1); dets, counts, _ = m.fetch(2)
    want = [len(ref_post.detect_from_heads(ref, f, nc, (416, 416), 0.1)[0]) for f in range(2)]
    print("   detections gpu", counts.tolist(), "oracle", want)
    ms = m.time_layers(2, 3)
    print("   forward bs2: sum of layer times %.3f ms" % ms.sum())
    m.close()
