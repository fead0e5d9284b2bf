#!/usr/bin/env python
"""Developer tool: run one synthetic model on the GPU and print per-layer error vs the fp32 CPU oracle."""
import argparse
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fastdet_b200 import _native, modelgen  # noqa: E402
from oracle import ref_graph, ref_post  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--arch", default="tiny")
    ap.add_argument("--classes", type=int, default=80)
    ap.add_argument("--size", type=int, default=416)
    ap.add_argument("--batch", type=int, default=2)
    ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("--unfolded", action="store_true")
    ap.add_argument("--thr", type=float, default=0.1)
    a = ap.parse_args()
    opts = modelgen.ExportOptions(fold_bn=not a.unfolded)
    onnx = modelgen.build_onnx(a.arch, a.classes, a.size, a.seed, opts)
    frames = np.stack([modelgen.synthetic_frame(100 + i, a.size) for i in range(a.batch)])
    m = _native.Model(onnx, a.classes, (a.size, a.size), device=0)
    t0 = time.time()
    m.preprocess(frames, a.batch, (a.size, a.size))
    m.forward(a.batch)
    heads = m.heads(a.batch)
    print(f"gpu forward (first call, incl. graph capture): {time.time() - t0:.3f}s")
    exe = ref_graph.GraphExecutor(onnx)
    x = np.concatenate([ref_post.normalise(f) for f in frames])
    t0 = time.time()
    vals = exe.run(x, all_values=True)
    print(f"oracle forward: {time.time() - t0:.2f}s")
    worst = 0.0
    for i, L in enumerate(m.layers()):
        ref = vals.get(L["out_name"])
        if ref is None:
            print(i, L["name"], "no oracle value for", L["out_name"])
            continue
        got = m.layer_output(i, a.batch)
        err = np.abs(got - ref).max()
        rel = err / max(np.abs(ref).max(), 1e-9)
        rms = np.sqrt(np.mean((got - ref) ** 2)) / max(np.sqrt(np.mean(ref ** 2)), 1e-9)
        worst = max(worst, rel)
        print(f"{i:3d} kind={L['kind']} {L['name']:<10} {L['out_name']:<14} c={L['c']:<5} hw={L['h']:<4} "
              f"max|ref|={np.abs(ref).max():8.3f} rms(ref)={np.sqrt(np.mean(ref**2)):7.3f} maxerr/max={rel:.2e} rms_rel={rms:.2e}")
    for h, (g, r) in enumerate(zip(heads, exe.run(x))):
        print(f"head{h}: max|gpu-ref|/max|ref| = {np.abs(g - r).max() / np.abs(r).max():.3e}  rms_rel = "
              f"{np.sqrt(np.mean((g - r) ** 2)) / np.sqrt(np.mean(r ** 2)):.3e}")
    ref_heads = exe.run(x)
    m.postprocess(a.batch, a.thr)
    dets, counts, total = m.fetch(a.batch)
    for f in range(a.batch):
        ref_res, _, _ = ref_post.detect_from_heads(ref_heads, f, a.classes, (a.size, a.size), a.thr)
        print(f"frame {f}: gpu {counts[f]} detections (total {total[f]}), oracle {len(ref_res)}")
        for d, r in list(zip(dets[f, :counts[f]], ref_res))[:5]:
            print("   gpu", int(d["klass"]), f"{d['conf']:.4f} {d['x']:.2f} {d['y']:.2f} {d['w']:.2f} {d['h']:.2f}",
                  "| ref", r[0], f"{r[1]:.4f} {r[2]:.2f} {r[3]:.2f} {r[4]:.2f} {r[5]:.2f}")
    print("worst layer rel err", worst)


if __name__ == "__main__":
    main()
