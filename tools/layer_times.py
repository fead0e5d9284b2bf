#!/usr/bin/env python
"""Developer tool: per-layer device time of the forward pass, with each layer's tensor / HBM roofline."""
import argparse, json, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fastdet_b200 import _native, modelgen

ap = argparse.ArgumentParser()
ap.add_argument("--arch", default="full"); ap.add_argument("--classes", type=int, default=80)
ap.add_argument("--size", type=int, default=416); ap.add_argument("--batch", type=int, default=64)
ap.add_argument("--reps", type=int, default=10); ap.add_argument("--seed", type=int, default=2)
ap.add_argument("--json", default="")
a = ap.parse_args()
peaks = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))
PF, BW = peaks["bf16_tflops"] * 1e12, peaks["hbm_gbs"] * 1e9
m = _native.Model(modelgen.build_onnx(a.arch, a.classes, a.size, a.seed), a.classes, (a.size, a.size), device=0)
n = a.batch
frames = np.stack([modelgen.synthetic_frame(100 + i, a.size) for i in range(4)])[np.arange(n) % 4]
m.preprocess(np.ascontiguousarray(frames), n, (a.size, a.size)); m.forward(n)
ms = m.time_layers(n, a.reps)
L = m.layers()
rows = []
tot_fl = 0.0
prev = (3, a.size, a.size)
print(f"{'#':>3} {'name':<9} {'cin':>5}->{'c':<5} k s {'hw':>4} {'bn':>3} {'ms':>8} {'TFLOP/s':>8} {'%tens':>6} {'GB/s':>7} {'%hbm':>5} {'t_roof':>7} {'eff':>5}")
X = m.exec_info(n)
for i, (l, t) in enumerate(zip(L, ms)):
    if X[i]["kernel_name"] == "fused_next":  # no launch of its own: its work and FLOPs are in the next row (conv_stem_kernel)
        print(f"{i:3d} {l['name']:<9} {l['cin']:5d}->{l['c']:<5d} {l['ksize']} {l['stride']} {l['h']:4d}  (computed inside the next layer's kernel)")
        rows.append(dict(i=i, name=l["name"], ms=0.0, flops=0.0, bytes=0, t_roof_ms=0.0))
        continue
    fl = l["flops"] * n + (L[i - 1]["flops"] * n if X[i]["kernel_name"] in ("stem", "block") else 0.0)
    hw_out = l["h"] * l["w"]
    if l["kind"] in (0, 1):
        ho = l["h"] // (2 if l["upsample2x"] else 1)
        hin = ho * l["stride"]
        in_b = n * hin * hin * l["cin"] * (1 if l["kind"] == 0 else 2)
        if X[i]["kernel_name"] == "stem":
            in_b = n * hin * hin * 3  # the u8 frames
        out_b = n * hw_out * l["c"] * (4 if l["out_fp32"] else 2)
        w_b = l["c"] * l["cin"] * l["ksize"] ** 2 * 2
        res_b = out_b if l["has_residual"] else 0
        if X[i]["kernel_name"] == "block":  # reads the block's input once (it is also the residual), writes the output once
            in_b, res_b = out_b, 0
        byts = in_b + out_b + w_b + res_b
    else:
        byts = n * hw_out * l["c"] * 2 * 2
    t_roof = max(fl / PF, byts / BW) * 1e3
    rows.append(dict(i=i, name=l["name"], ms=float(t), flops=fl, bytes=byts, t_roof_ms=t_roof))
    tot_fl += fl
    print(f"{i:3d} {l['name']:<9} {l['cin']:5d}->{l['c']:<5d} {l['ksize']} {l['stride']} {l['h']:4d} {l['block_n']:3d} {t:8.4f} "
          f"{fl / t * 1e-9 if t else 0:8.1f} {100 * fl / t * 1e3 / PF if t else 0:6.1f} {byts / t * 1e-6:7.0f} {100 * byts / t * 1e3 / BW:5.1f} {t_roof:7.4f} {100 * t_roof / t:5.1f}")
T = float(ms.sum())
print(f"total {T:.3f} ms for batch {n}: {n / T * 1e3:.0f} frames/s (layers timed alone), {tot_fl / T * 1e-9:.1f} TFLOP/s = {100 * tot_fl / T * 1e3 / PF:.1f}% of {PF*1e-12:.0f} TF burst peak; "
      f"sum of per-layer rooflines {sum(r['t_roof_ms'] for r in rows):.3f} ms")
if a.json:
    json.dump(dict(arch=a.arch, batch=n, size=a.size, total_ms=T, rows=rows), open(a.json, "w"), indent=1)
