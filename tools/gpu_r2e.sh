#!/bin/bash
tag=${1:-r02e}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "full_heads_and_layers or batch_invariance or rsu_heads or tiny_heads" 2>&1 | tail -30
timeout 300 python tools/serve_bench.py --gpus 1 --streams 8 --seconds 2 --models full 2>&1 | tail -3 | cut -c1-600
timeout 300 python tools/serve_bench.py --gpus 1 --streams 64 --seconds 3 2>&1 | tail -3 | cut -c1-600
for lb in 0 64 128; do
python - <<PY
import numpy as np
from fastdet_b200 import _native, modelgen
from oracle import ref_graph, ref_post
_native.set_option("latency_bn", $lb)
data = modelgen.build_onnx("full", 80, 416, 2)
m = _native.Model(data, 80, (416,416), device=0)
exe = ref_graph.GraphExecutor(data)
for n in (1, 2, 4):
    frames = np.stack([modelgen.synthetic_frame(100+i, 416) for i in range(n)])
    m.preprocess(frames, n, (416,416)); m.forward(n)
    x = np.concatenate([ref_post.normalise(f) for f in frames])
    vals = exe.run(x, all_values=True)
    info = m.exec_info(n)
    worst = []
    for i, L in enumerate(m.layers()):
        out = m.layer_output(i, n); ref = vals[L["out_name"]]
        rms = float(np.sqrt(np.mean((out-ref)**2))/np.sqrt(np.mean(ref**2)))
        if not rms < 2e-2: worst.append((i, L["name"], info[i]["kernel_name"], info[i]["block_n"], info[i]["split_k"], round(rms,4)))
    print("latency_bn=$lb n=%d bad layers:" % n, worst[:6])
PY
done
