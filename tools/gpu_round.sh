#!/bin/bash
# usage (GPU box): bash tools/gpu_round.sh <tag>   -- tests, dev-harness check, bench with/without PDL, layer table
tag=${1:-x}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/pytest_gpu_$tag.log
timeout 120 ./build/test_conv check 2>&1 | grep -E "FAIL|failing|rror" | head
python bench.py --steps 30 --warmup 3 --quick > gpurun_out/bench_${tag}_pdl.json 2> gpurun_out/bench_${tag}_err.log; tail -3 gpurun_out/bench_${tag}_err.log
FASTDET_NO_PDL=1 python bench.py --steps 30 --warmup 3 --quick > gpurun_out/bench_${tag}_nopdl.json 2>> gpurun_out/bench_${tag}_err.log
python - <<PY
import json
for k in ("pdl","nopdl"):
    try:
        d=json.loads(open("gpurun_out/bench_${tag}_%s.json"%k).read().strip().splitlines()[-1])
        print(k, d["value"], d["ms_per_step"], "fwd", d["roofline"]["forward_ms_per_step"], "frac", d["roofline"]["frac"], "e2e", d["e2e"], d["clocks"])
    except Exception as e: print(k, "failed", e)
PY
python tools/layer_times.py --reps 10 --json gpurun_out/layers_$tag.json > gpurun_out/layers_$tag.txt 2>&1; tail -1 gpurun_out/layers_$tag.txt
