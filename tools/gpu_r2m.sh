#!/bin/bash
tag=${1:-r02m}
mkdir -p gpurun_out
timeout 600 python tools/tile_deps_ab.py 2>&1 | tee gpurun_out/tile_deps_ab_$tag.txt
timeout 900 python -m pytest tests/test_gpu_production_paths.py -m gpu -q -x 2>&1 | tail -15
timeout 600 python bench.py --steps 30 --warmup 3 --quick > gpurun_out/bench_$tag.json 2> gpurun_out/bench_${tag}_err.log; tail -3 gpurun_out/bench_${tag}_err.log; python - <<PY
import json
d=json.load(open("gpurun_out/bench_$tag.json"))
print("value",d["value"],"e2e",d["e2e"]["value"],"sync",d["e2e"]["synchronous_fd_detect"],"roofline",d["roofline"]["frac"],d["roofline"]["forward_ms_per_batch"],"parity",d["parity_in_run"],"clocks",d["clocks"])
PY
