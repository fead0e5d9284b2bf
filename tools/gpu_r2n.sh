#!/bin/bash
tag=${1:-r02n}
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q 2>&1 | tail -12 > gpurun_out/pytest_gpu_$tag.log; tail -6 gpurun_out/pytest_gpu_$tag.log
timeout 300 ./build/test_conv check > gpurun_out/test_conv_$tag.log 2>&1; grep -E "FAIL|failing|rror" gpurun_out/test_conv_$tag.log | head -5
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python bench.py --steps 30 --warmup 3 > gpurun_out/bench_$tag.json 2> gpurun_out/bench_${tag}_err.log; tail -3 gpurun_out/bench_${tag}_err.log; python - <<PY
import json
d=json.load(open("gpurun_out/bench_$tag.json"))
print("value",d["value"],"e2e",d["e2e"]["value"],"sync",d["e2e"]["synchronous_fd_detect"],"jpeg",d["e2e"]["from_jpeg"]["value"],"roofline",d["roofline"]["frac"],d["roofline"]["forward_ms_per_batch"],"parity",d["parity_in_run"]["ok"],"clocks",d["clocks"],"bs1",d.get("bs1_latency_ms"))
PY
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 2>> gpurun_out/bench_${tag}_err.log | cut -c1-200
