#!/bin/bash
tag=${1:-r02c}
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q -s 2>&1 | tail -150 > gpurun_out/pytest_gpu_$tag.log; grep -E "passed|failed|FAILED|ERROR|oracle:|floor" gpurun_out/pytest_gpu_$tag.log | tail -40
timeout 600 python tools/halo_ab.py 2>&1 | tee gpurun_out/halo_ab_$tag.txt
timeout 600 python tools/latency_sweep.py 2>&1 | tee gpurun_out/latency_sweep_$tag.txt
timeout 600 python bench.py --steps 30 --warmup 3 > gpurun_out/bench_$tag.json 2> gpurun_out/bench_${tag}_err.log; tail -3 gpurun_out/bench_${tag}_err.log; python - <<PY
import json
d=json.load(open("gpurun_out/bench_$tag.json"))
print("value",d["value"],"e2e",d["e2e"]["value"],"sync",d["e2e"]["synchronous_fd_detect"],"roofline",d["roofline"]["frac"],d["roofline"]["forward_ms_per_batch"],"parity",d["parity_in_run"]["ok"],"clocks",d["clocks"],"bs1",d.get("bs1_latency_ms"))
PY
timeout 300 python tools/layer_times.py --reps 10 --json gpurun_out/layers_416_$tag.json > gpurun_out/layers_416_$tag.txt 2>&1; tail -1 gpurun_out/layers_416_$tag.txt
timeout 300 python bench.py --config rsu --steps 20 --warmup 3 --quick > gpurun_out/benchrsu_$tag.json 2>> gpurun_out/bench_${tag}_err.log; cut -c1-200 gpurun_out/benchrsu_$tag.json
timeout 300 python bench.py --config 608 --steps 5 --warmup 3 --quick > gpurun_out/bench608_$tag.json 2>> gpurun_out/bench_${tag}_err.log; cut -c1-200 gpurun_out/bench608_$tag.json
tail -5 gpurun_out/bench_${tag}_err.log
