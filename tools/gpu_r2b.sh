#!/bin/bash
# round-2 GPU pass B: full GPU test suite, chunk sweep, bench, layer tables, other bench configs
tag=${1:-r02b}
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q -s 2>&1 | tail -120 > gpurun_out/pytest_gpu_$tag.log; grep -E "passed|failed|FAILED|ERROR|oracle:" gpurun_out/pytest_gpu_$tag.log | tail -40
timeout 600 python tools/chunk_sweep.py 2>&1 | tee gpurun_out/chunk_sweep_$tag.txt | cut -c1-400
timeout 300 python tools/chunk_sweep.py --interleave 1 --frames 2,4 2>&1 | tee -a gpurun_out/chunk_sweep_$tag.txt | cut -c1-400
timeout 600 python bench.py --steps 30 --warmup 3 > gpurun_out/bench_$tag.json 2> gpurun_out/bench_${tag}_err.log; tail -3 gpurun_out/bench_${tag}_err.log; python - <<PY
import json
d=json.load(open("gpurun_out/bench_$tag.json"))
print("value",d["value"],"e2e",d["e2e"]["value"],"sync",d["e2e"]["synchronous_fd_detect"],"roofline",d["roofline"]["frac"],d["roofline"]["forward_ms_per_batch"],"parity",d["parity_in_run"]["ok"],"clocks",d["clocks"],"bs1",d.get("bs1_latency_ms"))
PY
timeout 300 python tools/layer_times.py --reps 10 --json gpurun_out/layers_416_$tag.json > gpurun_out/layers_416_$tag.txt 2>&1; head -11 gpurun_out/layers_416_$tag.txt; tail -1 gpurun_out/layers_416_$tag.txt
timeout 300 python tools/layer_times.py --size 608 --batch 64 --reps 5 --json gpurun_out/layers_608_$tag.json > gpurun_out/layers_608_$tag.txt 2>&1; tail -1 gpurun_out/layers_608_$tag.txt
timeout 300 python bench.py --config 608 --steps 5 --warmup 3 --quick > gpurun_out/bench608_$tag.json 2>> gpurun_out/bench_${tag}_err.log; cut -c1-300 gpurun_out/bench608_$tag.json
timeout 300 python bench.py --config serve --gpus 1 --seconds 5 > gpurun_out/serve1_$tag.json 2>> gpurun_out/bench_${tag}_err.log; cat gpurun_out/serve1_$tag.json | cut -c1-1200
tail -5 gpurun_out/bench_${tag}_err.log
