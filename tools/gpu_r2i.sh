#!/bin/bash
tag=${1:-r02i}
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q 2>&1 | tail -15 > gpurun_out/pytest_gpu_$tag.log; tail -6 gpurun_out/pytest_gpu_$tag.log
timeout 300 ./build/test_conv check > gpurun_out/test_conv_$tag.log 2>&1; grep -E "FAIL|failing|rror" gpurun_out/test_conv_$tag.log | head -5
timeout 600 python bench.py --steps 30 --warmup 3 > gpurun_out/bench_$tag.json 2> gpurun_out/bench_${tag}_err.log; tail -3 gpurun_out/bench_${tag}_err.log; python - <<PY
import json
d=json.load(open("gpurun_out/bench_$tag.json"))
print("value",d["value"],"e2e",d["e2e"]["value"],"sync",d["e2e"]["synchronous_fd_detect"],"jpeg",d["e2e"]["from_jpeg"]["value"],"roofline",d["roofline"]["frac"],d["roofline"]["forward_ms_per_batch"],"parity",d["parity_in_run"]["ok"],"clocks",d["clocks"],"bs1",d.get("bs1_latency_ms"), "pre", d["roofline_pre"]["frac"], "post", d["roofline_post"]["ms"])
PY
timeout 300 python bench.py --impl reference --steps 10 --warmup 2 > gpurun_out/benchref_$tag.json 2>> gpurun_out/bench_${tag}_err.log; cut -c1-160 gpurun_out/benchref_$tag.json
timeout 300 python bench.py --config tiny-cpu > gpurun_out/tinycpu_$tag.json 2>> gpurun_out/bench_${tag}_err.log; cut -c1-900 gpurun_out/tinycpu_$tag.json
timeout 300 python tools/layer_times.py --reps 10 --json gpurun_out/layers_416_$tag.json > gpurun_out/layers_416_$tag.txt 2>&1; tail -1 gpurun_out/layers_416_$tag.txt
timeout 300 python tools/layer_times.py --size 608 --batch 64 --reps 5 --json gpurun_out/layers_608_$tag.json > gpurun_out/layers_608_$tag.txt 2>&1; tail -1 gpurun_out/layers_608_$tag.txt
timeout 300 python tools/latency_sweep.py 2>&1 | tee gpurun_out/latency_sweep_$tag.txt | cut -c1-300
bash tools/profile_round.sh $tag 2>&1 | tail -8
