#!/bin/bash
# One GPU pass: driver-style tests, kernel checkers, smoke, the bench line (+ reference arm), per-layer table.
tag=${1:-pass}
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q --tb=line 2>&1 | cut -c1-400 > gpurun_out/pytest_gpu_$tag.log; tail -4 gpurun_out/pytest_gpu_$tag.log
timeout 300 ./build/test_conv check > gpurun_out/test_conv_$tag.log 2>&1; grep -E "FAIL|failing|rror" gpurun_out/test_conv_$tag.log | head -5
timeout 300 ./build/test_stem all > gpurun_out/test_stem_$tag.log 2>&1; grep -E "FAIL|failing|rror|time n" gpurun_out/test_stem_$tag.log | cut -c1-200 | head -8
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python bench.py --steps 30 --warmup 3 > gpurun_out/bench_$tag.json 2> gpurun_out/bench_${tag}_err.log; tail -3 gpurun_out/bench_${tag}_err.log; python - <<PY
import json
d=json.load(open("gpurun_out/bench_$tag.json"))
print("value",d["value"],"e2e",d["e2e"]["value"],"sync",d["e2e"]["synchronous_fd_detect"],"jpeg",d["e2e"]["from_jpeg"]["value"],"roofline",d["roofline"]["frac"],d["roofline"]["forward_ms_per_batch"],"parity",d["parity_in_run"],"clocks",d["clocks"],"bs1",d.get("bs1_latency_ms",{}).get("p50"),"pre",d["roofline_pre"]["ms"],d["roofline_pre"]["frac"])
PY
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 2>> gpurun_out/bench_${tag}_err.log | cut -c1-200
timeout 300 python tools/layer_times.py --json gpurun_out/layers_416_$tag.json > gpurun_out/layers_416_$tag.txt 2>&1; tail -1 gpurun_out/layers_416_$tag.txt
timeout 300 python tools/stem_ab.py > gpurun_out/stem_ab_$tag.txt 2>&1; tail -3 gpurun_out/stem_ab_$tag.txt
