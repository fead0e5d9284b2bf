#!/bin/bash
tag=${1:-r02l}
mkdir -p gpurun_out
for inf in 1 2; do
  timeout 300 python tools/serve_bench.py --gpus 8 --streams 64 --seconds 10 --inflight $inf > gpurun_out/serve8_inf${inf}_$tag.json 2> gpurun_out/serve8_inf${inf}_${tag}_err.log; tail -2 gpurun_out/serve8_inf${inf}_${tag}_err.log; cut -c1-600 gpurun_out/serve8_inf${inf}_$tag.json
done
timeout 300 python tools/serve_bench.py --gpus 8 --streams 256 --seconds 8 --inflight 2 > gpurun_out/serve8_256_$tag.json 2>> gpurun_out/serve8_inf2_${tag}_err.log; cut -c1-600 gpurun_out/serve8_256_$tag.json
timeout 300 python tools/serve_bench.py --gpus 8 --streams 1024 --seconds 8 --inflight 2 > gpurun_out/serve8_1024_$tag.json 2>> gpurun_out/serve8_inf2_${tag}_err.log; cut -c1-600 gpurun_out/serve8_1024_$tag.json
