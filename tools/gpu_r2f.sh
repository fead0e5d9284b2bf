#!/bin/bash
tag=${1:-r02f}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_server_dispatch.py -m gpu -q -x 2>&1 | tail -15
for n in 2 1; do
  timeout 300 python bench.py --config serve --gpus $n --seconds 8 > gpurun_out/serve${n}_$tag.json 2> gpurun_out/serve${n}_${tag}_err.log; tail -2 gpurun_out/serve${n}_${tag}_err.log
  cut -c1-120 gpurun_out/serve${n}_$tag.json; python -c "
import json
d=json.load(open('gpurun_out/serve${n}_$tag.json')); print('serve n=$n fps',d['value'],'lat',d['latency_ms'],'mean_batch',d['mean_batch'],'per_dev',d['frames_per_device'])"
done
