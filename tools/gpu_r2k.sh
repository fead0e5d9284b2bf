#!/bin/bash
tag=${1:-r02k}
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_server_dispatch.py -m gpu -q 2>&1 | tail -4
for n in 8 4; do
  timeout 400 python bench.py --config serve --gpus $n --seconds 15 > gpurun_out/serve${n}_$tag.json 2> gpurun_out/serve${n}_${tag}_err.log; tail -2 gpurun_out/serve${n}_${tag}_err.log
  python -c "
import json
d=json.load(open('gpurun_out/serve${n}_$tag.json')); print('serve n=$n fps',d['value'],'lat',d['latency_ms'],'mean_batch',d['mean_batch'],'per_dev',d['frames_per_device'],'load_s',d['model_load_seconds'])"
done
