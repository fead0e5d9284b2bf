#!/bin/bash
# round-2 multi-GPU pass (gpurun --gpus 8): config 5 (serve) at 1/2/4/8 GPUs, config 4 (608 sharded) at 2/4/8, headline scaling 2/4/8
tag=${1:-r02d}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
nproc
run_tr() {  # n_gpus, extra bench args..., output file
  local n=$1; shift; local out=$1; shift
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + n)) bench.py --gpus $n "$@" > $out 2>> gpurun_out/multi_${tag}_err.log
  cut -c1-260 $out
}
for n in 8 4 2 1; do
  timeout 300 python bench.py --config serve --gpus $n --seconds 10 > gpurun_out/serve${n}_$tag.json 2>> gpurun_out/multi_${tag}_err.log
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/serve${n}_$tag.json"))
    print("serve n=$n fps",d["value"],"lat",d["latency_ms"],"mean_batch",d["mean_batch"],"per_dev",d["frames_per_device"])
except Exception as e: print("serve $n failed", e)
PY
done
for n in 8 4 2; do run_tr $n gpurun_out/bench608_${n}_$tag.json --config 608 --steps 10 --warmup 3 --quick; done
timeout 300 python bench.py --config 608 --steps 10 --warmup 3 --quick > gpurun_out/bench608_1_$tag.json 2>> gpurun_out/multi_${tag}_err.log; cut -c1-200 gpurun_out/bench608_1_$tag.json
for n in 8 4 2; do run_tr $n gpurun_out/bench_${n}_$tag.json --steps 50 --warmup 3 --quick; done
timeout 300 python bench.py --steps 50 --warmup 3 --quick > gpurun_out/bench_1_$tag.json 2>> gpurun_out/multi_${tag}_err.log; cut -c1-200 gpurun_out/bench_1_$tag.json
tail -5 gpurun_out/multi_${tag}_err.log
