#!/bin/bash
# 2-GPU sanity pass (gpurun --gpus 2): the multi-GPU dispatcher test that a 1-GPU box skips, config 5 on 2 GPUs, the headline and
# config 4 under torchrun at N = 2
tag=${1:-two}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -4
timeout 600 python -m pytest tests/test_server_dispatch.py -m gpu -q --tb=line 2>&1 | tail -3
timeout 300 python bench.py --config serve --gpus 2 --seconds 8 > gpurun_out/serve2_$tag.json 2> gpurun_out/two_${tag}_err.log; cut -c1-300 gpurun_out/serve2_$tag.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29502 bench.py --gpus 2 --steps 50 --warmup 3 --quick > gpurun_out/bench_2_$tag.json 2>> gpurun_out/two_${tag}_err.log; cut -c1-260 gpurun_out/bench_2_$tag.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29503 bench.py --gpus 2 --config 608 --steps 10 --warmup 3 --quick > gpurun_out/bench608_2_$tag.json 2>> gpurun_out/two_${tag}_err.log; cut -c1-260 gpurun_out/bench608_2_$tag.json
tail -3 gpurun_out/two_${tag}_err.log
