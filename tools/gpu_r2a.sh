#!/bin/bash
# round-2 GPU pass A: tests, kernel checker, bench baseline, layer tables, ncu of the early-layer / 1x1 kernels
tag=${1:-r02a}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv,noheader | head -1
timeout 1500 python -m pytest tests -m gpu -q -x -s 2>&1 | tail -60 > gpurun_out/pytest_gpu_$tag.log; tail -25 gpurun_out/pytest_gpu_$tag.log
timeout 300 ./build/test_conv check > gpurun_out/test_conv_$tag.log 2>&1; grep -E "FAIL|failing|rror|strip" gpurun_out/test_conv_$tag.log | head -20
timeout 600 python bench.py --steps 30 --warmup 3 > gpurun_out/bench_$tag.json 2> gpurun_out/bench_${tag}_err.log; tail -3 gpurun_out/bench_${tag}_err.log; cut -c1-1500 gpurun_out/bench_$tag.json
timeout 300 python tools/layer_times.py --reps 10 --json gpurun_out/layers_416_$tag.json > gpurun_out/layers_416_$tag.txt 2>&1; tail -1 gpurun_out/layers_416_$tag.txt
timeout 300 python tools/layer_times.py --size 608 --batch 64 --reps 5 --json gpurun_out/layers_608_$tag.json > gpurun_out/layers_608_$tag.txt 2>&1; tail -1 gpurun_out/layers_608_$tag.txt
CMD="python bench.py --steps 2 --warmup 3 --quick --no-parity"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"conv0_ws|conv_halo" -c 8 -o gpurun_out/prof_early_$tag $CMD > gpurun_out/ncu_early_$tag.log 2>&1; tail -2 gpurun_out/ncu_early_$tag.log
# 1x1 layers of the three stages + one stride-2 3x3: launches 9..12 (conv10 s2, conv11 1x1 swap, conv12 strip), 26..28, 43..45 of the conv_tc sequence
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s 4 -c 3 -o gpurun_out/prof_tc52_$tag $CMD > gpurun_out/ncu_tc52_$tag.log 2>&1; tail -1 gpurun_out/ncu_tc52_$tag.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s 21 -c 3 -o gpurun_out/prof_tc26_$tag $CMD > gpurun_out/ncu_tc26_$tag.log 2>&1; tail -1 gpurun_out/ncu_tc26_$tag.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s 38 -c 3 -o gpurun_out/prof_tc13_$tag $CMD > gpurun_out/ncu_tc13_$tag.log 2>&1; tail -1 gpurun_out/ncu_tc13_$tag.log
ls -la gpurun_out | tail -20
