#!/bin/bash
# usage (GPU box): bash tools/profile_list.sh <tag>  -- plain bench run (must exit 0), then the ncu launch list of the same command
tag=${1:-x}
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --quick"
$CMD > gpurun_out/plain_$tag.log 2> gpurun_out/plain_${tag}_err.log || { echo "plain run failed"; tail -5 gpurun_out/plain_${tag}_err.log; exit 1; }
tail -c 600 gpurun_out/plain_$tag.log
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 2000 --csv \
    --log-file gpurun_out/launches_$tag.csv $CMD > gpurun_out/ncu_list_$tag.log 2>&1
tail -2 gpurun_out/ncu_list_$tag.log
