#!/bin/bash
# usage (GPU box): bash tools/profile_round.sh <tag>
#   1. plain bench run (must exit 0), 2. ncu launch list of the same command (time + DRAM bytes per launch),
#   3. ncu --set full capture of three launches of the CTA-pair conv kernel (the hot 3x3 layers) with source.
tag=${1:-x}
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --quick"
$CMD > gpurun_out/plain_$tag.log 2> gpurun_out/plain_${tag}_err.log || { echo "plain run failed"; tail -5 gpurun_out/plain_${tag}_err.log; exit 1; }
tail -c 600 gpurun_out/plain_$tag.log
# warm-up: 3 bench warm-ups + graph capture; skip everything before the two timed steps (77 launches each)
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 2000 --csv \
    --log-file gpurun_out/launches_$tag.csv $CMD > gpurun_out/ncu_list_$tag.log 2>&1
tail -2 gpurun_out/ncu_list_$tag.log
ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernelILi256ELb1 -s 60 -c 3 \
    -o gpurun_out/prof_convpair_$tag $CMD > gpurun_out/ncu_full_$tag.log 2>&1
tail -2 gpurun_out/ncu_full_$tag.log
