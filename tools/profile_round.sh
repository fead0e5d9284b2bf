#!/bin/bash
# usage (GPU box): bash tools/profile_round.sh <tag>
#   1. plain bench run (must exit 0), 2. ncu launch list of the same command (time + DRAM bytes per launch),
#   3. ncu --set full capture of four consecutive conv_tc launches (hot 3x3 layers on the CTA-pair kernel + a 1x1) with source.
# then, on the build box: python tools/summarize_profiles.py <tag> gpurun_out/layers_<x>.json
tag=${1:-x}
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --quick --no-parity"
$CMD > gpurun_out/plain_$tag.log 2> gpurun_out/plain_${tag}_err.log || { echo "plain run failed"; tail -5 gpurun_out/plain_${tag}_err.log; exit 1; }
tail -c 600 gpurun_out/plain_$tag.log
# warm-up: 3 bench warm-ups + graph capture; skip everything before the two timed steps (77 launches each)
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 2000 --csv \
    --log-file gpurun_out/launches_$tag.csv $CMD > gpurun_out/ncu_list_$tag.log 2>&1
tail -2 gpurun_out/ncu_list_$tag.log
# (ncu matches the base name: template arguments cannot be selected; -s lands on 3x3 / 1x1 layers of the 52x52 stage)
ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s 236 -c 4 \
    -o gpurun_out/prof_convtc_$tag $CMD > gpurun_out/ncu_full_$tag.log 2>&1
tail -2 gpurun_out/ncu_full_$tag.log
