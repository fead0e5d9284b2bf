#!/bin/bash
# usage (GPU box): bash tools/profile_round.sh <tag>
#   1. plain bench run (must exit 0), 2. ncu launch list of the same command (time + DRAM bytes per launch),
#   3. ncu --set full captures with source: the fused stem + the fused residual block (one launch each), and a 52x52 strip 3x3 +
#      the 1x1 after it on conv_tc_kernel.
# then, on the build box: python tools/summarize_profiles.py <tag> gpurun_out/layers_<x>.json
tag=${1:-x}
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --quick --no-parity"
$CMD > gpurun_out/plain_$tag.log 2> gpurun_out/plain_${tag}_err.log || { echo "plain run failed"; tail -5 gpurun_out/plain_${tag}_err.log; exit 1; }
tail -c 300 gpurun_out/plain_$tag.log
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 2000 --csv \
    --log-file gpurun_out/launches_$tag.csv $CMD > gpurun_out/ncu_list_$tag.log 2>&1
tail -2 gpurun_out/ncu_list_$tag.log | cut -c1-200
# (ncu matches the base name: template arguments cannot be selected; conv_tc launches per pass: 69, the 6th is conv12 = strip 3x3 at 52x52)
ncu --set full --clock-control none --import-source on -k "regex:conv_(stem|block)_kernel" -s 8 -c 2 \
    -o gpurun_out/prof_stemblock_$tag $CMD > gpurun_out/ncu_full_sb_$tag.log 2>&1
tail -2 gpurun_out/ncu_full_sb_$tag.log | cut -c1-200
ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s 212 -c 2 \
    -o gpurun_out/prof_convtc_$tag $CMD > gpurun_out/ncu_full_$tag.log 2>&1
tail -2 gpurun_out/ncu_full_$tag.log | cut -c1-200
