#!/bin/bash
# usage (GPU box): bash tools/profile_full.sh <tag>  -- plain bench run (must exit 0), then ONE ncu --set full capture of four
# consecutive conv_tc launches of the 52x52 stage (3x3 layers on the CTA-pair kernel + 1x1 layers), with source
tag=${1:-x}
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --quick"
$CMD > gpurun_out/plain2_$tag.log 2> gpurun_out/plain2_${tag}_err.log || { echo "plain run failed"; tail -5 gpurun_out/plain2_${tag}_err.log; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s 236 -c 4 \
    -o gpurun_out/prof_convtc_$tag $CMD > gpurun_out/ncu_full_$tag.log 2>&1
tail -2 gpurun_out/ncu_full_$tag.log
