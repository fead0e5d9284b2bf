#!/bin/bash
# usage (GPU box): bash tools/profile_halo.sh <tag> -- plain bench run, then ONE ncu --set full capture of the two
# conv_halo_kernel<64,1> launches (conv7, conv9) of a timed forward, with source
tag=${1:-x}
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --quick"
$CMD > gpurun_out/plain3_$tag.log 2> gpurun_out/plain3_${tag}_err.log || { echo "plain run failed"; tail -5 gpurun_out/plain3_${tag}_err.log; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:conv_halo_kernel -s 14 -c 2 \
    -o gpurun_out/prof_halo_$tag $CMD > gpurun_out/ncu_halo_$tag.log 2>&1
tail -2 gpurun_out/ncu_halo_$tag.log
